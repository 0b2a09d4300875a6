"""In-tree build of the CUDA library (sm_100a only) with plain nvcc.

    python -m turbo_decoder_cuda_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    turbo_decoder_cuda_b200/lib/libtdb200.so         C ABI (include/tdb200.h) + kernels
    turbo_decoder_cuda_b200/lib/libtdb200_compat.so  reference-signature C++ wrappers (compat/)
    turbo_decoder_cuda_b200/lib/tdb200_burst         C++ multi-GPU caller of the C ABI (compat/tdb200_burst.cpp)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
COMPAT = os.path.join(PKG, "compat")
LIB = os.path.join(PKG, "lib")
OBJ = os.path.join(LIB, "obj")
INCLUDE = os.path.join(ROOT, "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]

# (source, extra flags)
KERNEL_TUS = [
    ("tdb200_api.cu", []),
    ("tdb200_ref64.cu", ["-fmad=false"]),
    ("tdb200_fast.cu", []),
    # the throughput kernel: one translation unit per channel-LLR type, compiled in parallel
    ("tdb200_fast_inst_f32.cu", ["-Xptxas", "-v"]),
    ("tdb200_fast_inst_f64.cu", ["-Xptxas", "-v"]),
    ("tdb200_fast_inst_s8.cu", ["-Xptxas", "-v"]),
    ("tdb200_fast_inst_f16.cu", ["-Xptxas", "-v"]),
    ("tdb200_fast_inst_crc_f32.cu", []),
    ("tdb200_fast_inst_crc_f64.cu", []),
    ("tdb200_fast_inst_crc_s8.cu", []),
    ("tdb200_fast_inst_crc_f16.cu", []),
    ("tdb200_fast_inst_lm_f32.cu", ["-Xptxas", "-v"]),
    ("tdb200_fast_inst_lm_f64.cu", []),
    ("tdb200_fast_inst_lm_s8.cu", []),
    ("tdb200_fast_inst_lm_f16.cu", []),
    ("tdb200_fast_inst_sym1.cu", []),
    ("tdb200_fast_inst_sym2.cu", []),
    ("tdb200_fast_inst_lm_sym1.cu", []),
    ("tdb200_fast_inst_lm_sym2.cu", []),
    ("tdb200_f32.cu", ["-fmad=false", "-Xptxas", "-v"]),
    ("tdb200_encode.cu", []),
    ("tdb200_modem.cu", ["-fmad=false"]),
    ("tdb200_ratematch.cu", ["-fmad=false"]),
    ("tdb200_crc.cu", []),
    ("tdb200_ubench.cu", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    log = []
    jobs = []
    for src, extra in KERNEL_TUS:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((src, [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]))

    def _compile(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for src, r in pool.map(_compile, jobs):
                log.append(r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    so = os.path.join(LIB, "libtdb200.so")
    if force or _stale(so, objs):
        # the shared CUDA runtime (found through the rpath, or already mapped by the host process): the shipped
        # artefacts then carry only the runtime symbols they import, not a private copy of the whole runtime
        cudart_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64")
        cmd = [nvcc] + ARCH + ["-shared", "-o", so] + objs + ["-cudart", "shared", "-Xlinker", "-rpath", "-Xlinker", cudart_dir]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    compat_src = os.path.join(COMPAT, "ittc_compat.cpp")
    compat_so = os.path.join(LIB, "libtdb200_compat.so")
    if os.path.exists(compat_src) and (force or _stale(compat_so, [compat_src, so] + hdrs)):
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", INCLUDE, compat_src, "-o", compat_so,
               "-L", LIB, "-ltdb200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("compat build failed:\n%s\n%s" % (r.stdout, r.stderr))
    burst_src = os.path.join(COMPAT, "tdb200_burst.cpp")
    burst_exe = os.path.join(LIB, "tdb200_burst")
    if os.path.exists(burst_src) and (force or _stale(burst_exe, [burst_src, so] + hdrs)):
        # a plain C++ caller of the C ABI (one host thread per GPU); nvcc only to find the CUDA runtime
        cmd = [nvcc, "-O2", "-std=c++17", "-I", INCLUDE, burst_src, "-o", burst_exe, "-L", LIB, "-ltdb200", "-cudart", "shared",
               "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-Xlinker", "-rpath", "-Xlinker",
               os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64"), "-Xcompiler", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("tdb200_burst build failed:\n%s\n%s" % (r.stdout, r.stderr))
    # measurement aid: the box's pinned host-to-device ceiling with N concurrent streams (tools/h2d_control.cpp)
    ctl_src = os.path.join(ROOT, "tools", "h2d_control.cpp")
    ctl_exe = os.path.join(LIB, "h2d_control")
    if os.path.exists(ctl_src) and (force or _stale(ctl_exe, [ctl_src])):
        cmd = [nvcc, "-O2", "-std=c++17", "-x", "cu", ctl_src, "-o", ctl_exe, "-cudart", "shared", "-Xlinker", "-rpath", "-Xlinker",
               os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64"), "-Xcompiler", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("h2d_control build failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        sys.stderr.write("".join(log))
    if any(log):
        with open(os.path.join(LIB, "ptxas.log"), "w") as f:
            f.write("".join(log))
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
