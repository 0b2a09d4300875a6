"""CPU tests of the host side: the C-ABI library loads and exports everything include/tdb200.h
declares (no compute without a GPU), the synthetic-traffic generator equals the oracle encoder,
the fixed-point specification model relates to the reference-derived max-log oracle, and the
codeblock sharding used by bench.py --gpus N works under a 2-rank gloo group."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle_lib import ALGO_MAXLOG, FxParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from turbo_decoder_cuda_b200 import build
    build.build()
    from turbo_decoder_cuda_b200 import load_library
    return load_library()


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "tdb200.h")).read()
    declared = set(re.findall(r"\b(tdb200_[a-z0-9_]+)\s*\(", hdr))
    assert {"tdb200_create", "tdb200_destroy", "tdb200_decode_batch", "tdb200_siso_batch",
            "tdb200_default_config", "tdb200_lte_qpp_params", "tdb200_get_plan",
            "tdb200_last_error", "tdb200_status_string"} <= declared
    for name in declared:
        assert hasattr(lib, name), "libtdb200.so does not export %s" % name


def test_compat_exports_reference_signatures(lib):
    """compat/ re-exports the reference's C++ entry points (mangled exactly as g++ mangles ITTC/main.h)."""
    so = os.path.join(ROOT, "turbo_decoder_cuda_b200", "lib", "libtdb200_compat.so")
    if not os.path.exists(so):
        pytest.skip("compat library not built")
    c = ctypes.CDLL(so)
    for sym in ("_Z13TurboDecodingPdPii", "_Z15Log_MAP_decoderPdS_iS_i", "_Z15TurboCodingInitv", "_Z18TurboCodingReleasev",
                "_Z10rate_matchPiiS_i", "_Z13de_rate_matchPdS_ii"):   # declared-only in the reference (main.h:23-24)
        assert hasattr(c, sym), sym


def test_lte_table_matches_oracle(lib, oracle):
    from turbo_decoder_cuda_b200 import TdbError, lte_qpp_params
    for K in oracle.lte_sizes():
        assert lte_qpp_params(K) == oracle.lte_params(K)
    with pytest.raises(TdbError):
        lte_qpp_params(6145)


def test_no_cpu_fallback(lib):
    """Without a CUDA device creation fails loudly (status NO_DEVICE); with one this test is moot."""
    import torch
    from turbo_decoder_cuda_b200 import TdbError, TurboDecoder
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(TdbError) as e:
        TurboDecoder(6144)
    assert e.value.status == 3 and "no CPU path" in str(e.value)


def test_synth_matches_oracle_encoder(oracle):
    import torch
    from turbo_decoder_cuda_b200 import synth
    for K in (40, 1008, 6144):
        pi = oracle.qpp(K)
        assert np.array_equal(synth.qpp_permutation(K).numpy(), pi)
        bits, llr = synth.make_batch(K, 3, 1.0, seed=5)
        coded = synth.turbo_encode(bits, torch.from_numpy(pi).long())
        for c in range(3):
            assert np.array_equal(coded[c].numpy(), oracle.encode(bits[c].numpy().astype(np.int32), pi).astype(np.uint8))
        assert abs(synth.sigma_from_ebn0(1.0, K) - oracle.sigma(1.0, K)) < 1e-12
        assert llr.shape == (3, 3 * K + 12) and llr.dtype == torch.float32
    # LLR sign convention: positive = bit 1; at high SNR the hard slice of the systematic LLRs is the data
    bits, llr = synth.make_batch(512, 2, 20.0, seed=1)
    assert np.array_equal((llr[:, 0:3 * 512:3] > 0).numpy().astype(np.uint8), bits.numpy())


def _fx(K, n_iter, L, G, q2=3, F=3):
    return FxParams(K=K, n_iter=n_iter, sub_len=L, warmup=G, frac_bits=F, llr_clip=127,
                    ext_clip=(1 << (F + 6)) - 1, ext_scale_q2=q2, early_term=0)


def test_fixed_point_model_is_exact_maxlog_when_unsegmented(oracle):
    """One sub-block, no scaling: the int model must reproduce the fp64 max-log oracle (reference
    Log_MAP_decoder with max* -> max) exactly on LLRs that are multiples of 1/8 -- every add/max is
    exact in both.  This ties the GPU specification to the reference-derived oracle."""
    K, n_iter = 512, 2
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, 4, 0.0, seed=17)
    q = np.clip(np.rint(llr * 8), -127, 127) / 8.0
    for c in range(4):
        bits, le, it, ovf = oracle.fx_decode(q[c].astype(np.float32), pi, _fx(K, n_iter, K, 0, q2=4), want_le=True)
        ob, _, o2, ole = oracle.decode(q[c], pi, n_iter, algo=ALGO_MAXLOG, want_llr=True)
        assert ovf == 0 and it == n_iter
        assert np.abs(ole[:K]).max() < 127.0           # no clamp active, else the comparison is void
        assert np.array_equal(bits, ob[-1])
        assert np.array_equal(le[pi], np.rint(ole[:K] * 8).astype(np.int32))


def test_fixed_point_model_subblocks_and_guard(oracle):
    """Segmentation only perturbs boundary metrics: at 1.5 dB every geometry decodes the block, and a
    guard of 16 steps restores what pure next-iteration initialisation loses in the waterfall."""
    K = 6144
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, 2, 1.5, seed=3)
    for L, G in ((6144, 0), (48, 16), (48, 0), (96, 32), (24, 24)):
        for c in range(2):
            b, _, _, ovf = oracle.fx_decode(llr[c].astype(np.float32), pi, _fx(K, 8, L, G))
            assert ovf == 0 and np.array_equal(b, bits[c]), (L, G)
    # L must be a multiple of 8 dividing K: the model rejects any other plan
    assert oracle.fx_decode(llr[0].astype(np.float32), pi, _fx(K, 8, 50, 0))[2] == -1


def test_logmap_fixed_point_model_tracks_the_reference_decoder(oracle):
    """The integer specification of TDB200_ALGO_LOGMAP_S16 (oracle/turbo_oracle_fx.c, logmap = 1: max* with the linear
    correction on shared quarter-differences) against the reference's fp64 Log-MAP (the restatement of
    ITTC/log_map.cpp:898-1047 with E_algorithm :779-801) on the same frames in the waterfall (K = 6144, 0.4 dB, where the
    max-log decoder of the same arithmetic fails most frames): no int16 overflow, nearly the same set of frames
    decoded, and far ahead of max-log.  The statistically meaningful version of this statement (16 384 frames per
    point, per iteration) is profiles/r02_bler_paired_*.json, made on the GPU by tools/bler_paired.py."""
    from concurrent.futures import ThreadPoolExecutor
    K, n, eb = 6144, 24, 0.4
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n, eb, seed=31337)
    llr32 = llr.astype(np.float32)
    lm = FxParams(K=K, n_iter=8, sub_len=48, warmup=24, frac_bits=4, llr_clip=127, ext_clip=511, ext_scale_q2=4, logmap=1, lm_upper_off=1)
    ml = FxParams(K=K, n_iter=8, sub_len=48, warmup=16, frac_bits=3, llr_clip=127, ext_clip=511, ext_scale_q2=3)
    with ThreadPoolExecutor(8) as pool:
        ref = list(pool.map(lambda c: (oracle.decode(llr[c], pi, 8)[-1] != bits[c]).any(), range(n)))
        a = list(pool.map(lambda c: oracle.fx_decode(llr32[c], pi, lm), range(n)))
        b = list(pool.map(lambda c: oracle.fx_decode(llr32[c], pi, ml), range(n)))
    assert all(r[3] == 0 for r in a), "int16 range exceeded"
    e_ref = int(sum(ref))
    e_lm = sum(int((r[0] != bits[c]).any()) for c, r in enumerate(a))
    e_ml = sum(int((r[0] != bits[c]).any()) for c, r in enumerate(b))
    assert e_lm <= e_ref + 2, (e_ref, e_lm, e_ml)
    assert e_ml >= e_lm + 2, (e_ref, e_lm, e_ml)
    # and the correction is what makes the difference: the same model without it is the max-log decoder
    lm0 = FxParams(K=K, n_iter=8, sub_len=48, warmup=24, frac_bits=4, llr_clip=127, ext_clip=511, ext_scale_q2=4, logmap=0)
    assert oracle.fx_decode(llr32[0], pi, lm0)[3] == 0


def test_early_termination_model(oracle):
    K = 1024
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, 3, 2.5, seed=9)
    p = _fx(K, 8, 32, 16)
    p.early_term = 1
    for c in range(3):
        b, _, it, _ = oracle.fx_decode(llr[c].astype(np.float32), pi, p)
        assert 2 <= it < 8 and np.array_equal(b, bits[c])


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from turbo_decoder_cuda_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 1000003
    lo, hi = shard.shard_range(n, world, rank)
    got = [None] * world
    dist.all_gather_object(got, (lo, hi))
    t = shard.max_over_ranks(float(rank + 1) * 1.5, dist if world > 1 else None)
    total = shard.sum_over_ranks(hi - lo, dist if world > 1 else None)
    q.put((rank, got, t, total))
    dist.destroy_process_group()


def test_codeblock_sharding_two_ranks_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    for rank, got, t, total in res:
        assert got[0] == (0, 500002) and got[1] == (500002, 1000003)   # disjoint, covering, balanced
        assert t == 3.0 and total == 1000003


def test_shard_range_properties():
    from turbo_decoder_cuda_b200 import shard
    for n in (0, 1, 7, 4096, 1000000):
        for w in (1, 2, 4, 8):
            r = [shard.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_reference_main_links_against_compat(lib, tmp_path):
    """Link-level drop-in: the reference's own main.cpp + modanddem.cpp + the
    encoder half of log_map.cpp link against libtdb200_compat with the two decode symbols taken from
    us.  Needs the reference tree, so it runs in the dev container only (no GPU needed to LINK)."""
    ref = "/root/reference/ITTC"
    if not os.path.exists(os.path.join(ref, "main.cpp")):
        pytest.skip("reference tree not present")
    import subprocess
    libdir = os.path.join(ROOT, "turbo_decoder_cuda_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libtdb200_compat.so")):
        pytest.skip("compat library not built")
    exe = str(tmp_path / "turbo_sim")
    objs = []
    for src, extra in (("main.cpp", []), ("modanddem.cpp", []),
                       ("log_map.cpp", ["-DTurboDecoding=ittc_cpu_TurboDecoding", "-DLog_MAP_decoder=ittc_cpu_Log_MAP_decoder"])):
        o = str(tmp_path / (src + ".o"))
        subprocess.check_call(["g++", "-O1", "-w", "-c", os.path.join(ref, src), "-o", o] + extra)
        objs.append(o)
    subprocess.check_call(["g++", "-o", exe] + objs + ["-L", libdir, "-ltdb200_compat", "-ltdb200", "-Wl,-rpath," + libdir])
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    assert "_Z13TurboDecodingPdPii" in syms, "main.cpp's TurboDecoding call must resolve to the compat library"


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs without a GPU (it times the reference's CPU decoder) and prints one
    JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "Gbit/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["ber"] == 0.0   # two codeblocks at 1.0 dB decode cleanly


def test_segmentation_known_cases_and_abi(oracle, lib):
    import ctypes as C
    assert oracle.segmentation(6144) == dict(C=1, K_plus=6144, K_minus=0, C_plus=1, C_minus=0, F=0, L=0)
    # the largest LTE transport block (75376 bits + CRC24A): 13 blocks of 5824
    assert oracle.segmentation(75376 + 24) == dict(C=13, K_plus=5824, K_minus=5760, C_plus=13, C_minus=0, F=0, L=24)
    assert oracle.segmentation(6145) == dict(C=2, K_plus=3136, K_minus=3072, C_plus=1, C_minus=1, F=15, L=24)

    class Seg(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("C", "K_plus", "K_minus", "C_plus", "C_minus", "F", "L")]
    lib.tdb200_segmentation.argtypes = [C.c_int, C.POINTER(Seg)]
    sizes = oracle.lte_sizes()
    for B in list(range(25, 400)) + list(range(6000, 6400)) + [12000, 12257, 30000, 61664, 75400, 100000, 150000]:
        s = Seg()
        assert lib.tdb200_segmentation(B, C.byref(s)) == 0      # host arithmetic: no device needed
        got = {n: getattr(s, n) for n, _ in Seg._fields_}
        assert got == oracle.segmentation(B), B
        Bp = B + got["C"] * got["L"]
        assert got["C_plus"] * got["K_plus"] + got["C_minus"] * got["K_minus"] == Bp + got["F"]
        assert got["K_plus"] in sizes and (got["C_minus"] == 0 or got["K_minus"] in sizes) and 0 <= got["F"] < 64 * got["C"]
    assert lib.tdb200_segmentation(0, C.byref(Seg())) != 0


def test_header_is_plain_c99(lib, tmp_path):
    """include/tdb200.h is the C ABI: it must compile as strict C99 (no C++-isms, no torch / CUDA types) and a C program
    linked against the library can call the host-only entry points without a GPU."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "tdb200.h"\n'
                   'int main(void) { tdb200_seg_info s; tdb200_config c; int f1, f2;\n'
                   '  if (tdb200_segmentation(75400, &s) != TDB200_OK || s.C != 13 || s.K_plus != 5824) return 1;\n'
                   '  if (tdb200_default_config(&c, 6144) != TDB200_OK || c.n_iter != 8) return 2;\n'
                   '  if (tdb200_lte_qpp_params(6144, &f1, &f2) != TDB200_OK || f1 != 263 || f2 != 480) return 3;\n'
                   '  return tdb200_lte_qpp_params(6145, &f1, &f2) == TDB200_OK ? 4 : 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "turbo_decoder_cuda_b200", "lib")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-ltdb200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_hashed_max_star_table_equals_the_linear_scan(oracle):
    """csrc/tdb200_ref64.cu looks the max* correction up through a table hashed by the exponent and the top three
    mantissa bits of |x-y| (entry = the one breakpoint inside that interval and the value below it; the value from it on
    is the next entry's).  Restated here in numpy and compared with the oracle's scan of the reference's 16-entry table
    (E_algorithm, ITTC/log_map.cpp:779-801) on the breakpoints, their neighbours in the last place, the interval
    boundaries and a dense sweep -- the two must agree for every double."""
    idx = np.array([0.0, 0.08824, 0.19587, 0.31026, 0.43275, 0.56508, 0.70963, 0.86972,
                    1.0502, 1.2587, 1.5078, 1.8212, 2.2522, 2.9706, 3.6764, 4.3758])
    val = np.array([0.69315, 0.65, 0.6, 0.55, 0.5, 0.45, 0.4, 0.35, 0.3, 0.25, 0.2, 0.15, 0.1, 0.05, 0.025, 0.0])
    n_ent = 58
    bp = np.full(n_ent + 2, 1e300)
    below = np.zeros(n_ent + 2)
    below[0] = val[0]
    for e in range(1, n_ent - 1):
        b, s = (e - 1) >> 3, (e - 1) & 7
        base = 2.0 ** (b - 4)
        lo = base * (1.0 + 0.125 * s)
        hi = lo + base * 0.125
        k = int(np.searchsorted(idx, lo, side="right")) - 1
        below[e] = val[k]
        inside = [t for t in range(1, 16) if lo <= idx[t] < hi]
        assert len(inside) <= 1, "two breakpoints in one interval"      # what the hash relies on
        assert all(idx[t] != lo for t in inside), "breakpoint on an interval boundary"
        if inside:
            bp[e] = idx[inside[0]]

    def hashed(d):
        hi_word = (np.abs(d).view(np.uint64) >> np.uint64(32)).astype(np.int64)
        e = np.clip((hi_word >> 17) - ((1023 - 4) * 8 - 1), 0, n_ent - 1)
        return np.where(np.abs(d) < bp[e], below[e], below[e + 1])

    pts = [np.linspace(0.0, 9.0, 200001), idx, np.nextafter(idx, 10.0), np.nextafter(idx, -10.0)]
    edges = np.array([2.0 ** (b - 4) * (1 + s / 8.0) for b in range(8) for s in range(8)])
    pts += [edges, np.nextafter(edges, 10.0), np.nextafter(edges, -10.0), np.array([1e-300, 0.06249999, 8.0, 1e3, 2e20])]
    d = np.abs(np.concatenate(pts))
    want = np.array([oracle.max_star(0.0, float(x)) - max(0.0, float(x)) for x in d])   # max*(0, d) = d + corr(d)
    got = hashed(d)
    # max*(0,d) - d carries the rounding of the sum for large d; compare where the scan is exact, and the sum otherwise
    assert np.array_equal(d + got, np.array([oracle.max_star(0.0, float(x)) for x in d]))
    small = d < 1.0
    assert np.allclose(got[small], want[small], rtol=0, atol=1e-15)
