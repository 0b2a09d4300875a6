"""ctypes bindings for the parity checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`Oracle`  -> oracle/_build/libtdoracle.so  (C restatement, oracle/turbo_oracle.c)
`RefLib`  -> oracle/_ref/libittc_ref.so    (the reference's own ITTC sources compiled in place,
                                           oracle/ref_harness.cpp); optional.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libtdoracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libittc_ref.so")

ALGO_LOGMAP_LUT = 1
ALGO_MAXLOG = 2

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    """Compile oracle/'s C restatement (and oracle/_ref when /root/reference is present)."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in sorted(os.listdir(ORACLE_DIR)) if f.endswith((".c", ".h"))]
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=subprocess.DEVNULL)
    ref_src = "/root/reference/ITTC/log_map.cpp"
    harness = os.path.join(ORACLE_DIR, "ref_harness.cpp")
    if os.path.exists(ref_src) and (force or not os.path.exists(REF_SO)
                                    or os.path.getmtime(harness) > os.path.getmtime(REF_SO)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "ref"], stdout=subprocess.DEVNULL)


class FxParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "K", "n_iter", "sub_len", "warmup", "frac_bits", "llr_clip", "ext_clip",
        "ext_scale_q2", "early_term", "et_threshold", "crc_poly",
        "logmap", "lm_t4", "lm_upper", "lm_tt", "lm_tc", "lm_warm_maxlog", "lm_upper_off", "lm_t4_lam", "lm_exact")]


class F32Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("K", "n_iter", "sub_len", "warmup", "logmap", "early_term")] + \
               [(n, C.c_float) for n in ("ext_scale", "ext_clamp", "et_threshold")]


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.tdo_gen_trellis.argtypes = [_ip, _ip, _ip, _ip]
        L.tdo_qpp_index.argtypes = [C.c_int, C.c_int, C.c_int, _ip]
        L.tdo_lte_qpp_params.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tdo_lte_size_at.argtypes = [C.c_int]
        L.tdo_turbo_encode.argtypes = [_ip, C.c_int, _ip, _ip]
        L.tdo_channel_llr.argtypes = [_ip, C.c_int, C.c_double, C.c_ulonglong, C.c_ulonglong, _dp]
        L.tdo_sigma_from_ebn0.argtypes = [C.c_double, C.c_int]
        L.tdo_sigma_from_ebn0.restype = C.c_double
        L.tdo_max_star_lut.argtypes = [C.c_double, C.c_double]
        L.tdo_max_star_lut.restype = C.c_double
        L.tdo_siso.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double]
        L.tdo_turbo_decode.argtypes = [_dp, C.c_int, _ip, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tdo_turbo_decode_batch.argtypes = [_dp, C.c_int, C.c_int, _ip, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int]
        L.tdo_turbo_decode_batch.restype = C.c_double
        L.tdo_fx_decode.argtypes = [_fp, _ip, C.POINTER(FxParams), _ip, C.c_void_p, C.c_void_p]
        L.tdo_fx_decode.restype = C.c_int
        L.tdo_f32_decode.argtypes = [_fp, _ip, C.POINTER(F32Params), _ip, C.c_void_p, C.c_void_p]
        L.tdo_f32_decode.restype = C.c_int
        L.tdo_rm_geometry.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tdo_rm_circular_buffer.argtypes = [C.c_int, _ip]
        L.tdo_rm_k0.argtypes = [C.c_int, C.c_int, C.c_int]
        L.tdo_rm_selection.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip]
        L.tdo_rm_selection_f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip]
        L.tdo_rm_circular_buffer_f.argtypes = [C.c_int, C.c_int, _ip]
        L.tdo_rate_match_f.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip]
        L.tdo_rate_dematch_f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.tdo_rate_match.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int, _ip]
        L.tdo_rate_dematch.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.tdo_rate_dematch_f32.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        L.tdo_crc24.argtypes = [np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), C.c_int, C.c_uint]
        L.tdo_crc24.restype = C.c_uint
        L.tdo_segmentation.argtypes = [C.c_int, _ip]
        L.tdo_modulate.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp]
        L.tdo_demap_f64.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, _dp]
        L.tdo_demap_f32.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_float, _fp]
        L.tdo_quant_s8.argtypes = [_fp, C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")]

    # ---- TS 36.212 rate matching (turbo_oracle_rm.c)
    def rm_geometry(self, K):
        R, Kpi, ND = C.c_int(), C.c_int(), C.c_int()
        Kw = self.lib.tdo_rm_geometry(K, C.byref(R), C.byref(Kpi), C.byref(ND))
        return {"R": R.value, "Kpi": Kpi.value, "ND": ND.value, "Kw": Kw}

    def rm_circular_buffer(self, K):
        w = np.zeros(self.rm_geometry(K)["Kw"], np.int32)
        self.lib.tdo_rm_circular_buffer(K, w)
        return w

    def rm_k0(self, K, rv, Ncb=0):
        return self.lib.tdo_rm_k0(K, rv, Ncb)

    def rm_selection(self, K, E, rv, Ncb=0):
        sel = np.zeros(max(E, 1), np.int32)
        if self.lib.tdo_rm_selection(K, E, rv, Ncb, sel):
            raise ValueError("rate matching: empty circular buffer")
        return sel[:E]

    def rate_match(self, coded, K, E, rv, Ncb=0):
        out = np.zeros(max(E, 1), np.int32)
        if self.lib.tdo_rate_match(np.ascontiguousarray(coded, np.int32), K, E, rv, Ncb, out):
            raise ValueError("rate matching: empty circular buffer")
        return out[:E]

    def rate_dematch(self, e_llr, K, rv, Ncb=0, into=None):
        """double accumulation for float64 input, float accumulation otherwise (as on the device)."""
        f32 = np.asarray(e_llr).dtype != np.float64
        dt = np.float32 if f32 else np.float64
        e = np.ascontiguousarray(e_llr, dt)
        llr = np.zeros(3 * K + 12, dt) if into is None else np.array(into, dt)
        fn = self.lib.tdo_rate_dematch_f32 if f32 else self.lib.tdo_rate_dematch
        if fn(e if e.size else np.zeros(1, dt), K, e.size, rv, Ncb, 0 if into is None else 1, llr):
            raise ValueError("rate matching: empty circular buffer")
        return llr

    # ... with F filler bits at the head of the code block (<NULL> in d0 / d1)
    def rm_circular_buffer_f(self, K, F):
        w = np.zeros(self.rm_geometry(K)["Kw"], np.int32)
        self.lib.tdo_rm_circular_buffer_f(K, F, w)
        return w

    def rm_selection_f(self, K, E, rv, Ncb, F):
        sel = np.zeros(max(E, 1), np.int32)
        if self.lib.tdo_rm_selection_f(K, E, rv, Ncb, F, sel):
            raise ValueError("rate matching: empty circular buffer")
        return sel[:E]

    def rate_match_f(self, coded, K, E, rv, Ncb, F):
        out = np.zeros(max(E, 1), np.int32)
        if self.lib.tdo_rate_match_f(np.ascontiguousarray(coded, np.int32), K, E, rv, Ncb, F, out):
            raise ValueError("rate matching: empty circular buffer")
        return out[:E]

    def rate_dematch_f(self, e_llr, K, rv, Ncb, F, fill=-100.0, into=None):
        e = np.ascontiguousarray(e_llr, np.float64)
        llr = np.zeros(3 * K + 12, np.float64) if into is None else np.array(into, np.float64)
        if self.lib.tdo_rate_dematch_f(e if e.size else np.zeros(1), K, e.size, rv, Ncb, F, 0 if into is None else 1, fill, llr):
            raise ValueError("rate matching: empty circular buffer")
        return llr

    # ---- CRC24 / segmentation (turbo_oracle_crc.c)
    CRC24A, CRC24B = 0x864CFB, 0x800063

    def crc24(self, bits, poly):
        bits = np.ascontiguousarray(bits, np.uint8).ravel()
        return int(self.lib.tdo_crc24(bits if bits.size else np.zeros(1, np.uint8), bits.size, poly))

    def segmentation(self, B):
        out = np.zeros(7, np.int32)
        if self.lib.tdo_segmentation(B, out):
            raise ValueError("segmentation: B=%d" % B)
        return dict(zip(("C", "K_plus", "K_minus", "C_plus", "C_minus", "F", "L"), map(int, out)))

    # ---- mapper / soft demapper (turbo_oracle_mod.c)
    def modulate(self, bits, M):
        bits = np.ascontiguousarray(bits, np.int32).ravel()
        si, sq = np.zeros(bits.size // M), np.zeros(bits.size // M)
        if self.lib.tdo_modulate(bits, bits.size, M, si, sq):
            raise ValueError("modulate: M=%d, %d bits" % (M, bits.size))
        return si, sq

    def demap_f64(self, si, sq, M, kf):
        si, sq = np.ascontiguousarray(si, np.float64).ravel(), np.ascontiguousarray(sq, np.float64).ravel()
        out = np.zeros(si.size * M)
        if self.lib.tdo_demap_f64(si, sq, si.size, M, kf, out):
            raise ValueError("demap: M=%d" % M)
        return out

    def demap_f32(self, si, sq, M, kf):
        si, sq = np.ascontiguousarray(si, np.float32).ravel(), np.ascontiguousarray(sq, np.float32).ravel()
        out = np.zeros(si.size * M, np.float32)
        if self.lib.tdo_demap_f32(si, sq, si.size, M, kf, out):
            raise ValueError("demap: M=%d" % M)
        return out

    def quant_s8(self, llr, frac_bits=3, clip=127):
        llr = np.ascontiguousarray(llr, np.float32).ravel()
        out = np.zeros(llr.size, np.int8)
        self.lib.tdo_quant_s8(llr, llr.size, frac_bits, clip, out)
        return out

    # ---- constants
    def trellis(self):
        no, ns, lo, ls = (np.zeros(n, np.int32) for n in (32, 16, 32, 16))
        self.lib.tdo_gen_trellis(no, ns, lo, ls)
        return no.reshape(8, 4), ns.reshape(8, 2), lo.reshape(8, 4), ls.reshape(8, 2)

    def lte_sizes(self):
        return [self.lib.tdo_lte_size_at(i) for i in range(self.lib.tdo_lte_num_sizes())]

    def lte_params(self, K):
        f1, f2 = C.c_int(), C.c_int()
        if self.lib.tdo_lte_qpp_params(K, C.byref(f1), C.byref(f2)) != 0:
            raise ValueError(f"K={K} is not an LTE turbo block size")
        return f1.value, f2.value

    def qpp(self, K, f1=None, f2=None):
        if f1 is None:
            f1, f2 = self.lte_params(K)
        pi = np.zeros(K, np.int32)
        self.lib.tdo_qpp_index(K, f1, f2, pi)
        return pi

    # ---- test-vector generation
    def encode(self, bits, pi):
        K = len(bits)
        coded = np.zeros(3 * K + 12, np.int32)
        self.lib.tdo_turbo_encode(np.ascontiguousarray(bits, np.int32), K, pi, coded)
        return coded

    def sigma(self, ebn0_db, K):
        return self.lib.tdo_sigma_from_ebn0(ebn0_db, K)

    def channel(self, coded, sigma, seed, stream=0):
        llr = np.zeros(len(coded), np.float64)
        self.lib.tdo_channel_llr(np.ascontiguousarray(coded, np.int32), len(coded), sigma, seed, stream, llr)
        return llr

    def make_batch(self, K, n_cb, ebn0_db, seed, pi=None):
        """(bits[n_cb,K] int32, llr[n_cb,3K+12] float64) -- seeded, reproducible."""
        pi = self.qpp(K) if pi is None else pi
        rng = np.random.default_rng(seed)
        bits = rng.integers(0, 2, size=(n_cb, K), dtype=np.int32)
        sig = self.sigma(ebn0_db, K)
        llr = np.zeros((n_cb, 3 * K + 12), np.float64)
        for c in range(n_cb):
            llr[c] = self.channel(self.encode(bits[c], pi), sig, seed, c)
        return bits, llr

    # ---- decode
    def max_star(self, x, y):
        return self.lib.tdo_max_star_lut(x, y)

    def siso(self, recs, La, algo=ALGO_LOGMAP_LUT, terminated=1, tempmax_floor=math.nan):
        T = len(La)
        out = np.zeros(T, np.float64)
        self.lib.tdo_siso(np.ascontiguousarray(recs, np.float64), np.ascontiguousarray(La, np.float64),
                          terminated, out, T, algo, tempmax_floor)
        return out

    def decode(self, llr, pi, n_iter, algo=ALGO_LOGMAP_LUT, want_llr=False):
        K = len(pi)
        T = K + 3
        bits = np.zeros((n_iter, K), np.int32)
        l1 = np.zeros(T) if want_llr else None
        l2 = np.zeros(T) if want_llr else None
        le = np.zeros(T) if want_llr else None
        self.lib.tdo_turbo_decode(np.ascontiguousarray(llr, np.float64), K, pi, n_iter, algo,
                                  _opt(bits), _opt(l1), _opt(l2), _opt(le))
        return (bits, l1, l2, le) if want_llr else bits

    def decode_batch(self, llr, pi, n_iter, algo=ALGO_LOGMAP_LUT, n_threads=1):
        llr = np.ascontiguousarray(llr, np.float64)
        n_cb, K = llr.shape[0], len(pi)
        bits = np.zeros((n_cb, K), np.int32)
        secs = self.lib.tdo_turbo_decode_batch(llr, n_cb, K, pi, n_iter, algo, _opt(bits), n_threads)
        return bits, secs

    def fx_decode(self, llr_f32, pi, params, want_le=False):
        K = len(pi)
        bits = np.zeros(K, np.int32)
        le = np.zeros(K, np.int32) if want_le else None
        ovf = C.c_int(0)
        it = self.lib.tdo_fx_decode(np.ascontiguousarray(llr_f32, np.float32), pi, C.byref(params),
                                    bits, _opt(le), C.cast(C.byref(ovf), C.c_void_p))
        return bits, le, it, ovf.value


    def f32_decode(self, llr_f32, pi, params, want_soft=False):
        """fp32 windowed Log-MAP / max-log model: (bits, llr, le, iterations); soft values in SISO-2 order."""
        K = len(pi)
        bits = np.zeros(K, np.int32)
        llr = np.zeros(K, np.float32) if want_soft else None
        le = np.zeros(K, np.float32) if want_soft else None
        it = self.lib.tdo_f32_decode(np.ascontiguousarray(llr_f32, np.float32), pi, C.byref(params), bits, _opt(llr), _opt(le))
        return bits, llr, le, it


class RefLib:
    """The reference's own code (oracle/_ref).  available() is False where it was never built."""

    @staticmethod
    def available():
        build_oracle()
        return os.path.exists(REF_SO)

    def __init__(self, K, f1, f2):
        build_oracle()
        L = self.lib = C.CDLL(REF_SO)
        self.K = K
        L.ref_init.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_get_qpp.argtypes = [_ip]
        L.ref_get_trellis.argtypes = [_ip, _ip, _ip, _ip]
        L.ref_encode.argtypes = [_ip, _ip]
        L.ref_max_star.argtypes = [C.c_double, C.c_double]
        L.ref_max_star.restype = C.c_double
        L.ref_channel.argtypes = [_ip, C.c_double, C.c_uint, _dp]
        L.ref_turbo_decoding.argtypes = [_dp, _ip]
        L.ref_siso.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int]
        L.ref_decode_iters.argtypes = [_dp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_decode_batch.argtypes = [_dp, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_decode_batch.restype = C.c_double
        L.ref_module.argtypes = [_ip, _dp, _dp, C.c_int, C.c_int]
        L.ref_demodule.argtypes = [_dp, _dp, C.c_int, _dp, C.c_double, C.c_int]
        L.ref_init(K, f1, f2)

    def qpp(self):
        pi = np.zeros(self.K, np.int32)
        self.lib.ref_get_qpp(pi)
        return pi

    def trellis(self):
        no, ns, lo, ls = (np.zeros(n, np.int32) for n in (32, 16, 32, 16))
        self.lib.ref_get_trellis(no, ns, lo, ls)
        return no.reshape(8, 4), ns.reshape(8, 2), lo.reshape(8, 4), ls.reshape(8, 2)

    def encode(self, bits):
        coded = np.zeros(3 * self.K + 12, np.int32)
        self.lib.ref_encode(np.ascontiguousarray(bits, np.int32), coded)
        return coded

    def max_star(self, x, y):
        return self.lib.ref_max_star(x, y)

    def channel(self, coded, sigma, seed):
        llr = np.zeros(3 * self.K + 12, np.float64)
        self.lib.ref_channel(np.ascontiguousarray(coded, np.int32), sigma, seed, llr)
        return llr

    def module(self, bits, M):
        """The reference's module(): bits -> (I, Q) constellation points."""
        bits = np.ascontiguousarray(bits, np.int32).ravel()
        si, sq = np.zeros(bits.size // M), np.zeros(bits.size // M)
        self.lib.ref_module(bits, si, sq, bits.size, M)
        return si, sq

    def demodule(self, si, sq, M, kf):
        """The reference's demodule(): max-log soft demapper."""
        si, sq = np.array(si, np.float64).ravel(), np.array(sq, np.float64).ravel()
        out = np.zeros(si.size * M)
        self.lib.ref_demodule(si, sq, si.size, out, kf, M)
        return out

    def turbo_decoding(self, llr):
        """The reference's TurboDecoding(): 15 iterations; returns bits[15,K]."""
        buf = np.array(llr, np.float64, copy=True)
        out = np.zeros(15 * self.K, np.int32)
        self.lib.ref_turbo_decoding(buf, out)
        return out.reshape(15, self.K)

    def siso(self, recs, La, terminated=1):
        T = len(La)
        out = np.zeros(T, np.float64)
        self.lib.ref_siso(np.array(recs, np.float64), np.array(La, np.float64), terminated, out, T)
        return out

    def decode(self, llr, n_iter, want_llr=False):
        K, T = self.K, self.K + 3
        bits = np.zeros((n_iter, K), np.int32)
        l1 = np.zeros(T) if want_llr else None
        l2 = np.zeros(T) if want_llr else None
        le = np.zeros(T) if want_llr else None
        self.lib.ref_decode_iters(np.ascontiguousarray(llr, np.float64), n_iter,
                                  _opt(bits), _opt(l1), _opt(l2), _opt(le))
        return (bits, l1, l2, le) if want_llr else bits

    def decode_batch(self, llr, n_iter, n_threads=1):
        llr = np.ascontiguousarray(llr, np.float64)
        n_cb = llr.shape[0]
        bits = np.zeros((n_cb, self.K), np.int32)
        secs = self.lib.ref_decode_batch(llr, n_cb, n_iter, _opt(bits), n_threads)
        return bits, secs
