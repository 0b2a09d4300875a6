"""Python mirror of the C ABI in include/tdb200.h (ctypes; no compute happens in Python).

The reference's interface for this path is a pair of C++ free functions
(`TurboDecoding`, `Log_MAP_decoder`, ITTC/main.h:20, ITTC/log_map.cpp:898,1146) configured by
globals; `TurboDecoder` keeps their argument meaning (LLR layout 3K+12 in the reference's
multiplex order, hard decisions per iteration, natural order) for a batch of codeblocks.
Buffers may be numpy arrays (host) or torch CUDA tensors (device, zero-copy).

There is no CPU fallback: if the CUDA library is missing or no GPU is present, construction
raises.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libtdb200.so")

ALGO_LOGMAP_F64 = 0
ALGO_MAXLOG_S16 = 1
ALGO_LOGMAP_F32 = 2
ALGO_MAXLOG_F32 = 3
ALGO_LINLOGMAP_F32 = 4
ALGO_LOGMAP_S16 = 5
ALGO_NAMES = {"logmap_f64": 0, "maxlog_s16": 1, "logmap_f32": 2, "maxlog_f32": 3, "linlogmap_f32": 4, "logmap_s16": 5}

LLR_F64, LLR_F32, LLR_S8, LLR_F16 = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1


class TdbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("tdb200 status %d: %s" % (status, msg))
        self.status = status


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "K", "f1", "f2", "n_iter", "algo", "sub_block", "warmup", "early_term", "et_threshold", "ext_scale_q2",
        "frac_bits", "ext_clip", "device", "max_batch")]


class Outputs(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("bits_iters", C.c_void_p), ("iters_used", C.c_void_p),
                ("llr_siso1", C.c_void_p), ("llr_siso2", C.c_void_p), ("ext_siso2", C.c_void_p)]


class SegInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("C", "K_plus", "K_minus", "C_plus", "C_minus", "F", "L")]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "K", "f1", "f2", "n_iter", "algo", "sub_block", "n_sub_blocks", "warmup", "cb_per_cta",
        "threads_per_cta", "smem_bytes", "max_batch", "sm_count", "kernel_launches_last_call")]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m turbo_decoder_cuda_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.tdb200_last_error.restype = C.c_char_p
    L.tdb200_status_string.restype = C.c_char_p
    L.tdb200_status_string.argtypes = [C.c_int]
    L.tdb200_default_config.argtypes = [C.POINTER(Config), C.c_int]
    L.tdb200_lte_qpp_params.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tdb200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.tdb200_destroy.argtypes = [C.c_void_p]
    L.tdb200_destroy.restype = None
    L.tdb200_ubench_issue_rate.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.tdb200_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(Outputs), C.c_void_p]
    L.tdb200_siso_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int, C.c_int, C.c_void_p]
    L.tdb200_get_plan.argtypes = [C.c_void_p, C.POINTER(PlanInfo)]
    L.tdb200_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tdb200_channel_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.c_uint64, C.c_void_p]
    L.tdb200_modulate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]
    L.tdb200_awgn_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                    C.c_double, C.c_uint64, C.c_void_p]
    L.tdb200_demap_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.tdb200_decode_symbols_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_double, C.POINTER(Outputs), C.c_void_p]
    L.tdb200_modulate_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                       C.c_int, C.c_void_p]
    L.tdb200_demap_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.c_size_t, C.c_int, C.c_double, C.c_void_p]
    L.tdb200_set_filler_bits.argtypes = [C.c_void_p, C.c_int]
    L.tdb200_rate_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tdb200_rate_dematch_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p]
    L.tdb200_decode_rm_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(Outputs), C.c_void_p]
    L.tdb200_crc24_attach_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tdb200_crc24_check_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tdb200_segmentation.argtypes = [C.c_int, C.POINTER(SegInfo)]
    _lib = L
    return L


def _check(status):
    if status != 0:
        raise TdbError(status, load_library().tdb200_last_error().decode())


CRC24A, CRC24B = 0, 1


def segmentation(B):
    """tdb200_segmentation: code-block segmentation of a transport block of B bits (CRC24A included),
    TS 36.212 5.1.2 -> dict(C, K_plus, K_minus, C_plus, C_minus, F, L).  Host arithmetic only."""
    info = SegInfo()
    _check(load_library().tdb200_segmentation(int(B), C.byref(info)))
    return {n: getattr(info, n) for n, _ in SegInfo._fields_}


def lte_qpp_params(K):
    f1, f2 = C.c_int(), C.c_int()
    _check(load_library().tdb200_lte_qpp_params(K, C.byref(f1), C.byref(f2)))
    return f1.value, f2.value


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr_of(x):
    """(pointer, mem_space) of a numpy array or a torch tensor."""
    if x is None:
        return None, None
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    assert x.flags["C_CONTIGUOUS"]
    return x.ctypes.data, MEM_HOST


_LLR_TYPES = {"float64": LLR_F64, "float32": LLR_F32, "int8": LLR_S8, "float16": LLR_F16}


class TurboDecoder:
    """Batched iterative PCCC decoder handle (tdb200_create / tdb200_destroy)."""

    def __init__(self, K, n_iter=8, algo="maxlog_s16", f1=0, f2=0, sub_block=0, warmup=0,
                 early_term=False, ext_scale_q2=0, frac_bits=0, ext_clip=0, device=0, max_batch=0, et_threshold=0):
        L = load_library()
        cfg = Config()
        _check(L.tdb200_default_config(C.byref(cfg), K))
        cfg.f1, cfg.f2, cfg.n_iter = f1, f2, n_iter
        cfg.algo = ALGO_NAMES[algo] if isinstance(algo, str) else int(algo)
        # early_term: False/0 off, True/1 decisions + magnitude, "crc24b"/2 and "crc24a"/3 the CRC stopping rule
        early_term = {"crc24b": 2, "crc24a": 3}.get(early_term, early_term)
        cfg.sub_block, cfg.warmup, cfg.early_term = sub_block, warmup, int(early_term)
        cfg.ext_scale_q2, cfg.frac_bits, cfg.device, cfg.max_batch = ext_scale_q2, frac_bits, device, max_batch
        cfg.ext_clip, cfg.et_threshold = ext_clip, et_threshold
        h = C.c_void_p()
        _check(L.tdb200_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._L = L
        self.K, self.T, self.n_iter, self.algo, self.device = K, K + 3, n_iter, cfg.algo, device
        self.llr_len = 3 * K + 12

    def close(self):
        if getattr(self, "_h", None):
            self._L.tdb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def issue_rate(self, mix=2):
        """Measured issue rate of the add-compare-select instruction mix on this decoder's GPU (thread-ops/clk/SM)."""
        v = C.c_double(0.0)
        _check(self._L.tdb200_ubench_issue_rate(self.device, int(mix), C.byref(v)))
        return v.value

    def plan(self):
        info = PlanInfo()
        _check(self._L.tdb200_get_plan(self._h, C.byref(info)))
        return {n: getattr(info, n) for n, _ in PlanInfo._fields_}

    @property
    def float_dtype(self):
        return np.float64 if self.algo == ALGO_LOGMAP_F64 else np.float32

    def decode_raw(self, llr_ptr, llr_type, mem, n_cb, bits=None, bits_iters=None, iters_used=None,
                   llr_siso1=None, llr_siso2=None, ext_siso2=None, stream=0):
        """Thin call-through with raw pointers (used by bench.py's timed loops)."""
        o = Outputs(bits, bits_iters, iters_used, llr_siso1, llr_siso2, ext_siso2)
        _check(self._L.tdb200_decode_batch(self._h, llr_ptr, llr_type, mem, n_cb, C.byref(o), stream))

    def decode(self, llr, want=("bits",), stream=0):
        """Decode llr[n_cb, 3K+12] (numpy -> host path, torch.cuda tensor -> device path).

        `want` selects outputs among bits, bits_iters, iters_used, llr_siso1, llr_siso2, ext_siso2;
        returns a dict of arrays of the same kind (numpy / torch) as the input.
        """
        n_cb = int(llr.shape[0])
        assert int(llr.shape[1]) == self.llr_len, "llr must be [n_cb, 3K+12]"
        tname = str(llr.dtype).replace("torch.", "")
        if tname not in _LLR_TYPES:
            raise TypeError("llr dtype %s not supported" % tname)
        ptr, mem = _ptr_of(llr)
        K, T = self.K, self.T
        shapes = {"bits": ((n_cb, K), "uint8"), "bits_iters": ((n_cb, self.n_iter, K), "int32"),
                  "iters_used": ((n_cb,), "int32"),
                  "llr_siso1": ((n_cb, T), None), "llr_siso2": ((n_cb, T), None), "ext_siso2": ((n_cb, T), None)}
        outs = {}
        for name in want:
            shape, dt = shapes[name]
            if _is_torch(llr):
                import torch
                tdt = getattr(torch, dt) if dt else (torch.float64 if self.algo == ALGO_LOGMAP_F64 else torch.float32)
                outs[name] = torch.empty(shape, dtype=tdt, device=llr.device)
            else:
                outs[name] = np.empty(shape, dtype=dt or self.float_dtype)
        ptrs = {k: _ptr_of(v)[0] for k, v in outs.items()}
        self.decode_raw(ptr, _LLR_TYPES[tname], mem, n_cb, stream=stream, **ptrs)
        return outs

    def encode(self, bits, stream=0):
        """tdb200_encode_batch: bits [n_cb, K] uint8 -> coded [n_cb, 3K+12] uint8 (the TurboEnCoding replacement)."""
        n_cb = int(bits.shape[0])
        assert int(bits.shape[1]) == self.K and str(bits.dtype).endswith("uint8")
        bp, mem = _ptr_of(bits)
        if _is_torch(bits):
            import torch
            coded = torch.empty((n_cb, self.llr_len), dtype=torch.uint8, device=bits.device)
        else:
            coded = np.empty((n_cb, self.llr_len), dtype=np.uint8)
        _check(self._L.tdb200_encode_batch(self._h, bp, _ptr_of(coded)[0], mem, n_cb, stream))
        return coded

    def channel(self, coded, sigma, seed, dtype="float32", stream=0):
        """tdb200_channel_batch: BPSK + AWGN + LLR = 2r/sigma^2 on coded [n_cb, 3K+12] uint8."""
        n_cb = int(coded.shape[0])
        assert int(coded.shape[1]) == self.llr_len
        cp, mem = _ptr_of(coded)
        if _is_torch(coded):
            import torch
            llr = torch.empty((n_cb, self.llr_len), dtype=getattr(torch, dtype), device=coded.device)
        else:
            llr = np.empty((n_cb, self.llr_len), dtype=dtype)
        _check(self._L.tdb200_channel_batch(self._h, cp, _ptr_of(llr)[0], _LLR_TYPES[dtype], mem, n_cb,
                                            float(sigma), int(seed), stream))
        return llr

    # ---- higher-order mapping (ITTC/modanddem.cpp module()/demodule()); symbols are planar (I, Q)
    def _like(self, ref, shape, dtype):
        if _is_torch(ref):
            import torch
            return torch.empty(shape, dtype=getattr(torch, dtype), device=ref.device)
        return np.empty(shape, dtype=dtype)

    def modulate(self, coded, modulation, dtype="float32", stream=0):
        """tdb200_modulate_flat: bits [n_cb, n] uint8 -> (I, Q), each [n_cb, n / modulation]
        (n = 3K+12 for whole coded blocks, E for rate-matched ones)."""
        n_cb, n = int(coded.shape[0]), int(coded.shape[1])
        assert n % modulation == 0 and str(coded.dtype).endswith("uint8")
        si, sq = self._like(coded, (n_cb, n // modulation), dtype), self._like(coded, (n_cb, n // modulation), dtype)
        cp, mem = _ptr_of(coded)
        _check(self._L.tdb200_modulate_flat(self._h, cp, _ptr_of(si)[0], _ptr_of(sq)[0], _LLR_TYPES[dtype], mem, n_cb * n,
                                            int(modulation), stream))
        return si, sq

    def awgn(self, x, sigma, seed, stream=0):
        """tdb200_awgn_batch: x + sigma * N(0,1), same kind and dtype as x."""
        tname = str(x.dtype).replace("torch.", "")
        y = self._like(x, tuple(x.shape), tname)
        xp, mem = _ptr_of(x)
        n = int(np.prod(tuple(x.shape)))
        _check(self._L.tdb200_awgn_batch(self._h, xp, _ptr_of(y)[0], _LLR_TYPES[tname], mem, n, float(sigma), int(seed), stream))
        return y

    def demap(self, sym_i, sym_q, modulation, kf, dtype="float32", stream=0):
        """tdb200_demap_flat: received symbols [n_cb, ns] -> llr [n_cb, ns * modulation] of `dtype`
        (float64: the reference's demodule() bit for bit; int8: the s16 decoder's channel values)."""
        n_cb, ns = int(sym_i.shape[0]), int(sym_i.shape[1])
        assert tuple(sym_i.shape) == tuple(sym_q.shape)
        tname = str(sym_i.dtype).replace("torch.", "")
        llr = self._like(sym_i, (n_cb, ns * modulation), dtype)
        ip, mem = _ptr_of(sym_i)
        _check(self._L.tdb200_demap_flat(self._h, ip, _ptr_of(sym_q)[0], _LLR_TYPES[tname], _ptr_of(llr)[0], _LLR_TYPES[dtype],
                                         mem, n_cb * ns * modulation, int(modulation), float(kf), stream))
        return llr

    def decode_symbols_raw(self, i_ptr, q_ptr, sym_type, mem, n_cb, modulation, kf, bits=None, iters_used=None, stream=0):
        """Thin call-through with raw pointers (timed loops)."""
        o = Outputs(bits, None, iters_used, None, None, None)
        _check(self._L.tdb200_decode_symbols_batch(self._h, i_ptr, q_ptr, sym_type, mem, n_cb, int(modulation), float(kf),
                                                   C.byref(o), stream))

    def decode_symbols(self, sym_i, sym_q, modulation, kf, want=("bits",), stream=0):
        """tdb200_decode_symbols_batch: demodule() + TurboDecoding() in one call."""
        n_cb = int(sym_i.shape[0])
        assert int(sym_i.shape[1]) * modulation == self.llr_len and tuple(sym_i.shape) == tuple(sym_q.shape)
        tname = str(sym_i.dtype).replace("torch.", "")
        K, T = self.K, self.T
        shapes = {"bits": ((n_cb, K), "uint8"), "bits_iters": ((n_cb, self.n_iter, K), "int32"),
                  "iters_used": ((n_cb,), "int32"),
                  "llr_siso1": ((n_cb, T), None), "llr_siso2": ((n_cb, T), None), "ext_siso2": ((n_cb, T), None)}
        fl = "float64" if self.algo == ALGO_LOGMAP_F64 else "float32"
        outs = {name: self._like(sym_i, shapes[name][0], shapes[name][1] or fl) for name in want}
        o = Outputs(**{k: _ptr_of(v)[0] for k, v in outs.items()})
        ip, mem = _ptr_of(sym_i)
        _check(self._L.tdb200_decode_symbols_batch(self._h, ip, _ptr_of(sym_q)[0], _LLR_TYPES[tname], mem, n_cb,
                                                   int(modulation), float(kf), C.byref(o), stream))
        return outs

    # ---- TS 36.212 rate matching (the reference's declared-only rate_match / de_rate_match)
    def set_filler_bits(self, F):
        """tdb200_set_filler_bits: the first F information bits of every code block of this handle are filler bits
        (<NULL> in d0 / d1: not transmitted; the soft inverse writes a confident 0 there)."""
        _check(self._L.tdb200_set_filler_bits(self._h, int(F)))

    def rate_match(self, coded, E, rv=0, ncb=0, stream=0):
        """tdb200_rate_match_batch: coded [n_cb, 3K+12] uint8 -> transmitted bits [n_cb, E]."""
        n_cb = int(coded.shape[0])
        assert int(coded.shape[1]) == self.llr_len and str(coded.dtype).endswith("uint8")
        e = self._like(coded, (n_cb, E), "uint8")
        cp, mem = _ptr_of(coded)
        _check(self._L.tdb200_rate_match_batch(self._h, cp, _ptr_of(e)[0], mem, n_cb, int(E), int(rv), int(ncb), stream))
        return e

    def rate_dematch(self, e_llr, rv=0, ncb=0, into=None, stream=0):
        """tdb200_rate_dematch_batch: e_llr [n_cb, E] -> llr [n_cb, 3K+12] of the same dtype;
        `into` (same kind, shape [n_cb, 3K+12]) is combined in place (HARQ) and returned."""
        n_cb, E = int(e_llr.shape[0]), int(e_llr.shape[1])
        tname = str(e_llr.dtype).replace("torch.", "")
        llr = into if into is not None else self._like(e_llr, (n_cb, self.llr_len), tname)
        ep, mem = _ptr_of(e_llr)
        _check(self._L.tdb200_rate_dematch_batch(self._h, ep, _ptr_of(llr)[0], _LLR_TYPES[tname], mem, n_cb, E, int(rv), int(ncb),
                                                 0 if into is None else 1, stream))
        return llr

    def decode_rm(self, e_llr, rv=0, ncb=0, want=("bits",), stream=0):
        """tdb200_decode_rm_batch: de_rate_match() + TurboDecoding() in one call."""
        n_cb, E = int(e_llr.shape[0]), int(e_llr.shape[1])
        tname = str(e_llr.dtype).replace("torch.", "")
        K, T = self.K, self.T
        shapes = {"bits": ((n_cb, K), "uint8"), "bits_iters": ((n_cb, self.n_iter, K), "int32"),
                  "iters_used": ((n_cb,), "int32"),
                  "llr_siso1": ((n_cb, T), None), "llr_siso2": ((n_cb, T), None), "ext_siso2": ((n_cb, T), None)}
        fl = "float64" if self.algo == ALGO_LOGMAP_F64 else "float32"
        outs = {name: self._like(e_llr, shapes[name][0], shapes[name][1] or fl) for name in want}
        o = Outputs(**{k: _ptr_of(v)[0] for k, v in outs.items()})
        ep, mem = _ptr_of(e_llr)
        _check(self._L.tdb200_decode_rm_batch(self._h, ep, _LLR_TYPES[tname], mem, n_cb, E, int(rv), int(ncb), C.byref(o), stream))
        return outs

    # ---- transport-block stage: CRC24A / CRC24B (TS 36.212 5.1.1)
    def crc24_attach(self, bits, which=CRC24B, stream=0):
        """tdb200_crc24_attach_batch: overwrite the last 24 bits of every row of bits [n_rows, n] with the
        CRC of the rest (in place); rows are code blocks (n = K) or transport blocks (any n > 24)."""
        assert str(bits.dtype).endswith("uint8")
        bp, mem = _ptr_of(bits)
        _check(self._L.tdb200_crc24_attach_batch(self._h, bp, int(bits.shape[1]), int(which), mem, int(bits.shape[0]), stream))
        return bits

    def crc24_check(self, bits, which=CRC24B, want_remainder=False, stream=0):
        """tdb200_crc24_check_batch: ok [n_rows] uint8 (and the 24-bit remainders) of bits [n_rows, n]."""
        n_cb = int(bits.shape[0])
        assert str(bits.dtype).endswith("uint8")
        ok = self._like(bits, (n_cb,), "uint8")
        rem = self._like(bits, (n_cb,), "int32") if want_remainder else None
        bp, mem = _ptr_of(bits)
        _check(self._L.tdb200_crc24_check_batch(self._h, bp, int(bits.shape[1]), int(which), _ptr_of(ok)[0], _ptr_of(rem)[0], mem, n_cb, stream))
        return (ok, rem) if want_remainder else ok

    def siso(self, recs, La, terminated=1, stream=0):
        """One BCJR pass (tdb200_siso_batch), the Log_MAP_decoder replacement; doubles only."""
        n_cb = int(La.shape[0])
        rp, mem = _ptr_of(recs)
        lp, _ = _ptr_of(La)
        if _is_torch(La):
            import torch
            out = torch.empty_like(La)
        else:
            out = np.empty_like(La)
        op, _ = _ptr_of(out)
        _check(self._L.tdb200_siso_batch(self._h, rp, lp, terminated, op, mem, n_cb, stream))
        return out
