"""GPU parity of the fp64 reference-order mode (TDB200_ALGO_LOGMAP_F64) against the oracle.

north_star bar: a-posteriori / extrinsic LLRs within 1e-3 absolute of the reference CPU Log-MAP
on identical channel LLRs, identical hard decisions.  This mode performs the reference's fp64
operations in the reference's order, so the tolerance used here is far tighter (1e-6).
All calls go through the C ABI (ctypes).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LLR_TOL = 1e-6  # north_star allows 1e-3


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.parametrize("K,ebn0,n_cb,n_iter", [
    (6144, 1.0, 3, 8),   # BASELINE configs[0]: the reference's own CPU-runnable case
    (6144, 0.4, 2, 8),   # waterfall: not converged early, LLRs small -> most sensitive to LUT flips
    (6144, 0.0, 2, 4),
    (40, 2.0, 9, 6),     # smallest LTE block, ragged batch (not a multiple of 4)
    (512, 1.5, 5, 5),
    (1008, 1.0, 4, 3),
])
def test_decode_matches_oracle(oracle, K, ebn0, n_cb, n_iter):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n_cb, ebn0, seed=1234 + K)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f64", max_batch=8)
    want = ("bits", "bits_iters", "iters_used", "llr_siso1", "llr_siso2", "ext_siso2")
    # device path (zero-copy torch tensors) and host path (numpy) must agree with each other
    out_d = dec.decode(torch.from_numpy(llr).cuda(), want=want)
    torch.cuda.synchronize()
    out_h = dec.decode(llr, want=want)
    for c in range(n_cb):
        ob, o1, o2, ole = oracle.decode(llr[c], pi, n_iter, want_llr=True)
        for out in (out_h, {k: v.cpu().numpy() for k, v in out_d.items()}):
            assert np.array_equal(out["bits_iters"][c], ob), "hard decisions differ from the oracle"
            assert np.array_equal(out["bits"][c], ob[-1].astype(np.uint8))
            assert out["iters_used"][c] == n_iter
            assert np.abs(out["llr_siso1"][c] - o1).max() < LLR_TOL
            assert np.abs(out["llr_siso2"][c] - o2).max() < LLR_TOL
            assert np.abs(out["ext_siso2"][c] - ole).max() < LLR_TOL
    assert dec.plan()["kernel_launches_last_call"] >= 1


def test_llrs_are_bit_identical_to_the_oracle(oracle):
    """Stronger than the 1e-3 bar of north_star and the 1e-6 of the test above: the kernel performs the reference's
    floating-point operations on the same operands in the same order (re-created alpha windows and the hashed max*
    table included), so every a-posteriori and extrinsic LLR of the waterfall case equals the CPU restatement bit for bit
    (tools/llr_parity_report.py does the same against the compiled reference: profiles/r02_llr_parity_report.txt)."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_iter = 6144, 8
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, 5, 0.4, seed=77)   # ragged: one warp of four codeblocks and one of one
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f64", max_batch=8)
    out = dec.decode(torch.from_numpy(llr).cuda(), want=("bits_iters", "llr_siso1", "llr_siso2", "ext_siso2"))
    for c in range(5):
        ob, o1, o2, ole = oracle.decode(llr[c], pi, n_iter, want_llr=True)
        assert np.array_equal(out["bits_iters"][c].cpu().numpy(), ob)
        for name, ref in (("llr_siso1", o1), ("llr_siso2", o2), ("ext_siso2", ole)):
            assert np.array_equal(out[name][c].cpu().numpy(), ref), name


def test_llr_input_types(oracle):
    """float32 input is widened exactly; the decode of float32-rounded LLRs matches the oracle on them."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_iter = 256, 4
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, 4, 1.0, seed=5)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f64")
    out = dec.decode(llr32, want=("bits_iters", "llr_siso2"))
    for c in range(4):
        ob, _, o2, _ = oracle.decode(llr32[c].astype(np.float64), pi, n_iter, want_llr=True)
        assert np.array_equal(out["bits_iters"][c], ob)
        assert np.abs(out["llr_siso2"][c] - o2).max() < LLR_TOL


def test_siso_matches_oracle(oracle):
    """tdb200_siso_batch == Log_MAP_decoder on arbitrary (recs, La), terminated and not."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K = 1024
    T = K + 3
    rng = np.random.default_rng(3)
    n_cb = 5
    recs = rng.normal(0, 2.0, size=(n_cb, 2 * T))
    La = rng.normal(0, 3.0, size=(n_cb, T))
    dec = TurboDecoder(K, n_iter=1, algo="logmap_f64")
    for term in (1, 0):
        got = dec.siso(recs, La, terminated=term)
        for c in range(n_cb):
            ref = oracle.siso(recs[c], La[c], terminated=term)
            assert np.abs(got[c] - ref).max() < LLR_TOL


def test_empty_batch_and_bad_args():
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TdbError, TurboDecoder
    dec = TurboDecoder(40, n_iter=2, algo="logmap_f64")
    out = dec.decode(np.zeros((0, 132)), want=("bits",))
    assert out["bits"].shape == (0, 40)
    with pytest.raises(TdbError):
        TurboDecoder(41, algo="logmap_f64")          # not an LTE size and no f1/f2 given
    with pytest.raises(TdbError):
        TurboDecoder(40, f1=4, f2=10, algo="logmap_f64")  # not a permutation
    with pytest.raises(TdbError):
        TurboDecoder(40, n_iter=0, algo="logmap_f64")


def test_compat_layer_is_the_reference_call(oracle):
    """The reference's own entry points, called the way ITTC/main.cpp:221 calls them (through the mangled
    symbols of libtdb200_compat.so): TurboDecoding(double*, int*, int) fills flow_decoded for every
    iteration and halves the caller's buffer; Log_MAP_decoder(...) is one BCJR pass."""
    _torch_cuda()
    import ctypes as C
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "turbo_decoder_cuda_b200", "lib", "libtdb200_compat.so")
    if not os.path.exists(so):
        pytest.skip("compat library not built")
    os.environ["TDB200_COMPAT_ITERS"] = "6"
    lib = C.CDLL(so)
    K, n_iter = 512, 6
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, 1, 1.5, seed=3)
    buf = np.array(llr[0], np.float64, copy=True)
    out = np.zeros(n_iter * K, np.int32)
    dp, ip = np.ctypeslib.ndpointer(np.float64), np.ctypeslib.ndpointer(np.int32)
    f = getattr(lib, "_Z13TurboDecodingPdPii")
    f.argtypes = [dp, ip, C.c_int]
    f.restype = None
    f(buf, out, 3 * K + 12)
    ob, o1, o2, le = oracle.decode(llr[0], pi, n_iter, want_llr=True)
    assert np.array_equal(out.reshape(n_iter, K), ob)
    assert np.array_equal(buf, llr[0] * 0.5), "the reference halves flow_for_decode in place (log_map.cpp:1202-1205)"
    # one SISO pass on the demultiplexed half-LLRs
    T = K + 3
    recs = np.zeros(2 * T)
    recs[0:2 * K:2] = 0.5 * llr[0][0:3 * K:3]
    recs[1:2 * K:2] = 0.5 * llr[0][1:3 * K:3]
    recs[2 * K:] = 0.5 * llr[0][3 * K:3 * K + 6]
    La = np.zeros(T)
    got = np.zeros(T)
    g = getattr(lib, "_Z15Log_MAP_decoderPdS_iS_i")
    g.argtypes = [dp, dp, C.c_int, dp, C.c_int]
    g.restype = None
    g(recs, La, 1, got, T)
    want = oracle.siso(recs, La, terminated=1)
    assert np.abs(got - want).max() < LLR_TOL
    # the declared-only pair rate_match / de_rate_match (ITTC/main.h:23-24, call sites main.cpp:196,204)
    coded = oracle.encode(bits[0].astype(np.int32), pi)
    E = 2 * K
    tx = np.zeros(E, np.int32)
    rmf = getattr(lib, "_Z10rate_matchPiiS_i")
    rmf.argtypes = [ip, C.c_int, ip, C.c_int]
    rmf.restype = None
    rmf(np.ascontiguousarray(coded, np.int32), 3 * K + 12, tx, E)
    assert np.array_equal(tx, oracle.rate_match(coded, K, E, 0))
    rx = np.random.default_rng(5).standard_normal(E)
    back = np.zeros(3 * K + 12)
    dmf = getattr(lib, "_Z13de_rate_matchPdS_ii")
    dmf.argtypes = [dp, dp, C.c_int, C.c_int]
    dmf.restype = None
    dmf(rx, back, E, 3 * K + 12)
    assert np.array_equal(back, oracle.rate_dematch(rx, K, 0))


def test_error_paths_and_concurrent_handles(oracle):
    """Argument checking returns status codes with a message (never a crash), and two handles with
    different plans can live and run side by side (one stream each)."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TdbError, TurboDecoder
    for kw in (dict(K=6145), dict(K=6144, n_iter=0), dict(K=6144, sub_block=50), dict(K=6144, sub_block=48, warmup=12),
               dict(K=6144, frac_bits=9), dict(K=6144, ext_clip=62), dict(K=6144, et_threshold=48), dict(K=6144, algo=7),
               dict(K=6144, f1=2, f2=2), dict(K=6144, device=99)):
        with pytest.raises(TdbError) as e:
            TurboDecoder(**kw)
        assert e.value.status in (1, 2) and str(e.value)
    d1 = TurboDecoder(6144, n_iter=4, algo="maxlog_s16")
    d2 = TurboDecoder(512, n_iter=4, algo="maxlog_s16", sub_block=16, warmup=8)
    with pytest.raises(TdbError):
        d1.decode(torch.zeros((2, 3 * 6144 + 12), device="cuda"), want=("llr_siso1",))   # fp64-mode output
    b1, l1 = oracle.make_batch(6144, 4, 1.5, seed=1)
    b2, l2 = oracle.make_batch(512, 6, 2.0, seed=2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    x1, x2 = torch.from_numpy(l1.astype(np.float32)).cuda(), torch.from_numpy(l2.astype(np.float32)).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        o1 = d1.decode(x1, want=("bits",), stream=s1.cuda_stream)
        o2 = d2.decode(x2, want=("bits",), stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(o1["bits"].cpu().numpy(), b1.astype(np.uint8))
    assert np.array_equal(o2["bits"].cpu().numpy(), b2.astype(np.uint8))
    d1.close(); d2.close()
    d1.close()  # idempotent
