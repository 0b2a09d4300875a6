"""GPU parity tests of the fp32 sub-block-parallel decoders (TDB200_ALGO_MAXLOG_F32 / LOGMAP_F32),
through the C ABI, against their plain-C specification oracle/turbo_oracle_f32.c and against the
fp64 restatement of the reference (oracle/turbo_oracle.c)."""
import numpy as np
import pytest

from oracle_lib import F32Params

pytestmark = pytest.mark.gpu

# The Log-MAP variant evaluates ln(1+e^-d) with the hardware ex2/lg2 approximations (about 2 ulp each)
# where the model calls exp2f/log2f; over 4-8 iterations of a converging block the a-posteriori
# values stay within this absolute tolerance (they are of order 10..100).
LOGMAP_TOL = 2e-2


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _params(K, n_iter, L, G, logmap, et=0):
    return F32Params(K=K, n_iter=n_iter, sub_len=L, warmup=G, logmap=logmap, early_term=et,
                     ext_scale=1.0 if logmap else 0.75, ext_clamp=1.0e6, et_threshold=8.0)


@pytest.mark.parametrize("K,L,G,n_cb,n_iter,ebn0", [
    (6144, 0, 0, 3, 6, 0.8),     # auto plan (L=48, G=16)
    (6144, 96, 32, 2, 3, 0.6),
    (6144, 48, 0, 2, 3, 0.6),    # next-iteration initialisation only
    (512, 16, 16, 3, 5, 1.5),
    (1008, 0, 0, 3, 4, 1.0),     # P = 21: not a warp multiple
    (40, 40, 0, 5, 6, 2.0),      # single sub-block: the unsegmented recursion
])
@pytest.mark.parametrize("algo,lm", [("maxlog_f32", 0), ("linlogmap_f32", 2)])
def test_maxlog_and_linear_logmap_f32_bit_exact(oracle, K, L, G, n_cb, n_iter, ebn0, algo, lm):
    """Neither variant touches a special-function unit, so every float operation is IEEE and the
    plain-C model reproduces the kernel bit for bit."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=5 + K + L)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo=algo, sub_block=L, warmup=G)
    plan = dec.plan()
    assert plan["cb_per_cta"] == 1
    for x in (torch.from_numpy(llr32).cuda(), llr32):   # device path, host path
        out = dec.decode(x, want=("bits", "llr_siso2", "ext_siso2", "iters_used"))
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
        for c in range(n_cb):
            b, l, le, it = oracle.f32_decode(llr32[c], pi, _params(K, n_iter, plan["sub_block"], plan["warmup"], lm), want_soft=True)
            assert np.array_equal(out["bits"][c], b.astype(np.uint8))
            assert np.array_equal(out["llr_siso2"][c][:K], l), "a-posteriori values must be bit-identical"
            assert np.array_equal(out["ext_siso2"][c][:K], le)
            assert out["iters_used"][c] == n_iter


@pytest.mark.parametrize("K,L,G,n_iter,ebn0", [(6144, 0, 0, 6, 1.0), (6144, 128, 32, 4, 1.2), (1024, 32, 16, 5, 1.5)])
def test_logmap_f32_matches_model(oracle, K, L, G, n_iter, ebn0):
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    n_cb = 3
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=17 + K)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f32", sub_block=L, warmup=G)
    plan = dec.plan()
    out = dec.decode(llr32, want=("bits", "llr_siso2", "ext_siso2"))
    for c in range(n_cb):
        b, l, le, _ = oracle.f32_decode(llr32[c], pi, _params(K, n_iter, plan["sub_block"], plan["warmup"], 1), want_soft=True)
        got = out["llr_siso2"][c][:K]
        assert np.abs(got - l).max() < LOGMAP_TOL
        assert np.abs(out["ext_siso2"][c][:K] - le).max() < LOGMAP_TOL
        firm = np.abs(l) > LOGMAP_TOL       # identical hard decisions wherever |LLR| exceeds the tolerance
        assert np.array_equal((got >= 0)[firm], (l >= 0)[firm])
        assert np.array_equal(out["bits"][c][pi][firm], b[pi][firm].astype(np.uint8))


def test_unsegmented_logmap_f32_tracks_reference_logmap(oracle):
    """One sub-block (L = K) is the unsegmented recursion: its decisions equal those of the fp64
    restatement of the reference (which tabulates the same correction in 16 steps) on blocks that
    converge, and the a-posteriori LLRs agree to the size of the LUT steps."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 1024, 4, 6
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n_cb, 2.0, seed=23)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f32", sub_block=K, warmup=0)
    assert dec.plan()["n_sub_blocks"] == 1
    out = dec.decode(llr.astype(np.float32), want=("bits", "llr_siso2"))
    for c in range(n_cb):
        ob, _, o2, _ = oracle.decode(llr[c], pi, n_iter, want_llr=True)
        assert np.array_equal(out["bits"][c], ob[-1].astype(np.uint8))
        assert np.array_equal(out["bits"][c], bits[c].astype(np.uint8))
        rel = np.abs(out["llr_siso2"][c][:K] - o2[:K]) / (1.0 + np.abs(o2[:K]))
        assert rel.max() < 0.1


@pytest.mark.parametrize("algo,logmap", [("maxlog_f32", 0), ("logmap_f32", 1)])
def test_early_termination_f32(oracle, algo, logmap):
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 2048, 6, 8
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n_cb, 1.6, seed=41)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo=algo, early_term=True)
    plan = dec.plan()
    out = dec.decode(llr32, want=("bits", "iters_used"))
    its = []
    for c in range(n_cb):
        b, _, _, it = oracle.f32_decode(llr32[c], pi, _params(K, n_iter, plan["sub_block"], plan["warmup"], logmap, et=1))
        its.append(it)
        if not logmap:
            assert np.array_equal(out["bits"][c], b.astype(np.uint8))
    if not logmap:
        assert out["iters_used"].tolist() == its
    assert out["iters_used"].min() >= 2 and out["iters_used"].max() < n_iter
    assert np.array_equal(out["bits"], bits.astype(np.uint8))


@pytest.mark.parametrize("algo,lm", [("maxlog_f32", 0), ("linlogmap_f32", 2), ("logmap_f32", 1)])
def test_extreme_llrs_stay_finite(oracle, algo, lm):
    """Channel values far beyond the binary16 range (the shared-memory format of these modes) are clamped to +-65504
    on the way in: no infinity reaches the recursions, the outputs stay finite and equal the model's, and a strongly
    received codeword decodes to itself."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_iter = 1024, 4
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, 3, 3.0, seed=23)
    llr32 = llr.astype(np.float32)
    llr32[0] *= 1e6                       # |LLR| up to ~1e7: +-inf as binary16 without the clamp
    llr32[1, ::5] = np.float32(3.0e38)
    llr32[1, 1::5] = np.float32(-3.0e38)
    dec = TurboDecoder(K, n_iter=n_iter, algo=algo)
    plan = dec.plan()
    out = dec.decode(llr32, want=("bits", "llr_siso2"))
    assert np.isfinite(out["llr_siso2"]).all()
    assert np.array_equal(out["bits"][0], bits[0].astype(np.uint8))
    for c in range(3):
        b, l, _, _ = oracle.f32_decode(llr32[c], pi, _params(K, n_iter, plan["sub_block"], plan["warmup"], lm), want_soft=True)
        if lm != 1:     # the exact Log-MAP mode uses the hardware ex2/lg2: tolerance-based (test_logmap_f32_matches_model)
            assert np.array_equal(out["bits"][c], b.astype(np.uint8))
            assert np.array_equal(out["llr_siso2"][c][:K], l)
