"""GPU parity of TDB200_ALGO_LOGMAP_S16 -- Log-MAP (max* with the linear correction) in the packed 16-bit
arithmetic of the throughput kernel -- through the C ABI.

Integer work -> the bar is BIT-EXACT: hard decisions, extrinsics and iteration counts must equal the int32
specification oracle/turbo_oracle_fx.c (logmap = 1) on the same seeded inputs, for compile-time and run-time
geometries, both fixed-point formats, every input type, ragged batches and saturated inputs.  How this mode
relates to the reference's fp64 Log-MAP (ITTC/log_map.cpp:898-1047 with E_algorithm :779-801) is statistical
and is tested in test_gpu_ber.py (same frames through both decoders).
"""
import numpy as np
import pytest

from oracle_lib import FxParams

pytestmark = pytest.mark.gpu


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def lm_params(K, n_iter, L, G, F=4, ext_clip=511, early_term=0, et_threshold=0):
    return FxParams(K=K, n_iter=n_iter, sub_len=L, warmup=G, frac_bits=F, llr_clip=127, ext_clip=ext_clip,
                    ext_scale_q2=4, early_term=early_term, et_threshold=et_threshold, logmap=1, lm_upper_off=1)


def _check(oracle, dec, llr_in, llr_f32, pi, prm, n_cb, K, iters=None):
    out = dec.decode(llr_in, want=("bits", "ext_siso2", "llr_siso2", "iters_used"))
    out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    scale = float(1 << prm.frac_bits)
    for c in range(n_cb):
        bits, le, it, ovf = oracle.fx_decode(llr_f32[c], pi, prm, want_le=True)
        assert ovf == 0, "int16 range exceeded in the specification model"
        assert np.array_equal(out["bits"][c], bits.astype(np.uint8)), "hard decisions differ (cb %d)" % c
        got = np.rint(out["ext_siso2"][c][:K] * scale).astype(np.int32)
        assert np.array_equal(got, le[pi]), "extrinsics differ (cb %d)" % c
        lam = out["llr_siso2"][c][:K]
        assert np.array_equal((lam >= 0).astype(np.uint8), bits[pi].astype(np.uint8))
        assert out["iters_used"][c] == (prm.n_iter if iters is None else iters[c])


@pytest.mark.parametrize("K,L,G,n_cb,n_iter,ebn0", [
    (6144, 0, 0, 5, 8, 0.4),     # auto plan (L=48, G=24): compile-time geometry, odd batch, waterfall
    (6144, 48, 32, 2, 4, 0.4),   # the longer guard
    (6144, 48, 16, 2, 4, 0.4),   # the faster guard
    (6144, 96, 32, 2, 3, 0.4),   # run-time geometry
    (5120, 40, 32, 3, 3, 0.8),
    (4096, 32, 16, 3, 3, 0.8),
    (6144, 24, 24, 2, 2, 0.8),   # guard == sub-block length
    (40, 40, 0, 7, 6, 2.0),      # single sub-block: the unsegmented recursion
    (40, 8, 8, 4, 4, 2.0),
    (512, 16, 16, 4, 5, 1.5),
    (1008, 0, 0, 3, 4, 1.0),     # P = 21: not a warp multiple, several pairs per CTA
    (2048, 64, 16, 2, 4, 1.0),
])
def test_bit_exact_vs_fixed_point_model(oracle, K, L, G, n_cb, n_iter, ebn0):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=177 + K + L)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16", sub_block=L, warmup=G)
    plan = dec.plan()
    if L == 0 and K == 6144:
        assert (plan["sub_block"], plan["warmup"]) == (48, 24)
    prm = lm_params(K, n_iter, plan["sub_block"], plan["warmup"])
    _check(oracle, dec, torch.from_numpy(llr32).cuda(), llr32, pi, prm, n_cb, K)   # device path
    _check(oracle, dec, llr32, llr32, pi, prm, n_cb, K)                             # host path


@pytest.mark.parametrize("F,ec", [(3, 511), (4, 255), (3, 1023)])
def test_fixed_point_formats(oracle, F, ec):
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 1024, 4, 5
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, 1.0, seed=31)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16", frac_bits=F, ext_clip=ec)
    plan = dec.plan()
    _check(oracle, dec, llr32, llr32, pi, lm_params(K, n_iter, plan["sub_block"], plan["warmup"], F, ec), n_cb, K)


def test_input_types(oracle):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 768, 3, 4
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, 1.2, seed=8)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16")
    plan = dec.plan()
    prm = lm_params(K, n_iter, plan["sub_block"], plan["warmup"])
    _check(oracle, dec, llr, llr32, pi, prm, n_cb, K)
    q = np.clip(np.rint(llr32 * 16.0), -127, 127).astype(np.int8)     # 4 fractional bits
    _check(oracle, dec, q, (q.astype(np.float32) / 16.0), pi, prm, n_cb, K)
    h = llr32.astype(np.float16)
    _check(oracle, dec, h, h.astype(np.float32), pi, prm, n_cb, K)
    _check(oracle, dec, torch.from_numpy(h).cuda(), h.astype(np.float32), pi, prm, n_cb, K)


def test_extreme_llrs_do_not_overflow(oracle):
    """Saturated, erased (0 / NaN) and non-codeword inputs stay inside int16 and stay bit-exact."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_iter = 6144, 8
    pi = oracle.qpp(K)
    rng = np.random.default_rng(5)
    bits, llr = oracle.make_batch(K, 4, 3.0, seed=9)
    llr32 = llr.astype(np.float32)
    llr32[0] *= 1e4
    llr32[1, ::7] = 0.0
    llr32[1, 5::11] = np.nan
    llr32[2] = (rng.integers(0, 2, llr32.shape[1]) * 2 - 1) * 1e3
    llr32[3] = 0.0
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16")
    plan = dec.plan()
    prm = lm_params(K, n_iter, plan["sub_block"], plan["warmup"])
    out = dec.decode(llr32, want=("bits", "ext_siso2"))
    for c in range(4):
        b, le, it, ovf = oracle.fx_decode(np.nan_to_num(llr32[c], nan=0.0), pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(out["bits"][c], b.astype(np.uint8))
        assert np.array_equal(np.rint(out["ext_siso2"][c][:K] * 16).astype(np.int32), le[pi])
    assert np.array_equal(out["bits"][0], bits[0].astype(np.uint8))


@pytest.mark.parametrize("K,ebn0", [(6144, 1.0), (6144, 0.45), (1024, 2.0)])
def test_early_termination(oracle, K, ebn0):
    """Per-codeblock stopping rule (decisions unchanged and every |a-posteriori| >= threshold): iteration counts and
    delivered decisions as the model's, per codeblock."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    n_cb, n_iter = 6, 8
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n_cb, ebn0, seed=4242 + K)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16", early_term=1, max_batch=n_cb)
    plan = dec.plan()
    T = 1 << (4 + 3)
    out = dec.decode(llr32, want=("bits", "iters_used"))
    res = [oracle.fx_decode(llr32[c], pi, lm_params(K, n_iter, plan["sub_block"], plan["warmup"], early_term=1, et_threshold=T))
           for c in range(n_cb)]
    assert list(out["iters_used"]) == [r[2] for r in res]
    for c in range(n_cb):   # a block delivers the decisions it stopped with, whatever its lane mate does
        assert np.array_equal(out["bits"][c], res[c][0].astype(np.uint8))


@pytest.mark.parametrize("algo,K,early", [("logmap_s16", 6144, 0), ("maxlog_s16", 6144, 0), ("logmap_s16", 1008, 0),
                                          ("maxlog_s16", 512, 0), ("logmap_s16", 6144, 1), ("maxlog_s16", 6144, 1)])
def test_per_iteration_decisions(oracle, algo, K, early):
    """bits_iters of the throughput decoders: row k is what the same decoder delivers with n_iter = k + 1 -- the
    reference's flow_decoded + K*iteration (ITTC/log_map.cpp:1261-1264), which main.cpp:224-237 counts errors on.
    With early termination rows past the stop repeat the delivered decisions."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    n_cb, n_iter = 5, 6
    ebn0 = 0.6 if K == 6144 else 1.5
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=99 + K)
    llr32 = llr.astype(np.float32)
    x = torch.from_numpy(llr32).cuda()
    dec = TurboDecoder(K, n_iter=n_iter, algo=algo, early_term=early)
    out = dec.decode(x, want=("bits", "bits_iters", "iters_used"))
    rows = out["bits_iters"].cpu().numpy()
    bits = out["bits"].cpu().numpy()
    used = out["iters_used"].cpu().numpy()
    assert rows.shape == (n_cb, n_iter, K) and rows.dtype == np.int32
    assert np.array_equal(rows[:, -1, :].astype(np.uint8), bits)
    ran = [int(u) for u in used]
    for k in range(n_iter):
        d = TurboDecoder(K, n_iter=k + 1, algo=algo)
        b = d.decode(x, want=("bits",))["bits"].cpu().numpy()
        for c in range(n_cb):
            if k < ran[c]:
                assert np.array_equal(rows[c, k].astype(np.uint8), b[c]), "iteration %d cb %d" % (k + 1, c)
            else:
                assert np.array_equal(rows[c, k].astype(np.uint8), bits[c])
        d.close()
    # host path delivers the same slab
    out_h = dec.decode(llr32, want=("bits_iters",))
    assert np.array_equal(out_h["bits_iters"], rows)


def test_every_lte_block_size_logmap(oracle):
    """BASELINE configs[3], correctness side, for the Log-MAP kernels: all 188 LTE block sizes decode bit-exactly against the
    integer model with the library's own plan for that K (own plan table, csrc/tdb200_plan_table_lm.h; compile-time,
    run-time-P and fully run-time instantiations)."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    sizes = oracle.lte_sizes()
    n_iter = 2
    kinds = set()
    for K in sizes:
        pi = oracle.qpp(K)
        _, llr = oracle.make_batch(K, 3, 1.5, seed=K + 7)      # odd batch: one CTA runs half empty
        llr32 = llr.astype(np.float32)
        dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_s16", max_batch=4)
        plan = dec.plan()
        assert plan["sub_block"] * plan["n_sub_blocks"] == K
        kinds.add((plan["sub_block"], plan["warmup"]))
        out = dec.decode(llr32, want=("bits",))
        prm = lm_params(K, n_iter, plan["sub_block"], plan["warmup"])
        b, _, _, ovf = oracle.fx_decode(llr32[2], pi, prm)
        assert ovf == 0 and np.array_equal(out["bits"][2], b.astype(np.uint8)), "K=%d" % K
        dec.close()
    assert len(kinds) >= 8   # the plans really differ


@pytest.mark.parametrize("K,early,n_cb", [(1312, False, 1801), (656, True, 1801), (512, False, 2401)])
def test_packed_pairs_large_batch_logmap(oracle, K, early, n_cb):
    """Several codeblock pairs per CTA in the Log-MAP kernels (a large batch is what makes the planner pack them): bit-exact,
    with and without early termination, on a sample of the batch."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, 32, 1.5, seed=K)
    llr32 = np.tile(llr.astype(np.float32), ((n_cb + 31) // 32, 1))[:n_cb]
    dec = TurboDecoder(K, n_iter=4, algo="logmap_s16", early_term=early, max_batch=n_cb)
    plan = dec.plan()
    out = dec.decode(torch.from_numpy(llr32).cuda(), want=("bits", "iters_used"))
    got, its = out["bits"].cpu().numpy(), out["iters_used"].cpu().numpy()
    prm = lm_params(K, 4, plan["sub_block"], plan["warmup"], early_term=1 if early else 0, et_threshold=128 if early else 0)
    for c in (0, 1, 31, 32, n_cb - 1):
        b, _, it, ovf = oracle.fx_decode(llr32[c], pi, prm)
        assert ovf == 0 and np.array_equal(got[c], b.astype(np.uint8)), "cb %d" % c
        assert int(its[c]) == it
