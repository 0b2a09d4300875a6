"""GPU parity of the throughput mode (TDB200_ALGO_MAXLOG_S16) through the C ABI.

Integer work -> the bar is BIT-EXACT: hard decisions and extrinsics must equal the int32
specification oracle/turbo_oracle_fx.c on the same seeded inputs, for every geometry
(sub-block length L, guard G), fixed-point format and input type, including ragged batches
(odd n_cb: the second int16 lane of the last CTA is a duplicate whose outputs are dropped).
The statistical relation of this mode to the reference Log-MAP is tested in test_gpu_ber.py.
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


def _fx_params(K, n_iter, L, G, F=3, q2=3, ext_clip=0):
    return FxParams(K=K, n_iter=n_iter, sub_len=L, warmup=G, frac_bits=F,
                    llr_clip=min((1 << (F + 4)) - 1, 127), ext_clip=ext_clip or (1 << (F + 6)) - 1,
                    ext_scale_q2=q2, early_term=0)


def _check(oracle, dec, llr_in, llr_f32, pi, prm, n_cb, K):
    out = dec.decode(llr_in, want=("bits", "ext_siso2", "llr_siso2", "iters_used"))
    out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    scale = float(1 << prm.frac_bits)
    for c in range(n_cb):
        bits, le, it, ovf = oracle.fx_decode(llr_f32[c], pi, prm, want_le=True)
        assert ovf == 0, "int16 range exceeded in the specification model"
        assert np.array_equal(out["bits"][c], bits.astype(np.uint8)), "hard decisions differ (cb %d)" % c
        got = np.rint(out["ext_siso2"][c][:K] * scale).astype(np.int32)
        assert np.array_equal(got, le[pi]), "extrinsics differ (cb %d)" % c
        # a-posteriori sign must agree with the delivered decision
        lam = out["llr_siso2"][c][:K]
        assert np.array_equal((lam >= 0).astype(np.uint8), bits[pi].astype(np.uint8))
        assert out["iters_used"][c] == prm.n_iter


@pytest.mark.parametrize("K,L,G,n_cb,n_iter,ebn0", [
    (6144, 0, 0, 5, 8, 0.6),     # auto plan (L=48, G=16): BASELINE configs[1] geometry, odd batch
    (6144, 48, 0, 2, 4, 0.4),    # next-iteration initialisation only
    (6144, 96, 32, 2, 3, 0.4),
    (6144, 32, 8, 3, 3, 0.8),
    (6144, 24, 24, 2, 2, 0.8),   # guard == sub-block length
    (40, 40, 0, 7, 6, 2.0),      # single sub-block: exact unsegmented max-log
    (40, 8, 8, 4, 4, 2.0),
    (512, 16, 16, 4, 5, 1.5),
    (1008, 0, 0, 3, 4, 1.0),     # P = 21: not a warp multiple
    (2048, 64, 16, 2, 4, 1.0),
])
def test_bit_exact_vs_fixed_point_model(oracle, K, L, G, n_cb, n_iter, ebn0):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=77 + K + L)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16", sub_block=L, warmup=G)
    plan = dec.plan()
    prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
    _check(oracle, dec, torch.from_numpy(llr32).cuda(), llr32, pi, prm, n_cb, K)   # device path
    _check(oracle, dec, llr32, llr32, pi, prm, n_cb, K)                             # host path


@pytest.mark.parametrize("F,q2,ec", [(4, 3, 0), (3, 4, 0), (2, 3, 0), (3, 3, 63), (3, 3, 1023), (3, 4, 255)])
def test_fixed_point_formats(oracle, F, q2, ec):
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 1024, 4, 5
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, 1.0, seed=31)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16", frac_bits=F, ext_scale_q2=q2, ext_clip=ec)
    plan = dec.plan()
    _check(oracle, dec, llr32, llr32, pi, _fx_params(K, n_iter, plan["sub_block"], plan["warmup"], F, q2, ec), n_cb, K)


def test_input_types(oracle):
    """float64 input is rounded to float32 first; int8 input is taken as already quantised; float16 is exact in float32."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 768, 3, 4
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, 1.2, seed=8)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16")
    plan = dec.plan()
    prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
    _check(oracle, dec, llr, llr32, pi, prm, n_cb, K)
    q = np.clip(np.rint(llr32 * 8.0), -127, 127).astype(np.int8)
    _check(oracle, dec, q, (q.astype(np.float32) / 8.0), pi, prm, n_cb, K)
    h = llr32.astype(np.float16)                       # binary16 transport: the model sees the rounded values
    _check(oracle, dec, h, h.astype(np.float32), pi, prm, n_cb, K)
    import torch
    _check(oracle, dec, torch.from_numpy(h).cuda(), h.astype(np.float32), pi, prm, n_cb, K)


def test_extreme_llrs_do_not_overflow(oracle):
    """Saturated, erased (0 / NaN) and sign-alternating inputs stay inside int16 and stay bit-exact."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_iter = 6144, 8
    pi = oracle.qpp(K)
    rng = np.random.default_rng(5)
    bits, llr = oracle.make_batch(K, 4, 3.0, seed=9)
    llr32 = llr.astype(np.float32)
    llr32[0] *= 1e4                      # everything saturates at the channel clip
    llr32[1, ::7] = 0.0                  # erasures
    llr32[1, 5::11] = np.nan
    llr32[2] = (rng.integers(0, 2, llr32.shape[1]) * 2 - 1) * 1e3   # saturated garbage: no codeword
    llr32[3] = 0.0                       # all erased
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16")
    plan = dec.plan()
    prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
    out = dec.decode(llr32, want=("bits", "ext_siso2"))
    for c in range(4):
        b, le, it, ovf = oracle.fx_decode(np.nan_to_num(llr32[c], nan=0.0), pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(out["bits"][c], b.astype(np.uint8))
        assert np.array_equal(np.rint(out["ext_siso2"][c][:K] * 8).astype(np.int32), le[pi])
    assert np.array_equal(out["bits"][0], bits[0].astype(np.uint8))


@pytest.mark.parametrize("K,ebn0", [(6144, 1.2), (6144, 0.5), (1024, 2.0), (40, 3.0)])
def test_early_termination(oracle, K, ebn0):
    """Hard-decision-aided stop: iters_used and the delivered bits equal the model's stopping iteration and its
    decisions at that iteration for every codeblock (the two codeblocks of a lane pair leave the SM together, but a
    stopped block's decisions are frozen)."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    n_cb, n_iter = 7, 8
    pi = oracle.qpp(K)
    _, llr = oracle.make_batch(K, n_cb, ebn0, seed=91 + K)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16", early_term=True)
    plan = dec.plan()
    out = dec.decode(llr32, want=("bits", "iters_used"))
    prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
    prm.early_term = 1
    prm.et_threshold = 1 << (prm.frac_bits + 3)  # the library default: |LLR| >= 8
    res = [oracle.fx_decode(llr32[c], pi, prm) for c in range(n_cb)]
    its = [r[2] for r in res]
    assert out["iters_used"].tolist() == its
    for c in range(n_cb):   # a block delivers the decisions it stopped with, whatever its lane mate does
        assert np.array_equal(out["bits"][c], res[c][0].astype(np.uint8))
    assert min(its) >= 2


def test_decodes_clean_codewords(oracle):
    """Size-independent property at the full BASELINE size: at 2 dB every codeblock of a 64-block
    batch decodes to the transmitted bits."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K = 6144
    bits, llr = oracle.make_batch(K, 64, 2.0, seed=4)
    dec = TurboDecoder(K, n_iter=8, algo="maxlog_s16")
    out = dec.decode(torch.from_numpy(llr.astype(np.float32)).cuda(), want=("bits",))
    assert np.array_equal(out["bits"].cpu().numpy(), bits.astype(np.uint8))


def test_every_lte_block_size(oracle):
    """BASELINE configs[3], correctness side: all 188 LTE block sizes (K = 40 ... 6144) decode
    bit-exactly against the integer model with the library's own plan for that K."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    sizes = oracle.lte_sizes()
    assert len(sizes) == 188 and sizes[0] == 40 and sizes[-1] == 6144
    n_iter = 3
    for K in sizes:
        pi = oracle.qpp(K)
        bits, llr = oracle.make_batch(K, 3, 1.5, seed=K)       # odd batch: one CTA runs half empty
        llr32 = llr.astype(np.float32)
        dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16", max_batch=4)
        plan = dec.plan()
        assert plan["sub_block"] * plan["n_sub_blocks"] == K
        out = dec.decode(llr32, want=("bits",))
        prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
        for c in (0, 2):
            b, _, _, ovf = oracle.fx_decode(llr32[c], pi, prm)
            assert ovf == 0 and np.array_equal(out["bits"][c], b.astype(np.uint8)), "K=%d cb %d" % (K, c)
        dec.close()


@pytest.mark.parametrize("K,early,n_cb", [(1312, False, 1801), (1312, True, 1801), (656, False, 1801), (656, True, 1801),
                                          (2624, False, 1801), (3136, True, 1801), (512, False, 2401), (512, True, 2401),
                                          (960, True, 1801)])
def test_packed_pairs_large_batch(oracle, K, early, n_cb):
    """Sub-block counts just above a warp multiple (P = 41; P = 56 where shared memory is the limit) share a CTA between two or three
    codeblock pairs once the batch is large enough to fill the device; the result stays bit-exact
    (decisions, and per-pair stopping with early termination), also for the odd last codeblock."""
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    n_iter = 5
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, n_cb, 1.6, seed=K + 5)
    llr32 = llr.astype(np.float32)
    dec = TurboDecoder(K, n_iter=n_iter, algo="maxlog_s16", early_term=early, max_batch=4096)
    plan = dec.plan()
    assert plan["cb_per_cta"] >= 4, plan
    out = dec.decode(llr32, want=("bits", "iters_used"))
    prm = _fx_params(K, n_iter, plan["sub_block"], plan["warmup"])
    if early:
        prm.early_term = 1
        prm.et_threshold = 1 << (prm.frac_bits + 3)
    for c in (0, 1, 4, 5, 900, 901, n_cb - 3, n_cb - 2, n_cb - 1):
        b, _, it, _ = oracle.fx_decode(llr32[c], pi, prm)
        assert out["iters_used"][c] == it, "cb %d" % c
        assert np.array_equal(out["bits"][c], b.astype(np.uint8)), "cb %d" % c
    # everything else: at 1.6 dB nearly every block of these sizes decodes to the transmitted bits
    wrong = (out["bits"] != bits.astype(np.uint8)).any(axis=1)
    assert wrong.mean() < 0.05
    dec.close()


def test_int8_input_large_batch_and_alignment(oracle):
    """8-bit channel values with more codeblocks than resident CTAs (the next-row L2 prefetch runs;
    byte rows are only 4-byte aligned) and a row-offset view of the buffer; a misaligned device
    pointer is refused, not faulted on."""
    import torch
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import TdbError
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    K, n_cb = 6144, 1500
    dec = TurboDecoder(K, n_iter=4, max_batch=2048)
    from turbo_decoder_cuda_b200 import synth
    bits, llr = synth.make_batch(K, 64, 1.5, seed=5, device="cuda")
    q = torch.clamp(torch.round(llr * 8), -127, 127).to(torch.int8)
    big = q.repeat((n_cb + 63) // 64 + 1, 1)[:n_cb + 1].contiguous()
    want = dec.decode(q)["bits"]
    got = dec.decode(big[:n_cb])["bits"]
    assert torch.equal(got[:64], want) and torch.equal(got[640:704], want)
    odd = dec.decode(big[1:n_cb + 1])["bits"]          # starts 18444 bytes in: 4-byte aligned only
    assert torch.equal(odd[63:127], want)
    flat = big.view(-1)
    with pytest.raises(TdbError):
        dec.decode_raw(flat.data_ptr() + 1, 2, 1, 4, bits=got.data_ptr())
