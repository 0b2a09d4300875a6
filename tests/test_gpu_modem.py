"""GPU tests of the mapper / soft demapper either side of the decode path:
tdb200_modulate_batch / tdb200_demap_batch / tdb200_decode_symbols_batch against the oracle's
restatement of module()/demodule() (ITTC/modanddem.cpp:175,674), against the golden vectors the
reference itself produced (tests/golden/modem_golden.npz) and, where the reference build travelled
(oracle/_ref), against the reference's demodule() + decoder chain."""
import os

import numpy as np
import pytest

from oracle_lib import ALGO_LOGMAP_LUT, FxParams, RefLib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODS = (1, 2, 3, 4, 6)


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _traffic(oracle, K, n_cb, M, sigma, seed, dtype=np.float32):
    """bits -> oracle encoder -> oracle mapper -> seeded Gaussian noise (values representable in dtype)."""
    rng = np.random.default_rng(seed)
    pi = oracle.qpp(K)
    bits = rng.integers(0, 2, size=(n_cb, K), dtype=np.uint8)
    coded = np.stack([oracle.encode(b.astype(np.int32), pi) for b in bits]).astype(np.uint8)
    si, sq = oracle.modulate(coded.ravel(), M)
    ri = (si + sigma * rng.standard_normal(si.size)).astype(dtype).reshape(n_cb, -1)
    rq = (sq + sigma * rng.standard_normal(sq.size)).astype(dtype).reshape(n_cb, -1)
    return bits, coded, ri, rq, pi


@pytest.mark.parametrize("M", MODS)
def test_modulate_is_the_reference_table(oracle, M):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 512, 3
    dec = TurboDecoder(K, max_batch=8)
    rng = np.random.default_rng(M)
    coded = rng.integers(0, 2, size=(n_cb, 3 * K + 12), dtype=np.uint8)
    si, sq = oracle.modulate(coded.ravel(), M)
    for dtype in ("float64", "float32", "float16"):
        di, dq = dec.modulate(torch.from_numpy(coded).cuda(), M, dtype=dtype)
        hi, hq = dec.modulate(coded, M, dtype=dtype)
        for got_i, got_q in ((di.cpu().numpy(), dq.cpu().numpy()), (hi, hq)):
            assert got_i.shape == (n_cb, (3 * K + 12) // M)
            assert np.array_equal(got_i.ravel(), si.astype(dtype)) and np.array_equal(got_q.ravel(), sq.astype(dtype))
    if M == 6:   # all 64 index values, as the reference mapped them
        g = np.load(os.path.join(GOLD, "modem_golden.npz"))
        row = np.zeros(3 * K + 12, np.uint8)
        row[:384] = g["all_bits_6"]
        di, dq = dec.modulate(torch.from_numpy(row[None]).cuda(), 6, dtype="float64")
        assert np.array_equal(di.cpu().numpy()[0, :64], g["all_si_6"]) and np.array_equal(dq.cpu().numpy()[0, :64], g["all_sq_6"])


@pytest.mark.parametrize("M", MODS)
def test_demap_f64_bit_identical_to_demodule(oracle, M):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 40, 2          # 3K+12 = 132 LLRs per codeblock; the golden block is 264 = two rows
    dec = TurboDecoder(K, algo="logmap_f64", max_batch=4)
    g = np.load(os.path.join(GOLD, "modem_golden.npz"))
    ri, rq = g["ri_%d" % M].reshape(n_cb, -1), g["rq_%d" % M].reshape(n_cb, -1)
    kf = float(g["kf_%d" % M])
    got = dec.demap(torch.from_numpy(ri).cuda(), torch.from_numpy(rq).cuda(), M, kf, dtype="float64").cpu().numpy()
    assert np.array_equal(got.ravel(), g["llr_%d" % M]), "fp64 demapper vs the reference's demodule() output"
    assert np.array_equal(dec.demap(ri, rq, M, kf, dtype="float64").ravel(), g["llr_%d" % M]), "host-memory path"
    # a larger seeded case against the restatement (and the live reference where it travelled)
    K2 = 512
    dec2 = TurboDecoder(K2, algo="logmap_f64", max_batch=4)
    _, _, xi, xq, _ = _traffic(oracle, K2, 3, M, 0.5, seed=7 + M, dtype=np.float64)
    got = dec2.demap(torch.from_numpy(xi).cuda(), torch.from_numpy(xq).cuda(), M, 2.0, dtype="float64").cpu().numpy()
    want = oracle.demap_f64(xi.ravel(), xq.ravel(), M, 2.0)
    assert np.array_equal(got.ravel(), want)
    if RefLib.available():
        assert np.array_equal(want, RefLib(40, 3, 10).demodule(xi.ravel(), xq.ravel(), M, 2.0))


@pytest.mark.parametrize("M", MODS)
def test_demap_f32_family_bit_exact_vs_model(oracle, M):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 1024, 4
    dec = TurboDecoder(K, max_batch=8)
    sigma = {1: 0.9, 2: 0.7, 3: 0.45, 4: 0.3, 6: 0.15}[M]
    _, _, ri, rq, _ = _traffic(oracle, K, n_cb, M, sigma, seed=30 + M)
    ri[0, :4] = [0.0, 1e6, -1e6, np.float32(1e-30)]      # far outside / degenerate inputs
    kf = np.float32(1.0 / (2.0 * sigma * sigma))
    want = oracle.demap_f32(ri.ravel(), rq.ravel(), M, kf)
    ti, tq = torch.from_numpy(ri).cuda(), torch.from_numpy(rq).cuda()
    got = dec.demap(ti, tq, M, float(kf), dtype="float32").cpu().numpy()
    assert np.array_equal(got.ravel(), want), "fp32 demapper vs its C model"
    got16 = dec.demap(ti, tq, M, float(kf), dtype="float16").cpu().numpy()
    with np.errstate(over="ignore"):
        assert np.array_equal(got16.ravel(), want.astype(np.float16))
    got8 = dec.demap(ti, tq, M, float(kf), dtype="int8").cpu().numpy()
    assert np.array_equal(got8.ravel(), oracle.quant_s8(want))
    assert np.array_equal(dec.demap(ri, rq, M, float(kf), dtype="int8"), got8), "host-memory path"
    # other symbol types: half symbols are exact in float, double symbols are rounded to float first
    with np.errstate(over="ignore"):
        h_i, h_q = ri.astype(np.float16), rq.astype(np.float16)
        want_h = oracle.demap_f32(h_i.astype(np.float32).ravel(), h_q.astype(np.float32).ravel(), M, kf)
    got_h = dec.demap(torch.from_numpy(h_i).cuda(), torch.from_numpy(h_q).cuda(), M, float(kf), dtype="float32").cpu().numpy()
    assert np.array_equal(got_h.ravel(), want_h, equal_nan=True)
    got_d = dec.demap(ti.double(), tq.double(), M, float(kf), dtype="float32").cpu().numpy()
    assert np.array_equal(got_d.ravel(), want)
    # and the fp32 metric is the reference's metric up to float rounding
    ref64 = oracle.demap_f64(ri.ravel()[8:], rq.ravel()[8:], M, float(kf))
    assert np.abs(want[8 * M:] - ref64).max() <= 1e-4 * max(1.0, np.abs(ref64).max())


@pytest.mark.parametrize("M", MODS)
def test_decode_symbols_equals_demap_then_decode(oracle, M):
    """s16 decoder: symbols in -> same decisions and extrinsics as the integer model run on the
    fp32-demapped LLRs; host and device memory; more codeblocks than one pipeline chunk."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 6144, 6
    ebn0 = {1: 1.6, 2: 1.6, 3: 4.0, 4: 4.0, 6: 6.0}[M]
    rate = K / (3.0 * K + 12.0)
    sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 / (rate * M))          # ITTC/main.cpp:174
    bits, _, ri, rq, pi = _traffic(oracle, K, n_cb, M, sigma, seed=50 + M)
    kf = 1.0 / (2.0 * sigma * sigma)
    dec = TurboDecoder(K, n_iter=6, max_batch=8)
    plan = dec.plan()
    ti, tq = torch.from_numpy(ri).cuda(), torch.from_numpy(rq).cuda()
    out = dec.decode_symbols(ti, tq, M, kf, want=("bits", "ext_siso2"))
    llr32 = oracle.demap_f32(ri.ravel(), rq.ravel(), M, np.float32(kf)).reshape(n_cb, -1)
    prm = FxParams(K=K, n_iter=6, sub_len=plan["sub_block"], warmup=plan["warmup"], frac_bits=3, llr_clip=127,
                   ext_clip=511, ext_scale_q2=3, early_term=0, et_threshold=64)
    got_bits, got_le = out["bits"].cpu().numpy(), out["ext_siso2"].cpu().numpy()
    for c in range(n_cb):
        want_bits, le, it, ovf = oracle.fx_decode(llr32[c], pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(got_bits[c], want_bits.astype(np.uint8)), "cb %d" % c
        assert np.array_equal(np.rint(got_le[c][:K] * 8).astype(np.int32), le[pi]), "cb %d" % c
    assert np.mean(got_bits != bits) < 1e-3, "operating point should decode (almost) cleanly"
    # the two-call form and the host-memory pipeline give the same decisions
    two = dec.decode(dec.demap(ti, tq, M, kf, dtype="int8"))["bits"].cpu().numpy()
    assert np.array_equal(two, got_bits)
    small = TurboDecoder(K, n_iter=6, max_batch=2)   # chunks of 2: exercises the slot ring with symbols
    assert np.array_equal(small.decode_symbols(ri, rq, M, kf)["bits"], got_bits)
    assert np.array_equal(small.decode_symbols(ti, tq, M, kf)["bits"].cpu().numpy(), got_bits)
    # half-precision symbols: a quarter of the float-LLR bytes per LLR at M = 1, 1/24 at M = 6
    h = dec.decode_symbols(ti.half(), tq.half(), M, kf)["bits"].cpu().numpy()
    assert np.mean(h != bits) < 1e-3


@pytest.mark.parametrize("M", (2, 4, 6))
def test_reference_chain_demodule_then_turbo_decoding(oracle, M):
    """fp64 mode: symbols -> demodule() -> TurboDecoding(), bit-identical to the reference chain."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 512, 3, 4
    sigma = {2: 0.55, 4: 0.28, 6: 0.14}[M]
    bits, _, ri, rq, pi = _traffic(oracle, K, n_cb, M, sigma, seed=70 + M, dtype=np.float64)
    kf = 1.0 / (2.0 * sigma * sigma)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f64", max_batch=2)
    out = dec.decode_symbols(torch.from_numpy(ri).cuda(), torch.from_numpy(rq).cuda(), M, kf,
                             want=("bits_iters", "llr_siso2"))
    host = dec.decode_symbols(ri, rq, M, kf, want=("bits_iters", "llr_siso2"))
    ref = RefLib(K, *oracle.lte_params(K)) if RefLib.available() else None
    for c in range(n_cb):
        llr = oracle.demap_f64(ri[c], rq[c], M, kf)
        want_bits, _, l2, _ = oracle.decode(llr, pi, n_iter, algo=ALGO_LOGMAP_LUT, want_llr=True)
        assert np.array_equal(out["bits_iters"].cpu().numpy()[c], want_bits)
        assert np.array_equal(out["llr_siso2"].cpu().numpy()[c], l2), "a-posteriori LLRs, bit for bit"
        assert np.array_equal(host["bits_iters"][c], want_bits) and np.array_equal(host["llr_siso2"][c], l2)
        if ref is not None:
            rb, _, rl2, _ = ref.decode(ref.demodule(ri[c], rq[c], M, kf), n_iter, want_llr=True)
            assert np.array_equal(rb, want_bits) and np.array_equal(rl2, l2)


def test_awgn_statistics_and_determinism():
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    dec = TurboDecoder(512, max_batch=4)
    x = torch.zeros(1 << 20, dtype=torch.float32, device="cuda") + 0.25
    a, b, c = dec.awgn(x, 0.5, seed=3), dec.awgn(x, 0.5, seed=3), dec.awgn(x, 0.5, seed=4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    n = (a - 0.25) / 0.5
    N = n.numel()
    assert abs(float(n.mean())) < 5 / N ** 0.5
    assert abs(float(n.var()) - 1.0) < 5 * (2.0 / N) ** 0.5
    assert abs(float((n ** 4).mean()) - 3.0) < 0.05
    assert torch.equal(dec.awgn(x, 0.0, seed=1), x)
    h = dec.awgn(np.full(1001, 0.25, np.float64), 0.5, seed=3)      # host memory, ragged length, doubles
    assert h.dtype == np.float64 and abs(h.mean() - 0.25) < 0.1 and h.std() > 0.4


def test_monte_carlo_chain_on_device_64qam():
    """encode -> modulate -> AWGN -> decode_symbols, everything on the device (the main.cpp loop
    at MODULATION = 6): clean at 6.5 dB, broken at 2 dB (below the 64QAM BICM capacity limit for 2 bit/symbol)."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, M = 6144, 64, 6
    dec = TurboDecoder(K, max_batch=64)
    bits = torch.randint(0, 2, (n_cb, K), dtype=torch.uint8, device="cuda")
    si, sq = dec.modulate(dec.encode(bits), M)
    rate = K / (3.0 * K + 12.0)
    fer = {}
    for ebn0 in (6.5, 2.0):
        sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 / (rate * M))
        ri, rq = dec.awgn(si, sigma, seed=1), dec.awgn(sq, sigma, seed=2)
        out = dec.decode_symbols(ri, rq, M, 1.0 / (2 * sigma * sigma))["bits"]
        fer[ebn0] = float((out != bits).any(dim=1).float().mean())
    assert fer[6.5] == 0.0 and fer[2.0] > 0.9, fer


def test_modem_error_paths():
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import TdbError
    dec = TurboDecoder(40, max_batch=4)
    from turbo_decoder_cuda_b200.decoder import _check
    x = np.zeros(132, np.float32)
    with pytest.raises(TdbError):                # 5 bits per symbol is not a reference modulation
        _check(dec._L.tdb200_demap_batch(dec._h, x.ctypes.data, x.ctypes.data, 1, x.ctypes.data, 1, 0, 1, 5, 1.0, None))
    s = np.zeros((1, 66), np.float32)
    with pytest.raises(TdbError):
        dec.decode_symbols(s, s, 2, -1.0)        # Kf must be positive
    with pytest.raises(TdbError):
        dec.demap(s.astype(np.int8), s.astype(np.int8), 2, 1.0)   # symbols cannot be int8
    assert dec.decode_symbols(np.zeros((0, 66), np.float32), np.zeros((0, 66), np.float32), 2, 1.0)["bits"].shape == (0, 40)


@pytest.mark.parametrize("algo", ["maxlog_s16", "logmap_s16"])
@pytest.mark.parametrize("M", [1, 2])
def test_fused_demapper_equals_two_call_form(oracle, algo, M, monkeypatch):
    """BPSK / QPSK float symbols into the packed decoders are demapped inside the decoder's load stage
    (ITTC/modanddem.cpp:189-260 fused in front of the branch metrics): decisions and extrinsics equal demapping to the
    8-bit hand-over format and decoding that, from device and from host memory, for an odd batch, several chunks and a
    short block with several pairs per CTA; the unfused path (TDB200_NO_FUSED_DEMAP) gives the same."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    for K, n_cb in ((6144, 5), (1008, 7), (40, 9)):
        rate = K / (3.0 * K + 12.0)
        ebn0 = 1.0 if K == 6144 else 2.5
        sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 / (rate * M))
        bits, _, ri, rq, pi = _traffic(oracle, K, n_cb, M, sigma, seed=70 + M + K)
        kf = 1.0 / (2.0 * sigma * sigma)
        dec = TurboDecoder(K, n_iter=5, algo=algo, max_batch=4)
        ti, tq = torch.from_numpy(ri).cuda(), torch.from_numpy(rq).cuda()
        fused = dec.decode_symbols(ti, tq, M, kf, want=("bits", "ext_siso2"))
        two = dec.decode(dec.demap(ti, tq, M, kf, dtype="int8"), want=("bits", "ext_siso2"))
        assert torch.equal(fused["bits"], two["bits"]) and torch.equal(fused["ext_siso2"], two["ext_siso2"]), (K, "device")
        host = dec.decode_symbols(ri, rq, M, kf, want=("bits", "ext_siso2"))
        assert np.array_equal(host["bits"], two["bits"].cpu().numpy()) and np.array_equal(host["ext_siso2"], two["ext_siso2"].cpu().numpy())
        monkeypatch.setenv("TDB200_NO_FUSED_DEMAP", "1")
        unfused = dec.decode_symbols(ti, tq, M, kf, want=("bits",))
        monkeypatch.delenv("TDB200_NO_FUSED_DEMAP")
        assert torch.equal(unfused["bits"], two["bits"])
        # a device view that breaks the 16-byte alignment falls back to the separate demapper, same result
        if M == 1:
            pad = torch.zeros(ti.numel() + 1, device="cuda")
            pad[1:] = ti.ravel()
            off = pad[1:].view_as(ti)
            assert torch.equal(dec.decode_symbols(off, tq, M, kf, want=("bits",))["bits"], two["bits"])
        dec.close()
