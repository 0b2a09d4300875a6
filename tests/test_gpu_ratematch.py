"""GPU tests of the TS 36.212 rate-matching stage around the decode path:
tdb200_rate_match_batch / tdb200_rate_dematch_batch / tdb200_decode_rm_batch against the oracle's
literal restatement of the specification (oracle/turbo_oracle_rm.c).  The device builds its
permutation in closed form, the oracle fills the padded matrices and runs the selection loop, so
agreement over every block size is a check of both.  (The reference only declares this stage,
ITTC/main.h:23-24: parity with 3GPP vectors is unpinned, see the oracle's header.)"""
import numpy as np
import pytest

from oracle_lib import ALGO_LOGMAP_LUT, FxParams

pytestmark = pytest.mark.gpu


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def test_every_block_size_matches_the_specification_restatement(oracle):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    rng = np.random.default_rng(2)
    for K in oracle.lte_sizes():
        NL = 3 * K + 12
        Kw = oracle.rm_geometry(K)["Kw"]
        rv = int(rng.integers(0, 4))
        E = int(rng.integers(1, 3 * NL))                      # puncturing through double repetition
        ncb = 0 if rng.random() < 0.6 else int(rng.integers(Kw // 3, Kw))
        dec = TurboDecoder(K, n_iter=1, max_batch=2)
        coded = rng.integers(0, 2, size=(2, NL), dtype=np.uint8)
        got = dec.rate_match(torch.from_numpy(coded).cuda(), E, rv, ncb).cpu().numpy()
        for c in range(2):
            assert np.array_equal(got[c], oracle.rate_match(coded[c], K, E, rv, ncb).astype(np.uint8)), (K, E, rv, ncb)
        e = rng.standard_normal((2, E)).astype(np.float32)
        back = dec.rate_dematch(torch.from_numpy(e).cuda(), rv, ncb).cpu().numpy()
        for c in range(2):
            assert np.array_equal(back[c], oracle.rate_dematch(e[c], K, rv, ncb)), (K, E, rv, ncb)
        dec.close()


def test_filler_bits_match_the_specification_restatement(oracle):
    """tdb200_set_filler_bits: the F filler positions of d0 / d1 are <NULL> (TS 36.212 5.1.3.2.1) -- never selected, and
    the soft inverse writes the fixed confident 0 there; device == oracle for random (K, F, rv, E, N_cb), all four soft
    types, and a first code block with filler bits decodes through rate matching."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TdbError, TurboDecoder
    rng = np.random.default_rng(11)
    for K in (40, 104, 512, 1056, 3136, 6144):
        NL = 3 * K + 12
        Kw = oracle.rm_geometry(K)["Kw"]
        dec = TurboDecoder(K, n_iter=1, max_batch=2)
        for F in (int(rng.integers(1, min(K, 64))), int(rng.integers(1, K))):
            rv = int(rng.integers(0, 4))
            E = int(rng.integers(1, 3 * NL))
            ncb = 0 if rng.random() < 0.5 else int(rng.integers(Kw // 2, Kw))
            dec.set_filler_bits(F)
            coded = rng.integers(0, 2, size=(2, NL), dtype=np.uint8)
            got = dec.rate_match(torch.from_numpy(coded).cuda(), E, rv, ncb).cpu().numpy()
            e = rng.standard_normal((2, E))
            back = dec.rate_dematch(torch.from_numpy(e).cuda(), rv, ncb).cpu().numpy()
            back32 = dec.rate_dematch(torch.from_numpy(e.astype(np.float32)).cuda(), rv, ncb).cpu().numpy()
            back8 = dec.rate_dematch(torch.from_numpy(np.clip(np.rint(e * 8), -127, 127).astype(np.int8)).cuda(), rv, ncb).cpu().numpy()
            for c in range(2):
                assert np.array_equal(got[c], oracle.rate_match_f(coded[c], K, E, rv, ncb, F).astype(np.uint8)), (K, F, E, rv, ncb)
                assert np.array_equal(back[c], oracle.rate_dematch_f(e[c], K, rv, ncb, F)), (K, F, E, rv, ncb)
                fill = np.zeros(NL, bool)
                fill[0:3 * F:3] = fill[1:3 * F:3] = True
                assert (back32[c][fill] == -100.0).all() and (back8[c][fill] == -127).all()
        dec.set_filler_bits(0)
        coded = rng.integers(0, 2, size=(1, NL), dtype=np.uint8)
        assert np.array_equal(dec.rate_match(torch.from_numpy(coded).cuda(), NL, 0, 0).cpu().numpy()[0],
                              oracle.rate_match(coded[0], K, NL, 0, 0).astype(np.uint8))
        with pytest.raises(TdbError):
            dec.set_filler_bits(K)
        dec.close()
    # a first code block with F = 40 filler zeros, rate 1/2, through the whole chain: the fillers are never sent
    K, F = 6144, 40
    dec = TurboDecoder(K, n_iter=8, max_batch=4)
    dec.set_filler_bits(F)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    bits = torch.randint(0, 2, (4, K), dtype=torch.uint8, device="cuda", generator=g)
    bits[:, :F] = 0
    E = 2 * K
    tx = dec.rate_match(dec.encode(bits), E, 0, 0)
    sigma = 0.7
    rx = (2.0 * tx.float() - 1.0) + sigma * torch.randn(tx.shape, device="cuda", generator=g)
    out = dec.decode_rm(rx * (2.0 / sigma ** 2), 0, 0)
    assert torch.equal(out["bits"], bits)


def test_dematch_types_combining_and_host_path(oracle):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 1024, 3
    NL = 3 * K + 12
    dec = TurboDecoder(K, max_batch=4)
    rng = np.random.default_rng(4)
    E = NL + 500                                              # one full wrap plus a repeated head
    e = (4 * rng.standard_normal((n_cb, E))).astype(np.float32)
    te = torch.from_numpy(e).cuda()
    f32 = dec.rate_dematch(te, 1).cpu().numpy()
    assert np.array_equal(dec.rate_dematch(e, 1), f32), "host-memory path"
    f64 = dec.rate_dematch(te.double(), 1).cpu().numpy()
    for c in range(n_cb):
        assert np.array_equal(f64[c], oracle.rate_dematch(e[c].astype(np.float64), K, 1))
    h = dec.rate_dematch(te.half(), 1).cpu().numpy()
    for c in range(n_cb):   # half in: float sums of the half values, rounded to half once
        assert np.array_equal(h[c], oracle.rate_dematch(e[c].astype(np.float16).astype(np.float32), K, 1).astype(np.float16))
    q = np.clip(np.rint(e * 8), -127, 127).astype(np.int8)
    s8 = dec.rate_dematch(torch.from_numpy(q).cuda(), 1).cpu().numpy()
    sel = oracle.rm_selection(K, E, 1)
    for c in range(n_cb):   # 8-bit in: integer sums, saturated once at the end
        want = np.zeros(NL, np.int64)
        np.add.at(want, sel, q[c].astype(np.int64))
        assert np.array_equal(s8[c], np.clip(want, -127, 127).astype(np.int8))
    # HARQ: a retransmission with rv = 2 combined into the buffer of the first transmission
    e2 = (4 * rng.standard_normal((n_cb, 900))).astype(np.float32)
    buf = dec.rate_dematch(te, 1)
    both = dec.rate_dematch(torch.from_numpy(e2).cuda(), 2, into=buf)
    assert both.data_ptr() == buf.data_ptr()
    for c in range(n_cb):
        assert np.array_equal(both.cpu().numpy()[c], oracle.rate_dematch(e2[c], K, 2, into=f32[c]))
    hb = dec.rate_dematch(e2, 2, into=f32.copy())
    assert np.array_equal(hb, both.cpu().numpy()), "host-memory path with accumulation"
    # E = 0: nothing received, everything erased
    z = dec.rate_dematch(torch.zeros((n_cb, 0), dtype=torch.float32, device="cuda"), 0)
    assert z.shape == (n_cb, NL) and not z.any()


@pytest.mark.parametrize("rate_num,rate_den,ebn0", [(1, 2, 2.5), (3, 4, 4.5), (1, 5, 0.8)])
def test_decode_rm_equals_dematch_then_decode(oracle, rate_num, rate_den, ebn0):
    """s16 decoder fed rate-matched LLRs: punctured (rate 1/2, 3/4) and repeated (rate 1/5) blocks.
    Same decisions and extrinsics as the integer model run on the de-rate-matched float LLRs."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, rv = 6144, 5, 0
    E = K * rate_den // rate_num
    pi = oracle.qpp(K)
    rng = np.random.default_rng(60 + rate_den)
    bits = rng.integers(0, 2, size=(n_cb, K), dtype=np.uint8)
    dec = TurboDecoder(K, n_iter=8, max_batch=8)
    coded = dec.encode(torch.from_numpy(bits).cuda())
    tx = dec.rate_match(coded, E, rv).cpu().numpy()
    sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 * rate_den / rate_num)
    r = ((2.0 * tx - 1.0) + sigma * rng.standard_normal(tx.shape)).astype(np.float32)
    e_llr = (2.0 * r / np.float32(sigma * sigma)).astype(np.float32)
    te = torch.from_numpy(e_llr).cuda()
    out = dec.decode_rm(te, rv, want=("bits", "ext_siso2"))
    got_bits, got_le = out["bits"].cpu().numpy(), out["ext_siso2"].cpu().numpy()
    plan = dec.plan()
    prm = FxParams(K=K, n_iter=8, sub_len=plan["sub_block"], warmup=plan["warmup"], frac_bits=3, llr_clip=127,
                   ext_clip=511, ext_scale_q2=3, early_term=0, et_threshold=64)
    for c in range(n_cb):
        llr = oracle.rate_dematch(e_llr[c], K, rv)
        want_bits, le, it, ovf = oracle.fx_decode(llr, pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(got_bits[c], want_bits.astype(np.uint8)), "cb %d" % c
        assert np.array_equal(np.rint(got_le[c][:K] * 8).astype(np.int32), le[pi]), "cb %d" % c
    assert np.array_equal(got_bits, bits), "operating point should decode cleanly"
    two = dec.decode(dec.rate_dematch(te, rv))["bits"].cpu().numpy()
    assert np.array_equal(two, got_bits)
    small = TurboDecoder(K, n_iter=8, max_batch=2)   # chunks of 2 through the slot ring
    assert np.array_equal(small.decode_rm(e_llr, rv)["bits"], got_bits), "host-memory pipeline"
    assert np.array_equal(small.decode_rm(te, rv)["bits"].cpu().numpy(), got_bits)
    for dt in (torch.float16, torch.float64):   # other input types: equal to their own two-call form
        a = dec.decode_rm(te.to(dt), rv)["bits"]
        b = dec.decode(dec.rate_dematch(te.to(dt), rv))["bits"]
        assert torch.equal(a, b)
    q = torch.clamp(torch.round(te * 8), -127, 127).to(torch.int8)
    assert torch.equal(dec.decode_rm(q, rv)["bits"], dec.decode(dec.rate_dematch(q, rv))["bits"])


def test_reference_decoder_behind_rate_matching(oracle):
    """fp64 mode: de_rate_match() + TurboDecoding() equals the oracle's decoder on the oracle's
    de-rate-matched doubles, bit for bit (LLRs included); two redundancy versions combined."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb, n_iter = 512, 3, 4
    pi = oracle.qpp(K)
    rng = np.random.default_rng(77)
    dec = TurboDecoder(K, n_iter=n_iter, algo="logmap_f64", max_batch=2)
    bits = rng.integers(0, 2, size=(n_cb, K), dtype=np.uint8)
    coded = dec.encode(bits)
    E = 700                       # rate 0.73 on the first transmission: too little on its own at this SNR
    sigma = 0.9
    llrs = []
    for rv in (0, 2):
        tx = dec.rate_match(coded, E, rv)
        llrs.append(2.0 * ((2.0 * tx - 1.0) + sigma * rng.standard_normal(tx.shape)) / sigma ** 2)
    out = dec.decode_rm(torch.from_numpy(llrs[0]).cuda(), 0, want=("bits_iters", "llr_siso2"))
    for c in range(n_cb):
        want_bits, _, l2, _ = oracle.decode(oracle.rate_dematch(llrs[0][c], K, 0), pi, n_iter, algo=ALGO_LOGMAP_LUT, want_llr=True)
        assert np.array_equal(out["bits_iters"].cpu().numpy()[c], want_bits)
        assert np.array_equal(out["llr_siso2"].cpu().numpy()[c], l2)
    first_errors = int((out["bits_iters"].cpu().numpy()[:, -1] != bits).sum())
    buf = dec.rate_dematch(llrs[0], 0)
    buf = dec.rate_dematch(llrs[1], 2, into=buf)
    comb = dec.decode(buf, want=("bits_iters",))["bits_iters"][:, -1]
    for c in range(n_cb):
        both = oracle.rate_dematch(llrs[1][c], K, 2, into=oracle.rate_dematch(llrs[0][c], K, 0))
        assert np.array_equal(comb[c], oracle.decode(both, pi, n_iter)[-1])
    assert int((comb != bits).sum()) == 0 and first_errors > 0, "incremental redundancy has to help"


def test_rate_matching_error_paths():
    _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import TdbError
    dec = TurboDecoder(40, max_batch=4)
    coded = np.zeros((1, 132), np.uint8)
    with pytest.raises(TdbError):
        dec.rate_match(coded, 100, rv=4)
    with pytest.raises(TdbError):
        dec.rate_match(coded, 100, rv=0, ncb=-5)
    with pytest.raises(TdbError):
        dec.rate_match(coded, 100, rv=0, ncb=1)       # the first buffer entry is <NULL>: nothing to send
    assert dec.rate_match(coded, 0).shape == (1, 0)
    assert dec.decode_rm(np.zeros((0, 77), np.float32))["bits"].shape == (0, 40)


@pytest.mark.parametrize("M,E", [(2, 12290), (4, 9164), (6, 8190), (3, 12291)])
def test_rate_matched_and_modulated_chain(oracle, M, E):
    """The whole transmit / receive chain around the decoder on the device: encode -> rate match ->
    map (rows of E bits, not a multiple of 12: the flat mapper / demapper and its tail path) -> AWGN ->
    soft demap -> decode from rate-matched LLRs."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 6144, 5
    dec = TurboDecoder(K, n_iter=8, max_batch=8)
    g = torch.Generator(device="cuda")
    g.manual_seed(M)
    bits = torch.randint(0, 2, (n_cb, K), dtype=torch.uint8, device="cuda", generator=g)
    tx = dec.rate_match(dec.encode(bits), E, 0)
    si, sq = dec.modulate(tx, M)
    assert tuple(si.shape) == (n_cb, E // M)
    want_i, want_q = oracle.modulate(tx.cpu().numpy().ravel(), M)
    assert np.array_equal(si.cpu().numpy().ravel(), want_i.astype(np.float32)) and np.array_equal(sq.cpu().numpy().ravel(), want_q.astype(np.float32))
    ebn0 = {2: 3.0, 3: 5.5, 4: 6.5, 6: 10.0}[M]
    sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 / ((K / E) * M))
    ri, rq = dec.awgn(si, sigma, seed=1), dec.awgn(sq, sigma, seed=2)
    kf = 1.0 / (2 * sigma * sigma)
    e_llr = dec.demap(ri, rq, M, kf)
    want = oracle.demap_f32(ri.cpu().numpy().ravel(), rq.cpu().numpy().ravel(), M, np.float32(kf))
    assert np.array_equal(e_llr.cpu().numpy().ravel(), want), "flat demapper incl. the tail of each call"
    e8 = dec.demap(ri, rq, M, kf, dtype="int8")
    assert np.array_equal(e8.cpu().numpy().ravel(), oracle.quant_s8(want))
    # an unaligned output buffer takes the symbol-by-symbol path: same values
    buf = torch.empty(n_cb * E + 1, dtype=torch.float32, device="cuda")
    from turbo_decoder_cuda_b200.decoder import _check
    _check(dec._L.tdb200_demap_flat(dec._h, ri.data_ptr(), rq.data_ptr(), 1, buf.data_ptr() + 4, 1, 1, n_cb * E, M, kf, None))
    assert torch.equal(buf[1:], e_llr.view(-1))
    out = dec.decode_rm(e_llr, 0)["bits"]
    assert torch.equal(out, bits)
    assert torch.equal(dec.decode_rm(e8, 0)["bits"], out), "8-bit hand-over: same decisions"
