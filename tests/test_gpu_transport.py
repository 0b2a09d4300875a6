"""GPU tests of the transport-block stage above the decode path: CRC24A / CRC24B on
the device against the bit-serial oracle (pinned to the catalogue check values, test_oracle.py), and
the segmentation -> encode -> channel -> decode -> CRC check -> concatenation round trip."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


@pytest.mark.parametrize("K", [40, 104, 512, 1056, 6144])
def test_crc24_attach_and_check(oracle, K):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import CRC24A, CRC24B
    dec = TurboDecoder(K, max_batch=8)
    rng = np.random.default_rng(K)
    n = 7
    bits = rng.integers(0, 2, size=(n, K), dtype=np.uint8)
    bits[0] = 0
    bits[1] = 1
    for which, poly in ((CRC24A, oracle.CRC24A), (CRC24B, oracle.CRC24B)):
        t = torch.from_numpy(bits.copy()).cuda()
        dec.crc24_attach(t, which)
        got = t.cpu().numpy()
        host = dec.crc24_attach(bits.copy(), which)
        for r in range(n):
            c = oracle.crc24(bits[r, :K - 24], poly)
            want = np.array([(c >> (23 - i)) & 1 for i in range(24)], np.uint8)
            assert np.array_equal(got[r, :K - 24], bits[r, :K - 24]) and np.array_equal(got[r, K - 24:], want), (K, which, r)
        assert np.array_equal(host, got), "host-memory path"
        ok, rem = dec.crc24_check(t, which, want_remainder=True)
        assert ok.cpu().numpy().all() and not rem.cpu().numpy().any()
        bad = got.copy()
        bad[2, 5] ^= 1                       # one flipped payload bit
        bad[3, K - 1] ^= 1                   # one flipped CRC bit
        ok, rem = dec.crc24_check(torch.from_numpy(bad).cuda(), which, want_remainder=True)
        assert ok.cpu().numpy().tolist() == [1, 1, 0, 0, 1, 1, 1]
        for r in range(n):
            assert int(rem.cpu().numpy()[r]) == oracle.crc24(bad[r], poly)
        assert np.array_equal(dec.crc24_check(bad, which), ok.cpu().numpy()), "host-memory path"


def test_crc24_on_transport_block_rows(oracle):
    """Rows of arbitrary length (the largest LTE transport block, and an odd length)."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import CRC24A
    dec = TurboDecoder(40, max_batch=2)
    rng = np.random.default_rng(1)
    for B in (75400, 25, 1001):
        tb = rng.integers(0, 2, size=(3, B), dtype=np.uint8)
        t = torch.from_numpy(tb.copy()).cuda()
        dec.crc24_attach(t, CRC24A)
        got = t.cpu().numpy()
        for r in range(3):
            c = oracle.crc24(tb[r, :B - 24], oracle.CRC24A)
            assert np.array_equal(got[r, B - 24:], np.array([(c >> (23 - i)) & 1 for i in range(24)], np.uint8)), B
        assert dec.crc24_check(t, CRC24A).cpu().numpy().all()


@pytest.mark.parametrize("A,algo", [(1000, "maxlog_s16"), (6120, "maxlog_s16"), (6121, "maxlog_s16"), (20000, "maxlog_s16"),
                                    (75376, "maxlog_s16"), (6121, "logmap_s16"), (75376, "logmap_s16")])
def test_transport_block_round_trip(oracle, A, algo):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200.transport import TransportBlockCodec
    n_tb = 3
    tb = TransportBlockCodec(A, n_iter=8, early_term=True, algo=algo)
    seg = oracle.segmentation(A + 24)
    assert tb.seg == seg
    g = torch.Generator(device="cuda")
    g.manual_seed(A)
    payload = torch.randint(0, 2, (n_tb, A), dtype=torch.uint8, device="cuda", generator=g)
    blocks = tb.segment(payload)
    assert sum(b.shape[0] for _, b in blocks) == n_tb * seg["C"]
    # block contents against a plain numpy restatement of 5.1.2 for transport block 0
    p0 = payload[0].cpu().numpy()
    c = oracle.crc24(p0, oracle.CRC24A)
    stream = np.concatenate([np.zeros(seg["F"], np.uint8), p0, [(c >> (23 - i)) & 1 for i in range(24)]]).astype(np.uint8)
    pos = 0
    for K, b in blocks:
        cnt = b.shape[0] // n_tb
        for r in range(cnt):
            row = b[r].cpu().numpy()      # transport block 0 owns the first `cnt` rows of the group
            take = K - seg["L"]
            assert np.array_equal(row[:take], stream[pos:pos + take]), (A, K, r)
            if seg["L"]:
                cb = oracle.crc24(row[:take], oracle.CRC24B)
                assert np.array_equal(row[take:], np.array([(cb >> (23 - i)) & 1 for i in range(24)], np.uint8))
            pos += take
    assert pos == stream.size
    # through the code and a clean-enough channel
    coded = tb.encode(blocks)
    sigma = 0.75
    llrs = [(K, tb.dec[K].channel(cw, sigma, seed=K)) for K, cw in coded]
    out, tb_ok, cb_ok = tb.decode(llrs)
    assert torch.equal(out, payload) and bool(tb_ok.all()) and bool(cb_ok.all())
    assert tuple(cb_ok.shape) == (n_tb, seg["C"])
    # erase one code block of transport block 1: its CRC24B and the block's CRC24A must both fail
    if seg["C"] > 1:
        K, l = llrs[-1]
        cnt = l.shape[0] // n_tb
        l = l.clone()
        l[1 * cnt + cnt - 1] = 0.01 * torch.randn_like(l[0])
        out2, tb_ok2, cb_ok2 = tb.decode(llrs[:-1] + [(K, l)])
        assert tb_ok2.cpu().numpy().tolist() == [1, 0, 1]
        want = np.ones((n_tb, seg["C"]), np.uint8)
        want[1, -1] = 0
        assert np.array_equal(cb_ok2.cpu().numpy(), want)
        assert torch.equal(out2[0], payload[0]) and torch.equal(out2[2], payload[2])


@pytest.mark.parametrize("K,which,ebn0", [(6144, "crc24b", 1.0), (5824, "crc24b", 1.2), (1056, "crc24a", 1.8), (104, "crc24a", 3.5)])
def test_crc_stopping_rule_logmap_s16_bit_exact(oracle, K, which, ebn0):
    """The CRC stopping rule in the Log-MAP kernels (a run-time mode of their one pass): decisions and iteration counts
    equal the integer model's (logmap = 1, early_term = 2) block by block, for compile-time geometry (6144), the
    run-time-P instantiations (5824, 1056) and the fully run-time kernel (104)."""
    torch = _torch_cuda()
    from oracle_lib import FxParams
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import CRC24A, CRC24B
    n_cb, n_iter = 5, 8
    dec = TurboDecoder(K, n_iter=n_iter, early_term=which, max_batch=8, algo="logmap_s16")
    hda = TurboDecoder(K, n_iter=n_iter, early_term=True, max_batch=8, algo="logmap_s16")
    g = torch.Generator(device="cuda")
    g.manual_seed(K + 1)
    bits = torch.randint(0, 2, (n_cb, K), dtype=torch.uint8, device="cuda", generator=g)
    no_crc = bits[2].clone()
    dec.crc24_attach(bits, CRC24B if which == "crc24b" else CRC24A)
    bits[2] = no_crc                                  # this one does not divide by the generator
    llr = dec.channel(dec.encode(bits), oracle.sigma(ebn0, K), seed=9)
    out = dec.decode(llr, want=("bits", "iters_used"))
    got_bits, got_it = out["bits"].cpu().numpy(), out["iters_used"].cpu().numpy()
    plan = dec.plan()
    poly = oracle.CRC24B if which == "crc24b" else oracle.CRC24A
    prm = FxParams(K=K, n_iter=n_iter, sub_len=plan["sub_block"], warmup=plan["warmup"], frac_bits=4, llr_clip=127,
                   ext_clip=511, ext_scale_q2=4, early_term=2, et_threshold=128, crc_poly=poly, logmap=1, lm_upper_off=1)
    pi = oracle.qpp(K)
    h_llr = llr.cpu().numpy()
    for c in range(n_cb):
        want_bits, _, it, ovf = oracle.fx_decode(h_llr[c], pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(got_bits[c], want_bits.astype(np.uint8)), "cb %d" % c
        assert int(got_it[c]) == it, "cb %d: %d vs %d iterations" % (c, got_it[c], it)
    assert got_it[2] == n_iter
    assert np.array_equal(got_bits, bits.cpu().numpy()), "operating point decodes cleanly"
    it_hda = hda.decode(llr, want=("iters_used",))["iters_used"].cpu().numpy()
    keep = np.arange(n_cb) != 2
    assert got_it[keep].mean() <= it_hda[keep].mean(), "the CRC rule never needs more iterations than decisions + magnitude"


@pytest.mark.parametrize("K,which,ebn0", [(6144, "crc24b", 1.0), (5824, "crc24b", 1.2), (1056, "crc24a", 1.8), (512, "crc24a", 2.2), (104, "crc24a", 3.5)])
def test_crc_stopping_rule_bit_exact(oracle, K, which, ebn0):
    """early_term = CRC: decisions and the iteration count equal the integer model's, block by block; a block
    whose payload carries no valid CRC runs all iterations and comes out as without early termination."""
    torch = _torch_cuda()
    from oracle_lib import FxParams
    from turbo_decoder_cuda_b200 import TurboDecoder
    from turbo_decoder_cuda_b200.decoder import CRC24A, CRC24B
    n_cb, n_iter = 7, 8
    dec = TurboDecoder(K, n_iter=n_iter, early_term=which, max_batch=8)
    plain = TurboDecoder(K, n_iter=n_iter, max_batch=8)
    hda = TurboDecoder(K, n_iter=n_iter, early_term=True, max_batch=8)
    g = torch.Generator(device="cuda")
    g.manual_seed(K)
    bits = torch.randint(0, 2, (n_cb, K), dtype=torch.uint8, device="cuda", generator=g)
    no_crc = bits[3].clone()
    dec.crc24_attach(bits, CRC24B if which == "crc24b" else CRC24A)
    bits[3] = no_crc                                  # this one does not divide by the generator
    llr = dec.channel(dec.encode(bits), oracle.sigma(ebn0, K), seed=5)
    out = dec.decode(llr, want=("bits", "iters_used"))
    got_bits, got_it = out["bits"].cpu().numpy(), out["iters_used"].cpu().numpy()
    plan = dec.plan()
    poly = oracle.CRC24B if which == "crc24b" else oracle.CRC24A
    prm = FxParams(K=K, n_iter=n_iter, sub_len=plan["sub_block"], warmup=plan["warmup"], frac_bits=3, llr_clip=127,
                   ext_clip=511, ext_scale_q2=3, early_term=2, et_threshold=64, crc_poly=poly)
    pi = oracle.qpp(K)
    h_llr = llr.cpu().numpy()
    for c in range(n_cb):
        want_bits, _, it, ovf = oracle.fx_decode(h_llr[c], pi, prm, want_le=True)
        assert ovf == 0
        assert np.array_equal(got_bits[c], want_bits.astype(np.uint8)), "cb %d" % c
        assert int(got_it[c]) == it, "cb %d: %d vs %d iterations" % (c, got_it[c], it)
    assert got_it[3] == n_iter and np.array_equal(got_bits[3], plain.decode(llr)["bits"].cpu().numpy()[3])
    assert np.array_equal(got_bits, bits.cpu().numpy()), "operating point decodes cleanly"
    it_hda = hda.decode(llr, want=("iters_used",))["iters_used"].cpu().numpy()
    keep = np.arange(n_cb) != 3
    assert got_it[keep].mean() <= it_hda[keep].mean(), "the CRC rule never needs more iterations than decisions + magnitude"
    # host-memory path and an odd batch of one
    assert np.array_equal(dec.decode(h_llr[:1])["bits"], got_bits[:1])
    from turbo_decoder_cuda_b200.decoder import TdbError
    with pytest.raises(TdbError):
        dec.decode(llr, want=("bits", "ext_siso2"))   # a block may stop after SISO-1: no SISO-2 extrinsic to report
