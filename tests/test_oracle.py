"""CPU tests of the oracle (oracle/turbo_oracle.c): pinned against the golden vectors generated
from the reference's own code (tests/golden/make_golden.py), against known answers of the code
(SURVEY.md 4 / Appendix B), and -- where oracle/_ref was built -- against the reference live."""
import glob
import os

import numpy as np
import pytest

from oracle_lib import ALGO_MAXLOG, RefLib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ------------------------------------------------------------------ constants
def test_trellis_literals(oracle):
    # dumped from the running reference == turboDecoderBianJieZhi.cu:208-226
    no, ns, lo, ls = oracle.trellis()
    assert ns[:, 0].tolist() == [0, 4, 5, 1, 2, 6, 7, 3]
    assert ns[:, 1].tolist() == [4, 0, 1, 5, 6, 2, 3, 7]
    assert ls[:, 0].tolist() == [0, 3, 4, 7, 1, 2, 5, 6]
    assert ls[:, 1].tolist() == [1, 2, 5, 6, 0, 3, 4, 7]
    assert no[:, 1].tolist() == [-1, -1, 1, 1, 1, 1, -1, -1]
    assert (no[:, 3] == -no[:, 1]).all() and (no[:, 0] == -1).all() and (no[:, 2] == 1).all()


def test_qpp_known_values_and_table(oracle):
    pi = oracle.qpp(6144)
    assert pi[:8].tolist() == [0, 743, 2446, 5109, 2588, 1027, 426, 785]  # TurboDecoder.cu:60
    assert oracle.lte_params(6144) == (263, 480) and oracle.lte_params(2688) == (127, 504)  # main.cpp:17-19,36-37
    sizes = oracle.lte_sizes()
    expect = list(range(40, 512, 8)) + list(range(512, 1024, 16)) + list(range(1024, 2048, 32)) + list(range(2048, 6145, 64))
    assert sizes == expect and len(sizes) == 188
    for K in sizes:
        assert np.array_equal(np.sort(oracle.qpp(K)), np.arange(K)), "pi not a bijection for K=%d" % K


def test_impulse_encode(oracle):
    K = 6144
    bits = np.zeros(K, np.int32)
    bits[0] = 1
    c = oracle.encode(bits, oracle.qpp(K))
    assert "".join(map(str, c[:12])) == "111011011011"
    assert "".join(map(str, c[-12:])) == "000111000111"
    z = oracle.encode(np.zeros(K, np.int32), oracle.qpp(K))
    assert not z.any()


# ------------------------------------------------------------------ golden vectors from the reference
def test_max_star_lut_golden(oracle):
    g = np.load(os.path.join(GOLD, "max_star_lut.npz"))
    for x, y in zip(g["x"], g["y"]):
        assert oracle.max_star(0.0, float(x)) == y
        assert oracle.max_star(float(x), 0.0) == y        # symmetric
    assert oracle.max_star(1.0, 1.0) == 1.0 + 0.69315
    assert oracle.max_star(-3.0, 5.0) == 5.0              # |d| >= 4.3758 -> no correction


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "k*.npz"))))
def test_decode_matches_reference_golden(oracle, path):
    g = np.load(path)
    K, n_iter = int(g["K"]), int(g["n_iter"])
    pi = oracle.qpp(K, int(g["f1"]), int(g["f2"]))
    llr = g["llr"].astype(np.float64)
    bits, l1, l2, le = oracle.decode(llr, pi, n_iter, want_llr=True)
    ref_bits = np.unpackbits(g["bits"], axis=1)[:, :K]
    assert np.array_equal(bits, ref_bits)
    # identical operation order -> rounding-noise agreement (exactly 0 for the large blocks; the
    # reference's uninitialised tempmax[] leaves ~1e-13 on K=40, see oracle/ref_harness.cpp)
    for a, b in ((l1, g["llr1"]), (l2, g["llr2"]), (le, g["le"])):
        assert np.abs(a - b).max() < 1e-9
    tx = np.unpackbits(g["tx_bits"])[:K]
    assert np.array_equal(bits[-1], tx), "fixture decodes cleanly by the last iteration"


# ------------------------------------------------------------------ live reference (dev container only)
@pytest.mark.skipif(not RefLib.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference(oracle):
    K = 1024
    f1, f2 = oracle.lte_params(K)
    r = RefLib(K, f1, f2)
    pi = oracle.qpp(K)
    assert np.array_equal(pi, r.qpp())
    for a, b in zip(oracle.trellis(), r.trellis()):
        assert np.array_equal(a, b)
    bits_tx, llr = oracle.make_batch(K, 3, 0.8, seed=21)
    for c in range(3):
        assert np.array_equal(oracle.encode(bits_tx[c], pi), r.encode(bits_tx[c]))
        rb, r1, r2, rle = r.decode(llr[c], 6, want_llr=True)
        ob, o1, o2, ole = oracle.decode(llr[c], pi, 6, want_llr=True)
        assert np.array_equal(rb, ob)
        assert max(np.abs(r1 - o1).max(), np.abs(r2 - o2).max(), np.abs(rle - ole).max()) < 1e-9
        assert np.array_equal(r.turbo_decoding(llr[c])[:6], rb)   # the reference's own 15-iteration entry point
    rng = np.random.default_rng(0)
    T = K + 3
    recs, La = rng.normal(0, 2, 2 * T), rng.normal(0, 3, T)
    for term in (1, 0):
        assert np.abs(r.siso(recs, La, term) - oracle.siso(recs, La, terminated=term, tempmax_floor=0.0)).max() < 1e-9


# ------------------------------------------------------------------ properties
def test_noiseless_and_erasure_properties(oracle):
    K = 512
    pi = oracle.qpp(K)
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, K, dtype=np.int32)
    coded = oracle.encode(bits, pi)
    llr = (2.0 * coded - 1.0) * 4.0
    out = oracle.decode(llr, pi, 2)
    assert np.array_equal(out[0], bits) and np.array_equal(out[1], bits)
    # erase every parity-2 value and 10 % of the rest: the code still recovers the block
    llr[2:3 * K:3] = 0.0
    llr[rng.random(llr.size) < 0.1] = 0.0
    assert np.array_equal(oracle.decode(llr, pi, 8)[-1], bits)
    # all-zero input: LLR == 0 everywhere -> every decision is 1 (decision(): LLR<0 -> 0 else 1, :869-877)
    assert oracle.decode(np.zeros(3 * K + 12), pi, 1)[0].all()


def test_maxlog_is_logmap_without_correction(oracle):
    """max-log SISO output differs from the LUT Log-MAP by a bounded amount (|corr| <= 0.69315 per max*)."""
    rng = np.random.default_rng(2)
    T = 259
    recs, La = rng.normal(0, 1.5, 2 * T), rng.normal(0, 1.0, T)
    a = oracle.siso(recs, La)
    b = oracle.siso(recs, La, algo=ALGO_MAXLOG)
    assert 0 < np.abs(a - b).max() < 8.0
    assert (np.sign(a) == np.sign(b)).mean() > 0.9


def test_batch_threads_agree(oracle):
    K = 256
    pi = oracle.qpp(K)
    bits, llr = oracle.make_batch(K, 6, 1.5, seed=3)
    b1, _ = oracle.decode_batch(llr, pi, 4, n_threads=1)
    b4, _ = oracle.decode_batch(llr, pi, 4, n_threads=4)
    assert np.array_equal(b1, b4)
    for c in range(6):
        assert np.array_equal(oracle.decode(llr[c], pi, 4)[-1], b1[c])


# ---- mapper / soft demapper: oracle/turbo_oracle_mod.c against the reference's
#      module()/demodule() -- golden vectors always, the live reference build where it exists
MODS = (1, 2, 3, 4, 6)


@pytest.mark.parametrize("M", MODS)
def test_modem_matches_reference_golden(oracle, M):
    g = np.load(os.path.join(GOLD, "modem_golden.npz"))
    si, sq = oracle.modulate(g["bits_%d" % M], M)
    assert np.array_equal(si, g["si_%d" % M]) and np.array_equal(sq, g["sq_%d" % M])
    llr = oracle.demap_f64(g["ri_%d" % M], g["rq_%d" % M], M, float(g["kf_%d" % M]))
    assert np.array_equal(llr, g["llr_%d" % M]), "demodule() restatement must be bit-identical"
    # the fp32 model of the device demapper: same metric, float rounding only
    l32 = oracle.demap_f32(g["ri_%d" % M], g["rq_%d" % M], M, float(g["kf_%d" % M]))
    assert np.abs(l32 - llr).max() <= 2e-5 * max(1.0, np.abs(llr).max())
    # and its 8-bit hand-over differs from quantising the reference's doubles by at most one step
    q = oracle.quant_s8(l32).astype(int)
    qref = np.clip(np.rint(llr * 8), -127, 127).astype(int)
    assert np.abs(q - qref).max() <= 1 and np.mean(q != qref) < 0.01
    if M == 6:
        si, sq = oracle.modulate(g["all_bits_6"], 6)
        assert np.array_equal(si, g["all_si_6"]) and np.array_equal(sq, g["all_sq_6"])


@pytest.mark.skipif(not RefLib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("M", MODS)
def test_modem_vs_live_reference(oracle, M):
    ref = RefLib(40, 3, 10)
    rng = np.random.default_rng(100 + M)
    bits = rng.integers(0, 2, 12 * 500).astype(np.int32)
    si, sq = oracle.modulate(bits, M)
    ri, rq = ref.module(bits, M)
    assert np.array_equal(si, ri) and np.array_equal(sq, rq)
    for sigma in (0.05, 0.3, 1.5):   # also far outside the constellation
        xi = si + sigma * rng.standard_normal(si.size)
        xq = sq + sigma * rng.standard_normal(si.size)
        kf = 1 / (2 * sigma ** 2)
        assert np.array_equal(oracle.demap_f64(xi, xq, M, kf), ref.demodule(xi, xq, M, kf))
    # noiseless: the hard decision of every LLR is the transmitted bit
    assert np.array_equal(oracle.demap_f64(si, sq, M, 1.0) > 0, bits == 1)


# ---- TS 36.212 rate matching: oracle/turbo_oracle_rm.c.  The reference only declares
#      rate_match()/de_rate_match() (ITTC/main.h:23-24), so these are structural checks of the literal
#      restatement of the specification -- parity with 3GPP test vectors is UNPINNED (none offline).
def test_rate_matching_structure(oracle):
    assert oracle.rm_geometry(40) == {"R": 2, "Kpi": 64, "ND": 20, "Kw": 192}
    assert oracle.rm_geometry(6144) == {"R": 193, "Kpi": 6176, "ND": 28, "Kw": 18528}
    for K in oracle.lte_sizes():
        g = oracle.rm_geometry(K)
        w = oracle.rm_circular_buffer(K)
        assert np.array_equal(np.sort(w[w >= 0]), np.arange(3 * K + 12)), "every coded bit sits in the buffer exactly once"
        assert np.count_nonzero(w < 0) == 3 * g["ND"]
        assert [oracle.rm_k0(K, rv) for rv in range(4)] == [g["R"] * (2 * -(-g["Kw"] // (8 * g["R"])) * rv + 2) for rv in range(4)]
        # systematic part first: the first K_pi entries hold d0 = the systematic bits and four tail bits
        s = w[:g["Kpi"]]
        s = s[s >= 0]
        tails = {3 * K + 0, 3 * K + 3, 3 * K + 6, 3 * K + 9}         # x_K, z_K+1, x'_K, z'_K+1
        assert all((v % 3 == 0 and v < 3 * K) or v in tails for v in s) and s.size == K + 4
        # then d1 and d2 interlaced: parity 1 of step i next to parity 2 of a neighbouring step
        p = w[g["Kpi"]:]
        assert all(v < 0 or v >= 3 * K or v % 3 == 1 for v in p[0::2]) and all(v < 0 or v >= 3 * K or v % 3 == 2 for v in p[1::2])


def test_rate_matching_round_trips(oracle):
    rng = np.random.default_rng(3)
    K = 512
    NL = 3 * K + 12
    coded = rng.integers(0, 2, NL).astype(np.int32)
    for rv in range(4):
        full = oracle.rm_selection(K, NL, rv)
        assert np.array_equal(np.sort(full), np.arange(NL)), "one wrap sends every bit once, whatever the start"
        assert np.array_equal(oracle.rate_match(coded, K, NL, rv), coded[full])
    # puncturing: what was not sent comes back as 0, what was sent comes back unchanged
    E = K + 300
    sel = oracle.rm_selection(K, E, 0)
    e = rng.standard_normal(E)
    back = oracle.rate_dematch(e, K, 0)
    assert np.array_equal(back[sel], e) and np.count_nonzero(back) == E
    # repetition: the second wrap lands on the same positions and is summed
    E = NL + 100
    sel = oracle.rm_selection(K, E, 1)
    assert np.array_equal(sel[NL:], sel[:100])
    e = rng.standard_normal(E)
    back = oracle.rate_dematch(e, K, 1)
    want = np.zeros(NL)
    np.add.at(want, sel, e)
    assert np.allclose(back, want, rtol=0, atol=1e-12)
    # HARQ: a second transmission with another rv accumulates into the first
    e2 = rng.standard_normal(700)
    both = oracle.rate_dematch(e2, K, 2, into=back)
    np.add.at(want, oracle.rm_selection(K, 700, 2), e2)
    assert np.allclose(both, want, rtol=0, atol=1e-12)
    # limited soft buffer: positions beyond N_cb are never sent
    Ncb = 1200
    sel = oracle.rm_selection(K, 5000, 0, Ncb)
    w = oracle.rm_circular_buffer(K)
    assert set(sel) == set(w[:Ncb][w[:Ncb] >= 0])


def test_rate_matched_block_decodes(oracle):
    """rate 1/2 by puncturing (E = 2K) at 2.5 dB: encode -> rate match -> BPSK/AWGN -> de-rate-match -> decode."""
    K, rv = 512, 0
    pi = oracle.qpp(K)
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, K).astype(np.int32)
    coded = oracle.encode(bits, pi)
    E = 2 * K
    e = oracle.rate_match(coded, K, E, rv)
    sigma = 10 ** (-2.5 / 20) * np.sqrt(0.5 / 0.5)
    r = (2.0 * e - 1.0) + sigma * rng.standard_normal(E)
    llr = oracle.rate_dematch(2 * r / sigma ** 2, K, rv)
    out = oracle.decode(llr, pi, 8)
    assert np.array_equal(out[-1], bits)


# ---- CRC24A / CRC24B and code-block segmentation: oracle/turbo_oracle_crc.c
def test_crc24_catalogue_check_values(oracle):
    msg = np.unpackbits(np.frombuffer(b"123456789", np.uint8))
    assert oracle.crc24(msg, oracle.CRC24A) == 0xCDE703      # CRC-24/LTE-A
    assert oracle.crc24(msg, oracle.CRC24B) == 0x23EF52      # CRC-24/LTE-B
    rng = np.random.default_rng(0)
    for poly in (oracle.CRC24A, oracle.CRC24B):
        a = rng.integers(0, 2, 1000).astype(np.uint8)
        c = oracle.crc24(a, poly)
        full = np.concatenate([a, [(c >> (23 - i)) & 1 for i in range(24)]]).astype(np.uint8)
        assert oracle.crc24(full, poly) == 0, "payload || CRC divides by the generator"
        full[17] ^= 1
        assert oracle.crc24(full, poly) != 0
        assert oracle.crc24(np.concatenate([np.zeros(77, np.uint8), a]), poly) == c, "leading zeros (filler bits) are free"
