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
#      restatement of the specification.  Pinned by a hand-derived known answer for K = 40 (below); no 3GPP
#      conformance vector and no independent implementation exists offline.
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


def test_rate_matching_known_answer_k40_derived_from_ts36212(oracle):
    """A known answer worked out BY HAND from the text of TS 36.212 (v8.8.0) for the smallest block, K = 40 -- not
    produced by any code in this repository.  It pins the sub-block interleaver, the parity interlacing, the tail-bit
    multiplexing and the redundancy-version start points of oracle/turbo_oracle_rm.c (and through the bit-exact GPU
    tests, of build_rm_table in csrc/tdb200_ratematch.cu).

    Derivation.
    5.1.3.2.2 (trellis termination): each stream d(i) has D = K + 4 = 44 bits; d0[k] = x_k, d1[k] = z_k, d2[k] = z'_k for
      k < K, then  d0[K..K+3] = x_K, z_{K+1}, x'_K, z'_{K+1};  d1[K..K+3] = z_K, x_{K+2}, z'_K, x'_{K+2};
      d2[K..K+3] = x_{K+1}, z_{K+2}, x'_{K+1}, z'_{K+2}.
      In the reference's multiplex order (ITTC/log_map.cpp:566-578) x_k, z_k, z'_k sit at 3k, 3k+1, 3k+2 and the tail
      (x_K z_K x_{K+1} z_{K+1} x_{K+2} z_{K+2} | x'_K z'_K x'_{K+1} z'_{K+1} x'_{K+2} z'_{K+2}) at 3K .. 3K+11.
    5.1.4.1.1 (sub-block interleaver): C = 32 columns, R = ceil(44/32) = 2 rows, K_pi = 64, N_D = 64 - 44 = 20 <NULL>s
      in front: y[k] = <NULL> for k < 20, y[20+k] = d[k].  Written row by row (row 0 = y[0..31], row 1 = y[32..63]),
      columns permuted by P = <0,16,8,24,4,20,12,28,2,18,10,26,6,22,14,30,1,17,9,25,5,21,13,29,3,19,11,27,7,23,15,31>,
      read column by column:  v0 = y[0], y[32], y[16], y[48], y[8], y[40], y[24], y[56], y[4], y[36], y[20], y[52], ...
      i.e. in terms of d0:      N,   d12,   N,     d28,   N,    d20,   d4,    d36,   N,    d16,   d0,    d32,  ...
      d2 uses v2[k] = y[pi(k)], pi(k) = (P(floor(k/R)) + C*(k mod R) + 1) mod K_pi:
      pi = 1, 33, 17, 49, 9, 41, 25, 57, ...  ->  v2 = N, d2[13], N, d2[29], N, d2[21], d2[5], d2[37], ...
    5.1.4.1.2 (bit collection): w[k] = v0[k] (k < 64), w[64+2k] = v1[k], w[64+2k+1] = v2[k]; K_w = 192.
      Bit selection starts at k0 = R * (2 * ceil(N_cb / (8R)) * rv + 2) = 2 * (24 rv + 2) = 4, 52, 100, 148 and skips <NULL>s.
    """
    K = 40
    assert oracle.rm_geometry(K) == {"R": 2, "Kpi": 64, "ND": 20, "Kw": 192}
    assert [oracle.rm_k0(K, rv) for rv in range(4)] == [4, 52, 100, 148]
    N = -1
    d0 = lambda k: 3 * k if k < K else {40: 3 * K + 0, 41: 3 * K + 3, 42: 3 * K + 6, 43: 3 * K + 9}[k]     # noqa: E731
    d1 = lambda k: 3 * k + 1 if k < K else {40: 3 * K + 1, 41: 3 * K + 4, 42: 3 * K + 7, 43: 3 * K + 10}[k]  # noqa: E731
    d2 = lambda k: 3 * k + 2 if k < K else {40: 3 * K + 2, 41: 3 * K + 5, 42: 3 * K + 8, 43: 3 * K + 11}[k]  # noqa: E731
    w = oracle.rm_circular_buffer(K)
    # systematic part, first 24 entries (columns 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26)
    assert list(w[:24]) == [N, d0(12), N, d0(28), N, d0(20), d0(4), d0(36), N, d0(16), d0(0), d0(32),
                            N, d0(24), d0(8), d0(40), N, d0(14), N, d0(30), N, d0(22), d0(6), d0(38)]
    # the last column read is 31: y[31] = d0[11], y[63] = d0[43] = z'_{K+1}
    assert list(w[62:64]) == [d0(11), d0(43)]
    # interlaced parity part: (v1[k], v2[k]) pairs for k = 0..7
    assert list(w[64:80]) == [N, N, d1(12), d2(13), N, N, d1(28), d2(29), N, N, d1(20), d2(21), d1(4), d2(5), d1(36), d2(37)]
    # the wrap of pi(): k = 63 -> P(31) + 32 + 1 = 64 = 0 (mod 64): the very last entry is y2[0], a <NULL>
    assert w[191] == N and w[190] == d1(43)
    # rv 0 starts at w[4]: the first 20 transmitted bits (hand-listed from the columns above, <NULL>s skipped)
    sel = oracle.rm_selection(K, 20, 0)
    assert list(sel) == [d0(20), d0(4), d0(36), d0(16), d0(0), d0(32), d0(24), d0(8), d0(40), d0(14), d0(30), d0(22),
                         d0(6), d0(38), d0(18), d0(2), d0(34), d0(26), d0(10), d0(42)]
    # rv 1 starts at w[52] = column P(26) = 11, row 0: y[11] is <NULL>, so the first bit sent is y[43] = d0[23],
    # followed by column 27: y[27] = d0[7], y[59] = d0[39]
    assert list(oracle.rm_selection(K, 3, 1)) == [d0(23), d0(7), d0(39)]
    # rv 2 starts at w[100] = parity pair k = 18 (column P(9) = 18, row 0): v1 = y1[18] <NULL>, v2 = y2[pi(18)] with
    # pi(18) = 18 + 0 + 1 = 19 <NULL>; pair k = 19 (row 1): v1 = y1[50] = d1[30], v2 = y2[18 + 32 + 1 = 51] = d2[31]
    assert list(oracle.rm_selection(K, 2, 2)) == [d1(30), d2(31)]


def test_rate_matching_filler_bits_k40_by_hand(oracle):
    """F filler bits (TS 36.212 5.1.2; 5.1.3.2.1: d0_k = d1_k = <NULL> for k < F, d2 is transmitted).  Continues the hand
    derivation of the K = 40 case above with F = 8: the <NULL>s of d0[0..7] and d1[0..7] join the 20 dummy ones of each
    stream, so in the listing of w[0..23] above d0(4), d0(0), d0(6) disappear, and so does d1(4) among the first parity
    pairs -- d2(5) next to it stays."""
    K, F = 40, 8
    N = -1
    d0 = lambda k: 3 * k if k < K else {40: 3 * K + 0, 41: 3 * K + 3, 42: 3 * K + 6, 43: 3 * K + 9}[k]     # noqa: E731
    d1 = lambda k: 3 * k + 1                                                                                # noqa: E731
    d2 = lambda k: 3 * k + 2                                                                                # noqa: E731
    w = oracle.rm_circular_buffer_f(K, F)
    assert list(w[:24]) == [N, d0(12), N, d0(28), N, d0(20), N, d0(36), N, d0(16), N, d0(32),
                            N, d0(24), d0(8), d0(40), N, d0(14), N, d0(30), N, d0(22), N, d0(38)]
    assert list(w[64:80]) == [N, N, d1(12), d2(13), N, N, d1(28), d2(29), N, N, d1(20), d2(21), N, d2(5), d1(36), d2(37)]
    # rv 0 from w[4]: d0(20), [d0(4) is a filler], d0(36), d0(16), [d0(0)], d0(32), d0(24), d0(8), ...
    assert list(oracle.rm_selection_f(K, 8, 0, 0, F)) == [d0(20), d0(36), d0(16), d0(32), d0(24), d0(8), d0(40), d0(14)]
    for K2, F2 in ((40, 8), (512, 40), (6144, 56), (1024, 1023)):
        g = oracle.rm_geometry(K2)
        w0, wf = oracle.rm_circular_buffer(K2), oracle.rm_circular_buffer_f(K2, F2)
        gone = set(w0[w0 != wf])                                   # what the filler bits removed from the buffer
        assert gone == {3 * k for k in range(F2)} | {3 * k + 1 for k in range(F2)}
        assert np.count_nonzero(wf < 0) == 3 * g["ND"] + 2 * F2
        # one full wrap sends every remaining bit exactly once, and no filler position
        n_left = 3 * K2 + 12 - 2 * F2
        sel = oracle.rm_selection_f(K2, n_left, 0, 0, F2)
        assert len(set(sel)) == n_left and not (set(sel) & gone)
        # soft inverse: filler positions (systematic and parity 1) hold the fixed value, everything else as without fillers
        e = np.random.default_rng(K2).standard_normal(n_left)
        back = oracle.rate_dematch_f(e, K2, 0, 0, F2, fill=-100.0)
        assert all(back[3 * k] == -100.0 and back[3 * k + 1] == -100.0 for k in range(F2))
        assert np.array_equal(back[sel], e)
    assert np.array_equal(oracle.rm_circular_buffer_f(40, 0), oracle.rm_circular_buffer(40))


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
