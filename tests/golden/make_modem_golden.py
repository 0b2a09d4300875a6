"""Generate tests/golden/modem_golden.npz from the reference's own mapper and soft demapper.

Run in the dev container only (needs /root/reference, compiled in place into oracle/_ref by
oracle/Makefile):      python tests/golden/make_modem_golden.py
For each modulation M (the reference's modu_index: 1 BPSK, 2 QPSK, 3 8PSK, 4 16QAM, 6 64QAM):
  bits_M          [264]      seeded random bits (264 = 22 groups of 12: a multiple of every M)
  si_M, sq_M      [264/M]    what module() (ITTC/modanddem.cpp:175) returned for them
  ri_M, rq_M      [264/M]    the same symbols plus seeded Gaussian noise, stored as float32-
                             representable values (so the fp32 device path reads the same numbers)
  llr_M           [264]      what demodule() (modanddem.cpp:674) returned for (ri, rq) at Kf = kf_M
Every constellation point appears at least once for M <= 4; for 64QAM a second block `all_*_6`
holds all 64 index values in order.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefLib  # noqa: E402

SIGMA = {1: 0.8, 2: 0.6, 3: 0.4, 4: 0.3, 6: 0.15}


def main():
    assert RefLib.available(), "oracle/_ref is not built (needs /root/reference)"
    r = RefLib(40, 3, 10)
    rng = np.random.default_rng(20260)
    out = {}
    for M in (1, 2, 3, 4, 6):
        bits = rng.integers(0, 2, 264).astype(np.int32)
        si, sq = r.module(bits, M)
        sg = SIGMA[M]
        ri = (si + sg * rng.standard_normal(si.size)).astype(np.float32).astype(np.float64)
        rq = (sq + sg * rng.standard_normal(sq.size)).astype(np.float32).astype(np.float64)
        kf = 1.0 / (2.0 * sg * sg)
        llr = r.demodule(ri, rq, M, kf)
        out.update({"bits_%d" % M: bits.astype(np.uint8), "si_%d" % M: si, "sq_%d" % M: sq,
                    "ri_%d" % M: ri, "rq_%d" % M: rq, "llr_%d" % M: llr, "kf_%d" % M: np.float64(kf)})
    idx = np.arange(64)
    bits = ((idx[:, None] >> np.arange(5, -1, -1)) & 1).astype(np.int32).ravel()
    si, sq = r.module(bits, 6)
    out.update({"all_bits_6": bits.astype(np.uint8), "all_si_6": si, "all_sq_6": sq})
    path = os.path.join(HERE, "modem_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
