"""Generate the golden vectors of tests/golden/ from the reference's own code.

Run in the dev container only (needs /root/reference, compiled in place into oracle/_ref by
oracle/Makefile):      python tests/golden/make_golden.py
Each fixture holds float32-representable channel LLRs (so they store compactly and convert to the
reference's double exactly) and what the UNMODIFIED reference decoder returned for them:
  bits  [n_iter, K]  hard decisions per iteration  (TurboDecoding's flow_decoded, log_map.cpp:1264)
  llr1, llr2, le     last-iteration a-posteriori (SISO-1 / SISO-2) and extrinsic LLRs, taken by
                     re-stating the TurboDecoding loop around the reference's Log_MAP_decoder
                     (oracle/ref_harness.cpp: ref_decode_iters); the first 8 rows of `bits` are
                     also cross-checked against the reference's own 15-iteration TurboDecoding().
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, RefLib  # noqa: E402

CASES = [  # name, K, Eb/N0 dB, n_iter, seed
    ("k40_2p0dB", 40, 2.0, 8, 11),
    ("k512_1p0dB", 512, 1.0, 8, 12),
    ("k6144_1p0dB", 6144, 1.0, 8, 13),   # BASELINE configs[0]
    ("k6144_0p4dB", 6144, 0.4, 8, 14),   # waterfall region
]


def main():
    assert RefLib.available(), "oracle/_ref is not built (needs /root/reference)"
    o = Oracle()
    for name, K, ebn0, n_iter, seed in CASES:
        f1, f2 = o.lte_params(K)
        r = RefLib(K, f1, f2)
        bits_tx, llr = o.make_batch(K, 1, ebn0, seed=seed)
        llr32 = llr[0].astype(np.float32)
        bits, l1, l2, le = r.decode(llr32.astype(np.float64), n_iter, want_llr=True)
        full = r.turbo_decoding(llr32.astype(np.float64))
        assert np.array_equal(full[:n_iter], bits)
        # guard against the reference's uninitialised tempmax[] (log_map.cpp:925,989) having picked up
        # heap garbage in this process (oracle/ref_harness.cpp explains): a sane run agrees with the
        # C restatement to rounding noise
        ob, o1, o2, ole = o.decode(llr32.astype(np.float64), o.qpp(K), n_iter, want_llr=True)
        assert np.array_equal(ob, bits) and np.abs(o2 - l2).max() < 1e-9 and np.abs(ole - le).max() < 1e-9
        np.savez_compressed(os.path.join(HERE, name + ".npz"), K=K, f1=f1, f2=f2, ebn0=ebn0, n_iter=n_iter,
                            tx_bits=np.packbits(bits_tx[0].astype(np.uint8)), llr=llr32,
                            bits=np.packbits(bits.astype(np.uint8), axis=1), llr1=l1, llr2=l2, le=le)
        print(name, "bit errors per iteration:", (bits != bits_tx[0]).sum(1))
    # max* LUT: the reference's E_algorithm on a grid, including every breakpoint
    r = RefLib(40, 3, 10)
    xs = np.concatenate([np.linspace(-6, 6, 241), [0.08824, 0.19587, 0.31026, 0.43275, 0.56508, 0.70963, 0.86972,
                                                   1.0502, 1.2587, 1.5078, 1.8212, 2.2522, 2.9706, 3.6764, 4.3758]])
    ys = np.array([r.max_star(0.0, float(x)) for x in xs])
    np.savez_compressed(os.path.join(HERE, "max_star_lut.npz"), x=xs, y=ys)


if __name__ == "__main__":
    main()
