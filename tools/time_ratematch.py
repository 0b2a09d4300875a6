"""Device time of the TS 36.212 rate-matching kernels and of decoding from rate-matched LLRs at the
BASELINE block size (K = 6144, 4096 codeblocks, 8 iterations), CUDA events on the launching stream.

    python tools/time_ratematch.py [--json out.json]

Per code rate: rate_match (bit gather), rate_dematch (float in / float out and float in / 8-bit
hand-over to the decoder; algorithmic bytes = E values read + 3K+12 values written per codeblock),
and decode_rm against decode on already de-rate-matched LLRs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder  # noqa: E402

K, N, IT = 6144, 4096, 8
NL = 3 * K + 12


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    a = ap.parse_args()
    dec = TurboDecoder(K, n_iter=IT, max_batch=N)
    bits = torch.randint(0, 2, (N, K), dtype=torch.uint8, device="cuda")
    coded = dec.encode(bits)
    rows = []
    for name, E, ebn0 in (("1/3", NL, 1.6), ("1/2", 2 * K, 2.5), ("3/4", 4 * K // 3, 4.5), ("1/5 (repetition)", 5 * K, 1.6)):
        rate = K / E
        sigma = 10 ** (-ebn0 / 20) * np.sqrt(0.5 / rate)
        tx = dec.rate_match(coded, E, 0)
        ms_rm = timed(lambda: dec.rate_match(coded, E, 0))
        r = dec.awgn(tx.float() * 2 - 1, sigma, seed=1)
        e_llr = r * (2.0 / (sigma * sigma))
        ms_dm = timed(lambda: dec.rate_dematch(e_llr, 0))
        llr = dec.rate_dematch(e_llr, 0)
        ms_dec = timed(lambda: dec.decode(llr), reps=5)
        ms_drm = timed(lambda: dec.decode_rm(e_llr, 0), reps=5)
        out = dec.decode_rm(e_llr, 0)["bits"]
        row = {"rate": name, "E": E, "ebn0_db": ebn0, "fer": float((out != bits).any(dim=1).float().mean()),
               "rate_match_ms": ms_rm, "rate_dematch_f32_ms": ms_dm,
               "rate_dematch_f32_gb_s": N * 4 * (E + NL) / ms_dm / 1e6,
               "decode_ms": ms_dec, "decode_rm_ms": ms_drm, "dematch_to_s8_ms": ms_drm - ms_dec,
               "decode_rm_gbit_s": N * K / ms_drm / 1e6}
        rows.append(row)
        print(json.dumps(row))
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"K": K, "codeblocks": N, "iterations": IT, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
