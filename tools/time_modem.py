"""Device time of the mapper / demapper kernels and of symbol-input decoding at the BASELINE block
size (K = 6144, 4096 codeblocks, 8 iterations), CUDA events on the launching stream.

    python tools/time_modem.py [--json out.json]

Per modulation: the fp32 demapper to the decoder's 8-bit channel values (HBM-bound: algorithmic bytes
= two symbol planes read + one byte per LLR written), decode from device-resident symbols, and the
end-to-end rate from pinned HOST symbols (float and half) next to the rate from host float LLRs --
the PCIe bytes per LLR are what changes.
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
    rate = K / NL
    rows = []
    ebn0 = {1: 1.6, 2: 1.6, 3: 4.0, 4: 4.0, 6: 6.0}
    # reference point: float LLRs from host memory (the bench.py e2e leg)
    out_host = torch.empty((N, K), dtype=torch.uint8).pin_memory()
    for M in (1, 2, 3, 4, 6):
        sigma = 10 ** (-ebn0[M] / 20) * np.sqrt(0.5 / (rate * M))
        kf = 1.0 / (2 * sigma * sigma)
        ms_mod = timed(lambda: dec.modulate(coded, M))
        si, sq = dec.modulate(coded, M)
        ri, rq = dec.awgn(si, sigma, seed=1), dec.awgn(sq, sigma, seed=2)
        ms_awgn = timed(lambda: dec.awgn(si, sigma, seed=1))
        ms_dem = timed(lambda: dec.demap(ri, rq, M, kf, dtype="int8"))
        ms_dem64 = timed(lambda: dec.demap(ri, rq, M, kf, dtype="float64"), reps=3, warm=1)
        llr8 = dec.demap(ri, rq, M, kf, dtype="int8")
        ms_dec = timed(lambda: dec.decode(llr8), reps=5)
        ms_sym = timed(lambda: dec.decode_symbols(ri, rq, M, kf), reps=5)
        ber = float((dec.decode_symbols(ri, rq, M, kf)["bits"] != bits).float().mean())
        row = {"modulation": M, "ebn0_db": ebn0[M], "ber": ber,
               "modulate_ms": ms_mod, "awgn_ms_per_plane": ms_awgn,
               "demap_s8_ms": ms_dem, "demap_s8_gb_s": N * (2 * 4 * NL / M + NL) / ms_dem / 1e6,
               "demap_f64_ms": ms_dem64,
               "decode_from_s8_llr_ms": ms_dec, "decode_symbols_ms": ms_sym,
               "decode_symbols_gbit_s": N * K / ms_sym / 1e6}
        # end to end from pinned host symbols, float and half
        for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
            hi, hq = ri.to(dt).cpu().pin_memory(), rq.to(dt).cpu().pin_memory()
            st = {"f32": 1, "f16": 3}[name]   # TDB200_LLR_F32 / TDB200_LLR_F16
            ms = timed(lambda: dec.decode_symbols_raw(hi.data_ptr(), hq.data_ptr(), st, 0, N, M, kf, bits=out_host.data_ptr()), reps=5)
            row["e2e_host_symbols_%s_gbit_s" % name] = N * K / ms / 1e6
            row["h2d_bytes_per_cb_%s" % name] = 2 * hi.element_size() * NL // M
        rows.append(row)
        print(json.dumps(row))
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"K": K, "codeblocks": N, "iterations": IT, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
