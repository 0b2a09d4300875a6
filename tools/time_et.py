"""Early-termination rules compared on the same traffic (K = 6144, 4096 codeblocks carrying a CRC24B, at most
8 iterations): none, decisions + magnitude, CRC.  python tools/time_et.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder  # noqa: E402
from turbo_decoder_cuda_b200.decoder import CRC24B  # noqa: E402
from turbo_decoder_cuda_b200.synth import sigma_from_ebn0  # noqa: E402

K, N = 6144, 4096


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
    decs = {"none": TurboDecoder(K, max_batch=N), "decisions+magnitude": TurboDecoder(K, early_term=True, max_batch=N),
            "crc24b": TurboDecoder(K, early_term="crc24b", max_batch=N)}
    d0 = decs["none"]
    bits = torch.randint(0, 2, (N, K), dtype=torch.uint8, device="cuda")
    d0.crc24_attach(bits, CRC24B)
    coded = d0.encode(bits)
    rows = []
    for ebn0 in (0.6, 0.8, 1.0, 1.5, 2.0):
        llr = d0.channel(coded, sigma_from_ebn0(ebn0, K), seed=int(ebn0 * 10))
        row = {"ebn0_db": ebn0}
        for name, dec in decs.items():
            out = dec.decode(llr, want=("bits", "iters_used"))
            ms = timed(lambda: dec.decode(llr), reps=5)
            row[name] = {"gbit_s": N * K / ms / 1e6, "mean_iters": float(out["iters_used"].float().mean()),
                         "fer": float((out["bits"] != bits).any(dim=1).float().mean())}
        rows.append(row)
        print(json.dumps(row))
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"K": K, "codeblocks": N, "max_iterations": 8, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
