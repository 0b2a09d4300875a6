"""Throughput of the auto plan for every LTE block size (BASELINE configs[3]: all 188 K), 8 fixed iterations and with
the decisions + magnitude stopping rule at 1.5 dB; batch 16384 (65536 for K <= 512).
    python tools/sweep_all_sizes.py --json gpurun_out/all_sizes.json [--algo logmap_s16]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder, synth  # noqa: E402


def lte_sizes():
    return list(range(40, 512, 8)) + list(range(512, 1024, 16)) + list(range(1024, 2048, 32)) + list(range(2048, 6145, 64))


def timed(fn, reps=4):
    for _ in range(2):
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
    ap.add_argument("--json", default="gpurun_out/all_sizes.json")
    ap.add_argument("--algo", default="maxlog_s16", help="maxlog_s16 | logmap_s16")
    a = ap.parse_args()
    rows = []
    for K in lte_sizes():
        N = 16384 if K > 512 else 65536
        bits, llr = synth.make_batch(K, 2048, 1.5, seed=K, device="cuda")
        llr = llr.repeat(N // 2048, 1).contiguous()
        dec = TurboDecoder(K, n_iter=8, max_batch=N, algo=a.algo)
        det = TurboDecoder(K, n_iter=8, max_batch=N, early_term=True, algo=a.algo)
        plan = dec.plan()
        ms, ms_et = timed(lambda: dec.decode(llr)), timed(lambda: det.decode(llr))
        out = det.decode(llr, want=("bits", "iters_used"))
        rows.append({"K": K, "L": plan["sub_block"], "P": plan["n_sub_blocks"], "cb_per_cta": plan["cb_per_cta"],
                     "gbit_s": round(N * K / ms / 1e6, 2), "gbit_s_et": round(N * K / ms_et / 1e6, 2),
                     "mean_iters_et": round(float(out["iters_used"].float().mean()), 2),
                     "bit_errors_et": int((out["bits"][:2048] != bits).sum())})
        print(json.dumps(rows[-1]), flush=True)
        dec.close()
        det.close()
    with open(a.json, "w") as f:
        json.dump(rows, f)


if __name__ == "__main__":
    main()
