#!/usr/bin/env python
"""Block-error parity of the library's own sub-block plan against the (nearly) unsegmented recursion, for every LTE
block size (BASELINE configs[3]: K = 40 ... 6144, all 188 sizes).

For each K the SAME frames are decoded twice by the same decoder mode: once with the auto plan (what tdb200_create
picks: sub-block length L from the measured table, guard 16 / 24) and once with a reference plan whose windows are so
long that segmentation cannot matter -- one sub-block (L = K) up to K = 1024, the longest admissible sub-blocks
(at least 192 steps) with guard 32 above.  Two operating points per size are found by a coarse scan with the reference
plan (block-error rate just below 0.2 and the point 0.2 dB above it).  Reported per (K, Eb/N0): both block-error
rates, the frames only one plan gets wrong, and the Eb/N0 loss of the auto plan from the local slope of the reference
plan's own curve; an entry FAILS when that loss exceeds 0.05 dB with |z| > 3.

    python tools/plan_ber_parity.py --algo maxlog_s16 --json gpurun_out/plan_ber_parity_maxlog_s16.json [--sizes 88 104 ...]
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TdbError, TurboDecoder, synth  # noqa: E402


def lte_sizes():
    return list(range(40, 512, 8)) + list(range(512, 1024, 16)) + list(range(1024, 2048, 32)) + list(range(2048, 6145, 64))


def reference_plan(K):
    """(L, G) of the plan that stands in for the unsegmented recursion."""
    if K <= 1024:
        return K, 0
    best = None
    for L in range(8, K + 1, 8):
        if K % L == 0 and K // L <= 256 and 192 <= L <= 1536:
            best = L if best is None else max(best, L)
    if best is None:  # no long divisor: the longest there is
        best = max(L for L in range(8, K + 1, 8) if K % L == 0 and K // L <= 256)
    return best, min(32, best)


def block_errors(dec, llr, bits, chunk=8192):
    bad = []
    for c0 in range(0, llr.shape[0], chunk):
        out = dec.decode(llr[c0:c0 + chunk], want=("bits",))["bits"]
        bad.append((out != bits[c0:c0 + chunk]).any(dim=1))
    return torch.cat(bad)


def measure(K, algo, N=0):
    """One row of the table: the auto plan against the reference plan on the same frames at two operating points."""
    N = N or (32768 if K <= 2048 else 16384)
    Lr, Gr = reference_plan(K)
    auto = TurboDecoder(K, n_iter=8, algo=algo, max_batch=8192)
    ref = TurboDecoder(K, n_iter=8, algo=algo, max_batch=8192, sub_block=Lr, warmup=Gr)
    pa, pr = auto.plan(), ref.plan()
    same = (pa["sub_block"], pa["warmup"]) == (pr["sub_block"], pr["warmup"])
    # coarse scan with the reference plan: first Eb/N0 (0.2 dB grid) with a block-error rate below 0.2
    eb = 0.0 if K >= 2048 else (0.4 if K >= 512 else 1.0)
    while eb < 8.0:
        bits, llr = synth.make_batch(K, 2048, eb, seed=7 * K + int(round(eb * 10)), device="cuda")
        if float(block_errors(ref, llr, bits).float().mean()) < 0.2:
            break
        eb += 0.2
    pts = []
    for e in (eb, eb + 0.2):
        bits, llr = synth.make_batch(K, N, e, seed=1000 * K + int(round(e * 10)), device="cuda")
        br = block_errors(ref, llr, bits)
        ba = br if same else block_errors(auto, llr, bits)
        pts.append({"ebn0_db": round(e, 2), "frames": N, "fer_reference_plan": float(br.float().mean()), "fer_auto_plan": float(ba.float().mean()),
                    "only_reference_wrong": int((br & ~ba).sum()), "only_auto_wrong": int((ba & ~br).sum())})
    # Eb/N0 loss from the local slope of the reference plan's curve
    f1, f2 = pts[0]["fer_reference_plan"], pts[1]["fer_reference_plan"]
    slope = math.log(f1 / f2) / 0.2 if f1 > 0 and f2 > 0 and f1 > f2 else None   # nepers per dB
    for p in pts:
        d = p["only_auto_wrong"] - p["only_reference_wrong"]
        n = p["only_auto_wrong"] + p["only_reference_wrong"]
        p["z"] = d / math.sqrt(n) if n else 0.0
        p["loss_db"] = (math.log(p["fer_auto_plan"] / p["fer_reference_plan"]) / slope
                        if slope and p["fer_auto_plan"] > 0 and p["fer_reference_plan"] > 0 else None)
    worst = max((p["loss_db"] or 0.0) for p in pts)
    fail = any((p["loss_db"] or 0.0) > 0.05 and p["z"] > 3.0 for p in pts)
    auto.close()
    ref.close()
    return {"K": K, "algo": algo, "auto_plan": {"L": pa["sub_block"], "P": pa["n_sub_blocks"], "G": pa["warmup"]},
            "reference_plan": {"L": pr["sub_block"], "P": pr["n_sub_blocks"], "G": pr["warmup"]},
            "points": pts, "worst_loss_db": worst, "fail": fail}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="gpurun_out/plan_ber_parity.json")
    ap.add_argument("--algo", default="maxlog_s16")
    ap.add_argument("--sizes", type=int, nargs="*", default=None)
    ap.add_argument("--frames", type=int, default=0, help="frames per point (0: 32768 up to K = 2048, 16384 above)")
    a = ap.parse_args()
    rows = []
    for K in (a.sizes or lte_sizes()):
        try:
            row = measure(K, a.algo, a.frames)
        except TdbError as e:
            row = {"K": K, "error": str(e)}
            rows.append(row)
            print(json.dumps(row), flush=True)
            continue
        rows.append(row)
        pts = row["points"]
        print(json.dumps({"K": K, "auto": row["auto_plan"], "ref": row["reference_plan"],
                          "fer": [(p["ebn0_db"], round(p["fer_reference_plan"], 5), round(p["fer_auto_plan"], 5)) for p in pts],
                          "loss_db": [None if p["loss_db"] is None else round(p["loss_db"], 3) for p in pts], "fail": row["fail"]}), flush=True)
    summary = {"algo": a.algo, "sizes": len(rows), "failed": [r["K"] for r in rows if r.get("fail")],
               "max_loss_db": max((r.get("worst_loss_db") or 0.0) for r in rows) if rows else 0.0,
               "rule": "fail = Eb/N0 loss of the auto plan against the reference plan above 0.05 dB with McNemar z > 3, at either operating point"}
    print(json.dumps(summary))
    os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
    with open(a.json, "w") as f:
        json.dump({"summary": summary, "rows": rows}, f)


if __name__ == "__main__":
    main()
