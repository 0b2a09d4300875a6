#!/usr/bin/env python
"""Decode frames produced by the REFERENCE'S OWN encoder and channel (TurboEnCoding, module, AWGN with
its mgrns noise, demodule -- through oracle/_ref/libittc_ref.so, the reference's sources compiled in
place) with the fp64 reference-order kernel, and compare the per-iteration block-error rates with the
reference's published table (ITTC/result.txt:102-116).  This separates the decoder (ours) from the
channel simulator (the reference's mgrns is a 16-bit LCG feeding a 12-term CLT sum): with true
Gaussian noise (tools/bler_table.py) some cells of the table differ by a few sigma at high
statistical power; with the reference's own noise they must not.

    python tools/bler_refchannel.py --out gpurun_out/bler_refchannel.json --frames 16384 --ebn0 0.3 0.4 0.5
"""
import argparse
import json
import math
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
K = 6144


def _gen(job):
    from oracle_lib import Oracle, RefLib
    eb, lo, hi = job
    o = Oracle()
    ref = RefLib(K, *o.lte_params(K))
    sigma = o.sigma(eb, K)
    rng = np.random.default_rng(1000003 * int(round(eb * 10)) + lo)
    bits = rng.integers(0, 2, size=(hi - lo, K), dtype=np.int32)
    llr = np.empty((hi - lo, 3 * K + 12), np.float64)
    for i in range(hi - lo):
        llr[i] = ref.channel(ref.encode(bits[i]), sigma, seed=(lo + i) * 2654435761 % (2 ** 31))
    return bits.astype(np.uint8), llr


def main():
    import torch
    from oracle_lib import RefLib
    from turbo_decoder_cuda_b200 import TurboDecoder
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bler_refchannel.json")
    ap.add_argument("--frames", type=int, default=16384)
    ap.add_argument("--ebn0", type=float, nargs="+", default=[0.3, 0.4, 0.5])
    ap.add_argument("--algo", default="logmap_f64", help="logmap_f64 | logmap_s16 | maxlog_s16 (decoders that deliver per-iteration decisions)")
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--sub-block", type=int, default=0)
    args = ap.parse_args()
    if not RefLib.available():
        raise SystemExit("oracle/_ref/libittc_ref.so is missing (build it in the dev container: make -C oracle ref)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ittc_result_bler.json")))
    ref = gold["runs"][1]
    NIT = 8
    dec = TurboDecoder(K, n_iter=NIT, algo=args.algo, max_batch=2048, sub_block=args.sub_block, warmup=args.warmup)
    res = {"frames": args.frames, "algo": args.algo, "plan": {k: dec.plan()[k] for k in ("sub_block", "warmup")}, "points": []}
    workers = max(1, len(os.sched_getaffinity(0)))
    with mp.Pool(workers) as pool:
        for eb in args.ebn0:
            ci = gold["ebn0_db"].index(round(eb, 1))
            step = max(1, args.frames // (4 * workers))
            jobs = [(eb, lo, min(lo + step, args.frames)) for lo in range(0, args.frames, step)]
            fe = np.zeros(NIT, np.int64)
            for bits, llr in pool.imap(_gen, jobs):
                x = llr if args.algo == "logmap_f64" else llr.astype(np.float32)
                out = dec.decode(torch.from_numpy(x).cuda(), want=("bits_iters",))["bits_iters"]
                err = (out != torch.from_numpy(bits).cuda()[:, None, :].to(torch.int32)).any(dim=2)
                fe += err.sum(dim=0).cpu().numpy()
            row = {"ebn0_db": eb, "bler": [], "reference": [], "z": []}
            for it in range(NIT):
                p1, n1 = ref["bler"][it][ci], ref["frames"][ci]
                p2, n2 = fe[it] / args.frames, args.frames
                pp = (p1 * n1 + p2 * n2) / (n1 + n2)
                sd = math.sqrt(max(pp * (1 - pp), 1e-12) * (1.0 / n1 + 1.0 / n2))
                row["bler"].append(float(p2)); row["reference"].append(float(p1)); row["z"].append(float((p2 - p1) / sd))
            res["points"].append(row)
            print("%.1f dB  ours %s\n        ref  %s\n        z    %s" % (eb, ["%.4f" % v for v in row["bler"]],
                  ["%.4f" % v for v in row["reference"]], ["%+.1f" % v for v in row["z"]]), flush=True)
    zs = [abs(z) for r in res["points"] for z in r["z"]]
    res["cells"] = len(zs)
    res["cells_outside_95pct"] = int(sum(z > 1.96 for z in zs))
    res["cells_outside_3sigma"] = int(sum(z > 3 for z in zs))
    res["max_abs_z"] = float(max(zs)) if zs else 0.0
    print("cells %d, |z| > 1.96: %d, |z| > 3: %d, max |z| %.2f" % (len(zs), res["cells_outside_95pct"], res["cells_outside_3sigma"], res["max_abs_z"]))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    mp.set_start_method("spawn")
    main()
