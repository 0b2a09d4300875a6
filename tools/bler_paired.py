#!/usr/bin/env python
"""Paired block-error comparison: THE SAME FRAMES through the reference's CPU Log-MAP and through a GPU decoder mode.

For every Eb/N0 the frames come either from the reference's own encoder and channel (TurboEnCoding, module, AWGN
with its mgrns noise, demodule -- oracle/_ref, the reference's sources compiled in place) or from the oracle's
Gaussian channel; each frame is decoded
  * by the CPU restatement of TurboDecoding()/Log_MAP_decoder() (oracle/turbo_oracle.c, bit-identical to the
    compiled reference: tests/test_oracle.py) -- fp64, 16-step max* table, unsegmented, and
  * by the GPU decoder(s) named with --algo, through the C ABI, with per-iteration decisions (bits_iters).
Reported per (Eb/N0, iteration): both block-error rates on these frames, the discordant pairs (frames only one of
the two decoders gets wrong), McNemar's z, the Wilson 95 % interval of the reference's rate on these frames and
whether ours lies inside it, and -- for the 8-iteration row -- the Eb/N0 shift that would explain the ratio of the
two rates given the local slope of the reference's own curve (ITTC/result.txt:109).

    python tools/bler_paired.py --algo logmap_s16 --frames 8192 --ebn0 0.3 0.4 --out gpurun_out/bler_paired.json
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
K, NIT = 6144, 8


def _work(job):
    """Generate frames [lo, hi) at eb and decode them with the CPU reference restatement (all iterations)."""
    from oracle_lib import Oracle, RefLib
    eb, lo, hi, refchan = job
    o = Oracle()
    pi = o.qpp(K)
    sigma = o.sigma(eb, K)
    rng = np.random.default_rng(1000003 * int(round(eb * 100)) + lo)
    bits = rng.integers(0, 2, size=(hi - lo, K), dtype=np.int32)
    llr = np.empty((hi - lo, 3 * K + 12), np.float64)
    ref = RefLib(K, *o.lte_params(K)) if refchan else None
    err = np.zeros((hi - lo, NIT), np.bool_)
    for i in range(hi - lo):
        if refchan:
            llr[i] = ref.channel(ref.encode(bits[i]), sigma, seed=(lo + i) * 2654435761 % (2 ** 31))
        else:
            llr[i] = o.channel(o.encode(bits[i], pi), sigma, 4242 + int(round(eb * 100)), lo + i)
        b = o.decode(llr[i], pi, NIT)
        err[i] = (b != bits[i][None, :]).any(axis=1)
    return bits.astype(np.uint8), llr, err


def wilson(k, n, z=1.959964):
    if n == 0:
        return 0.0, 1.0
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return max(0.0, c - h), min(1.0, c + h)


def main():
    import torch
    from oracle_lib import RefLib
    from turbo_decoder_cuda_b200 import TurboDecoder
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bler_paired.json")
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--ebn0", type=float, nargs="+", default=[0.3, 0.4])
    ap.add_argument("--algo", nargs="+", default=["logmap_s16"])
    ap.add_argument("--gaussian", action="store_true", help="oracle's Gaussian channel instead of the reference's own")
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--sub-block", type=int, default=0)
    args = ap.parse_args()
    refchan = not args.gaussian
    if refchan and not RefLib.available():
        raise SystemExit("oracle/_ref/libittc_ref.so is missing (build it in the dev container: make -C oracle ref)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ittc_result_bler.json")))
    pub = gold["runs"][1]
    decs = {a: TurboDecoder(K, n_iter=NIT, algo=a, max_batch=1024, sub_block=args.sub_block, warmup=args.warmup) for a in args.algo}
    res = {"frames": args.frames, "channel": "reference (mgrns)" if refchan else "gaussian", "K": K, "n_iter": NIT,
           "plans": {a: {k: d.plan()[k] for k in ("sub_block", "warmup")} for a, d in decs.items()}, "points": []}
    workers = max(1, len(os.sched_getaffinity(0)))
    with mp.Pool(workers) as pool:
        for eb in args.ebn0:
            step = max(1, min(256, args.frames // (2 * workers)))
            jobs = [(eb, lo, min(lo + step, args.frames), refchan) for lo in range(0, args.frames, step)]
            ref_err = []
            our_err = {a: [] for a in args.algo}
            for bits, llr, err in pool.imap(_work, jobs):
                ref_err.append(err)
                tb = torch.from_numpy(bits).cuda()[:, None, :].to(torch.int32)
                for a, dec in decs.items():
                    x = torch.from_numpy(llr if a == "logmap_f64" else llr.astype(np.float32)).cuda()
                    out = dec.decode(x, want=("bits_iters",))["bits_iters"]
                    our_err[a].append((out != tb).any(dim=2).cpu().numpy())
            ref_err = np.concatenate(ref_err)
            n = ref_err.shape[0]
            row = {"ebn0_db": eb, "frames": n, "reference_bler": [float(v) for v in ref_err.mean(axis=0)], "algos": {}}
            ci = gold["ebn0_db"].index(round(eb, 1)) if round(eb, 1) in gold["ebn0_db"] else None
            if ci is not None:
                row["published_bler"] = [pub["bler"][it][ci] for it in range(NIT)]
            for a in args.algo:
                e = np.concatenate(our_err[a])
                cells = []
                for it in range(NIT):
                    kr, ku = int(ref_err[:, it].sum()), int(e[:, it].sum())
                    only_ref = int((ref_err[:, it] & ~e[:, it]).sum())
                    only_us = int((e[:, it] & ~ref_err[:, it]).sum())
                    lo, hi = wilson(kr, n)
                    z = (only_us - only_ref) / math.sqrt(only_us + only_ref) if only_us + only_ref else 0.0
                    cells.append({"iteration": it + 1, "reference": kr / n, "ours": ku / n, "only_reference_wrong": only_ref,
                                  "only_ours_wrong": only_us, "mcnemar_z": z, "reference_ci95": [lo, hi],
                                  "inside_ci95": bool(lo - 1e-12 <= ku / n <= hi + 1e-12)})
                row["algos"][a] = cells
                last = cells[-1]
                print("%.2f dB %-12s it8: reference %.5f  ours %.5f  (only ref wrong %d, only ours wrong %d, z=%+.2f, inside CI: %s); inside CI at %d/8 iterations"
                      % (eb, a, last["reference"], last["ours"], last["only_reference_wrong"], last["only_ours_wrong"],
                         last["mcnemar_z"], last["inside_ci95"], sum(c["inside_ci95"] for c in cells)), flush=True)
            res["points"].append(row)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    mp.set_start_method("spawn")
    main()
