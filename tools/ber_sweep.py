#!/usr/bin/env python
"""BER/FER sweep of decoder configurations on one GPU (BASELINE configs[2]: windows / guards vs the
unsegmented recursion).  Every configuration decodes the SAME seeded LLR batches.

    python tools/ber_sweep.py --out gpurun_out/ber_sweep.jsonl --ebn0 0.4 0.6 0.8 1.0 --n 20000 \
        --cfg name:key=val,key=val ...

Config keys are TurboDecoder keyword arguments (algo, sub_block, warmup, frac_bits, ext_clip,
ext_scale_q2, early_term, n_iter).  One JSON line per (Eb/N0, config):
bit errors, frame errors, counts, Wilson 95% interval of the FER.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def wilson(k, n, z=1.96):
    if n == 0:
        return (0.0, 1.0)
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return ((c - h) / d, (c + h) / d)


def parse_cfg(s):
    name, _, rest = s.partition(":")
    kw = {}
    for item in filter(None, rest.split(",")):
        k, v = item.split("=")
        kw[k] = v if k == "algo" else int(v)
    return name, kw


def main():
    import torch
    from turbo_decoder_cuda_b200 import TurboDecoder, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ber_sweep.jsonl")
    ap.add_argument("--K", type=int, default=6144)
    ap.add_argument("--ebn0", type=float, nargs="+", default=[0.4, 0.6, 0.8, 1.0])
    ap.add_argument("--n", type=int, default=20000, help="codeblocks per Eb/N0 point")
    ap.add_argument("--n-slow", type=int, default=2048, help="codeblocks per point for logmap_f64 configs")
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--cfg", nargs="+", required=True)
    args = ap.parse_args()

    dev = torch.device("cuda", 0)
    cfgs = [parse_cfg(c) for c in args.cfg]
    decs = []
    for name, kw in cfgs:
        kw = dict(kw)
        limit = kw.pop("frames", None)   # per-configuration cap on the frames per point (slow modes)
        kw.setdefault("n_iter", 8)
        kw.setdefault("algo", "maxlog_s16")
        kw.setdefault("max_batch", args.chunk)
        decs.append((name, dict(kw, frames=limit), TurboDecoder(args.K, **kw)))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for eb in args.ebn0:
            acc = {name: [0, 0, 0, 0.0, 0.0] for name, _, _ in decs}  # bit errs, frame errs, frames, iters, seconds
            done = 0
            ci = 0
            while done < args.n:
                n = min(args.chunk, args.n - done)
                bits, llr = synth.make_batch(args.K, n, eb, seed=args.seed * 100003 + ci, device=dev)
                for name, kw, dec in decs:
                    cap = kw.get("frames") or (args.n_slow if kw["algo"] == "logmap_f64" else args.n)
                    if acc[name][2] >= cap:
                        continue
                    m = min(n, cap - acc[name][2])
                    x = llr[:m].double() if kw["algo"] == "logmap_f64" else llr[:m]
                    want = ("bits", "iters_used") if kw.get("early_term") else ("bits",)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    out = dec.decode(x, want=want)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    err = (out["bits"] != bits[:m]).sum(dim=1)
                    a = acc[name]
                    a[0] += int(err.sum().item())
                    a[1] += int((err > 0).sum().item())
                    a[2] += m
                    a[3] += float(out["iters_used"].sum().item()) if "iters_used" in out else m * kw["n_iter"]
                    a[4] += dt
                done += n
                ci += 1
            for name, kw, dec in decs:
                be, fe, nf, its, secs = acc[name]
                lo, hi = wilson(fe, nf)
                plan = dec.plan()
                line = {"K": args.K, "ebn0_db": eb, "cfg": name, "params": {k: v for k, v in kw.items() if v is not None}, "sub_block": plan["sub_block"],
                        "guard": plan["warmup"], "frames": nf, "bit_errors": be, "frame_errors": fe,
                        "ber": be / (nf * args.K), "fer": fe / nf, "fer_ci95": [lo, hi],
                        "mean_iters": its / nf, "gbit_s_incl_sync": nf * args.K / secs / 1e9}
                f.write(json.dumps(line) + "\n")
                f.flush()
                print("%.2f dB %-22s frames %6d  BER %.3e  FER %.3e [%.2e, %.2e]  iters %.2f" %
                      (eb, name, nf, line["ber"], line["fer"], lo, hi, line["mean_iters"]), flush=True)


if __name__ == "__main__":
    main()
