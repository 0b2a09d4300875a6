#!/usr/bin/env python
"""Per-iteration block-error table against the reference's published one (ITTC/result.txt:102-116,
fixture tests/golden/ittc_result_bler.json): rows = iteration 1..8, columns = Eb/N0 0.0 ... 1.0 dB.

    python tools/bler_table.py --out gpurun_out/bler_table.json [--frames 65536]

Modes (algo[:sub_block:guard]): logmap_f64 (reference operation order; one decode yields every iteration's decisions),
logmap_f32 / linlogmap_f32 / maxlog_s16 (one decode per iteration count).  Every cell carries the
z-score of the difference to the reference's cell as two binomial estimates.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from turbo_decoder_cuda_b200 import TurboDecoder, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bler_table.json")
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--frames-f64", type=int, default=8192)
    ap.add_argument("--modes", nargs="+", default=["logmap_f64", "logmap_f32", "linlogmap_f32", "maxlog_s16"])
    args = ap.parse_args()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ittc_result_bler.json")))
    ref = gold["runs"][1]
    K, NIT = 6144, 8
    dev = torch.device("cuda", 0)
    res = {"reference": {"bler": [r for r in ref["bler"][:NIT]], "frames": ref["frames"]}, "ebn0_db": gold["ebn0_db"], "modes": {}}
    for spec in args.modes:
        mode, _, geo = spec.partition(":")          # algo[:sub_block:guard]
        kw = {}
        if geo:
            L, G = geo.split(":")
            kw = {"sub_block": int(L), "warmup": int(G)}
        nf = args.frames_f64 if mode == "logmap_f64" else args.frames
        table = [[0.0] * len(gold["ebn0_db"]) for _ in range(NIT)]
        if mode == "logmap_f64":
            decs = {NIT: TurboDecoder(K, n_iter=NIT, algo=mode, max_batch=2048)}
        else:
            decs = {it: TurboDecoder(K, n_iter=it, algo=mode, max_batch=4096, **kw) for it in range(1, NIT + 1)}
        for ci, eb in enumerate(gold["ebn0_db"]):
            fe = [0] * NIT
            done = 0
            while done < nf:
                n = min(4096 if mode != "logmap_f64" else 2048, nf - done)
                bits, llr = synth.make_batch(K, n, eb, seed=77000 + 131 * ci + done, device=dev,
                                             dtype=torch.float64 if mode == "logmap_f64" else torch.float32)
                if mode == "logmap_f64":
                    out = decs[NIT].decode(llr, want=("bits_iters",))["bits_iters"]      # [n, NIT, K] int32
                    err = (out != bits[:, None, :].to(torch.int32)).any(dim=2)            # [n, NIT]
                    for it in range(NIT):
                        fe[it] += int(err[:, it].sum().item())
                else:
                    for it in range(1, NIT + 1):
                        out = decs[it].decode(llr, want=("bits",))["bits"]
                        fe[it - 1] += int((out != bits).any(dim=1).sum().item())
                done += n
            for it in range(NIT):
                table[it][ci] = fe[it] / nf
            print(spec, eb, ["%.3g" % (fe[it] / nf) for it in range(NIT)], flush=True)
        z = [[0.0] * len(gold["ebn0_db"]) for _ in range(NIT)]
        for it in range(NIT):
            for ci in range(len(gold["ebn0_db"])):
                p1, n1 = ref["bler"][it][ci], ref["frames"][ci]
                p2, n2 = table[it][ci], nf
                pp = (p1 * n1 + p2 * n2) / (n1 + n2)
                sd = math.sqrt(max(pp * (1 - pp), 1e-12) * (1.0 / n1 + 1.0 / n2))
                z[it][ci] = (p2 - p1) / sd if sd > 0 else 0.0
        res["modes"][spec] = {"frames": nf, "bler": table, "z_vs_reference": z}
        for d in decs.values():
            d.close()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for mode, m in res["modes"].items():
        zz = [abs(v) for row in m["z_vs_reference"] for v in row]
        print("%-14s frames %6d  max |z| %.2f  cells with |z| > 3: %d of %d" % (mode, m["frames"], max(zz), sum(v > 3 for v in zz), len(zz)))


if __name__ == "__main__":
    main()
