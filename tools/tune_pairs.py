"""Codeblock pairs per CTA for sub-block counts that do not fill a warp multiple (TDB200_PAIRS_PER_CTA knob):
throughput of np = 1, 2, 3 for every LTE block size whose auto plan has 32 < P <= 64, 8 fixed iterations.
    python tools/tune_pairs.py --json gpurun_out/pairs_tuning.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder, synth  # noqa: E402
from tools.sweep_all_sizes import lte_sizes, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="gpurun_out/pairs_tuning.json")
    a = ap.parse_args()
    rows = []
    for K in lte_sizes():
        os.environ["TDB200_PAIRS_PER_CTA"] = "1"
        probe = TurboDecoder(K, n_iter=8, max_batch=4)
        P = probe.plan()["n_sub_blocks"]
        probe.close()
        if P <= 32 or P > 64:
            continue
        N = 16384 if K > 512 else 65536
        _, llr = synth.make_batch(K, 2048, 1.5, seed=K, device="cuda")
        llr = llr.repeat(N // 2048, 1).contiguous()
        row = {"K": K, "P": P}
        ref = None
        for np_ in (1, 2, 3):
            if np_ * P > 128:
                break
            os.environ["TDB200_PAIRS_PER_CTA"] = str(np_)
            dec = TurboDecoder(K, n_iter=8, max_batch=N)
            plan = dec.plan()
            ms = timed(lambda: dec.decode(llr))
            bits = dec.decode(llr, want=("bits",))["bits"]
            if ref is None:
                ref = bits
            row["np%d" % np_] = round(N * K / ms / 1e6, 2)
            row["np%d_same" % np_] = bool(torch.equal(bits, ref))
            row["np%d_cb_per_cta" % np_] = plan["cb_per_cta"]
            dec.close()
        rows.append(row)
        print(json.dumps(row), flush=True)
    del os.environ["TDB200_PAIRS_PER_CTA"]
    with open(a.json, "w") as f:
        json.dump(rows, f)


if __name__ == "__main__":
    main()
