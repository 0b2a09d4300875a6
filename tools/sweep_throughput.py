#!/usr/bin/env python
"""Block-length x batch-size throughput sweep (BASELINE configs[3]): LTE K = 40..6144, batches
1..65536 codeblocks, early termination on and off, device-resident LLRs, CUDA-event timing.

    python tools/sweep_throughput.py --out gpurun_out/sweep_k_batch.jsonl [--all-k]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from turbo_decoder_cuda_b200 import TurboDecoder, decoder as tdb, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep_k_batch.jsonl")
    ap.add_argument("--ks", type=int, nargs="+", default=[40, 104, 256, 512, 1024, 2048, 3072, 4096, 5120, 6144])
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 16, 256, 4096, 65536])
    ap.add_argument("--ebn0", type=float, default=1.5)
    ap.add_argument("--algo", default="maxlog_s16")
    ap.add_argument("--max-bytes", type=float, default=24e9, help="skip (K, batch) points whose LLRs exceed this")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for K in args.ks:
            base_n = 256
            bits0, llr0 = synth.make_batch(K, base_n, args.ebn0, seed=K, device=dev)
            for batch in args.batches:
                if batch * (3 * K + 12) * 4 > args.max_bytes:
                    continue
                reps = (batch + base_n - 1) // base_n
                llr = llr0.repeat(reps, 1)[:batch].contiguous()
                bits = bits0.repeat(reps, 1)[:batch]
                for et in (0, 1):
                    dec = TurboDecoder(K, n_iter=8, algo=args.algo, early_term=bool(et), max_batch=max(batch, 2))
                    out_bits = torch.empty((batch, K), dtype=torch.uint8, device=dev)
                    iters = torch.empty((batch,), dtype=torch.int32, device=dev)
                    st = torch.cuda.current_stream()

                    def step():
                        dec.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=out_bits.data_ptr(),
                                       iters_used=iters.data_ptr(), stream=st.cuda_stream)
                    for _ in range(3):
                        step()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    for _ in range(args.reps):
                        step()
                    e1.record(st)
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.reps
                    plan = dec.plan()
                    line = {"K": K, "batch": batch, "early_term": et, "algo": args.algo, "ebn0_db": args.ebn0,
                            "ms": ms, "gbit_s": batch * K / (ms * 1e-3) / 1e9, "us_per_call": 1e3 * ms,
                            "mean_iters": float(iters.float().mean().item()),
                            "bit_errors": int((out_bits != bits).sum().item()),
                            "sub_block": plan["sub_block"], "n_sub_blocks": plan["n_sub_blocks"], "guard": plan["warmup"]}
                    f.write(json.dumps(line) + "\n")
                    f.flush()
                    print("K %5d batch %6d et %d  %9.3f ms  %8.3f Gbit/s  iters %.2f  errs %d  (L=%d P=%d)" %
                          (K, batch, et, ms, line["gbit_s"], line["mean_iters"], line["bit_errors"], plan["sub_block"], plan["n_sub_blocks"]), flush=True)
                    dec.close()


if __name__ == "__main__":
    main()
