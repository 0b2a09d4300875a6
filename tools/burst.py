#!/usr/bin/env python
"""Codeblock-sharded burst decode (BASELINE configs[4]): ~10^6 codeblocks, K = 6144, 8 iterations,
over the GPUs of one box -- one process per GPU, no collective on the data path.

    python tools/burst.py --total 1000000                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/burst.py --total 1000000      # N GPUs

Each rank generates the LLRs of its shard on its own device, chunk by chunk (synth.py: random bits
-> (13,15) PCCC -> BPSK/AWGN -> LLR), and decodes them.  Generation is not timed; the decode of
every chunk is timed with CUDA events on the launching stream.  Rank 0 prints one JSON line:
aggregate Gbit/s (total bits / max over ranks of the summed decode time), BER, FER with a Wilson
interval, mean iterations.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return [(c - h) / d, (c + h) / d]


def main():
    import torch
    import torch.distributed as dist
    from turbo_decoder_cuda_b200 import TurboDecoder, decoder as tdb, shard, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1000000)
    ap.add_argument("--K", type=int, default=6144)
    ap.add_argument("--ebn0", type=float, default=1.0)
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--algo", default="maxlog_s16")
    ap.add_argument("--early-term", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K = args.K
    lo, hi = shard.shard_range(args.total, world, rank)
    dec = TurboDecoder(K, n_iter=8, algo=args.algo, device=local, early_term=bool(args.early_term), max_batch=args.chunk)
    out_bits = torch.empty((args.chunk, K), dtype=torch.uint8, device=dev)
    iters = torch.empty((args.chunk,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    ms_total, be, fe, its = 0.0, 0, 0, 0.0
    for ci, c0 in enumerate(range(lo, hi, args.chunk)):
        n = min(args.chunk, hi - c0)
        bits, llr = synth.make_batch(K, n, args.ebn0, seed=(c0 + 1) * 7919 + 13, device=dev, chunk=4096)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        dec.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, n, bits=out_bits.data_ptr(), iters_used=iters.data_ptr(),
                       stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms_total += e0.elapsed_time(e1)
        err = (out_bits[:n] != bits).sum(dim=1)
        be += int(err.sum().item())
        fe += int((err > 0).sum().item())
        its += float(iters[:n].sum().item())
    t = torch.tensor([ms_total, float(be), float(fe), its, float(hi - lo)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_max = float(mx[0])
    else:
        ms_max = ms_total
    if rank == 0:
        total = int(t[4])
        line = {"workload": "burst: %d codeblocks, K=%d, 8 iterations, Eb/N0 %.2f dB, %s%s" %
                            (total, K, args.ebn0, args.algo, ", early termination" if args.early_term else ""),
                "n_gpus": world, "codeblocks": total, "decode_ms_max_over_ranks": ms_max,
                "gbit_s": total * K / (ms_max * 1e-3) / 1e9, "bit_errors": int(t[1]), "frame_errors": int(t[2]),
                "ber": float(t[1]) / (total * K), "fer": float(t[2]) / total, "fer_ci95": wilson(int(t[2]), total),
                "mean_iters": float(t[3]) / total, "sharding": "contiguous codeblock ranges, no collective on the data path"}
        s = json.dumps(line)
        print(s, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
