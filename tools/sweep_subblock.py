"""Throughput of the s16 decoder against the sub-block length L for block sizes without compile-time
geometry (which L does the auto plan pick, which one is fastest?):  python tools/sweep_subblock.py [K ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TdbError, TurboDecoder, synth  # noqa: E402

Ks = [int(x) for x in sys.argv[1:]] or [5824, 5888, 5952, 6016, 6080, 4992, 3008, 2112, 1504, 1056]
N = 4096
for K in Ks:
    bits, llr = synth.make_batch(K, 512, 1.5, seed=1, device="cuda")
    llr = llr.repeat(N // 512, 1).contiguous()
    row = {"K": K}
    auto = TurboDecoder(K, n_iter=8, max_batch=N)
    row["auto_L"] = auto.plan()["sub_block"]
    for L in (0, 24, 32, 40, 48, 56, 64, 72, 80, 96, 104, 112, 128):
        if L and (K % L or K // L > 256):
            continue
        try:
            dec = auto if L == 0 else TurboDecoder(K, n_iter=8, max_batch=N, sub_block=L, warmup=16)
        except TdbError:
            continue
        for _ in range(2):
            out = dec.decode(llr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = dec.decode(llr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        err = int((out["bits"][:512] != bits).sum())
        row["L%d(P%d)" % (L, K // L) if L else "auto"] = round(N * K / ms / 1e6, 2)
        if err:
            row.setdefault("bit_errors", {})[L] = err
    print(json.dumps(row))
