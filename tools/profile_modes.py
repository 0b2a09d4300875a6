"""One launch of every kernel family at the BASELINE block size, for an ncu capture of the secondary
kernels (the throughput kernel has its own capture through bench.py):
    ncu --set full --clock-control none -o gpurun_out/prof_modes python tools/profile_modes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from turbo_decoder_cuda_b200 import TurboDecoder, synth  # noqa: E402

K = 6144
dev = torch.device("cuda", 0)
bits, llr = synth.make_batch(K, 2048, 1.0, seed=1, device=dev)        # encode_kernel + channel_kernel
for algo, n in (("maxlog_f32", 2048), ("linlogmap_f32", 2048), ("logmap_f32", 2048), ("logmap_f64", 512)):
    dec = TurboDecoder(K, n_iter=8, algo=algo, max_batch=n)
    x = llr[:n].double() if algo == "logmap_f64" else llr[:n]
    out = dec.decode(x, want=("bits",))
    torch.cuda.synchronize()
    print(algo, "bit errors", int((out["bits"] != bits[:n]).sum()))
    dec.close()
