"""Device-time of the encoder and channel kernels at the BASELINE block size (CUDA events).
    python tools/time_encode.py   # B200: encode 0.18 ms / 4096 codeblocks (561 GB/s of in+out bytes), channel 0.16 ms (2.4 TB/s)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder
K=6144; n=4096
dec=TurboDecoder(K, algo="maxlog_s16", max_batch=8)
bits=torch.randint(0,2,(n,K),dtype=torch.uint8,device="cuda")
for _ in range(3): c=dec.encode(bits)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): c=dec.encode(bits)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("encode: %.3f ms for %d codeblocks  -> %.1f Gbit/s info, %.0f GB/s of (K + 3K+12) bytes" % (ms, n, n*K/ms/1e6, n*(4*K+12)/ms/1e6))
for _ in range(3): l=dec.channel(c, 0.8, seed=1)
torch.cuda.synchronize()
e0.record()
for _ in range(10): l=dec.channel(c, 0.8, seed=1)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("channel: %.3f ms -> %.0f GB/s of (1 + 4) bytes per element" % (ms, n*(3*K+12)*5/ms/1e6))
