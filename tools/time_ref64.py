import sys, time, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle_lib import Oracle
from turbo_decoder_cuda_b200 import TurboDecoder
o = Oracle(); K = 6144
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bits, llr = o.make_batch(K, 8, 1.0, seed=1)
llr = np.tile(llr, (n // 8, 1))
d = torch.from_numpy(llr).cuda()
dec = TurboDecoder(K, n_iter=8, algo="logmap_f64", max_batch=n)
for _ in range(2):
    out = dec.decode(d, want=("bits",)); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); out = dec.decode(d, want=("bits",)); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("ref64: n_cb=%d  %.2f ms  -> %.3f Gbit/s ; bit errors %d" % (n, ms, n * K / ms / 1e6, int((out["bits"].cpu().numpy()[:8] != bits).sum())))
