import sys, time, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle_lib import Oracle
from turbo_decoder_cuda_b200 import TurboDecoder
o = Oracle(); K = 6144
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kw = {}
algo = "maxlog_s16"
for a in sys.argv[2:]:
    k, v = a.split("=")
    if k == "algo": algo = v
    else: kw[k] = int(v)
bits, llr = o.make_batch(K, 16, 1.0, seed=1)
llr = np.tile(llr.astype(np.float32), (n // 16, 1)); bits = np.tile(bits, (n // 16, 1))
d = torch.from_numpy(llr).cuda()
dec = TurboDecoder(K, n_iter=8, algo=algo, max_batch=n, **kw)
print(dec.plan())
for _ in range(3):
    out = dec.decode(d, want=("bits",)); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = dec.decode(d, want=("bits",)); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
print(algo + ": n_cb=%d  %.3f ms (min of 5; all %s) -> %.2f Gbit/s ; bit errors %d" % (n, ms, ["%.3f" % t for t in ts], n * K / ms / 1e6, int((out["bits"].cpu().numpy() != bits).sum())))
