"""Smallest end-to-end case for compute-sanitizer: each kernel family once, tiny batches.
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from turbo_decoder_cuda_b200 import TurboDecoder, synth  # noqa: E402

dev = torch.device("cuda", 0)
for K, algo, n in ((6144, "maxlog_s16", 3), (1008, "maxlog_s16", 3), (512, "logmap_f32", 2), (512, "maxlog_f32", 2), (256, "logmap_f64", 5)):
    bits, llr = synth.make_batch(K, n, 2.0, seed=K, device="cpu")
    x = llr.double() if algo == "logmap_f64" else llr
    dec = TurboDecoder(K, n_iter=3, algo=algo, early_term=(algo == "maxlog_s16"), max_batch=4)
    out = dec.decode(x.to(dev), want=("bits", "iters_used"))
    torch.cuda.synchronize()
    host = dec.decode(x.numpy(), want=("bits",))
    ok = np.array_equal(out["bits"].cpu().numpy(), host["bits"])
    print(K, algo, "device==host:", ok, "bit errors:", int((out["bits"].cpu() != bits).sum()))
    dec.close()
print("done")
