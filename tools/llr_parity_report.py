"""LLR parity report of the fp64 reference-order kernel at the BASELINE size: max |difference| of the
last-iteration a-posteriori / extrinsic LLRs against the oracle's restatement and, where oracle/_ref
travelled, against the reference itself (re-stated loop around its own Log_MAP_decoder).
    python tools/llr_parity_report.py        # needs a B200; measured: 0 and 0 (bit-identical)"""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle_lib import Oracle, RefLib
from turbo_decoder_cuda_b200 import TurboDecoder
o=Oracle(); K=6144; pi=o.qpp(K)
worst=0; worst_ref=0
ref = RefLib(K,*o.lte_params(K)) if RefLib.available() else None
for eb,seed in ((0.0,1),(0.4,2),(1.0,3)):
    bits,llr=o.make_batch(K,4,eb,seed=seed)
    dec=TurboDecoder(K,n_iter=8,algo="logmap_f64",max_batch=4)
    out=dec.decode(torch.from_numpy(llr).cuda(),want=("bits_iters","llr_siso1","llr_siso2","ext_siso2"))
    for c in range(4):
        ob,o1,o2,le=o.decode(llr[c],pi,8,want_llr=True)
        d=max(np.abs(out["llr_siso1"][c].cpu().numpy()-o1).max(), np.abs(out["llr_siso2"][c].cpu().numpy()-o2).max(), np.abs(out["ext_siso2"][c].cpu().numpy()-le).max())
        worst=max(worst,d)
        assert np.array_equal(out["bits_iters"][c].cpu().numpy(), ob)
        if ref is not None and c==0:
            rb, r1, r2, rle = ref.decode(llr[c], 8, want_llr=True)
            worst_ref=max(worst_ref, np.abs(out["llr_siso2"][c].cpu().numpy()-r2).max(), np.abs(out["ext_siso2"][c].cpu().numpy()-rle).max())
            assert np.array_equal(out["bits_iters"][c].cpu().numpy(), rb)
print("max |LLR diff| kernel vs oracle:", worst, " kernel vs compiled reference:", worst_ref)
