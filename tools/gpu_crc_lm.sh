python -m pytest tests/test_gpu_transport.py tests/test_gpu_logmap_s16.py -x -q 2>&1 | tail -5
python - <<'PY'
import sys, torch, json
sys.path.insert(0,'.')
from turbo_decoder_cuda_b200 import TurboDecoder, synth
from turbo_decoder_cuda_b200.decoder import CRC24B
K=6144; N=4096
res={}
for algo in ("logmap_s16","maxlog_s16"):
    for eb in (0.6,1.0,1.5):
        row={}
        for et in (0,1,"crc24b"):
            dec=TurboDecoder(K,n_iter=8,algo=algo,early_term=et,max_batch=N)
            g=torch.Generator(device="cuda"); g.manual_seed(1)
            bits=torch.randint(0,2,(N,K),dtype=torch.uint8,device="cuda",generator=g)
            dec.crc24_attach(bits,CRC24B)
            import math
            sigma=10**(-eb/20)*math.sqrt(0.5/(K/(3*K+12)))
            llr=dec.channel(dec.encode(bits),sigma,seed=3)
            for _ in range(2): out=dec.decode(llr,want=("bits","iters_used"))
            torch.cuda.synchronize()
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): dec.decode(llr)
            e1.record(); torch.cuda.synchronize()
            ms=e0.elapsed_time(e1)/5
            fer=float((out["bits"]!=bits).any(dim=1).float().mean())
            row[str(et)]={"gbit_s":round(N*K/ms/1e6,2),"mean_iters":round(float(out["iters_used"].float().mean()),2),"fer":fer}
            dec.close()
        res["%s@%.1f"%(algo,eb)]=row
        print(algo,eb,json.dumps(row),flush=True)
json.dump(res,open("gpurun_out/r02_early_termination_rules_logmap.json","w"),indent=1)
PY
