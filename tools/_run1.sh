python -m pytest tests/test_gpu_ratematch.py tests/test_gpu_modem.py -x -q 2>&1 | tail -3
python tools/time_ratematch.py --json gpurun_out/ratematch_timing.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    try: r=json.loads(l)
    except Exception: print(l[:300]); continue
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('rate','rate_match_ms','rate_dematch_f32_ms','rate_dematch_f32_gb_s','dematch_to_s8_ms','decode_rm_gbit_s')})
"
python tools/time_modem.py --json gpurun_out/modem_timing.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ('modulation','demap_s8_ms','decode_symbols_ms','decode_symbols_gbit_s','e2e_host_symbols_f32_gbit_s','e2e_host_symbols_f16_gbit_s')})
"
