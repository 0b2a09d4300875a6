python -m pytest tests/test_gpu_ratematch.py tests/test_gpu_ref64.py -x -q 2>&1 | tail -3
python tools/time_ratematch.py --json gpurun_out/ratematch_timing.json 2>&1 | tail -6
