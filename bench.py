#!/usr/bin/env python
"""bench.py -- headline benchmark of the turbo-decode hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU decoder on host cores

Metric: decoded information Gbit/s, K=6144, 8 iterations (BASELINE.json `metric`), on
BASELINE configs[1]: LTE K=6144 QPP(263,480), max-log-MAP, 8 iterations, 4096 codeblocks per GPU.
A "step" is one decode of one batch of synthetic channel LLRs (random bits -> (13,15) PCCC ->
BPSK/AWGN at --ebn0 -> LLR = 2r/sigma^2, made by turbo_decoder_cuda_b200/synth.py before timing).

  value  : whole-job Gbit/s with the LLR batch already resident in HBM (device pointers through
           the C ABI), timed with CUDA events on the launching stream, max over ranks.
  e2e    : the same metric through the same C-ABI call with HOST buffers (pinned): H2D of the LLRs
           and D2H of the hard decisions inside the timed region.
Codeblocks are independent: ranks decode disjoint shards, no collective on the data path ("weak").
Only the cpu_baseline leg and --impl reference touch oracle/ (as the thing timed there, per the
tier contract); the CUDA arm fails loudly if the CUDA library is missing.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 6144
N_ITER = 8
OPS_PER_INFO_BIT = 2 * N_ITER * 101 * (K + 3) / K  # ~101 add/max ops per trellis step per SISO


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_decoder():
    """(decode_batch(llr_f64, n_threads) -> (bits, seconds), kind) from oracle/ -- the reference's own
    code compiled in place when oracle/_ref exists, else the C restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, RefLib
    o = Oracle()
    f1, f2 = o.lte_params(K)
    if RefLib.available():
        r = RefLib(K, f1, f2)
        return (lambda llr, nt: r.decode_batch(llr, N_ITER, nt)), "reference"
    pi = o.qpp(K)
    return (lambda llr, nt: o.decode_batch(llr, pi, N_ITER, 1, nt)), "port"


def run_reference(args):
    """--impl reference: the reference CPU Log-MAP (fp64, 8 iterations) on all host cores, rank 0 only.
    Nothing of the product is imported here: frames, QPP parameters and the decoder all come from oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle()
    cores = host_cores()
    decode, kind = cpu_reference_decoder()
    # a bounded sample per step, sized from a calibration decode so that the whole run (all steps) stays
    # near two minutes of wall time whatever --steps is
    _, cal = o.make_batch(K, max(cores, 1), args.ebn0, seed=999)
    _, t_cal = decode(cal, cores)           # seconds for one codeword per core
    budget_s = 120.0
    n_sample = args.ref_sample or int(max(cores, min(args.batch, cores * budget_s / (max(args.steps, 1) * max(t_cal, 1e-3)))))
    bits, llr = o.make_batch(K, n_sample, args.ebn0, seed=1000)
    for _ in range(args.warmup):
        decode(llr[:max(cores, 1)], cores)
    t = 0.0
    errs = 0
    for _ in range(args.steps):
        out, secs = decode(llr, cores)
        t += secs
        errs += int((out != bits).sum())
    value = n_sample * K * args.steps / t / 1e9
    line = {
        "impl": "reference", "metric": "decoded info Gbit/s, K=6144, 8 iterations", "value": value, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "the reference's only CPU path for BASELINE configs[1]: LTE turbo K=6144 QPP(263,480), fp64 Log-MAP "
                               "with the 16-step max* table (ITTC/log_map.cpp), 8 iterations, a sample of %d codeblocks per step" % n_sample,
                   "K": K, "n_iter": N_ITER, "batch_per_gpu": n_sample, "ebn0_db": args.ebn0, "algo": "logmap_lut_f64",
                   "llr_input": "float64 [n_cb, 3K+12], reference multiplex order",
                   "parallelism": "one codeword per host thread, %d threads" % cores},
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind,
                         "sample": "%d codeblocks per step (K=6144, fp64 LUT Log-MAP, 8 iterations, one codeword per thread)" % n_sample},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ber": errs / float(n_sample * K * args.steps),
    }
    emit(line)


def workload_config(args, batch, plan):
    cfg = {"workload": "BASELINE configs[1]: LTE turbo K=6144 QPP(263,480), max-log-MAP, 8 iterations, "
                       "batch of %d codeblocks per GPU" % args.batch,
           "K": K, "n_iter": N_ITER, "batch_per_gpu": batch, "ebn0_db": args.ebn0, "algo": args.algo,
           "llr_input": "float32 [n_cb, 3K+12], reference multiplex order",
           "l2_policy": "inputs larger than L2 (%.0f MB of LLRs per step vs 126 MB)" % (batch * (3 * K + 12) * 4 / 1e6),
           "parallelism": "codeblock-sharded, no collective"}
    if plan:
        cfg.update({"sub_block": plan["sub_block"], "n_sub_blocks": plan["n_sub_blocks"], "guard": plan["warmup"],
                    "cb_per_cta": plan["cb_per_cta"], "smem_bytes_per_cta": plan["smem_bytes"]})
    return cfg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from turbo_decoder_cuda_b200 import TurboDecoder, decoder as tdb, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    batch = args.batch
    # synthetic traffic is made on the host (keeps the GPU launch list down to the decoder itself)
    bits, llr = synth.make_batch(K, batch, args.ebn0, seed=1000 + rank, device="cpu")
    bits, llr = bits.to(dev), llr.to(dev)
    dec = TurboDecoder(K, n_iter=N_ITER, algo=args.algo, device=local, max_batch=batch, sub_block=args.sub_block,
                       warmup=args.guard)
    plan = dec.plan()
    out_bits = torch.empty((batch, K), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step_dev():
        dec.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=out_bits.data_ptr(), stream=sp)

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    launches_per_step = dec.plan()["kernel_launches_last_call"]
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ber = float((out_bits != bits).sum().item()) / (batch * K)

    # ---- the same batch with early termination (informational; the headline is 8 fixed iterations)
    et_info = None
    crc_llr = crc_bits = None
    if args.algo == "maxlog_s16" and not args.no_early_term:
        dec_et = TurboDecoder(K, n_iter=N_ITER, algo=args.algo, device=local, max_batch=batch, sub_block=args.sub_block,
                              warmup=args.guard, early_term=True)
        et_iters = torch.empty((batch,), dtype=torch.int32, device=dev)
        et_bits = torch.empty((batch, K), dtype=torch.uint8, device=dev)

        def step_et():
            dec_et.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=et_bits.data_ptr(),
                              iters_used=et_iters.data_ptr(), stream=sp)
        for _ in range(3):
            step_et()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            step_et()
        g1.record(stream)
        torch.cuda.synchronize()
        et_info = {"_ms": g0.elapsed_time(g1) / args.steps,
                   "mean_iterations": float(et_iters.float().mean().item()),
                   "ber": float((et_bits != bits).sum().item()) / (batch * K),
                   "rule": "no hard decision changes and every |a-posteriori| >= 8.0, at most 8 iterations"}
        dec_et.close()
        # the CRC stopping rule needs blocks that end in a CRC24B: the same payloads with their last 24 bits
        # replaced, encoded and sent through the same channel (kernels of the library, outside any timed region)
        if args.sub_block == 0 and args.guard == 0:
            dec_crc = TurboDecoder(K, n_iter=N_ITER, algo=args.algo, device=local, max_batch=batch, early_term="crc24b")
            crc_bits = bits.clone()
            dec_crc.crc24_attach(crc_bits, tdb.CRC24B)
            crc_llr = dec_crc.channel(dec_crc.encode(crc_bits), synth.sigma_from_ebn0(args.ebn0, K), seed=1234 + rank)

            def step_crc():
                dec_crc.decode_raw(crc_llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=et_bits.data_ptr(),
                                   iters_used=et_iters.data_ptr(), stream=sp)
            for _ in range(3):
                step_crc()
            torch.cuda.synchronize()
            g0.record(stream)
            for _ in range(args.steps):
                step_crc()
            g1.record(stream)
            torch.cuda.synchronize()
            et_info["_crc_ms"] = g0.elapsed_time(g1) / args.steps
            et_info["crc24b_rule"] = {"mean_iterations_this_rank": float(et_iters.float().mean().item()),
                                      "ber_this_rank": float((et_bits != crc_bits).sum().item()) / (batch * K),
                                      "rule": "hard decisions of SISO-1 divide by the CRC24B generator (the last iteration counted is half-run)"}
            dec_crc.close()

    # ---- the Log-MAP decoder in the same packed arithmetic (TDB200_ALGO_LOGMAP_S16): the mode that sits on the
    #      reference's BER/FER curve (BASELINE configs[2]), same batch, device-resident, 8 fixed iterations
    lm_info = None
    if args.algo == "maxlog_s16" and not args.no_logmap:
        dec_lm = TurboDecoder(K, n_iter=N_ITER, algo="logmap_s16", device=local, max_batch=batch)
        lm_bits = torch.empty((batch, K), dtype=torch.uint8, device=dev)

        def step_lm():
            dec_lm.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=lm_bits.data_ptr(), stream=sp)
        for _ in range(3):
            step_lm()
        torch.cuda.synchronize()
        n_lm = max(1, args.steps // 4)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(n_lm):
            step_lm()
        g1.record(stream)
        torch.cuda.synchronize()
        lm_ms = g0.elapsed_time(g1) / n_lm
        pl = dec_lm.plan()
        lm_info = {"ms_per_step": lm_ms, "steps": n_lm, "ber": float((lm_bits != bits).sum().item()) / (batch * K),
                   "algo": "logmap_s16", "sub_block": pl["sub_block"], "guard": pl["warmup"], "frac_bits": 4,
                   "correction": "max* = max + max(0, 0.625 - |d|/4), all 30 max* per trellis step"}
        dec_lm.close()
        if crc_llr is not None:
            # the CRC stopping rule in the Log-MAP kernels, on the CRC-carrying frames of the leg above
            dec_lc = TurboDecoder(K, n_iter=N_ITER, algo="logmap_s16", device=local, max_batch=batch, early_term="crc24b")
            lc_iters = torch.empty((batch,), dtype=torch.int32, device=dev)

            def step_lc():
                dec_lc.decode_raw(crc_llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, batch, bits=lm_bits.data_ptr(),
                                  iters_used=lc_iters.data_ptr(), stream=sp)
            for _ in range(3):
                step_lc()
            torch.cuda.synchronize()
            g0.record(stream)
            for _ in range(n_lm):
                step_lc()
            g1.record(stream)
            torch.cuda.synchronize()
            lc_ms = g0.elapsed_time(g1) / n_lm
            lm_info["crc24b_rule"] = {"gbit_s_this_rank": batch * K / (lc_ms * 1e-3) / 1e9,
                                      "mean_iterations_this_rank": float(lc_iters.float().mean().item()),
                                      "ber_this_rank": float((lm_bits != crc_bits).sum().item()) / (batch * K)}
            dec_lc.close()
    crc_llr = None

    # ---- the fp64 reference-order decoder (TDB200_ALGO_LOGMAP_F64): the mode that meets the 1e-3 LLR bar (bit-identical
    #      to the reference's TurboDecoding), same batch, device-resident fp64 LLRs, 8 iterations
    f64_info = None
    if args.algo == "maxlog_s16" and not args.no_f64:
        dec64 = TurboDecoder(K, n_iter=N_ITER, algo="logmap_f64", device=local, max_batch=batch)
        f64_bits = torch.empty((batch, K), dtype=torch.uint8, device=dev)
        llr64 = llr.double()

        def step_f64():
            dec64.decode_raw(llr64.data_ptr(), tdb.LLR_F64, tdb.MEM_DEVICE, batch, bits=f64_bits.data_ptr(), stream=sp)
        step_f64()
        torch.cuda.synchronize()
        n64 = 2
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(n64):
            step_f64()
        g1.record(stream)
        torch.cuda.synchronize()
        f64_ms = g0.elapsed_time(g1) / n64
        f64_info = {"ms_per_step": f64_ms, "steps": n64, "gbit_s_this_rank": batch * K / (f64_ms * 1e-3) / 1e9,
                    "ber_this_rank": float((f64_bits != bits).sum().item()) / (batch * K), "algo": "logmap_f64",
                    "what": "unsegmented fp64 Log-MAP in the reference's operation order (LLRs bit-identical to the CPU reference); "
                            "bound by the latency of one trellis step and by shared-memory bandwidth (DESIGN.md 2.2)"}
        dec64.close()
        del llr64

    # ---- measured issue rate of the add-compare-select mix (the denominator of roofline.alu)
    rate_mix = dec.issue_rate(2)
    rate_alu = dec.issue_rate(0)

    # ---- end to end through host buffers (pinned): H2D + decode + D2H inside the timed region
    h_llr = torch.empty(llr.shape, dtype=llr.dtype, pin_memory=True)
    h_llr.copy_(llr)
    h_bits = torch.empty((batch, K), dtype=torch.uint8, pin_memory=True)

    def step_host():
        dec.decode_raw(h_llr.data_ptr(), tdb.LLR_F32, tdb.MEM_HOST, batch, bits=h_bits.data_ptr(), stream=sp)

    for _ in range(max(1, min(args.warmup, 3))):
        step_host()
    torch.cuda.synchronize()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    for _ in range(args.steps):
        step_host()
    f1.record(stream)
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), wall_ms)  # the host-buffer call blocks, so wall time is the honest clock
    e2e_ok = bool((h_bits.to(dev) == out_bits).all().item())

    # ---- informational: the same host-buffer call with binary16 LLRs (half the PCIe bytes)
    h16 = torch.empty(llr.shape, dtype=torch.float16, pin_memory=True)
    h16.copy_(llr)

    def step_host16():
        dec.decode_raw(h16.data_ptr(), tdb.LLR_F16, tdb.MEM_HOST, batch, bits=h_bits.data_ptr(), stream=sp)
    step_host16()
    torch.cuda.synchronize()
    n16 = max(1, args.steps // 4)
    t16 = time.perf_counter()
    for _ in range(n16):
        step_host16()
    torch.cuda.synchronize()
    ms_e2e16 = 1e3 * (time.perf_counter() - t16) / n16

    # ---- informational: 8-bit fixed-point LLRs (the decoder's own channel format: a quarter of the bytes)
    h8 = torch.empty(llr.shape, dtype=torch.int8, pin_memory=True)
    h8.copy_(torch.clamp(torch.round(llr * 8.0), -127, 127).to(torch.int8))

    def step_host8():
        dec.decode_raw(h8.data_ptr(), tdb.LLR_S8, tdb.MEM_HOST, batch, bits=h_bits.data_ptr(), stream=sp)
    step_host8()
    torch.cuda.synchronize()
    e2e8_ok = bool((h_bits.to(dev) == out_bits).all().item())  # same quantiser: same decisions as the float path
    t8 = time.perf_counter()
    for _ in range(n16):
        step_host8()
    torch.cuda.synchronize()
    ms_e2e8 = 1e3 * (time.perf_counter() - t8) / n16

    # per-rank host-to-device rate of the headline end-to-end leg (float32 LLRs), for the scaling diagnosis
    h2d_gbs = batch * (3 * K + 12) * 4 * args.steps / (ms_e2e * 1e-3) / 1e9
    h2d_per_rank = [h2d_gbs]
    et_ms = et_info.pop("_ms", None) if et_info else None
    crc_ms = et_info.pop("_crc_ms", None) if et_info else None
    lm_ms_all = lm_info["ms_per_step"] if lm_info else 0.0
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e16, ms_e2e8, et_ms or 0.0, crc_ms or 0.0, lm_ms_all], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e16, ms_e2e8, lm_ms_all = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[6])
        if et_ms is not None:
            et_ms = float(t[4])
        if crc_ms is not None:
            crc_ms = float(t[5])
        b = torch.tensor([ber, et_info["mean_iterations"] if et_info else 0.0, et_info["ber"] if et_info else 0.0,
                          lm_info["ber"] if lm_info else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        ber = float(b[0]) / world
        if et_info:
            et_info["mean_iterations"], et_info["ber"] = float(b[1]) / world, float(b[2]) / world
        if lm_info:
            lm_info["ber"] = float(b[3]) / world
        g = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g, torch.tensor([h2d_gbs], device=dev, dtype=torch.float64))
        h2d_per_rank = [float(x[0]) for x in g]
    if et_info:
        et_info["value"] = world * batch * K / (et_ms * 1e-3) / 1e9
        et_info["unit"] = "Gbit/s, all ranks (max time over ranks)"
        if crc_ms is not None:
            et_info["crc24b_rule"]["value"] = world * batch * K / (crc_ms * 1e-3) / 1e9
    if lm_info:
        lm_info["ms_per_step"] = lm_ms_all
        lm_info["value"] = world * batch * K / (lm_ms_all * 1e-3) / 1e9
        lm_info["unit"] = "Gbit/s, all ranks"

    if rank == 0:
        total_bits = float(world) * batch * K * args.steps
        value = total_bits / (ms * 1e-3) / 1e9
        e2e = total_bits / (ms_e2e * 1e-3) / 1e9
        hbm_peak, sm_max_mhz, peak_src = measured_peaks()
        kernel_ms = ms / args.steps / max(launches_per_step, 1)
        alg_bytes = batch * ((3 * K + 12) * 4 + K)  # LLRs read once + one byte per decision written
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        sm_count = plan["sm_count"]
        alu_peak = sm_count * 128 * sm_max_mhz * 1e6 * 2  # packed 16-bit lane-ops/s
        alu_peak_meas = sm_count * rate_mix * sm_max_mhz * 1e6 * 2  # the same with the measured issue rate of the mix
        alu_ach = (batch * K / (kernel_ms * 1e-3)) * OPS_PER_INFO_BIT
        line = {
            "metric": "decoded info Gbit/s, K=6144, 8 iterations", "value": value, "unit": "Gbit/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16",
            "data": "synthetic", "config": workload_config(args, batch, plan),
            "e2e": {"value": e2e, "unit": "Gbit/s", "h2d_bytes_per_step": batch * (3 * K + 12) * 4,
                    "d2h_bytes_per_step": batch * K, "matches_device_path": e2e_ok,
                    "h2d_gb_s_per_rank": h2d_per_rank,
                    "with_float16_llrs": {"value": world * batch * K / (ms_e2e16 * 1e-3) / 1e9, "unit": "Gbit/s, all ranks",
                                          "h2d_bytes_per_step": batch * (3 * K + 12) * 2},
                    "with_int8_llrs": {"value": world * batch * K / (ms_e2e8 * 1e-3) / 1e9, "unit": "Gbit/s, all ranks",
                                       "h2d_bytes_per_step": batch * (3 * K + 12), "matches_device_path": e2e8_ok,
                                       "note": "the decoder's own channel format (8-bit fixed point, ITTC/log_map.cpp:1283 dectobin intent): the documented host wire format"}},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "fast_s16_kernel", "kernel_ms": kernel_ms,
                         "note": "secondary bound: the path is ALU-bound by design (SURVEY.md 8d); see roofline.alu",
                         "alu": {"achieved": alu_ach / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s (packed 16-bit)",
                                 "frac": alu_ach / alu_peak, "ops_per_info_bit": OPS_PER_INFO_BIT,
                                 "peak_measured": alu_peak_meas / 1e12, "frac_of_measured": alu_ach / alu_peak_meas,
                                 "issue_rate_measured": {"viaddmnmx_viadd_mix": rate_mix, "viaddmnmx_alone": rate_alu,
                                                         "unit": "thread-ops/clk/SM (128 = one warp-instruction per clock per sub-partition)",
                                                         "how": "tdb200_ubench_issue_rate, run in this process before the timed region"}}},
            "ber": ber,
        }
        if et_info:
            line["early_termination"] = et_info
        if lm_info:
            # 30 max* per trellis step, each max + |difference| + correction + add instead of one max: 71 + 4 * 30 operations
            lm_ops = 2 * N_ITER * (71 + 4 * 30) * (K + 3) / K
            lm_ach = (world * batch * K / (lm_info["ms_per_step"] * 1e-3)) * lm_ops / world
            lm_info["roofline_alu"] = {"achieved": lm_ach / 1e12, "peak": alu_peak / 1e12, "peak_measured": alu_peak_meas / 1e12,
                                       "unit": "Tlane-op/s (packed 16-bit), per GPU", "frac": lm_ach / alu_peak,
                                       "frac_of_measured": lm_ach / alu_peak_meas, "ops_per_info_bit": lm_ops}
            line["logmap_s16"] = lm_info
        if f64_info:
            line["logmap_f64"] = f64_info
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            decode, kind = cpu_reference_decoder()
            n_sample = args.ref_sample or min(batch, 512 * cores)  # ~10-15 s of CPU work
            sample = llr[:n_sample].double().cpu().numpy()
            out, secs = decode(sample, cores)
            line["cpu_baseline"] = {
                "value": n_sample * K / secs / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
                "sample": "first %d codeblocks of the batch, fp64 LUT Log-MAP, 8 iterations, one codeword per thread, %.1f s" % (n_sample, secs),
                "ber": float((out != bits[:n_sample].cpu().numpy()).sum()) / (n_sample * K)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, when the
    # environment sets NCCL_DEBUG) is sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="codeblocks per GPU per step")
    ap.add_argument("--ebn0", type=float, default=1.0)
    ap.add_argument("--algo", default="maxlog_s16")
    ap.add_argument("--ref-sample", type=int, default=0, help="codeblocks per CPU-baseline step (0 = 4 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-early-term", action="store_true", help="skip the informational early-termination leg")
    ap.add_argument("--no-logmap", action="store_true", help="skip the Log-MAP (logmap_s16) leg")
    ap.add_argument("--no-f64", action="store_true", help="skip the fp64 reference-order (logmap_f64) leg")
    ap.add_argument("--sub-block", type=int, default=0, help="trellis steps per sub-block (0 = the library's plan)")
    ap.add_argument("--guard", type=int, default=0, help="warm-up steps across sub-block boundaries (with --sub-block)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
