#!/usr/bin/env python
"""Benchmark of the fused trajectory rollout (BASELINE.json metric: particle-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f16x3|tf32x3|fp32|tf32|bf16]

The headline number is measured in the fp32-grade mode f16x3 (drift network AND mixture-score contractions on tcgen05
tensor cores with the 3-pass fp16 (hi, lo) split of power-of-two scaled operands; parity-tested to the north-star
tolerance like tf32x3 and the fp32 SIMT anchor).  --fast-mode additionally times the reduced-precision bf16 kernel
and reports it separately with its tolerance.

A "step" is one complete rollout of the workload's particle batch through all grid times (the quantity the
reference times as eval/sample_time, solver/oc.py:148-158) followed by the estimator reduction; under
torchrun every rank integrates its own shard of particles (weak scaling: 65536 per GPU, RNG counter = global
particle index) and the only exchange is the all_gather of 8 doubles per rank.

Workload = BASELINE.json configs[1]: ManyModes d=50 (16 modes), RDS vp-ref (VP beta 0.1..20), diagonal-GMM
reference, target_informed (ScoreCtrl) drift, EI integrator, 200 steps, batch 65536, synthetic weights.
Prints ONE JSON line (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "many_modes_d50_M16_rds_gmmref_scorectrl_ei_K200_B65536"
B_PER_GPU, K_STEPS, DIM, MODES = 65536, 200, 50, 16
# algorithmic tensor FLOPs per particle-step: MLP 256 d + 16384, two diag-GMM scores 8 M d each
FLOPS_PER_PARTICLE_STEP = (256 * DIM + 16384) + 2 * 8 * MODES * DIM
METRIC, UNIT = "particle_steps_per_sec", "particle-steps/s"
CPU_SAMPLE_B = 16384
# Executed warp-instructions per particle-step of the f16x3 benchmark kernel, counted by ncu
# (profiles/r01_mix_summary.md: smsp__inst_executed.sum / (B K)): the SIMT work (mixture quadratic forms, erf GELU,
# Philox + Box-Muller, integrator) that bounds this path; used for the issue-slot figure in "roofline".
WARP_INSTR_PER_PARTICLE_STEP = {"f16x3": 3.5232e9 / (B_PER_GPU * K_STEPS)}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the same kernel (same capture)
DRAM_BYTES_PER_LAUNCH = {"f16x3": 16_284_672 + 768}
FAST_MODE_TOLERANCE = "log Z within 5e-2 abs, 99% of log-weights within 1e-2 rel (tests/test_rollout_parity_gpu.py)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.ready, self.recording = threading.Event(), False  # NVML init happens BEFORE the timed region

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)  # first calls are slow: make them before the timed region
            nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.ready.set()
            while not self._stop_evt.is_set():
                # polled continuously from before the warm-up (a first NVML query inside the timed region stalls the
                # launch path for tens of ms); only samples taken while `recording` are reported
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.recording:
                    self.samples.append(mhz)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_case(B):
    from tests.cases import case_ei_many_modes
    return case_ei_many_modes(K=K_STEPS, B=B, d=DIM, M=MODES)


def cpu_port_throughput(B_sample, threads, steps=1, warmup=0):
    """Times the oracle (torch-on-CPU restatement of the reference's loop, pinned to the reference by tests/golden)
    on a bounded sample of the workload: B_sample particles through all K_STEPS grid times."""
    import torch
    from oracle import rollout_oracle as O
    torch.set_num_threads(threads)
    case = build_case(B_sample)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B_sample, DIM, generator=g)
    noise = torch.randn(K_STEPS, B_sample, DIM, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.rollout(case["problem"], x0, noise)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return B_sample * K_STEPS / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python/torch and its
    source tree does not travel to the GPU box, so the arm runs the oracle port (oracle/rollout_oracle.py, pinned
    to the reference by tests/golden) with all host threads on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    B_sample = 8192
    value, dt = cpu_port_throughput(B_sample, threads, steps=max(1, min(args.steps, 5)), warmup=min(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"B={B_sample} of {B_PER_GPU}, K={K_STEPS}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"B={B_sample}, K={K_STEPS}, d={DIM}, torch {torch.__version__} CPU"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sde_sampler_lrds_b200 import _native as N
    from sde_sampler_lrds_b200.estimators import estimator_partials, gather_and_merge, metrics_from_partials
    from tests.product_builders import Built

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner at communicator
    # creation) are sent to stderr; the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rollout has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD if world > 1 else None

    B = B_PER_GPU
    case = build_case(B)
    built = Built(case, dev, args.precision)
    g = torch.Generator().manual_seed(1 + rank)
    x0_host = torch.randn(B, DIM, generator=g).pin_memory()
    x0 = x0_host.to(dev)
    offset = rank * B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(x_dev, seed):
        x, rnd, _ = built.simulate(x_dev, None, seed=seed, particle_offset=offset)
        part = estimator_partials(rnd, group=None) if group is None else None
        if group is not None:
            # the one exchange of the path: 8 doubles per rank
            r = rnd.reshape(-1)
            blocks = N.lib().lrds_estimator_blocks(r.numel())
            scratch = torch.empty(8 * (blocks + 1), device=dev, dtype=torch.float64)
            out = torch.empty(8, device=dev, dtype=torch.float64)
            N.check(N.lib().lrds_estimator_partials(N.ptr(r), r.numel(), N.ptr(out), N.ptr(scratch), N.stream_ptr(dev)))
            part = gather_and_merge(out, group)
        return x, rnd, part

    def barrier():
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=10)
    prev = None
    for w in range(args.warmup):
        # two generations of outputs stay alive, as in the timed loop: the caching allocator then owns every buffer
        # the timed iterations need (a cudaMalloc inside the timed region would stall the launch path for milliseconds)
        cur = step(x0, 1000 + w)
        prev = cur
    barrier()
    del prev, cur

    # ---- device-resident timing (value) + the dominant kernel alone (roofline) -----------------------------------
    sampler.recording = True
    launches0 = N.launch_count()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(args.steps)]
    last = None
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations (not timed)
        a, b, c = ev[i]
        a.record()
        x, rnd, _ = built.simulate(x0, None, seed=2000 + i, particle_offset=offset)
        b.record()
        r = rnd.reshape(-1)
        blocks = N.lib().lrds_estimator_blocks(r.numel())
        scratch = torch.empty(8 * (blocks + 1), device=dev, dtype=torch.float64)
        out = torch.empty(8, device=dev, dtype=torch.float64)
        N.check(N.lib().lrds_estimator_partials(N.ptr(r), r.numel(), N.ptr(out), N.ptr(scratch), N.stream_ptr(dev)))
        last = gather_and_merge(out, group, on_device=True) if group is not None else out
        c.record()
    barrier()
    launches = N.launch_count() - launches0
    sampler.recording = False
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
    kern_ms_each = [round(a.elapsed_time(b), 3) for a, b, _ in ev]
    t = torch.tensor([step_ms], device=dev, dtype=torch.float64)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B * K_STEPS / (ms_per_step * 1e-3)

    # ---- reduced-precision fast mode, reported separately ------------------------------------------------------
    fast = None
    if args.fast_mode:
        built_fast = Built(case, dev, "bf16")
        for w in range(3):
            built_fast.simulate(x0, None, seed=4000 + w, particle_offset=offset)
        barrier()
        fev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(3)]
        for i, (fa, fb) in enumerate(fev):
            flush.fill_(i & 0xFF)
            fa.record()
            _, rnd_f, _ = built_fast.simulate(x0, None, seed=2000 + i, particle_offset=offset)
            fb.record()
        barrier()
        fms = sum(fa.elapsed_time(fb) for fa, fb in fev) / len(fev)
        tf_ = torch.tensor([fms], device=dev, dtype=torch.float64)
        if group is not None:
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
        mf = metrics_from_partials(estimator_partials(rnd_f))
        fast = {"precision": "bf16", "value": world * B * K_STEPS / (float(tf_.item()) * 1e-3), "unit": UNIT,
                "kernel_ms": float(tf_.item()), "tolerance": FAST_MODE_TOLERANCE,
                "check": {"log_norm_const_is": mf["log_norm_const_is"], "elbo": mf["elbo"]}}

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    x_host_out = torch.empty(B, DIM).pin_memory()
    rnd_host_out = torch.empty(B, 1).pin_memory()
    def e2e_step(seed):
        xd = x0_host.to(dev, non_blocking=True)
        x, rnd, part = step(xd, seed)
        x_host_out.copy_(x, non_blocking=True)
        rnd_host_out.copy_(rnd, non_blocking=True)
        torch.cuda.synchronize(dev)

    for w in range(2):  # untimed: first touch of the pinned buffers, allocator entries of the staging tensors
        e2e_step(2900 + w)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(3000 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K_STEPS * args.steps / float(t.item())
    h2d = x0_host.numel() * 4
    d2h = (x_host_out.numel() + rnd_host_out.numel()) * 4
    clocks = sampler.stop()

    if rank != 0:
        if group is not None:
            dist.destroy_process_group()
        return

    peaks, how = measured_peaks()
    achieved = FLOPS_PER_PARTICLE_STEP * B * K_STEPS / (kern_ms * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops"])
    m = metrics_from_partials(last.cpu() if hasattr(last, "cpu") else last)
    threads = os.cpu_count() or 1
    cpu_value = cpu_dt = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_value, cpu_dt = cpu_port_throughput(CPU_SAMPLE_B, threads, steps=2, warmup=1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "particles_per_gpu": B, "grid_steps": K_STEPS, "dim": DIM, "modes": MODES,
                   "precision": args.precision, "noise": "in-kernel Philox4x32-10", "l2": "flushed between timed iterations",
                   "parallelism": f"particles sharded over {world} GPU(s), no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": DRAM_BYTES_PER_LAUNCH.get(args.precision), "kernel_ms": kern_ms, "kernel_ms_each": kern_ms_each, "peak_source": how + ", dense bf16 burst",
                     "flops_per_particle_step": FLOPS_PER_PARTICLE_STEP},
        "check": {"log_norm_const_is": m["log_norm_const_is"], "elbo": m["elbo"], "ess": m["effective_sample_size"]},
    }
    wi = WARP_INSTR_PER_PARTICLE_STEP.get(args.precision)
    if wi is not None and clocks.get("sm_mhz"):
        # explanatory: the kernel is bound by SIMT instruction issue, not by the tensor pipe
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issued = wi * B * K_STEPS / (kern_ms * 1e-3)
        slots = sms * 4 * clocks["sm_mhz"] * 1e6
        line["roofline"]["simt_issue"] = {"achieved": issued / 1e9, "peak": slots / 1e9, "unit": "G warp-instr/s",
                                          "frac": issued / slots, "warp_instr_per_particle_step": wi,
                                          "source": "ncu smsp__inst_executed.sum (profiles/r01_mix_summary.md) / live kernel time"}
    if fast is not None:
        line["fast_mode"] = fast
    if cpu_value is not None:
        line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"B={CPU_SAMPLE_B} of {B} particles, all K={K_STEPS} steps, d={DIM}; "
                                          f"{cpu_dt:.1f} s per rollout, torch {torch.__version__} CPU threads={threads}"}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if group is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "f16x3", "tf32", "bf16"])
    ap.add_argument("--fast-mode", action="store_true", help="also time the reduced-precision bf16 kernel")
    ap.add_argument("--no-fast-mode", action="store_true", help=argparse.SUPPRESS)  # accepted for older command lines
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
