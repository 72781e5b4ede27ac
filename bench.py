#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 (LB)MPC QP engine.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): LBMPC QP solves/sec (N=50, FP64).  Workload at every N: BASELINE.json configs[1],
"Batched 1024 initial conditions, Moore-Greitzer LBMPC N=50, FP64" per GPU (weak scaling: each rank solves
its own 1024-QP shard; QPs are independent, there is no data-path collective — NCCL only reduces the timing
and the result statistics).  A "step" is one lbmpc_solve_batch call over the rank's batch.

  value     device-resident throughput: inputs/outputs stay in HBM (device-pointer handle), each step timed
            with CUDA events on the launch stream, L2 flushed between steps (256 MiB memset, outside the events)
  e2e       the same metric through the host-pointer C-ABI call: pinned host buffers, H2D of the step's inputs
            and D2H of its results inside the timed region
  roofline  FP64-FMA roofline of the IPM kernel: algorithmic flops (iters x (1355 N + 108 n_g))
            / CUDA-event kernel time, against the DFMA peak measured on this device by lbmpc_measure_fp64_peak
  cpu_baseline  the CPU oracle port (same algorithm, C, POSIX threads over QPs) on the host cores
The reference arm (--impl reference) times that same CPU port: the reference's own solver is MATLAB fmincon /
CasADi-IPOPT, neither of which exists on this box (DESIGN.md "Reference arm").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "learning-based-mpc_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "LBMPC QP solves/sec (N=50, FP64)"
UNIT = "QP/s"
HORIZON = 50
BATCH_PER_GPU = 1024
FORM, VARIANT = "C", "LBMPC"
WORKLOAD = ("BASELINE configs[1]: batched 1024 initial conditions per GPU, Moore-Greitzer C-form LBMPC "
            "(robust rows on x_1: 24 polytope rows, 500 box rows), N=50, FP64, cold start")


def flops_per_iter(N, n_g):
    """SURVEY.md §8(d) closed form of the algorithmic flops of one interior-point iteration."""
    return 1355.0 * N + 108.0 * n_g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def problem_inputs(rank, world):
    from lbmpc_b200.dist import sample_initial_states, shard_range
    full = sample_initial_states(BATCH_PER_GPU * world, seed=0)      # config 2 distribution, seed 0
    lo, hi = shard_range(BATCH_PER_GPU * world, rank, world)
    return np.ascontiguousarray(full[lo:hi])


def cpu_port_throughput(dx0, min_seconds, threads):
    """QP/s of the CPU oracle port on `threads` host threads over repeated passes of the batch."""
    import lbmpc_b200
    from oracle_py import OracleProblem
    P = OracleProblem(FORM, VARIANT, lbmpc_b200.moore_greitzer_model(VARIANT), HORIZON)
    P.solve_batch(dx0[:64], nthreads=threads)                          # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        P.solve_batch(dx0, nthreads=threads)
        n += dx0.shape[0]
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            return n / dt, n, dt


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of the path (oracle port) on all host threads, rank 0 only."""
    if rank != 0:
        return
    dx0 = problem_inputs(0, 1)
    threads = os.cpu_count() or 1
    P_warm = max(args.warmup, 1)
    import lbmpc_b200
    from oracle_py import OracleProblem
    P = OracleProblem(FORM, VARIANT, lbmpc_b200.moore_greitzer_model(VARIANT), HORIZON)
    for _ in range(P_warm):
        P.solve_batch(dx0, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        P.solve_batch(dx0, nthreads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * dx0.shape[0] / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "horizon": HORIZON, "batch_per_step": int(dx0.shape[0])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} passes over the 1024-QP batch, {threads} POSIX threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference solver is MATLAB fmincon / CasADi-IPOPT (not installable here); this arm times the "
                    "repo's CPU port of the same path (oracle/lbmpc_oracle.c)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import lbmpc_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    mdl = lbmpc_b200.moore_greitzer_model(VARIANT)
    dx0_h = problem_inputs(rank, world)
    nb = dx0_h.shape[0]

    # ---------------- device-resident arm ----------------
    sol = lbmpc_b200.Solver(mdl, FORM, VARIANT, HORIZON, device=local_rank, device_pointers=True)
    dx0_d = torch.from_numpy(dx0_h).to(dev)
    out = sol.solve_batch(dx0_d, want_x=False)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    for _ in range(args.warmup):
        flush.zero_()
        sol.solve_batch(dx0_d, want_x=False, out=out)
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = sol.kernel_launches
    kern_ms = []
    for e0, e1 in ev:
        flush.zero_()
        e0.record(stream)
        sol.solve_batch(dx0_d, want_x=False, out=out)
        e1.record(stream)
        kern_ms.append(None)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(step_ms))
    launches = sol.kernel_launches - launches0
    kernel_ms_last = sol.last_kernel_ms
    clocks = sampler.stop()
    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy()

    # ---------------- end-to-end arm (host buffers, copies inside the timed region) ----------------
    hsol = lbmpc_b200.Solver(mdl, FORM, VARIANT, HORIZON, device=local_rank, max_batch=nb)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_in = pin((nb, 4), torch.float64)
    h_in.copy_(torch.from_numpy(dx0_h))
    h_uc, h_th, h_obj = pin((nb, HORIZON, 1), torch.float64), pin((nb, 1), torch.float64), pin((nb,), torch.float64)
    h_it, h_st = pin((nb,), torch.int32), pin((nb,), torch.int32)
    from lbmpc_b200.capi import _ptr

    def e2e_step():
        rc = hsol.lib.lbmpc_solve_batch(hsol.h, nb, _ptr(h_in), None, None, None, _ptr(h_uc), _ptr(h_th), None,
                                        _ptr(h_obj), _ptr(h_it), _ptr(h_st), None)
        if rc != 0:
            raise RuntimeError(hsol.lib.lbmpc_last_error().decode())
    for _ in range(args.warmup):
        e2e_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                                              # synchronous: returns after the D2H copies
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    h2d = nb * 4 * 8
    d2h = nb * (HORIZON * 8 + 8 + 8 + 4 + 4)
    assert np.array_equal(h_st.numpy(), status) and np.array_equal(h_it.numpy(), iters)

    # ---------------- reductions over ranks (max time, summed work) ----------------
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        st = lbmpc_b200.dist.reduce_stats(out["status"], out["iters"], out["obj"])
        n_total = nb * world
    else:
        n_total = nb
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_total * args.steps / (total_ms * 1e-3)
    e2e_value = n_total * args.steps / e2e_s

    # ---------------- roofline of the IPM kernel (rank 0's launch) ----------------
    n_g = 24
    alg_flops = float(iters.sum()) * flops_per_iter(HORIZON, n_g)
    k_ms = float(np.median(step_ms))                             # events bracket exactly one kernel launch
    peak_tf = lbmpc_b200.measure_fp64_peak(local_rank)
    achieved_tf = alg_flops / (k_ms * 1e-3) * 1e-12
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            traffic = json.load(f).get("ipm_kernel_lbmpc_n50_b1024", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    alg_bytes = nb * (32 + 8 * HORIZON + 24)
    roofline = {"bound": "fp64_fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                "peak_source": "lbmpc_measure_fp64_peak (register-resident DFMA chains) on this device; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flops_per_launch": alg_flops, "kernel_ms": k_ms,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (k_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                        "note": "HBM is not the bound: the iterate lives in shared memory"}}

    # ---------------- CPU baseline (oracle port) ----------------
    threads = os.cpu_count() or 1
    cpu_v, cpu_n, cpu_dt = cpu_port_throughput(dx0_h, 3.0, threads)
    hist = np.bincount(iters, minlength=1).tolist()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "horizon": HORIZON, "batch_per_gpu": BATCH_PER_GPU,
                       "l2": "flushed between timed steps (256 MiB memset outside the CUDA events)",
                       "parallelism": f"dp{world} (independent QP shards, no data-path collective)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_n} QPs ({cpu_n // nb} passes over the batch) in {cpu_dt:.1f} s, "
                                       f"{threads} POSIX threads"},
            "p50_us_per_solve": 1e3 * float(np.median(step_ms)) / nb,
            "iterations": {"mean": float(iters.mean()), "max": int(iters.max()), "hist": hist},
            "status_counts": np.bincount(status, minlength=4).tolist(),
            "slots_per_cta": sol.slots_per_cta, "kernel_ms_last_launch": kernel_ms_last}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
