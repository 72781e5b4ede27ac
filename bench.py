#!/usr/bin/env python3
"""bench.py — benchmark of the B200 (LB)MPC QP engine on the BASELINE.json configs.

    python bench.py --gpus N --steps K --warmup W [--config C]      (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W [--config C]

Metric (BASELINE.json): LBMPC QP solves/sec (FP64).  --config picks the BASELINE.json configs[] entry:
  1 (default, the headline)  batched 1024 initial conditions per GPU, Moore-Greitzer C-form LBMPC, N=50      weak scaling
  2  tracking LMPC with the 616-row terminal invariant set, batch 16384 per GPU, per-QP references, N=50       weak scaling
  3  long-horizon LBMPC N=200 with the learned-oracle correction (L2NW windows from train_data), batch 65536
     in TOTAL, split over the N GPUs                                                                           strong scaling
  4  Monte-Carlo closed loop, 125 000 scenarios per GPU (1 M on 8 GPUs) x T=100 steps, randomised disturbance,
     oracle + plant + solve per step on the GPU                                                                weak scaling
QPs / scenarios are independent: ranks own contiguous index shards, there is no data-path collective; NCCL gathers
the per-rank results to rank 0 (inside the e2e timing when N > 1) and reduces timings / statistics.

  value     device-resident throughput: inputs/outputs stay in HBM (device-pointer handle), each step timed with CUDA events
            on the launch stream, L2 flushed between steps (256 MiB memset, outside the events)
  e2e       the same metric with HOST buffers: config 1 through the host-pointer C-ABI call (pinned caller arrays are read
            and written in place by the kernel, zero-copy; `e2e_pageable` is the same call with pageable numpy arrays, which
            are staged), configs 2-4 pinned H2D copy of the step's inputs + device-pointer call + NCCL gather + D2H
  roofline  FP64-FMA roofline of the IPM kernel: algorithmic flops (iters x (1355 N + 108 n_g)) / CUDA-event
            kernel time, against the DFMA peak measured on this device by lbmpc_measure_fp64_peak
  cpu_baseline  the CPU oracle port (same algorithm, C, POSIX threads over QPs) on the host cores, bounded sample
The reference arm (--impl reference) times that same CPU port: the reference's own solver is MATLAB fmincon / CasADi-IPOPT,
neither of which exists on this box (DESIGN.md "Reference arm").  Both arms print the identical `config` object.
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

UNIT = "QP/s"
CONFIGS = {
    1: dict(metric="LBMPC QP solves/sec (N=50, FP64)", form="C", variant="LBMPC", N=50, batch=1024, ng=24, kind="solve",
            scaling="weak",
            workload="BASELINE configs[1]: batched 1024 initial conditions per GPU, Moore-Greitzer C-form LBMPC "
                     "(robust rows on x_1: 24 polytope rows, 500 box rows), N=50, FP64, cold start"),
    2: dict(metric="tracking-LMPC QP solves/sec (N=50, FP64)", form="C", variant="LMPC", N=50, batch=16384, ng=616, kind="solve",
            scaling="weak", ref=True,
            workload="BASELINE configs[2]: trackingMPC setpoint-tracking variant, C-form LMPC with the 616-row terminal invariant "
                     "set (term_set.mat), per-QP references, batch 16384 per GPU, N=50, FP64, cold start"),
    3: dict(metric="LBMPC QP solves/sec (N=200, FP64)", form="C", variant="LBMPC", N=200, batch=65536, ng=24, kind="oracle_solve",
            scaling="strong", q=100,
            workload="BASELINE configs[3]: long-horizon LBMPC N=200 with the learned-oracle correction (L2NW, q=100 windows of "
                     "train_data.mat, offsets along the u=u_eq rollout), batch 65536 in total split over the GPUs, FP64"),
    4: dict(metric="LBMPC closed-loop QP solves/sec (N=50, FP64)", form="C", variant="LBMPC", N=50, batch=125000, ng=24,
            kind="closed_loop", scaling="weak", T=100, q=100,
            workload="BASELINE configs[4]: Monte-Carlo closed loop, 125000 scenarios per GPU (1 M on 8 GPUs) x 100 steps, uniform "
                     "state disturbance, data window q=100, L2NW oracle + RK4 plant + solve per step on the GPU, N=50, FP64"),
}


def flops_per_iter(N, n_g):
    """SURVEY.md §8(d) closed form of the algorithmic flops of one interior-point iteration."""
    return 1355.0 * N + 108.0 * n_g


def config_object(cfg, gpus):
    """The `config` object both arms print (identical for the same --config / --gpus)."""
    per_gpu = cfg["batch"] if cfg["scaling"] == "weak" else cfg["batch"] // gpus
    c = {"workload": cfg["workload"], "horizon": cfg["N"], "batch_per_gpu": per_gpu, "batch_total": per_gpu * gpus,
         "l2": "flushed between timed steps (256 MiB memset outside the CUDA events)",
         "parallelism": f"dp{gpus} (independent QP shards, no data-path collective; NCCL gather of the results)"}
    if cfg["kind"] == "closed_loop":
        c["steps_per_scenario"] = cfg["T"]
    return c


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
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE, text=True)
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


def problem_inputs(cfg, rank, world):
    """Synthetic inputs of the rank's shard, deterministic in the GLOBAL index (the same QPs at every world size)."""
    from lbmpc_b200.dist import sample_initial_states, shard_range
    import lbmpc_b200
    cid = cfg["id"]
    total = cfg["batch"] * world if cfg["scaling"] == "weak" else cfg["batch"]
    lo, hi = shard_range(total, rank, world)
    seed = {1: 0, 2: 1, 3: 2, 4: 3}[cid]
    inp = {"dx0": np.ascontiguousarray(sample_initial_states(total, seed=seed)[lo:hi])}
    if cfg.get("ref"):      # per-QP tracking references on the steady-state manifold, half of them zero
        rng = np.random.default_rng(1)
        mdl = lbmpc_b200.moore_greitzer_model(cfg["variant"])
        xref = mdl["LAMBDA"][:, 0][None, :] * rng.uniform(-0.1, 0.1, (total, 1)) * (rng.random((total, 1)) < 0.5)
        inp["dx_ref"] = np.ascontiguousarray(xref[lo:hi])
    if cfg["kind"] == "oracle_solve":   # q-sample windows of the reference's training data at random offsets (seed 2)
        data = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))["casadi_train_data__data"]
        q = cfg["q"]
        offs = np.random.default_rng(2).integers(0, data.shape[1] - q, total)[lo:hi]
        idx = offs[:, None] + np.arange(q)[None, :]
        inp["X"] = np.ascontiguousarray(data[:3][:, idx].transpose(1, 2, 0))          # (nb, q, 3)
        inp["Y"] = np.ascontiguousarray(data[3:7][:, idx].transpose(1, 2, 0))
    if cfg["kind"] == "closed_loop":
        inp["x_init"] = np.ascontiguousarray(lbmpc_b200.X_WP[None, :] + inp["dx0"])
        inp["scenario0"] = lo
    return inp


def cpu_port_step(cfg, P, inp, threads, nmax=None):
    """One pass of the CPU oracle port over (a bounded sample of) the rank-0 inputs; returns the number of QPs solved."""
    n = inp["dx0"].shape[0] if nmax is None else min(nmax, inp["dx0"].shape[0])
    if cfg["kind"] == "solve":
        P.solve_batch(inp["dx0"][:n], None if "dx_ref" not in inp else inp["dx_ref"][:n], nthreads=threads)
        return n
    if cfg["kind"] == "oracle_solve":
        N = cfg["N"]
        d = np.stack([P.oracle_offsets(inp["dx0"][b], np.zeros(N), np.ascontiguousarray(inp["X"][b].T),
                                       np.ascontiguousarray(inp["Y"][b].T)) for b in range(n)])
        P.solve_batch(inp["dx0"][:n], None, d, nthreads=threads)
        return n
    import lbmpc_b200
    T = cfg["T"]
    wbar = np.array([0.02, 5e-4, 0.0, 0.0])
    from concurrent.futures import ThreadPoolExecutor
    def one(b):
        P.closed_loop(lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), inp["x_init"][b], T, q=cfg["q"], use_oracle=True, wbar=wbar, seed=7,
                      scenario=inp["scenario0"] + b)
    with ThreadPoolExecutor(threads) as ex:      # the C loop releases the GIL (ctypes)
        list(ex.map(one, range(n)))
    return n * T


def cpu_sample_size(cfg):
    """Bounded CPU sample per pass (about 1-3 s of work on 16-32 threads)."""
    return {"solve": None if cfg["batch"] <= 4096 else 4096, "oracle_solve": 1024, "closed_loop": 64}[cfg["kind"]]


def run_reference(args, cfg, rank, world):
    """Reference arm: the CPU restatement of the path (oracle port) on all host threads, rank 0 only."""
    if rank != 0:
        return
    import lbmpc_b200
    from oracle_py import OracleProblem
    inp = problem_inputs(cfg, 0, 1)
    threads = os.cpu_count() or 1
    P = OracleProblem(cfg["form"], cfg["variant"], lbmpc_b200.moore_greitzer_model(cfg["variant"]), cfg["N"])
    nmax = cpu_sample_size(cfg)
    for _ in range(max(args.warmup, 1)):
        cpu_port_step(cfg, P, inp, threads, nmax)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += cpu_port_step(cfg, P, inp, threads, nmax)
    dt = time.perf_counter() - t0
    value = n / dt
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_object(cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} passes over {n // args.steps} QPs of the rank-0 workload, {threads} POSIX threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference solver is MATLAB fmincon / CasADi-IPOPT (not installable here); this arm times the "
                    "repo's CPU port of the same path (oracle/lbmpc_oracle.c)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--kernel", default=None, choices=["auto", "warp", "cta", "stream", "mixed"],
                    help="force a thread mapping (experiments; 'mixed' reports dtype f32+f64 and is never the headline)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config], id=args.config)
    heavy = cfg["kind"] != "solve" or cfg["batch"] > 4096
    if args.steps is None:
        args.steps = 3 if heavy else 20
    if args.warmup is None:
        args.warmup = 3 if heavy else 5
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import lbmpc_b200
    from lbmpc_b200.capi import _ptr
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    mdl = lbmpc_b200.moore_greitzer_model(cfg["variant"])
    inp = problem_inputs(cfg, rank, world)
    nb, N, kind = inp["dx0"].shape[0], cfg["N"], cfg["kind"]
    total = cfg["batch"] * world if cfg["scaling"] == "weak" else cfg["batch"]
    qp_per_step = nb * (cfg["T"] if kind == "closed_loop" else 1)
    x_eq, u_eq, wbar = lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), np.array([0.02, 5e-4, 0.0, 0.0])

    # ---------------- device-resident arm ----------------
    sol = lbmpc_b200.Solver(mdl, cfg["form"], cfg["variant"], N, device=local_rank, device_pointers=True, max_batch=nb,
                            kernel=args.kernel)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in = {k: t(v) for k, v in inp.items() if isinstance(v, np.ndarray)}
    du0 = torch.zeros((nb, N, 1), dtype=torch.float64, device=dev) if kind == "oracle_solve" else None
    state = {"out": None}

    def dev_step(T=None):
        if kind == "solve":
            state["out"] = sol.solve_batch(d_in["dx0"], d_in.get("dx_ref"), want_x=False, out=state["out"])
        elif kind == "oracle_solve":
            d_off = sol.oracle_apply(d_in["dx0"], du0, d_in["X"], d_in["Y"])
            state["out"] = sol.solve_batch(d_in["dx0"], d_off=d_off, want_x=False, out=state["out"])
        else:
            state["out"] = sol.closed_loop(d_in["x_init"], T or cfg["T"], x_eq, u_eq, q=cfg["q"], use_oracle=True, wbar=wbar, seed=7,
                                           scenario0=inp["scenario0"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    for _ in range(args.warmup):
        flush.zero_()
        dev_step(T=2 if kind == "closed_loop" else None)     # closed loop: short warm-up runs (same kernels, 2 steps each)
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = sol.kernel_launches
    for e0, e1 in ev:
        flush.zero_()
        e0.record(stream)
        dev_step()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(step_ms))
    launches = sol.kernel_launches - launches0
    kernel_ms_last = sol.last_kernel_ms
    kernel_used = sol.last_kernel
    clocks = sampler.stop()
    out = state["out"]
    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy()

    # ---------------- end-to-end arm (host buffers, copies inside the timed region) ----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    e2e_extra = {}
    if kind == "solve" and "dx_ref" not in inp:
        # host-pointer C-ABI call: pinned caller arrays are mapped and accessed in place by the kernel
        hsol = lbmpc_b200.Solver(mdl, cfg["form"], cfg["variant"], N, device=local_rank, max_batch=nb, kernel=args.kernel)
        pe = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        h_in = pin(inp["dx0"])
        h_uc, h_th, h_obj = pe((nb, N, 1), torch.float64), pe((nb, 1), torch.float64), pe((nb,), torch.float64)
        h_it, h_st = pe((nb,), torch.int32), pe((nb,), torch.int32)

        res_host = {}
        if world == 1:
            def e2e_step():
                rc = hsol.lib.lbmpc_solve_batch(hsol.h, nb, _ptr(h_in), None, None, None, _ptr(h_uc), _ptr(h_th), None,
                                                _ptr(h_obj), _ptr(h_it), _ptr(h_st), None)
                if rc != 0:
                    raise RuntimeError(hsol.lib.lbmpc_last_error().decode())
            h2d = nb * 4 * 8
            d2h = nb * (N * 8 + 8 + 8 + 4 + 4)
            e2e_check = lambda: (np.array_equal(h_st.numpy(), status) and np.array_equal(h_it.numpy(), iters))
            e2e_path = "host-pointer lbmpc_solve_batch, pinned caller arrays accessed in place (zero-copy)"
        else:
            # several ranks: the results have to reach rank 0 over NCCL, so they stay on the device until the gather —
            # pinned inputs -> H2D -> device-pointer call -> ONE packed NCCL gather -> D2H of the gathered block on rank 0
            d_x = torch.empty_like(d_in["dx0"])   # device input buffer of the e2e arm (filled from the pinned host array every step)

            def e2e_step():
                d_x.copy_(h_in, non_blocking=True)
                o = sol.solve_batch(d_x, want_x=False, out=state["out"])
                res_host.update(lbmpc_b200.dist.gather_packed({"u0": o["uc"][:, 0, 0], "obj": o["obj"], "iters": o["iters"],
                                                               "status": o["status"]}, total, dst=0, to_host=True))
            h2d = nb * 4 * 8
            d2h = nb * 24
            e2e_check = lambda: True
            e2e_path = "pinned host inputs -> cudaMemcpyAsync H2D -> device-pointer C-ABI call -> one packed NCCL gather -> D2H on rank 0"
    else:
        # pinned host inputs -> H2D -> device-pointer calls -> NCCL gather to rank 0 -> D2H of the results
        h_in = {k: pin(v) for k, v in inp.items() if isinstance(v, np.ndarray) and k in ("dx0", "dx_ref", "X", "Y", "x_init")}
        res_host = {}

        def e2e_step():
            d = {k: v.to(dev, non_blocking=True) for k, v in h_in.items()}
            if kind == "solve":
                o = sol.solve_batch(d["dx0"], d.get("dx_ref"), want_x=False, out=state["out"])
                res = {"u0": o["uc"][:, 0, 0], "obj": o["obj"], "iters": o["iters"], "status": o["status"]}
            elif kind == "oracle_solve":
                o = sol.solve_batch(d["dx0"], d_off=sol.oracle_apply(d["dx0"], du0, d["X"], d["Y"]), want_x=False, out=state["out"])
                res = {"u0": o["uc"][:, 0, 0], "obj": o["obj"], "iters": o["iters"], "status": o["status"]}
            else:
                o = sol.closed_loop(d["x_init"], cfg["T"], x_eq, u_eq, q=cfg["q"], use_oracle=True, wbar=wbar, seed=7,
                                    scenario0=inp["scenario0"])
                res = {"x_final": o["x"][:, -1, :].contiguous(), "iters_sum": o["iters"].sum(1, dtype=torch.int32),
                       "status_max": o["status"].max(1).values}
            if world > 1:
                res_host.update(lbmpc_b200.dist.gather_packed(res, total, dst=0, to_host=True))   # one gather, one D2H on rank 0
            else:
                for k, v in res.items():
                    res_host[k] = v.cpu()                      # D2H of the results
        h2d = sum(v.numel() * v.element_size() for v in h_in.values())
        d2h = {"solve": nb * 24, "oracle_solve": nb * 24, "closed_loop": nb * 40}[kind]
        e2e_check = lambda: True
        e2e_path = "pinned host inputs -> cudaMemcpyAsync H2D -> device-pointer C-ABI calls -> NCCL gather -> D2H"
    for _ in range(2 if heavy else args.warmup):
        e2e_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    assert e2e_check()
    if kind == "solve" and "dx_ref" not in inp:
        # the same host-pointer call with PAGEABLE numpy arrays (what a MATLAB / plain numpy caller passes): staged copies
        t1 = time.perf_counter()
        for _ in range(args.steps):
            o_pg = hsol.solve_batch(inp["dx0"], want_x=False)
        e2e_extra["e2e_pageable"] = {"value": nb * args.steps / (time.perf_counter() - t1), "unit": UNIT,
                                     "note": "per rank; pageable numpy arrays through lbmpc_solve_batch (staging copies + output allocation)"}
        assert np.array_equal(o_pg["status"], status)
        # single-QP latency: one QP through the host-pointer call (p50 of 50 calls), the number p50_us_per_solve is NOT
        one = lbmpc_b200.Solver(mdl, cfg["form"], cfg["variant"], N, device=local_rank, max_batch=1)
        lat = []
        for i in range(60):
            t2 = time.perf_counter()
            one.solve_batch(inp["dx0"][i % nb:i % nb + 1], want_x=False)
            lat.append(time.perf_counter() - t2)
        e2e_extra["latency_us_single_qp"] = 1e6 * float(np.median(lat[10:]))

    # ---------------- reductions over ranks (max time, summed work) ----------------
    if world > 1:
        tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(tt[0]), float(tt[1])
        work = torch.tensor([float(qp_per_step)], dtype=torch.float64, device=dev)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        qp_total = float(work[0])
        stats = lbmpc_b200.dist.reduce_stats(out["status"].reshape(-1), out["iters"].reshape(-1),
                                             out["obj"] if "obj" in out else torch.zeros(out["status"].numel(), dtype=torch.float64, device=dev))
    else:
        qp_total = float(qp_per_step)
        stats = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = qp_total * args.steps / (total_ms * 1e-3)
    e2e_value = qp_total * args.steps / e2e_s

    # ---------------- roofline of the IPM kernel (rank 0's launch) ----------------
    n_g = cfg["ng"]
    alg_flops = float(iters.sum()) * flops_per_iter(N, n_g)
    # solve: the events bracket exactly one IPM launch; oracle_solve: the IPM kernel's own event time; closed loop: the whole
    # loop (oracle + plant kernels included: a lower bound of the IPM kernel's fraction)
    k_ms = {"solve": float(np.median(step_ms)), "oracle_solve": float(kernel_ms_last), "closed_loop": float(np.median(step_ms))}[kind]
    peak_tf = lbmpc_b200.measure_fp64_peak(local_rank)
    achieved_tf = alg_flops / (k_ms * 1e-3) * 1e-12
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            traffic = json.load(f).get(f"config{args.config}_{kernel_used}", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    alg_bytes = nb * (32 + 8 * N + 24 + (32 * N if kind == "oracle_solve" else 0) + (32 if cfg.get("ref") else 0))
    roofline = {"bound": "fp64_fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "kernel": f"ipm ({kernel_used} mapping)",
                "peak_source": "lbmpc_measure_fp64_peak (register-resident DFMA chains) on this device; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flops_per_launch": alg_flops, "kernel_ms": k_ms,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (k_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                        "note": "algorithmic I/O only; the warp / CTA mappings keep the iterate in shared memory, the stream mapping "
                                "streams it from HBM (profiles/)"}}

    if kernel_used in ("stream", "mixed") and kind != "closed_loop":
        # the stream mapping keeps the iterate in HBM: 178 doubles per stage and iteration by design — overhead, not
        # algorithmic traffic, reported so that the second bound of this mapping is visible next to the FP64 one
        design = float(iters.sum()) * 178 * 8 * (N + 1)
        roofline["hbm"]["stream_design_bytes_per_launch"] = design
        roofline["hbm"]["stream_design_gbs"] = design / (k_ms * 1e-3) * 1e-9
        if peaks.get("hbm_gbs"):
            roofline["hbm"]["stream_design_frac_of_peak"] = roofline["hbm"]["stream_design_gbs"] / peaks["hbm_gbs"]

    # ---------------- CPU baseline (oracle port, bounded sample) ----------------
    from oracle_py import OracleProblem
    threads = os.cpu_count() or 1
    P = OracleProblem(cfg["form"], cfg["variant"], mdl, N)
    nmax = cpu_sample_size(cfg)
    cpu_port_step(cfg, P, inp, threads, 64)
    cpu_n, t0 = 0, time.perf_counter()
    while True:
        cpu_n += cpu_port_step(cfg, P, inp, threads, nmax)
        cpu_dt = time.perf_counter() - t0
        if cpu_dt >= 3.0:
            break
    hist = np.bincount(iters.reshape(-1), minlength=1).tolist()
    st_counts = np.bincount(status.reshape(-1), minlength=4).tolist()
    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32+f64" if kernel_used == "mixed" else "f64", "data": "synthetic",
            "config": config_object(cfg, args.gpus),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "path": e2e_path},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_n / cpu_dt, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_n} QPs of the rank-0 workload in {cpu_dt:.1f} s, {threads} POSIX threads"},
            "p50_us_per_solve": 1e3 * float(np.median(step_ms)) / qp_per_step,
            "iterations": {"mean": float(iters.mean()), "max": int(iters.max()), "hist": hist},
            "status_counts": st_counts,
            "optimal_qp_per_s": value * st_counts[0] / max(1, sum(st_counts)),
            "kernel": kernel_used, "slots_per_cta": sol.slots_per_cta, "kernel_ms_last_launch": kernel_ms_last}
    line.update(e2e_extra)
    if stats is not None:
        line["all_rank_stats"] = stats
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
