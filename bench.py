#!/usr/bin/env python
"""bench.py -- CEM actions/sec of the planner hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cadm_b200|reference] [--config C2] [--m 1]

A "step" is one CEM decision (5 iterations x n candidates x p particles x h horizon steps through the ensemble
dynamics MLP, then top-k + refit) on synthetic start states and random-init weights.  Units per step =
iters * m * n * p * h (candidate x particle x horizon-step dynamics evaluations).

N = 1   workload = BASELINE.json configs[1]: HalfCheetah PE-TS (ens=5, part=20, cand=200, horizon=30), m = 1.
N > 1   one process per GPU (torchrun); candidates are sharded, 200 per GPU (weak scaling: n = 200 N), one NCCL
        all-gather of the per-candidate returns per CEM iteration.

value   device-resident: inputs already in HBM, CUDA-event time of K decisions (L2 flushed between decisions).
e2e     the same decisions through DynamicsModel.get_action() with HOST (NumPy) buffers: H2D of the inputs and
        D2H of the plan inside the timed region (wall clock around a call that synchronises).
--impl reference  times the CPU restatement of the reference's TF1.15 graph (oracle/ref_cpu.py; TensorFlow 1.15 is
        not installable here) on all host cores, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CEM actions/sec (cand x part x horizon steps/s)"
UNIT = "actions/s"
CAND_PER_GPU = 200


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


def ncu_traffic(kernel_name):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture of the same
    workload (profiles/*_ncu_summary.csv, written by tools/ncu_summary.py); None when no capture of that kernel is there."""
    import csv
    import glob
    key = kernel_name.split("(")[0]
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            if key not in vals[hdr.index("Kernel Name")]:
                continue
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(name)
                tot += float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
            return tot, os.path.relpath(path, ROOT)
        except (ValueError, IndexError, OSError):
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


# --------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the reference's own path restated op-for-op in PyTorch-CPU (oracle/ref_cpu.py)."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import cadm_oracle as orc
    from oracle.ref_cpu import RefCpuPlanner, cpu_info
    from cadm_b200.synth import CONFIGS, WORKLOAD_NAMES, synthetic_inputs
    from oracle.envs import get_env
    cfg = dict(CONFIGS[args.config])
    n = CAND_PER_GPU * world if args.config == "C2" else cfg["candidates"]
    env = get_env(cfg["env"])
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    E, p, h, m = cfg["ensemble"], cfg["particles"], cfg["horizon"], args.m
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    C = 10 if cfg["context"] else 0
    prm = orc.init_dynamics_params(rng, E, P + A + C, 200, D)
    prm.b_lv[...] = -6.0
    enc = orc.init_encoder_params(rng, E, (D + A) * 10) if cfg["context"] else None
    K = 10
    norm = orc.NormStats(np.zeros(P), np.ones(P), np.zeros(A), np.full(A, 0.6), np.zeros(D), np.full(D, 0.1),
                         np.zeros(D * K), np.ones(D * K), np.zeros(A * K), np.ones(A * K)).astype(np.float32)

    class _E:  # synthetic_inputs wants name/obs_dim/act_dim
        name, obs_dim, act_dim = cfg["env"], D, A
    inp = synthetic_inputs(_E, m, h, cfg["context"], seed=0)
    pl = RefCpuPlanner(prm, norm, cfg["env"], E, p, cfg["deterministic"], enc, threads=threads)
    # bounded sample: whole decisions while they are cheap, fewer CEM iterations otherwise (units scale with it)
    iters = 5 if n * p * m <= 8000 else max(1, int(5 * 8000 / (n * p * m)))
    units = iters * m * n * p * h
    call = lambda: pl.cem(inp["obs"], inp["init_mean"], inp["init_var"], n, inp.get("cp_obs"), inp.get("cp_act"), iters=iters)
    for _ in range(args.warmup):
        call()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        call()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = units * args.steps / total
    info = cpu_info()
    sample = f"{args.steps} x ({iters} of 5 CEM iterations of one decision, m={m}, n={n})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.config] + f", m={m}, n={n}",
                   "note": "CPU restatement of the TF1.15 graph (TF not installable here), PyTorch-CPU fp32, "
                           f"{info['model']}"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args, model, env, cfg, budget_s=20.0):
    """cpu_baseline object of the default arm: ref_cpu on this box's host cores, bounded sample (rank 0, N = 1)."""
    import numpy as np
    import torch
    from oracle import cadm_oracle as orc
    from oracle.ref_cpu import RefCpuPlanner
    from cadm_b200.synth import synthetic_inputs
    d = model._dyn
    prm = orc.DynamicsParams(list(d["W"]), list(d["b"]), d["W_mu"], d["b_mu"], d["W_lv"], d["b_lv"], d["max_logvar"], d["min_logvar"])
    enc = orc.EncoderParams(list(model._enc["W"]), list(model._enc["b"])) if model._enc is not None else None
    st = [np.asarray(s, np.float32) for s in model.get_normalization_stats()]
    norm = orc.NormStats(*st[:6], *(st[6:10] if len(st) > 6 else [None] * 4))
    threads = os.cpu_count()
    pl = RefCpuPlanner(prm, norm, cfg["env"], cfg["ensemble"], cfg["particles"], cfg["deterministic"], enc, threads=threads)
    m = args.m
    inp = synthetic_inputs(env, m, cfg["horizon"], cfg["context"], seed=0)
    n = cfg["candidates"]
    call = lambda: pl.cem(inp["obs"], inp["init_mean"], inp["init_var"], n, inp.get("cp_obs"), inp.get("cp_act"))
    call()
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 10):
        t0 = time.perf_counter()
        call()
        times.append(time.perf_counter() - t0)
    units = 5 * m * n * cfg["particles"] * cfg["horizon"]
    med = statistics.median(times)
    return {"value": units / med, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"median of {len(times)} whole decisions (1 warm-up), PyTorch-CPU fp32 restatement of the TF1.15 graph",
            "ms_per_decision": 1e3 * med}


# --------------------------------------------------------------------------------------------------------
def run_engine(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cadm_b200.parallel import ShardedCEMPlanner
    from cadm_b200.synth import WORKLOAD_NAMES, build_model, flops_per_unit, synthetic_inputs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_total = CAND_PER_GPU * world if args.config == "C2" else None
    if args.cand:
        n_total = args.cand                       # C5 sweep cell: global candidate count as given (strong scaling over GPUs)
    model, env, cfg = build_model(args.config, m_max=max(args.m, 1), candidates=n_total, particles=args.part or None, rank=rank,
                                  world=world, precision=args.precision, device=dev)
    if n_total is None:
        n_total = cfg["candidates"]
    eng = model.engine
    m, h, p, E = args.m, cfg["horizon"], cfg["particles"], cfg["ensemble"]
    inp = synthetic_inputs(env, m, h, cfg["context"], seed=0)
    dv = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
    planner = ShardedCEMPlanner(eng)
    units = 5 * m * n_total * p * h
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    out_mean = torch.empty_like(dv["init_mean"])
    out_var = torch.empty_like(dv["init_var"])

    def step(i, timed=None):
        seed = (1 << 32) | i
        if timed is not None:
            timed[0].record()
        if world == 1:
            eng.plan_cem_into(dv["obs"], dv["init_mean"], dv["init_var"], out_mean, out_var, dv.get("cp_obs"), dv.get("cp_act"), seed=seed)
        else:
            planner.plan(dv["obs"], dv["init_mean"], dv["init_var"], dv.get("cp_obs"), dv.get("cp_act"), seed=seed, logs=False)
        if timed is not None:
            timed[1].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    eng.set_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(float(i))                      # L2 flush, outside the timed events
        step(args.warmup + i, evs[i])
        kernel_ms.append(eng.last_rollout_ms())    # synchronises; sum over the 5 rollout launches of this decision
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    eng.set_timing(False)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- e2e: the public host API (NumPy in / NumPy out), H2D + D2H inside the timed region
    e2e = None
    if world == 1:
        for i in range(2):
            model.get_action(inp["obs"], *( (inp["cp_obs"], inp["cp_act"]) if cfg["context"] else ()), inp["init_mean"], inp["init_var"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if cfg["context"]:
                model.get_action(inp["obs"], inp["cp_obs"], inp["cp_act"], inp["init_mean"], inp["init_var"])
            else:
                model.get_action(inp["obs"], inp["init_mean"], inp["init_var"])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d = 4 * m * (env.obs_dim + 2 * h * env.act_dim + ((env.obs_dim + env.act_dim) * 10 if cfg["context"] else 0))
        d2h = 4 * m * h * env.act_dim
        e2e = {"value": units * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * e2e_s / args.steps}
        # the same control step with the sampler-side state on the device (PlannerSession): only the observation goes up and
        # the first action comes down; warm start, init_var and the history buffers stay in HBM
        from cadm_b200.samplers import PlannerSession
        sess = PlannerSession(model, m, state_diff=True)
        obs64 = inp["obs"].astype(np.float64)
        for i in range(2):
            a = sess.act(obs64)
            sess.observe(obs64 + 0.01, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            a = sess.act(obs64)
            sess.observe(obs64 + 0.01 * (i + 1), None)
        torch.cuda.synchronize()
        sess_s = time.perf_counter() - t0
        e2e["session"] = {"value": units * args.steps / sess_s, "unit": UNIT, "ms_per_step": 1e3 * sess_s / args.steps,
                          "h2d_bytes_per_step": 4 * m * env.obs_dim * 2, "d2h_bytes_per_step": 4 * m * env.act_dim,
                          "what": "PlannerSession.act + observe: warm start / init_var / history buffers resident on the device"}
    else:
        # multi-rank: same decision through the sharded planner fed from pinned host tensors
        pin = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items()}
        d2 = {k: torch.empty_like(v, device=dev) for k, v in pin.items()}           # staging buffers, allocated once
        out_host = torch.empty((m, h, env.act_dim), dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            for k, v in pin.items():
                d2[k].copy_(v, non_blocking=True)                                    # H2D of this step's inputs
            o = planner.plan(d2["obs"], d2["init_mean"], d2["init_var"], d2.get("cp_obs"), d2.get("cp_act"),
                             seed=(2 << 32) | i, logs=False)
            out_host.copy_(o["mean"], non_blocking=True)                             # D2H of the plan
            torch.cuda.synchronize()
            out_host.clamp_(-1, 1)                                                   # get_action clip (host side, as the reference)
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        h2d = 4 * m * (env.obs_dim + 2 * h * env.act_dim + ((env.obs_dim + env.act_dim) * 10 if cfg["context"] else 0))
        e2e = {"value": units * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4 * m * h * env.act_dim, "ms_per_step": 1e3 * e2e_s / args.steps}
    clocks = sampler.stop()

    if rank != 0:
        return
    peaks = load_peaks()
    In = env.proc_obs_dim + env.act_dim + (10 if cfg["context"] else 0)
    fpu = flops_per_unit(In, 200, env.obs_dim, cfg["deterministic"])
    # dominant kernel: the rollout kernel; one launch = one CEM iteration of this rank's candidates
    units_per_launch = m * (n_total // world) * p * h
    launch_ms = statistics.mean(kernel_ms) / 5.0
    achieved = units_per_launch * fpu / (launch_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic(eng.kernel_name) if (world == 1 and args.config == "C2" and m == 1 and not (args.cand or args.part)) else (None, None)
    line = {
        "metric": METRIC, "value": units * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.cand else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (WORKLOAD_NAMES[args.config] if not (args.cand or args.part) else
                                f"HalfCheetah PE-TS CEM sweep cell (ens={E}, part={p}, cand={n_total}, horizon={h})") + f", m={m}" + (f", n={n_total} sharded {world} x {n_total // world}" if world > 1 else ""),
                   "precision": args.precision, "kernel": eng.kernel_name,
                   "l2": "flushed between steps (256 MiB write outside the timed events); weights (2.7 MB) are L2-resident by design within a step",
                   "parallelism": (f"candidates sharded over {world} GPU(s), 1 all-gather of [m, n/G] returns per CEM iteration, "
                                   + ("fused into the particle-mean kernel over peer memory (NVLink stores + device flags)" if planner.fused else "NCCL"))
                   if world > 1 else "single GPU"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src, "kernel": eng.kernel_name,
                     "launch_ms": launch_ms, "flop_per_unit": fpu, "units_per_launch": units_per_launch,
                     "peak_source": f"{peaks['src']} bf16 dense burst (MEASURED_PEAKS.json)",
                     "kernel_share_of_step": statistics.mean(kernel_ms) / (total_ms / args.steps)},
        "wall_s_timed_loop": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args, model, env, cfg)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cadm_b200", choices=["cadm_b200", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--precision", default=os.environ.get("CADM_PRECISION", "tc3x"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cand", type=int, default=0, help="override the candidate count (C5 sweep cells)")
    ap.add_argument("--part", type=int, default=0, help="override the particle count (C5 sweep cells)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cadm_b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "cadm_b200":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, max(world, args.gpus))
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_engine(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
