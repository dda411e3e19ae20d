#!/usr/bin/env python
"""bench.py -- CEM actions/sec of the planner hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cadm_b200|reference] [--config C2] [--m 1]

A "step" is one CEM decision (5 iterations x n candidates x p particles x h horizon steps through the ensemble
dynamics MLP, then top-k + refit) on synthetic start states and random-init weights.  Units per step =
iters * m * n * p * h (candidate x particle x horizon-step dynamics evaluations).

N = 1   workload = BASELINE.json configs[1]: HalfCheetah PE-TS (ens=5, part=20, cand=200, horizon=30), m = 1.
N > 1   one process per GPU (torchrun); candidates are sharded with one exchange of the per-candidate returns per CEM
        iteration (fused into the refit kernel over peer memory; NCCL all-gather as the fallback).  `--scaling weak`
        (default): 200 candidates per GPU (n = 200 N) is the line's `value`; `--scaling strong`: the NAMED config
        (n = 200) split over the N GPUs.  Either way the line carries the other leg, BASELINE configs[3] (Ant + CaDM,
        n = 1000) split over the N GPUs, and the environment-sharded leg (10 environments per GPU, no exchange) under
        `legs`, so that one driver run per N records all of them.

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


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


def ncu_traffic(kernel_name, config="C2", m=1):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture of the same
    workload (profiles/r<round>_<kernel>_<config>_m<m>_ncu_summary.csv, written by tools/ncu_summary.py, latest round first);
    None when no capture of that kernel on that workload is there."""
    import csv
    import glob
    key = kernel_name.split("(")[0]
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_{config}_m{m}_ncu_summary.csv")), reverse=True):
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
    base_n = args.cand or cfg["candidates"]
    n = base_n if (world == 1 or args.scaling == "strong" or args.cand) else base_n * world      # the engine arm's primary leg
    if args.part:
        cfg["particles"] = args.part
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
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if (world == 1 or args.scaling == "strong" or args.cand) else "weak",
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
def measure_leg(args, rank, world, local_rank, config, m, n_total, particles, sharding, flush, timed_e2e=False, sampler=None):
    """One workload on `world` ranks: build the model, warm up, time args.steps decisions with CUDA events (L2 flushed between
    them, outside the events), max over ranks.  sharding: "candidates" (n_total split over the ranks, one exchange per CEM
    iteration) or "envs" (every rank plans its own m environments with all n_total candidates, no exchange at all)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from cadm_b200.synth import build_model, flops_per_unit, synthetic_inputs

    dev = torch.device("cuda", local_rank)
    by_env = sharding == "envs"
    model, env, cfg = build_model(config, m_max=max(m, 1), candidates=n_total, particles=particles or None, trained_like=not args.random_init_logvar,
                                  rank=0 if by_env else rank, world=1 if by_env else world, precision=args.precision, device=dev)
    eng = model.engine
    if by_env:
        eng.set_option("env_offset", rank * m)            # this rank's block of the m * world environments of the decision
    h, p, E = cfg["horizon"], cfg["particles"], cfg["ensemble"]
    inp = synthetic_inputs(env, m, h, cfg["context"], seed=rank if by_env else 0)
    dv = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
    planner = None if (by_env or world == 1) else model.sharded_planner()
    units = 5 * m * n_total * p * h * (world if by_env else 1)
    out_mean, out_var = torch.empty_like(dv["init_mean"]), torch.empty_like(dv["init_var"])

    def step(i, timed=None):
        seed = (1 << 32) | i
        if timed is not None:
            timed[0].record()
        if planner is None or planner.fused:
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
    # ---- the timed region: K decisions, one CUDA-event pair around each (L2 flushed between them, outside the events)
    launches0 = eng.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(float(i))                      # L2 flush, outside the timed events
        step(args.warmup + i, evs[i])
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- the same K decisions again with a CUDA-event pair around every rollout launch (recorded by the engine on the launching
    # stream): the per-launch duration the roofline uses.  Kept out of the timed region because an event between two kernels
    # serialises them -- the sample -> rollout -> refit chain is launched with programmatic dependent launch, which lets the next
    # kernel's prologue run under the previous kernel (csrc/common.cuh) -- so the first pass is what a user gets, this one is
    # the kernel alone.
    eng.set_timing(True)
    kernel_ms = []
    for i in range(args.steps):
        flush.fill_(float(i))
        step(args.warmup + i)
        kernel_ms.append(eng.last_rollout_ms())    # synchronises; sum over the 5 rollout launches of this decision
    barrier()
    eng.set_timing(False)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- e2e: the public host API (NumPy in / NumPy out), H2D + D2H inside the timed region -- DynamicsModel.get_action(), which
    # at world > 1 drives the candidate-sharded planner (every rank calls it with the same inputs)
    e2e = None
    if timed_e2e and not by_env:
        ga = (lambda: model.get_action(inp["obs"], inp["cp_obs"], inp["cp_act"], inp["init_mean"], inp["init_var"])) if cfg["context"] \
            else (lambda: model.get_action(inp["obs"], inp["init_mean"], inp["init_var"]))
        for i in range(2):
            ga()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            ga()
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        h2d = 4 * m * (env.obs_dim + 2 * h * env.act_dim + ((env.obs_dim + env.act_dim) * 10 if cfg["context"] else 0))
        e2e = {"value": units * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * m * h * env.act_dim,
               "ms_per_step": 1e3 * e2e_s / args.steps, "api": "DynamicsModel.get_action (NumPy in / NumPy out)"}
        if planner is None or planner.fused:
            # the same control step with the sampler-side state on the device (PlannerSession): only the observation goes up and
            # the first action comes down; warm start, init_var and the history buffers stay in HBM
            from cadm_b200.samplers import PlannerSession
            sess = PlannerSession(model, m, state_diff=True)
            obs64 = inp["obs"].astype(np.float64)
            for i in range(2):
                sess.act(obs64)
                sess.observe(obs64 + 0.01, None)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                sess.act(obs64)
                sess.observe(obs64 + 0.01 * (i + 1), None)
            barrier()
            sess_s = time.perf_counter() - t0
            e2e["session"] = {"value": units * args.steps / sess_s, "unit": UNIT, "ms_per_step": 1e3 * sess_s / args.steps,
                              "h2d_bytes_per_step": 4 * m * env.obs_dim * 2, "d2h_bytes_per_step": 4 * m * env.act_dim,
                              "what": "PlannerSession.act + observe: warm start / init_var / history buffers resident on the device"}
    if sampler is not None and sampler.proc is not None and len(sampler.rows) < 5:
        # the timed region of a small workload is over before nvidia-smi has produced a sample: keep the same decisions running
        # under the sampler for a moment so that the clocks line describes the GPU under THIS load
        t_end = time.perf_counter() + 0.8
        i = 0
        while time.perf_counter() < t_end:
            step(1000 + i)
            i += 1
        barrier()

    In = env.proc_obs_dim + env.act_dim + (10 if cfg["context"] else 0)
    fpu = flops_per_unit(In, 200, env.obs_dim, cfg["deterministic"])
    n_rank = n_total if by_env else n_total // world
    units_per_launch = m * n_rank * p * h            # one launch = one CEM iteration of this rank's rows
    launch_ms = statistics.mean(kernel_ms) / 5.0
    res = dict(value=units * args.steps / (total_ms * 1e-3), ms_per_step=total_ms / args.steps, launch_ms=launch_ms,
               kernel=eng.kernel_name, launches=int(launches), fpu=fpu, units_per_launch=units_per_launch,
               achieved=units_per_launch * fpu / (launch_ms * 1e-3) / 1e12,
               kernel_share=statistics.mean(kernel_ms) / (total_ms / args.steps), wall=wall, e2e=e2e,
               fused=bool(planner is not None and planner.fused), E=E, p=p, h=h, n=n_total, m=m)
    eng.close()
    return res


def leg_summary(r, what):
    return {"what": what, "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "rollout_launch_ms": r["launch_ms"],
            "kernel": r["kernel"], "kernel_share_of_step": r["kernel_share"], "achieved_tflops_per_gpu": r["achieved"]}


def run_engine(args, rank, world, local_rank):
    import torch
    from cadm_b200.synth import WORKLOAD_NAMES

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2
    sweep = bool(args.cand or args.part)
    base_n = args.cand or {"C1": 200, "C2": 200, "C3": 200, "C4": 1000}[args.config]
    strong = world == 1 or args.scaling == "strong" or sweep
    n_primary = base_n if strong else base_n * world
    if n_primary % world:
        raise SystemExit(f"{n_primary} candidates do not split over {world} ranks")

    sampler = ClockSampler(local_rank)
    sampler.start()
    prim = measure_leg(args, rank, world, local_rank, args.config, args.m, n_primary, args.part, "candidates", flush,
                       timed_e2e=True, sampler=sampler)
    clocks = sampler.stop()

    # ---- the other legs of a multi-GPU line (BASELINE.json: the NAMED configs at 1 / 2 / 4 / 8 GPUs are strong scaling)
    legs = {}
    if world > 1 and not sweep and not args.no_extra_legs:
        other = "weak" if strong else "strong"
        n_other = base_n * world if strong else base_n
        r = measure_leg(args, rank, world, local_rank, args.config, args.m, n_other, 0, "candidates", flush)
        legs[other] = leg_summary(r, f"{args.config}, n={n_other} candidates over {world} GPUs ({n_other // world} per GPU), m={args.m}")
        if args.config == "C2" and 1000 % world == 0:
            r = measure_leg(args, rank, world, local_rank, "C4", 1, 1000, 0, "candidates", flush)
            legs["strong_C4"] = leg_summary(r, f"BASELINE configs[3]: Ant PE-TS + CaDM, n=1000 candidates over {world} GPUs ({1000 // world} per GPU), m=1")
        r = measure_leg(args, rank, world, local_rank, args.config, 10, base_n, 0, "envs", flush)
        legs["env_sharded"] = leg_summary(r, f"{args.config}, m=10 environments PER GPU ({10 * world} in all), n={base_n}: environments sharded, "
                                             "no exchange during planning (cadm/samplers/sampler.py:107-120 plans 10-20 environments per call)")
    if rank != 0:
        return
    peaks = load_peaks()
    E, p, h = prim["E"], prim["p"], prim["h"]
    traffic, traffic_src = ncu_traffic(prim["kernel"], args.config, args.m) if (world == 1 and not sweep) else (None, None)
    name = WORKLOAD_NAMES[args.config] if not sweep else f"HalfCheetah PE-TS CEM sweep cell (ens={E}, part={p}, cand={n_primary}, horizon={h})"
    line = {
        "metric": METRIC, "value": prim["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": prim["ms_per_step"], "higher_is_better": True,
        # the label follows the requested mode at every N (at N = 1 both modes are the same run), so the driver's 1 -> 8 series is one mode
        "scaling": "weak" if (args.scaling == "weak" and not sweep) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name + f", m={args.m}" + (f", n={n_primary} sharded {world} x {n_primary // world}" if world > 1 else ""),
                   "scaling": ("strong: the named config's candidates split over the GPUs" if strong else
                               f"weak: {base_n} candidates per GPU (n = {n_primary})") if world > 1 else "single GPU",
                   "precision": args.precision, "kernel": prim["kernel"],
                   "weights": "random init of the reference (core/utils.py:636-641), " + ("b_logvar = 0 (noise-dominated rollouts)" if args.random_init_logvar
                                                                                          else "b_logvar = -6 (trained-like: small sampled noise; SURVEY 8d)"),
                   "l2": "flushed between steps (256 MiB write outside the timed events); weights (2.7 MB) are L2-resident by design within a step",
                   "kernel_timing": "roofline.launch_ms: CUDA events around every rollout launch, in a second pass over the same K decisions "
                                    "(an event between kernels serialises the programmatic-dependent-launch chain the timed pass runs with)",
                   "parallelism": (f"candidates sharded over {world} GPU(s), 1 all-gather of [m, n/G] returns per CEM iteration, "
                                   + ("fused into the refit kernel over peer memory (NVLink stores + device flags)" if prim["fused"] else "NCCL"))
                   if world > 1 else "single GPU"},
        "e2e": prim["e2e"],
        "gpu_launches": prim["launches"],
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": prim["achieved"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": prim["achieved"] / peaks["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src, "kernel": prim["kernel"],
                     "launch_ms": prim["launch_ms"], "flop_per_unit": prim["fpu"], "units_per_launch": prim["units_per_launch"],
                     "peak_source": f"{peaks['src']} bf16 dense burst (MEASURED_PEAKS.json)",
                     "kernel_share_of_step": prim["kernel_share"]},
        "wall_s_timed_loop": prim["wall"],
    }
    if legs:
        line["legs"] = legs
    if world == 1 and not args.no_cpu_baseline:
        from cadm_b200.synth import build_model
        model, env, cfg = build_model(args.config, m_max=max(args.m, 1), candidates=n_primary, particles=args.part or None,
                                      precision=args.precision, device=dev)
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: which leg is the line's `value` -- weak = the config's candidate count PER GPU, strong = the named "
                         "config split over the GPUs; the other leg, C4 strong and the environment-sharded leg go under `legs`")
    ap.add_argument("--no-extra-legs", action="store_true", help="N > 1: only the primary leg")
    ap.add_argument("--random-init-logvar", action="store_true",
                    help="keep the reference's random-init log-variance head (b_logvar = 0) instead of the trained-like b_logvar = -6 (SURVEY 8d: report both)")
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
