#!/usr/bin/env python
"""bench.py — snowpack-frequency DORT solves / second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU restatement of the reference's
                                                              # NumPy/SciPy DORT (oracle/) on the box's host cores

Default workload (BASELINE.json configs[1], SURVEY.md §8(d) cfg 2): IBA(exponential) + DORT, 20 layers, 32 streams, the
6 AMSR-E frequencies, theta = 55 deg, passive; S = 10 000 synthetic snowpacks per GPU (seed 2) -> 60 000 (snowpack x
frequency) solves per step and per GPU.  One "step" = one pass of the whole hot path over that batch.
``--workload cfg3|cfg4|cfg5`` runs the other BASELINE configs through exactly the same code (same JSON line, same
multi-GPU launch); ``--scaling strong`` splits ONE ensemble of ``--snowpacks`` members over the ranks.

  value  solves/s, inputs already resident in HBM (device pointers through smrtb200_solve_batch_device)
  e2e    solves/s through smrtb200_solve_batch_host: HOST buffers in and out, H2D/D2H inside the timed region
  roofline   dominant kernel vs the FP64 FMA peak measured on the box (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline   the CPU oracle on a bounded sample of the same workload, all host cores, numerical threads = 1,
                 with the parity of the GPU results over EVERY problem of that sample

With N > 1 (torchrun, one process per GPU) every rank solves its own members; the only communication is the final NCCL
all_gather of the results, inside the timed step.
"""
import os

# numerical-library threads are pinned BEFORE numpy / scipy are imported anywhere in this process tree: the CPU arm runs
# one process per core (how the reference itself parallelises: joblib processes with BLAS threads = 1,
# smrt/core/lib.py:655-666), and OpenBLAS reads these variables once, when the library is loaded
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ[_v] = "1"

import argparse  # noqa: E402
import json  # noqa: E402
import platform  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AMSRE = (6.925e9, 10.65e9, 18.7e9, 23.8e9, 36.5e9, 89e9)
CFG4_FREQS = (1.4135e9, 5.4e9, 6.925e9, 7.3e9, 9.6e9, 10.65e9, 13.5e9, 18.7e9, 23.8e9, 31.4e9, 36.5e9, 89.0e9)
UNIT = "solves/s"

# SURVEY.md §8(d) configurations.  snowpacks = members per GPU of the default (weak-scaling) run; full = the size
# BASELINE.json states for the configuration (whole job).
WORKLOADS = {
    "cfg2": dict(metric="snowpack-frequency DORT solves/sec (20-layer, 32-stream)",
                 desc="IBA(exponential)+DORT passive, 20 layers, 32 streams, 6 AMSR-E frequencies, theta=55deg",
                 layers=20, streams=32, freqs=AMSRE, seed=2, mode="P", m_max=0, snowpacks=10_000, full=10_000),
    "cfg3": dict(metric="snowpack-frequency DORT solves/sec (10-layer, 16-stream, active)",
                 desc="DMRT-QCA-shortrange(SHS)+DORT active, 10 layers, 16 streams, C/X/Ku backscatter at 40deg, m_max=2",
                 layers=10, streams=16, freqs=(5.4e9, 9.6e9, 13.5e9), seed=3, mode="A", m_max=2, snowpacks=50_000,
                 full=50_000),
    # the reference's DEFAULT radar configuration (DORT n_max_stream = 32, m_max = 2: blocks of 96 unknowns) on the cfg-3
    # snowpacks: not a BASELINE config, kept as a line of record for the 64 < h <= 128 kernels
    "active32": dict(metric="snowpack-frequency DORT solves/sec (10-layer, 32-stream, active)",
                     desc="DMRT-QCA-shortrange(SHS)+DORT active, 10 layers, 32 streams (reference default), C/X/Ku "
                          "backscatter at 40deg, m_max=2",
                     layers=10, streams=32, freqs=(5.4e9, 9.6e9, 13.5e9), seed=3, mode="A", m_max=2, snowpacks=5_000,
                     full=50_000),
    "cfg4": dict(metric="snowpack-frequency DORT solves/sec (50-layer, 64-stream)",
                 desc="IBA(exponential)+DORT passive, 50 layers, 64 streams, 12 frequencies, theta=55deg",
                 layers=50, streams=64, freqs=CFG4_FREQS, seed=4, mode="P", m_max=0, snowpacks=500, full=200_000),
    "cfg5": dict(metric="snowpack-frequency DORT solves/sec (30-layer sea ice, 32-stream)",
                 desc="IBA+DORT passive, 30-layer multi-year sea ice over ocean, 32 streams, 1.4 GHz, theta=40deg",
                 layers=30, streams=32, freqs=(1.4e9,), seed=5, mode="P", m_max=0, snowpacks=125_000, full=1_000_000),
}


# ---------------------------------------------------------------------------------------------------------------------
# synthetic ensembles (SURVEY.md §8(d) generators: one default_rng(seed), members drawn sequentially)
# ---------------------------------------------------------------------------------------------------------------------
def snow_members(S, seed, L, kind):
    rng = np.random.default_rng(seed)
    th = np.empty((S, L)); rho = np.empty((S, L)); T = np.empty((S, L)); p0 = np.empty((S, L))
    for s in range(S):
        th[s] = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
        if kind == "shs":
            rho[s] = rng.uniform(200, 400, L); T[s] = rng.uniform(240, 270, L); p0[s] = rng.uniform(1e-4, 3e-4, L)
        else:
            rho[s] = rng.uniform(150, 450, L); T[s] = rng.uniform(240, 272, L); p0[s] = rng.uniform(5e-5, 3e-4, L)
    return th, rho, T, p0


def sea_ice_members(S, seed=5, L=30):
    """cfg-5 generator: the arrays make_ice_column("multiyear", ...) would receive, member by member."""
    rng = np.random.default_rng(seed)
    th = np.empty((S, L)); T = np.empty((S, L)); sal = np.empty((S, L)); por = np.empty((S, 1)); pc = np.empty((S, 1))
    for s in range(S):
        H = rng.uniform(1, 3); dT = rng.uniform(10, 30); ss = rng.uniform(0.5, 1.5)
        por[s] = rng.uniform(0.02, 0.12); pc[s] = rng.uniform(0.5e-3, 1.5e-3)
        th[s] = H / L; T[s] = np.linspace(273.15 - dT, 273.15 - 1.8, L); sal[s] = np.linspace(2, 10, L) * 1e-3 * ss
    return th, T, sal, por, pc


def make_batch(workload, S, seed=None, lo=0, hi=None):
    """ProblemBatch of members [lo, hi) of the S-member ensemble of `workload` (frequency outermost)."""
    from smrt_b200.pack import pack_sea_ice_ensemble, pack_snow_ensemble

    w = WORKLOADS[workload]
    seed = w["seed"] if seed is None else seed
    hi = S if hi is None else hi
    sl = slice(lo, hi)
    if workload == "cfg5":
        th, T, sal, por, pc = sea_ice_members(S, seed, w["layers"])
        return pack_sea_ice_ensemble(w["freqs"][0], th[sl], T[sl], sal[sl], por[sl], pc[sl], theta_deg=40.0)
    if workload in ("cfg3", "active32"):
        th, rho, T, a = snow_members(S, seed, w["layers"], "shs")
        return pack_snow_ensemble(w["freqs"], th[sl], rho[sl], T[sl], microstructure="sticky_hard_spheres", radius=a[sl],
                                  stickiness=0.2, emmodel="dmrt_qca_shortrange", mode="A", theta_deg=40.0,
                                  theta_inc_deg=40.0, phi_deg=180.0)
    th, rho, T, pc = snow_members(S, seed, w["layers"], "exp")
    return pack_snow_ensemble(w["freqs"], th[sl], rho[sl], T[sl], corr_length=pc[sl], theta_deg=55.0)


def solver_options(workload):
    w = WORKLOADS[workload]
    return dict(n_max_stream=w["streams"], m_max=w["m_max"]) if w["mode"] == "A" else dict(n_max_stream=w["streams"])


def workload_config(workload, S, world=1, scaling="weak", note=None):
    w = WORKLOADS[workload]
    F = len(w["freqs"])
    cfg = {"workload": f"{w['desc']}, {S} synthetic snowpacks per GPU (SURVEY 8d {workload}, seed {w['seed']}) = "
                       f"{F * S} solves/step/GPU",
           "name": workload, "layers": w["layers"], "n_max_stream": w["streams"],
           "frequencies_ghz": [f / 1e9 for f in w["freqs"]], "snowpacks_per_gpu": S, "solves_per_step_per_gpu": F * S,
           "baseline_size_snowpacks": w["full"], "scaling": scaling,
           "l2_policy": "256 MiB buffer written between timed steps (L2 flush); the per-step layer workspace "
                        "(eigenvectors, GBs) is far larger than L2 anyway"}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic flops: upper bound with n = n_max_stream in every layer (the figure of record) and the
# same formula with the streams every layer actually keeps
# ---------------------------------------------------------------------------------------------------------------------
def _mode_sizes(mode, m_max):
    """(npol of the eigenproblems, npol of the boundary solves incl. the coherent pass of the active mode)"""
    if mode == "P":
        return [2], [2]
    eig = [2] + [3] * m_max
    return eig, [2] + eig


def f_alg_upper(workload):
    w = WORKLOADS[workload]
    eig_np, bnd_np = _mode_sizes(w["mode"], w["m_max"])
    L, n = w["layers"], w["streams"]
    eigen = sum(L * 31.0 * (p * n) ** 3 for p in eig_np)  # 4 h^3 (product) + 25 h^3 (eigenpairs) + 2 h^3 (E-)
    boundary = sum(L * (2.0 / 3.0 + 4.0 + 2.0) * (2 * p * n) ** 3 for p in bnd_np)  # LU + two N-rhs solves + GEMM
    return eigen, boundary


def f_alg_actual(workload, eps_eff, nlayer, max_problems=20000):
    """Mean over (a sample of) the problems of the same formulas with h_l = npol * n_l, n_l = the streams layer l keeps
    (streams.py:182-194: relsin = Re sqrt(eps* / eps_l) sqrt(1 - mu*^2) < 1)."""
    w = WORKLOADS[workload]
    n = w["streams"]
    eps = np.asarray(eps_eff)[:max_problems]
    nl = np.asarray(nlayer)[:max_problems]
    mu = np.sort(np.polynomial.legendre.leggauss(2 * n)[0])[::-1][:n]
    sin_star = np.sqrt(1.0 - mu ** 2)
    Bn, L = eps.shape
    valid = np.arange(L)[None, :] < nl[:, None]
    key_re = np.where(valid, eps.real, -np.inf)
    # complex argmax = lexicographic on (Re, Im) (streams.py:155)
    mx = key_re.max(axis=1, keepdims=True)
    key_im = np.where(valid & (key_re == mx), eps.imag, -np.inf)
    kstar = key_im.argmax(axis=1)
    eps_star = eps[np.arange(Bn), kstar]
    rindex = np.sqrt(eps_star[:, None] / np.where(valid, eps, 1.0)).real
    n_l = (rindex[:, :, None] * sin_star[None, None, :] < 1.0).sum(axis=2) * valid
    eig_np, bnd_np = _mode_sizes(w["mode"], w["m_max"])
    h3 = (n_l.astype(float) ** 3).sum(axis=1).mean()
    eigen = sum(31.0 * p ** 3 for p in eig_np) * h3
    boundary = sum((2.0 / 3.0 + 6.0) * (2 * p) ** 3 for p in bnd_np) * h3
    return eigen, boundary, float(n_l[valid].mean())


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.idx = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's NumPy/SciPy DORT) on every host core, one process per core
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return platform.processor() or "unknown"


def _oracle_init():
    """Pool initialiser (spawned process: numpy / scipy are imported HERE, after the environment was pinned)."""
    global _O, _TP_INFO
    from threadpoolctl import threadpool_info, threadpool_limits

    from oracle import dort_oracle

    _O = dort_oracle
    threadpool_limits(1)
    _TP_INFO = max([int(d.get("num_threads", 1)) for d in threadpool_info()] or [1])


def _oracle_solve(task):
    idx, problem = task
    t0 = time.perf_counter()
    try:
        values = np.asarray(_O.solve_problem(problem)["values"], dtype=float)
    except Exception:  # the oracle raises where the reference raises: counted, never compared
        values = None
    return idx, values, time.perf_counter() - t0, _TP_INFO


class CpuArm:
    """Process pool of oracle workers.  ``run(problems)`` returns (wall seconds, values per problem, busy seconds)."""

    def __init__(self, cores=None):
        import multiprocessing as mp

        self.cores = max(1, cores or os.cpu_count() or 1)
        # spawn: the workers import numpy / scipy themselves with the pinned environment (a forked worker would inherit
        # the parent's already initialised, multi-threaded BLAS pool)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_oracle_init)
        self.blas_threads = None

    def run(self, problems):
        tasks = list(enumerate(problems))
        t0 = time.perf_counter()
        res = self.pool.map(_oracle_solve, tasks, chunksize=1)
        wall = time.perf_counter() - t0
        values = [None] * len(tasks)
        busy = 0.0
        for idx, v, dt, tp in res:
            values[idx] = v
            busy += dt
            self.blas_threads = tp if self.blas_threads is None else max(self.blas_threads, tp)
        return wall, values, busy

    def close(self):
        self.pool.close()
        self.pool.join()


def sample_problems(workload, n_snow):
    """The first n_snow members of the workload's ensemble (identical to members 0..n_snow-1 of the GPU run: the
    generators draw member by member) as oracle problems, frequency outermost."""
    batch = make_batch(workload, n_snow)
    opts = solver_options(workload)
    return batch, [batch.to_problem(i, opts) for i in range(batch.B)]


def cpu_sample_size(workload, cores, budget_s):
    """Snowpacks of the bounded CPU sample: >= 8 solves per core, <= 200 snowpacks, sized for
    `budget_s` seconds at the per-core rates of BASELINE.md §2."""
    per_core = {"cfg2": 4.6, "cfg3": 7.1, "cfg4": 0.52, "cfg5": 4.8, "active32": 2.0}[workload]
    F = len(WORKLOADS[workload]["freqs"])
    by_budget = int(per_core * cores * budget_s / F)
    floor = -(-8 * cores // F)
    if workload == "cfg4":
        floor = -(-2 * cores // F)  # 2 s per solve: two solves per core keep the default run within minutes
    return int(max(2, min(200, max(by_budget, floor))))


def max_rel_err(gpu_values, cpu_values, S_gpu, n_snow, F):
    """max |x_gpu - x_cpu| / |x_cpu| over every problem of the CPU sample (problem (f, s) sits at f * S + s)."""
    worst, n = 0.0, 0
    for f in range(F):
        for s in range(n_snow):
            ref = cpu_values[f * n_snow + s]
            if ref is None:
                continue
            got = gpu_values[f * S_gpu + s].reshape(ref.shape)
            m = np.abs(ref) > 0
            if not np.all(np.isfinite(got)):
                return float("inf"), n
            worst = max(worst, float(np.max(np.abs(got[m] - ref[m]) / np.abs(ref[m]))))
            n += 1
    return worst, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    w = WORKLOADS[workload]
    F = len(w["freqs"])
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    warm = max(0, min(args.warmup, 1))
    n_snow = cpu_sample_size(workload, cores, budget_s=150.0 / (steps + warm))
    _, probs = sample_problems(workload, n_snow)
    arm = CpuArm(cores)
    arm.run(probs[:cores])  # imports, LAPACK initialisation
    for _ in range(warm):
        arm.run(probs)
    walls, busy = [], []
    for _ in range(steps):
        wall, _, b = arm.run(probs)
        walls.append(wall); busy.append(b)
    arm.close()
    value = len(probs) * steps / float(np.sum(walls))
    per_core = len(probs) * steps / float(np.sum(busy))
    sample = (f"{len(probs)} solves per step = the first {n_snow} snowpacks of the {workload} ensemble x {F} "
              f"frequencies (seed {w['seed']})")
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(walls) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(workload, n_snow, note="reference arm: CPU oracle port of the reference's NumPy/SciPy "
                                  "DORT (LAPACK dgees/dgeev + dgbsv, pocketfft), one spawned process per host core, "
                                  "numerical-library threads pinned to 1 before import and checked in the workers"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample,
                         "per_core": per_core, "blas_threads_per_process": arm.blas_threads, "cpu_model": _cpu_model(),
                         "load_balance": float(np.sum(busy) / (arm.cores * np.sum(walls)))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from smrt_b200 import capi
    from smrt_b200.device import DeviceBatch

    workload = args.workload
    w = WORKLOADS[workload]
    F = len(w["freqs"])
    n_streams = w["streams"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # keep stdout for the JSON line: NCCL's version / debug banner goes to stderr unless the caller chose a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S_arg = args.snowpacks if args.snowpacks else w["snowpacks"]
    if args.scaling == "strong":
        # ONE ensemble of S_arg members (the workload's own seed), contiguous shards
        from smrt_b200.dist import shard_bounds

        lo, hi = shard_bounds(S_arg, world, rank)
        batch = make_batch(workload, S_arg, lo=lo, hi=hi)
        S = hi - lo
        S_max = shard_bounds(S_arg, world, 0)[1]
        total_solves = F * S_arg
    else:
        # every rank its own members (weak scaling); rank 0 holds the workload's ensemble
        S = S_max = S_arg
        batch = make_batch(workload, S, seed=w["seed"] + 1000 * rank)
        total_solves = F * S * world
    sopt = solver_options(workload)
    opts = capi.make_options(batch, device=local_rank, **sopt)
    plan = capi.Plan(opts)
    dev = DeviceBatch(batch, n_streams, device=local_rank)
    bt = dev.struct()
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev.device)
    gathered = send = None
    if world > 1:
        # ragged shards (strong scaling) are padded to the largest one
        send = torch.zeros((F * S_max,) + tuple(dev.values.shape[1:]), dtype=torch.float64, device=dev.device)
        gathered = [torch.empty_like(send) for _ in range(world)]

    def step():
        plan.solve_device(bt, stream)
        if world > 1:
            send[:dev.values.shape[0]].copy_(dev.values)
            dist.all_gather(gathered, send)  # the only collective: final gather of the results

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = plan.launch_count
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    step_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (outside the event bracket)
        ev0.record()
        step()
        ev1.record()
        ev1.synchronize()
        step_ms.append(ev0.elapsed_time(ev1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = plan.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = total_solves * args.steps / (total_ms * 1e-3)

    # correctness guard: a bench number from wrong results is worthless
    status = dev.status.cpu().numpy()
    n_err = int(np.count_nonzero(status & capi.ST_ERR_MASK))
    tb = dev.values.cpu().numpy()
    eps_eff = dev.eps_eff.cpu().numpy().view(np.complex128)[..., 0]

    # end-to-end through the host-buffer entry point (H2D + kernels + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 3))
    plan.solve_host(batch)  # warm-up: allocates the pinned ring
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = plan.solve_host(batch)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = total_solves * e2e_steps / e2e_dt
    d2h_bytes = sum(a.nbytes for a in (out.values, out.ks, out.ka, out.eps_eff, out.n_streams, out.stream_angles,
                                       out.optical_depth, out.status))
    e2e_match = bool(np.array_equal(out.values, tb, equal_nan=True))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline -------------------------------------------------------------------------------------------------------
    # per-kernel durations are measured live with CUDA events in a separate pass with the chunks serialised on one
    # stream: in the throughput pass above the two slots overlap and a kernel's event interval includes queueing
    peak_tflops = capi.measure_fp64_peak(local_rank, 300.0)
    fe, fb = f_alg_upper(workload)
    fe_act, fb_act, mean_streams = f_alg_actual(workload, eps_eff, batch.nlayer)
    plan.close()
    splan = capi.Plan(capi.make_options(batch, device=local_rank, serialize=True, **sopt))
    eig_ms, bnd_ms, chunks = [], [], 1
    for it in range(3 if batch.B <= 100_000 else 2):
        flush.fill_(1)
        splan.solve_device(bt, stream)
        splan.sync_timing(stream)
        tm = splan.last_timing()
        if it > 0:
            eig_ms.append(tm["eigen_ms"]); bnd_ms.append(tm["boundary_ms"])
        chunks = tm["chunks"]
    workspace_gb = splan.workspace_bytes / 1e9
    splan.close()
    n_launch = chunks * len(eig_ms)
    eig_avg_ms = float(np.sum(eig_ms)) / max(n_launch, 1)
    bnd_avg_ms = float(np.sum(bnd_ms)) / max(n_launch, 1)
    solves_per_launch = batch.B / max(chunks, 1)
    dominant = "eigen_kernel" if np.sum(eig_ms) >= np.sum(bnd_ms) else "boundary_kernel"
    is_e = dominant == "eigen_kernel"
    t_dom = eig_avg_ms if is_e else bnd_avg_ms

    def tf(flops):
        return flops * solves_per_launch / (t_dom * 1e-3) / 1e12 if t_dom > 0 else 0.0

    achieved_upper, achieved_actual = tf(fe if is_e else fb), tf(fe_act if is_e else fb_act)
    # the figure of record is the SURVEY formula with n = n_max_stream; when the layers keep far fewer streams it
    # overcounts (fractions above 1 are possible) and the same formula with the actual streams is the honest one
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    # capture (profiles/traffic.json, written by tools/make_profile_summary.py), scaled to this run's solves per launch
    traffic = None
    if workload == "cfg2":
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dominant)
            traffic = float(tj["dram_bytes_per_solve"]) * solves_per_launch
        except Exception:
            pass
    per_gpu = value / world
    # `frac` is the figure of record of SURVEY 8(d) / BASELINE.md (F_alg with n = n_max_stream in every layer, the
    # definition the round-1 line used and ncu's executed-flop count confirms for cfg 2) unless that count exceeds the
    # peak, i.e. visibly overcounts (cfg 5 keeps 5-9 of 32 streams per layer): then the actual-stream figure is reported
    frac_record = achieved_upper / peak_tflops if peak_tflops else None
    frac_actual = achieved_actual / peak_tflops if peak_tflops else None
    use_actual = frac_record is not None and frac_record > 1.0
    roofline = {
        "bound": "fp64", "kernel": dominant, "achieved": achieved_actual if use_actual else achieved_upper,
        "peak": peak_tflops, "unit": "TFLOP/s", "frac": frac_actual if use_actual else frac_record, "traffic": traffic,
        "frac_basis": ("actual streams per layer (the n_max_stream count exceeds the peak: F_alg overcounts)"
                       if use_actual else "SURVEY 8(d) F_alg with n = n_max_stream in every layer (figure of record)"),
        "flops_model": "SURVEY 8(d): eigen 31 h^3, boundary (2/3 + 6) (2h)^3 per layer and mode, h = npol n; "
                       "`n_max_stream`: n = n_max_stream in every layer; `actual_streams`: n = the streams each layer "
                       "keeps (Jacobi executes more than 25 h^3 for the eigenpairs, so the executed flops of the eigen "
                       "kernel sit near the n_max_stream count: profiles/*_ncu_summary.txt)",
        "n_max_stream": {"achieved": achieved_upper, "frac": frac_record},
        "actual_streams": {"achieved": achieved_actual, "frac": frac_actual},
        "mean_streams_per_layer": mean_streams,
        "peak_source": "measured in this run: DFMA micro-kernel (smrtb200_measure_fp64_peak); MEASURED_PEAKS.json has "
                       "no FP64 entry",
        "algorithmic_flops_per_solve": {"eigen_kernel": fe_act, "boundary_kernel": fb_act,
                                        "upper_bound": {"eigen_kernel": fe, "boundary_kernel": fb}},
        "avg_launch_ms": {"eigen_kernel": eig_avg_ms, "boundary_kernel": bnd_avg_ms},
        "solves_per_launch": solves_per_launch,
        "note": "the boundary kernel eliminates with h x h blocks and executes ~5x fewer flops than the N x N count of "
                "SURVEY 8(d) charges it: only the dominant-kernel figure is a utilisation",
        "hbm": {"peak_gbs": peaks.get("hbm_gbs"),
                "note": "the path is FP64 / shared-memory bound, not HBM bound (profiles/traffic.json)"},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        n_snow = min(cpu_sample_size(workload, cores, budget_s=20.0), S)
        _, probs = sample_problems(workload, n_snow)
        arm = CpuArm(cores)
        arm.run(probs[:cores])
        wall, cpu_vals, busy = arm.run(probs)
        arm.close()
        err, n_cmp = max_rel_err(tb, cpu_vals, S, n_snow, F)
        cpu = {"value": len(probs) / wall, "unit": UNIT, "cores": arm.cores, "kind": "port",
               "sample": f"{len(probs)} solves (the first {n_snow} snowpacks x {F} frequencies of the same ensemble) in "
                         f"{wall:.1f} s; CPU oracle = restatement of the reference's NumPy/SciPy DORT, one spawned "
                         "process per core, numerical-library threads = 1 (checked in the workers)",
               "per_core": len(probs) / busy, "blas_threads_per_process": arm.blas_threads, "cpu_model": _cpu_model(),
               "load_balance": busy / (arm.cores * wall),
               # parity of the timed GPU results over EVERY problem of the sample
               "max_rel_err": err, "n_compared": n_cmp, "tolerance": 1e-6}

    line = {
        "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(workload, S, world, args.scaling),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(dev.h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "matches_device_path": e2e_match,
                "api": "smrtb200_solve_batch_host (ctypes, host buffers, pinned ring: H2D / kernels / D2H overlapped)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "errors": n_err, "wall_s": t_wall, "workspace_gb": workspace_gb, "per_gpu": per_gpu,
    }
    if args.api and world == 1:
        line["e2e_api"] = run_api(args, e2e_value)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_api(args, e2e_value):
    """Wall time of the public Python API on the contract workload: make_model("iba", "dort").run(amsre, snowpacks),
    Snowpack objects in, labelled Result out (packing, H2D, kernels, D2H, result assembly all inside)."""
    import smrt_b200
    from smrt_b200 import inputs

    w = WORKLOADS["cfg2"]
    S = args.snowpacks if args.snowpacks else w["snowpacks"]
    th, rho, T, pc = snow_members(S, w["seed"], w["layers"], "exp")
    t0 = time.perf_counter()
    snowpacks = [inputs.make_snowpack(th[s], "exponential", density=rho[s], temperature=T[s], corr_length=pc[s])
                 for s in range(S)]
    t_build = time.perf_counter() - t0
    sensor = inputs.amsre()
    model = smrt_b200.make_model("iba", "dort", rtsolver_options=dict(n_max_stream=w["streams"]))
    model.run(sensor, snowpacks[:64])  # warm-up (plan creation)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        res = model.run(sensor, snowpacks)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n = len(w["freqs"]) * S
    return {"value": n / best, "unit": UNIT, "wall_s": best, "solves": n, "snowpack_objects_build_s": t_build,
            "ratio_to_c_abi_e2e": (n / best) / e2e_value if e2e_value else None,
            "dims": list(res.data.dims), "api": "smrt_b200.make_model('iba','dort').run(amsre(), [Snowpack]*S)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--snowpacks", type=int, default=0,
                    help="synthetic snowpacks per GPU (weak scaling) or in total (--scaling strong); 0 = the workload's "
                         "default")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--api", action="store_true", help="also time the public Python API (make_model().run()) on cfg 2")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = the contract workload; cfg3 / cfg4 / cfg5: the other BASELINE configs")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner from C, and with 4 / 8
    # ranks it ignores NCCL_DEBUG_FILE) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the
    # whole run and Python's sys.stdout keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
