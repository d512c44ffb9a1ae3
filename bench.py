#!/usr/bin/env python
"""bench.py — snowpack-frequency DORT solves / second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU restatement of the reference's
                                                              # NumPy/SciPy DORT (oracle/) on the box's host cores

Workload (BASELINE.json configs[1], SURVEY.md §8(d) cfg 2): IBA(exponential) + DORT, 20 layers, 32 streams, the 6
AMSR-E frequencies, theta = 55 deg, passive; S = 10 000 synthetic snowpacks per GPU (seed 2) -> 60 000 (snowpack x
frequency) solves per step and per GPU.  One "step" = one pass of the whole hot path over that batch.

  value  solves/s, inputs already resident in HBM (device pointers through smrtb200_solve_batch_device)
  e2e    solves/s through smrtb200_solve_batch_host: HOST buffers in and out, H2D/D2H inside the timed region
  roofline   dominant kernel vs the FP64 FMA peak measured on the box (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline   the CPU oracle on a bounded sample of the same workload, all host cores

With N > 1 (torchrun, one process per GPU) every rank solves its own 10 000 snowpacks (weak scaling); the only
communication is the final NCCL all_gather of the brightness temperatures, inside the timed step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AMSRE = (6.925e9, 10.65e9, 18.7e9, 23.8e9, 36.5e9, 89e9)
N_LAYERS = 20
N_STREAMS = 32
SNOWPACKS_PER_GPU = 10_000
METRIC = "snowpack-frequency DORT solves/sec (20-layer, 32-stream)"
UNIT = "solves/s"


def synthetic_members(S, seed, L=N_LAYERS):
    """SURVEY.md §8(d) cfg 2 generator (one default_rng(seed), members drawn sequentially)."""
    rng = np.random.default_rng(seed)
    th = np.empty((S, L)); rho = np.empty((S, L)); T = np.empty((S, L)); pc = np.empty((S, L))
    for s in range(S):
        th[s] = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
        rho[s] = rng.uniform(150, 450, L)
        T[s] = rng.uniform(240, 272, L)
        pc[s] = rng.uniform(5e-5, 3e-4, L)
    return th, rho, T, pc


def make_batch(S, seed):
    from smrt_b200.pack import pack_snow_ensemble

    th, rho, T, pc = synthetic_members(S, seed)
    return pack_snow_ensemble(AMSRE, th, rho, T, corr_length=pc, theta_deg=55.0)


# Other BASELINE configs (parity-test cases; timed with --workload for the record, never the contract line) -----------
CFG4_FREQS = (1.4135e9, 5.4e9, 6.925e9, 7.3e9, 9.6e9, 10.65e9, 13.5e9, 18.7e9, 23.8e9, 31.4e9, 36.5e9, 89.0e9)
EXTRA_WORKLOADS = {
    # name: (description, layers, streams, default snowpacks, algorithmic GFLOP per solve from SURVEY.md 8(d))
    "cfg3": ("DMRT-QCA-SR + DORT active, 10 layers, 16 streams, C/X/Ku backscatter at 40 deg, m_max = 2 (seed 3)", 10, 16, 20000, 0.23),
    "cfg4": ("IBA(exponential) + DORT passive, 50 layers, 64 streams, 12 frequencies (seed 4)", 50, 64, 250, 8.84),
    "cfg5": ("IBA + DORT passive, 30-layer multi-year sea ice over ocean, 32 streams, 1.4 GHz at 40 deg (seed 5)", 30, 32,
             60000, 0.663),
}


def sea_ice_members(S, seed=5, L=30):
    """SURVEY.md 8(d) cfg-5 generator: the arrays make_ice_column("multiyear", ...) would receive, member by member."""
    rng = np.random.default_rng(seed)
    th = np.empty((S, L)); T = np.empty((S, L)); sal = np.empty((S, L)); por = np.empty((S, 1)); pc = np.empty((S, 1))
    for s in range(S):
        H = rng.uniform(1, 3); dT = rng.uniform(10, 30); ss = rng.uniform(0.5, 1.5)
        por[s] = rng.uniform(0.02, 0.12); pc[s] = rng.uniform(0.5e-3, 1.5e-3)
        th[s] = H / L; T[s] = np.linspace(273.15 - dT, 273.15 - 1.8, L); sal[s] = np.linspace(2, 10, L) * 1e-3 * ss
    return th, T, sal, por, pc


def make_extra_batch(name, S):
    from smrt_b200.pack import pack_snow_ensemble

    if name == "cfg5":
        from smrt_b200.pack import pack_sea_ice_ensemble

        th, T, sal, por, pc = sea_ice_members(S)
        return pack_sea_ice_ensemble(1.4e9, th, T, sal, por, pc, theta_deg=40.0)
    rng = np.random.default_rng({"cfg3": 3, "cfg4": 4}[name])
    L = EXTRA_WORKLOADS[name][1]
    th = np.empty((S, L)); rho = np.empty((S, L)); T = np.empty((S, L)); p0 = np.empty((S, L))
    for s in range(S):  # SURVEY.md 8(d) generators, members drawn sequentially
        th[s] = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
        if name == "cfg3":
            rho[s] = rng.uniform(200, 400, L); T[s] = rng.uniform(240, 270, L); p0[s] = rng.uniform(1e-4, 3e-4, L)
        else:
            rho[s] = rng.uniform(150, 450, L); T[s] = rng.uniform(240, 272, L); p0[s] = rng.uniform(5e-5, 3e-4, L)
    if name == "cfg3":
        return pack_snow_ensemble((5.4e9, 9.6e9, 13.5e9), th, rho, T, microstructure="sticky_hard_spheres", radius=p0,
                                  stickiness=0.2, emmodel="dmrt_qca_shortrange", mode="A", theta_deg=40.0,
                                  theta_inc_deg=40.0, phi_deg=180.0)
    return pack_snow_ensemble(CFG4_FREQS, th, rho, T, corr_length=p0, theta_deg=55.0)


def run_extra(args):
    """Device-resident throughput of another BASELINE config on one GPU (for DESIGN.md; not the bench contract)."""
    import torch

    from smrt_b200 import capi
    from smrt_b200.device import DeviceBatch

    desc, L, n, S_default, gflop = EXTRA_WORKLOADS[args.workload]
    S = args.snowpacks if args.snowpacks != SNOWPACKS_PER_GPU else S_default
    batch = make_extra_batch(args.workload, S)
    plan = capi.Plan(capi.make_options(batch, n_max_stream=n, m_max=2))
    dev = DeviceBatch(batch, n)
    bt = dev.struct()
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(max(args.warmup, 1)):
        plan.solve_device(bt, stream)
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ms, eig, bnd = [], [], []
    for _ in range(args.steps):
        ev0.record(); plan.solve_device(bt, stream); ev1.record(); ev1.synchronize()
        ms.append(ev0.elapsed_time(ev1))
        plan.sync_timing(stream); tm = plan.last_timing(); eig.append(tm["eigen_ms"]); bnd.append(tm["boundary_ms"])
    status = dev.status.cpu().numpy()
    vals = dev.values.cpu().numpy()
    rate = batch.B * args.steps / (sum(ms) * 1e-3)
    peak = capi.measure_fp64_peak(0, 300.0)
    print(json.dumps({"workload": args.workload, "description": desc, "snowpacks": S, "solves_per_step": batch.B,
                      "value": rate, "unit": UNIT, "ms_per_step": float(np.mean(ms)),
                      "eigen_ms_per_step": float(np.mean(eig)), "boundary_ms_per_step": float(np.mean(bnd)),
                      "algorithmic_gflop_per_solve": gflop, "fp64_peak_tflops": peak,
                      "whole_path_frac_of_fp64_peak": gflop * 1e9 * rate / 1e12 / peak if peak else None,
                      "errors": int(np.count_nonzero(status & capi.ST_ERR_MASK)),
                      "finite": bool(np.isfinite(vals[(status & capi.ST_ERR_MASK) == 0]).all()),
                      "workspace_gb": plan.workspace_bytes / 1e9}))


def f_alg_per_solve(L=N_LAYERS, n=N_STREAMS, npol=2):
    """algorithmic flops of one mode-solve, split between the two kernels."""
    N = 2 * npol * n
    h = N // 2
    eigen = L * 31.0 * h**3  # 4 h^3 (form the product) + 25 h^3 (eigenpairs) + 2 h^3 (recover E-)
    boundary = L * (2.0 / 3.0 + 4.0 + 2.0) * N**3  # LU + two N-rhs solves + one GEMM per block-elimination step
    return eigen, boundary  # sum = 10.54 L N^3


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
def _oracle_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import dort_oracle as O

    problems = args
    t0 = time.perf_counter()
    vals = [O.solve_problem(p)["values"] for p in problems]
    return time.perf_counter() - t0, vals


def cpu_reference_rate(n_snowpacks, seed, cores=None, repeat=1):
    """Oracle (CPU restatement of the reference's NumPy/SciPy DORT) on `cores` processes with numerical threads pinned
    to 1 — how the reference itself parallelises (joblib processes, smrt/core/lib.py:655-666)."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    batch = make_batch(n_snowpacks, seed)
    probs = [batch.to_problem(i, dict(n_max_stream=N_STREAMS)) for i in range(batch.B)]
    cores = max(1, min(cores, len(probs)))
    chunks = [probs[i::cores] for i in range(cores)]
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_worker, [c[:1] for c in chunks])  # warm-up: imports, LAPACK init
        for _ in range(repeat):
            t0 = time.perf_counter()
            pool.map(_oracle_worker, chunks)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return len(probs) / best, cores, len(probs), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_snow = max(2, min(4 * cores, 64) // 6 + 1)  # bounded sample: ~ a few solves per core
    rates, times = [], []
    steps = max(1, args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(n_snow, 2, cores)
    for _ in range(steps):
        rate, used, nsolves, dt = cpu_reference_rate(n_snow, 2, cores)
        rates.append(rate); times.append(dt)
    value = float(np.mean(rates))
    sample = f"{nsolves} solves per step ({n_snow} snowpacks of the cfg-2 ensemble x 6 frequencies, seed 2)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_snow, note="reference arm: CPU oracle port of the reference's NumPy/SciPy DORT "
                                  "(LAPACK dgees/dgeev + dgbsv), one process per host core, BLAS threads = 1"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(S, note=None):
    cfg = {"workload": f"IBA(exponential)+DORT passive, {N_LAYERS} layers, {N_STREAMS} streams, 6 AMSR-E frequencies, "
                       f"theta=55deg, {S} synthetic snowpacks per GPU (SURVEY 8d cfg 2, seed 2) = {6 * S} solves/step/GPU",
           "layers": N_LAYERS, "n_max_stream": N_STREAMS, "frequencies_ghz": [f / 1e9 for f in AMSRE],
           "snowpacks_per_gpu": S, "solves_per_step_per_gpu": 6 * S,
           "l2_policy": "256 MiB buffer written between timed steps (L2 flush); the per-step layer workspace "
                        "(eigenvectors, GBs) is far larger than L2 anyway"}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from smrt_b200 import capi
    from smrt_b200.device import DeviceBatch

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

    S = args.snowpacks
    batch = make_batch(S, seed=2 + 1000 * rank)  # every rank its own members (weak scaling); rank 0 = the cfg-2 ensemble
    opts = capi.make_options(batch, n_max_stream=N_STREAMS, device=local_rank)
    plan = capi.Plan(opts)
    dev = DeviceBatch(batch, N_STREAMS, device=local_rank)
    bt = dev.struct()
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev.device)
    gathered = [torch.empty_like(dev.values) for _ in range(world)] if world > 1 else None

    def step():
        plan.solve_device(bt, stream)
        if world > 1:
            dist.all_gather(gathered, dev.values)  # the only collective: final gather of the results

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
    step_ms, eig_ms, bnd_ms, chunks = [], [], [], 0
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (outside the event bracket)
        ev0.record()
        step()
        ev1.record()
        ev1.synchronize()
        step_ms.append(ev0.elapsed_time(ev1))
        plan.sync_timing(stream)
        tm = plan.last_timing()
        eig_ms.append(tm["eigen_ms"]); bnd_ms.append(tm["boundary_ms"]); chunks = tm["chunks"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = plan.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    solves_per_step = batch.B * world
    value = solves_per_step * args.steps / (total_ms * 1e-3)

    # correctness guard: a bench number from wrong results is worthless
    status = dev.status.cpu().numpy()
    n_err = int(np.count_nonzero(status & capi.ST_ERR_MASK))
    tb = dev.values.cpu().numpy()

    # end-to-end through the host-buffer entry point (H2D + kernels + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 3))
    plan.solve_host(batch)  # warm-up: allocates the pinned staging
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
    e2e_value = solves_per_step * e2e_steps / e2e_dt
    d2h_bytes = sum(a.nbytes for a in (out.values, out.ks, out.ka, out.eps_eff, out.n_streams, out.stream_angles,
                                       out.optical_depth, out.status))
    e2e_match = bool(np.array_equal(out.values, tb))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline -------------------------------------------------------------------------------------------------------
    # per-kernel durations are measured live with CUDA events in a separate pass with the chunks serialised on one
    # stream: in the throughput pass above the two slots overlap and a kernel's event interval includes queueing
    peak_tflops = capi.measure_fp64_peak(local_rank, 300.0)
    fe, fb = f_alg_per_solve()
    sopts = capi.make_options(batch, n_max_stream=N_STREAMS, device=local_rank, serialize=True)
    splan = capi.Plan(sopts)
    eig_ms, bnd_ms = [], []
    for it in range(3):
        flush.fill_(1)
        splan.solve_device(bt, stream)
        splan.sync_timing(stream)
        tm = splan.last_timing()
        if it > 0:
            eig_ms.append(tm["eigen_ms"]); bnd_ms.append(tm["boundary_ms"])
        chunks = tm["chunks"]
    launches_roof = splan.launch_count
    splan.close()
    n_launch = chunks * len(eig_ms)
    eig_avg_ms = float(np.sum(eig_ms)) / max(n_launch, 1)
    bnd_avg_ms = float(np.sum(bnd_ms)) / max(n_launch, 1)
    solves_per_launch = batch.B / max(chunks, 1)
    dominant = "eigen_kernel" if np.sum(eig_ms) >= np.sum(bnd_ms) else "boundary_kernel"
    f_dom = fe if dominant == "eigen_kernel" else fb
    t_dom = eig_avg_ms if dominant == "eigen_kernel" else bnd_avg_ms
    achieved = f_dom * solves_per_launch / (t_dom * 1e-3) / 1e12 if t_dom > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    # capture (profiles/traffic.json, written by tools/make_profile_summary.py), scaled to this run's solves per launch
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dominant)
        traffic = float(tj["dram_bytes_per_solve"]) * solves_per_launch
    except Exception:
        pass
    alg_bytes_per_solve = N_LAYERS * (2 * 64 * 64 + 64) * 8 * 2 + 20 * 14 * 8  # eigenvector workspace write + read
    roofline = {
        "bound": "fp64", "kernel": dominant, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops if peak_tflops else None, "traffic": traffic,
        "peak_source": "measured in this run: DFMA micro-kernel (smrtb200_measure_fp64_peak); MEASURED_PEAKS.json has "
                       "no FP64 entry",
        "algorithmic_flops_per_solve": {"eigen_kernel": fe, "boundary_kernel": fb, "total": fe + fb},
        "avg_launch_ms": {"eigen_kernel": eig_avg_ms, "boundary_kernel": bnd_avg_ms},
        "solves_per_launch": solves_per_launch,
        "whole_path": {"achieved": (fe + fb) * (value / world) / 1e12, "frac": (fe + fb) * (value / world) / 1e12 / peak_tflops
                       if peak_tflops else None},
        "hbm": {"algorithmic_bytes_per_solve": alg_bytes_per_solve,
                "achieved_gbs": alg_bytes_per_solve * (value / world) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                "note": "of measured (MEASURED_PEAKS.json); the path is FP64/shared-memory bound, not HBM bound"},
    }

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_snow = max(2, min(4 * cores, 64) // 6 + 1)
        rate, used, nsolves, dt = cpu_reference_rate(n_snow, 2, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{nsolves} solves ({n_snow} snowpacks x 6 frequencies of the same ensemble) in {dt:.1f} s; CPU "
                         "oracle = restatement of the reference's NumPy/SciPy DORT, one process per core, BLAS threads = 1"}
        # parity of the timed GPU results on that sample
        from oracle import dort_oracle as O
        ref = O.solve_problem(batch.to_problem(0, dict(n_max_stream=N_STREAMS)))["values"]
        cpu["max_rel_err_vs_oracle_member0"] = float(np.max(np.abs(tb[0] - ref) / np.abs(ref)))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(S),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(dev.h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "matches_device_path": e2e_match,
                "api": "smrtb200_solve_batch_host (ctypes, host buffers)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "errors": n_err, "wall_s": t_wall, "workspace_gb": plan.workspace_bytes / 1e9,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--snowpacks", type=int, default=SNOWPACKS_PER_GPU, help="synthetic snowpacks per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2"] + sorted(EXTRA_WORKLOADS),
                    help="cfg2 = the contract workload; cfg3 / cfg4 / cfg5: other BASELINE configs, one GPU, for the record")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner from C, and with 4 / 8
    # ranks it ignores NCCL_DEBUG_FILE) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the
    # whole run and Python's sys.stdout keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.workload != "cfg2":
        run_extra(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
