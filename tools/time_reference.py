#!/usr/bin/env python
"""Authoring-container check of the CPU arm: the UNMODIFIED reference (smrt from /root/reference, xarray stand-in) and
the oracle port (oracle/dort_oracle.py, what bench.py's cpu_baseline / --impl reference run on the GPU box, where the
reference tree does not exist) on the same cfg-2 snowpacks, same machine, numerical threads pinned to 1:
  (i)  one core   (reference: parallel_computation="none"; port: plain loop)
  (ii) all cores  (reference: joblib processes like JoblibParallelRunner; port: bench.CpuArm)
Writes profiles/<tag>_reference_vs_port.json.  usage: tools/time_reference.py [tag] [n_snowpacks]"""
import os

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ[_v] = "1"
import json
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "xarray_shim"), "/root/reference"]
import bench  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r04"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cores = os.cpu_count() or 1
w = bench.WORKLOADS["cfg2"]
th, rho, T, pc = bench.snow_members(S, w["seed"], w["layers"], "exp")


def run_reference_chunk(idx):
    import smrt
    from smrt import make_model, make_snowpack, sensor_list

    sps = [make_snowpack(th[s], "exponential", density=rho[s], temperature=T[s], corr_length=pc[s]) for s in idx]
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=w["streams"]))
    t0 = time.perf_counter()
    res = m.run(sensor_list.amsre(), sps, parallel_computation="none")
    return time.perf_counter() - t0, np.asarray(res.data.values)


if __name__ == "__main__":
    from joblib import Parallel, delayed

    out = {"workload": "cfg2", "snowpacks": S, "solves": 6 * S, "cores": cores, "cpu_model": bench._cpu_model()}
    run_reference_chunk([0])  # numba JIT, imports
    dt, ref_vals = run_reference_chunk(list(range(min(S, 6))))
    out["reference_one_core_solves_per_s"] = 6 * min(S, 6) / dt
    batch, probs = bench.sample_problems("cfg2", S)
    from oracle import dort_oracle as O
    O.solve_problem(probs[0])
    t0 = time.perf_counter()
    vals = [O.solve_problem(p)["values"] for p in probs[:36]]
    out["port_one_core_solves_per_s"] = 36 / (time.perf_counter() - t0)
    # all cores: the reference over joblib processes (what JoblibParallelRunner does), the port over bench.CpuArm
    chunks = [list(range(S))[i::cores] for i in range(cores)]
    with Parallel(n_jobs=cores) as par:
        par(delayed(run_reference_chunk)([0]) for _ in range(cores))  # warm every worker
        t0 = time.perf_counter()
        res = par(delayed(run_reference_chunk)(c) for c in chunks)
        out["reference_all_cores_solves_per_s"] = 6 * S / (time.perf_counter() - t0)
    arm = bench.CpuArm(cores)
    arm.run(probs[:cores])
    wall, pvals, busy = arm.run(probs)
    arm.close()
    out["port_all_cores_solves_per_s"] = len(probs) / wall
    # same numbers? reference values [F, S, 2, 1] vs the port's problems (frequency outermost)
    refall = np.empty((6, S, 2, 1))
    for c, (_, v) in zip(chunks, res):
        refall[:, c] = v
    port = np.array(pvals).reshape(6, S, 2, 1)
    out["max_rel_diff_port_vs_reference"] = float(np.max(np.abs(port - refall) / np.abs(refall)))
    out["port_over_reference_one_core"] = out["port_one_core_solves_per_s"] / out["reference_one_core_solves_per_s"]
    out["port_over_reference_all_cores"] = out["port_all_cores_solves_per_s"] / out["reference_all_cores_solves_per_s"]
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_reference_vs_port.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
