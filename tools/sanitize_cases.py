#!/usr/bin/env python
"""Small solves that walk every kernel instantiation, to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_cases.py [case ...]

Cases (truncated to 3 layers so that the 100x slower tools finish in seconds; parity is the GPU test-suite's business,
here only the code paths matter):
  passive16   16 streams passive: resident boundary kernel, two problems per SM
  passive32   32 streams passive: staged-operand boundary kernel (TMA bulk copies into [T | R])
  active16    16 streams active, m_max = 2 (cfg 3): three eigen records per layer, several right-hand sides
  active32    32 streams active, m_max = 2 (the reference's default radar setup: blocks of 96 unknowns)
  passive64   64 streams passive (cfg 4: blocks of 128 unknowns)
  passive128  128 streams passive (the reference's sea-ice example: global-scratch fallback)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from smrt_b200 import capi  # noqa: E402
from smrt_b200.pack import pack_snow_ensemble  # noqa: E402

CASES = {"passive16": ("P", 16), "passive32": ("P", 32), "active16": ("A", 16), "active32": ("A", 32),
         "passive64": ("P", 64), "passive128": ("P", 128)}


def run(name):
    mode, n = CASES[name]
    rng = np.random.default_rng(7)
    S, L = 3, 3
    th = np.concatenate((rng.uniform(0.05, 0.5, (S, L - 1)), np.full((S, 1), 1000.0)), axis=1)
    rho = rng.uniform(150, 450, (S, L)); T = rng.uniform(240, 272, (S, L)); pc = rng.uniform(5e-5, 3e-4, (S, L))
    kw = dict(mode="A", theta_inc_deg=40.0, theta_deg=40.0) if mode == "A" else dict(theta_deg=55.0)
    batch = pack_snow_ensemble([18.7e9, 36.5e9], th, rho, T, corr_length=pc, **kw)
    plan = capi.Plan(capi.make_options(batch, n_max_stream=n, m_max=2))
    out = plan.solve_host(batch)
    plan.close()
    ok = bool(np.isfinite(out.values).all()) and not np.any(out.status & capi.ST_ERR_MASK)
    print(f"{name}: {batch.B} problems, finite={ok}, first value {out.values.reshape(batch.B, -1)[0, 0]:.6f}", flush=True)
    return ok


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    sys.exit(0 if all([run(n) for n in names]) else 1)
