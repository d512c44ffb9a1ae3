"""Max relative error of the CUDA path against every reference-generated fixture (tests/golden), one line each.
Test tooling (imports tests/emu_util); used to judge numerical changes of the kernels on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from emu_util import load_golden, rel_err  # noqa: E402
from test_gpu_parity import ALL_GOLDEN, gpu_solve  # noqa: E402

worst = {0: 0.0, 1: 0.0}
for name in ALL_GOLDEN:
    d, batch, opts = load_golden(name)
    out = gpu_solve(batch, opts)
    err = rel_err(out.values, d["ref_values"], batch.mode)
    worst[batch.mode] = max(worst[batch.mode], err)
    print(f"{name:40s} mode={batch.mode} B={batch.B:3d} max_rel_err={err:.3e} status={np.unique(out.status)}")
print(f"WORST passive={worst[0]:.3e} active={worst[1]:.3e}")
