#!/usr/bin/env python
"""Profiling aid: dump the Jacobi operands M = C^T L (and X-, X+) of real cfg-2 layer eigenproblems, computed with the
NumPy model of the device algorithm (oracle/b200_algorithm.py), to cases.bin for the standalone kernel benchmarks."""
import os, sys, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import b200_algorithm as A

cases = []
orig = A.layer_eigen_symmetric
def hook(P_half, mu, weight, ks, ke, m, norm_rows=None, normalization=True):
    out = orig(P_half, mu, weight, ks, ke, m, norm_rows, normalization)
    # recompute the operands exactly like the model does
    npol = 2 if m == 0 else 3
    h = npol * len(mu)
    mu_a = np.repeat(mu, npol); w_a = np.repeat(weight, npol); coef = 0.5 if m == 0 else 0.25
    q = np.tile(np.array([1.0, 1.0, 2.0])[:npol], len(mu)); D = np.tile(np.array([1.0, 1.0, -1.0])[:npol], len(mu))
    norm = out[3]
    g = np.sqrt(norm * q * coef * w_a / mu_a)
    Gpp = g[:, None] * (P_half[:, :h] / q[:, None]) * g[None, :]
    Gpm = g[:, None] * (P_half[:, h:] * D[None, :] / q[:, None]) * g[None, :]
    Gpp = 0.5 * (Gpp + Gpp.T); Gpm = 0.5 * (Gpm + Gpm.T)
    Xm = np.diag(ke / mu_a) - Gpp + Gpm; Xp = np.diag(ke / mu_a) - Gpp - Gpm
    L = np.linalg.cholesky(Xm); C = np.linalg.cholesky(Xp)
    M = C.T @ L
    cases.append((h, M, Xm, Xp, np.sort(out[0])[::-1]))
    return out
A.layer_eigen_symmetric = hook
nsnow = int(sys.argv[1]) if len(sys.argv) > 1 else 2
full = len(sys.argv) > 2 and sys.argv[2] == "full"
batch = bench.make_batch("cfg2", nsnow)
for b in range(batch.B):
    A.solve_problem(batch.to_problem(b, dict(n_max_stream=32)))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cases.bin")
with open(path, "wb") as f:
    f.write(struct.pack("ii", len(cases), 1 if full else 0))
    for h, M, Xm, Xp, sig in cases:
        f.write(struct.pack("i", h))
        for arr in ((M, Xm, Xp) if full else (M,)):
            f.write(np.asfortranarray(arr).tobytes(order="F"))
        f.write(sig.tobytes())
print(len(cases), "cases, h in", min(c[0] for c in cases), max(c[0] for c in cases), "->", path)
