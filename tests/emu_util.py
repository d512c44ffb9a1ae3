"""Helpers to run the device code through the CPU SIMT emulator (tests/simt_emu) — test infrastructure."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from smrt_b200 import capi
from smrt_b200.pack import ProblemBatch

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "simt_emu")
GOLDEN = os.path.join(HERE, "golden")
_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
        _lib = C.CDLL(os.path.join(EMU_DIR, "libsmrt_emu.so"))
        _lib.emu_solve_batch.argtypes = [C.POINTER(capi.Options), C.POINTER(capi.Batch), C.c_int, C.c_void_p]
    return _lib


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    batch = ProblemBatch.from_fields({k[3:]: d[k] for k in d.files if k.startswith("in_")})
    opts = json.loads(str(d["rtsolver_options"]))
    return d, batch, opts


def emu_solve(batch, opts, threads=64):
    lib = emu_lib()
    o = capi.make_options(batch, n_max_stream=opts.get("n_max_stream", 32), m_max=opts.get("m_max", 2),
                          phase_normalization=opts.get("phase_normalization", "auto"),
                          prune_deep_snowpack=opts.get("prune_deep_snowpack"),
                          rayleigh_jeans_approximation=opts.get("rayleigh_jeans_approximation", False))
    out = capi.HostOutputs(batch, o.n_max_stream)
    keep = []
    bt = capi.host_batch_struct(batch, out, keep)
    diag = (C.c_int * 2)()
    rc = lib.emu_solve_batch(C.byref(o), C.byref(bt), threads, C.cast(diag, C.c_void_p))
    assert rc == 0
    out.jacobi_sweeps, out.eigenproblems = diag[0], diag[1]
    return out


def rel_err(values, ref, mode):
    """max relative error on Tb (passive) or on the (V,H)x(V,H) intensity block (active)"""
    if mode == 0:
        den = np.where(np.abs(ref) > 0, np.abs(ref), 1.0)
        return float(np.max(np.abs(values - ref) / den))
    v, r = values[..., 0:2, 0:2, :], ref[..., 0:2, 0:2, :]
    scale = np.abs(r).max()
    if scale == 0:
        return float(np.abs(v).max())
    m = np.abs(r) > 1e-12 * scale
    return float((np.abs(v - r)[m] / np.abs(r)[m]).max()) if m.any() else 0.0
