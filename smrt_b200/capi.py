"""ctypes binding of the C ABI in ``include/smrt_dort_b200.h`` (``libsmrt_dort_b200.so``).

The library is built in-tree by ``smrt_b200/csrc/Makefile`` (``__graft_entry__.build()``) for sm_100a.  There is no
fallback: if the shared object is missing or cannot be loaded, ``load_library()`` raises ``SMRTError``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .error import SMRTError
from .pack import MODE_ACTIVE, MODE_PASSIVE, ProblemBatch

ABI_VERSION = 4

NORM_OFF, NORM_ON, NORM_FORCED = 0, 1, 2
ST_OK, ST_NORMALIZATION, ST_EIGEN, ST_SINGULAR, ST_INPUT, ST_SUBSTRATE = 0, 1, 2, 3, 4, 5
ST_ERR_MASK = 15
ST_WARN_SHALLOW = 16

STATUS_MESSAGES = {
    ST_NORMALIZATION: "The re-normalization of the phase function exceeds the predefined threshold of 30%. This is "
                      "likely because of a too large grain size or a bug in the phase function. You can deactivate "
                      "this check using phase_normalization=\"forced\" as an option of the dort solver, or return NaN "
                      "instead with rtsolver_options=dict(error_handling='nan').",
    ST_EIGEN: "The diagonalization failed in DORT (single scattering albedo > 1 in a layer, too large grains for the "
              "emmodel, or an almost diagonal matrix).",
    ST_SINGULAR: "The boundary-condition system of DORT is singular.",
    ST_SUBSTRATE: "Reflectivity may be outside validity range. ksigma should be << 1",
    ST_INPUT: "Invalid layer input for the DORT solver (fewer than 2 streams in a layer, or the sticky hard sphere "
              "parameter t has no solution: revise the stickiness).",
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Options(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("device", C.c_int), ("mode", C.c_int), ("n_max_stream", C.c_int),
        ("m_max", C.c_int), ("max_layers", C.c_int), ("max_batch", C.c_int), ("n_theta", C.c_int),
        ("n_inc", C.c_int), ("normalization", C.c_int), ("rayleigh_jeans", C.c_int),
        ("prune_deep_snowpack", C.c_double), ("chunk", C.c_int), ("reserved", C.c_int),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("B", C.c_int),
        ("frequency", C.c_void_p), ("nlayer", C.c_void_p), ("thickness", C.c_void_p), ("temperature", C.c_void_p),
        ("frac_volume", C.c_void_p), ("eps_bg", C.c_void_p), ("eps_sc", C.c_void_p), ("emmodel", C.c_void_p),
        ("ms_kind", C.c_void_p), ("ms_p0", C.c_void_p), ("ms_p1", C.c_void_p), ("interface_kind", C.c_void_p),
        ("dense_snow_correction", C.c_void_p), ("substrate_kind", C.c_void_p), ("substrate_eps", C.c_void_p),
        ("substrate_temperature", C.c_void_p), ("substrate_params", C.c_void_p), ("atmosphere", C.c_void_p),
        ("inclusion", C.c_void_p), ("interface_params", C.c_void_p),
        ("theta", C.c_void_p), ("theta_inc", C.c_void_p),
        ("phi", C.c_double),
        ("values", C.c_void_p), ("ks", C.c_void_p), ("ka", C.c_void_p), ("eps_eff", C.c_void_p),
        ("n_streams_out", C.c_void_p), ("stream_angles", C.c_void_p), ("optical_depth", C.c_void_p),
        ("status", C.c_void_p),
    ]


INPUT_FIELDS = [  # (Batch field, ProblemBatch attribute, numpy dtype)
    ("frequency", "frequency", np.float64), ("nlayer", "nlayer", np.int32), ("thickness", "thickness", np.float64),
    ("temperature", "temperature", np.float64), ("frac_volume", "frac_volume", np.float64),
    ("eps_bg", "eps_bg", np.complex128), ("eps_sc", "eps_sc", np.complex128), ("emmodel", "emmodel", np.int32),
    ("ms_kind", "ms_kind", np.int32), ("ms_p0", "ms_p0", np.float64), ("ms_p1", "ms_p1", np.float64),
    ("interface_kind", "interface", np.int32), ("dense_snow_correction", "dense_snow_correction", np.int32),
    ("substrate_kind", "substrate_kind", np.int32), ("substrate_eps", "substrate_eps", np.complex128),
    ("substrate_temperature", "substrate_temperature", np.float64),
    ("substrate_params", "substrate_params", np.float64), ("atmosphere", "atmosphere", np.float64),
    ("inclusion", "inclusion", np.float64), ("interface_params", "interface_params", np.float64),
]

EXPORTED_SYMBOLS = [
    "smrtb200_abi_version", "smrtb200_last_error", "smrtb200_device_count", "smrtb200_plan_create",
    "smrtb200_plan_destroy", "smrtb200_plan_workspace_bytes", "smrtb200_plan_launch_count",
    "smrtb200_solve_batch_device", "smrtb200_solve_batch_host", "smrtb200_plan_sync_timing",
    "smrtb200_plan_last_timing",
    "smrtb200_measure_fp64_peak",
]

_LIB = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libsmrt_dort_b200.so")


def load_library():
    """Load libsmrt_dort_b200.so (built in-tree).  Raises SMRTError if it is missing — there is no CPU fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise SMRTError(f"the CUDA library {path} is not built: run `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (or `make -C smrt_b200/csrc`). There is no CPU fallback for the DORT path.")
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise SMRTError(f"cannot load {path}: {e}")
    lib.smrtb200_abi_version.restype = C.c_int
    lib.smrtb200_last_error.restype = C.c_char_p
    lib.smrtb200_device_count.argtypes = [_ip]
    lib.smrtb200_plan_create.argtypes = [C.POINTER(Options), C.POINTER(C.c_void_p)]
    lib.smrtb200_plan_destroy.argtypes = [C.c_void_p]
    lib.smrtb200_plan_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    lib.smrtb200_plan_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    lib.smrtb200_solve_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p]
    lib.smrtb200_solve_batch_host.argtypes = [C.c_void_p, C.POINTER(Batch)]
    lib.smrtb200_plan_sync_timing.argtypes = [C.c_void_p, C.c_void_p]
    lib.smrtb200_plan_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), _ip]
    lib.smrtb200_measure_fp64_peak.argtypes = [C.c_int, C.c_float, _dp]
    if lib.smrtb200_abi_version() != ABI_VERSION:
        raise SMRTError("libsmrt_dort_b200.so has a different ABI version: rebuild it")
    _LIB = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        msg = lib.smrtb200_last_error()
        raise SMRTError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def normalization_code(phase_normalization) -> int:
    """DORT option phase_normalization (reference dort.py:94-103, 240-242) -> ABI code."""
    if phase_normalization == "forced":
        return NORM_FORCED
    if phase_normalization in ("auto", True):
        return NORM_ON  # IBA and the DMRT models respect the reciprocity principle
    if phase_normalization in (False, None):
        return NORM_OFF
    raise SMRTError(f"invalid phase_normalization option: {phase_normalization!r}")


def make_options(batch: ProblemBatch, *, n_max_stream=32, m_max=2, phase_normalization="auto",
                 prune_deep_snowpack=None, rayleigh_jeans_approximation=False, device=0, max_batch=None,
                 chunk=0, serialize=False) -> Options:
    if prune_deep_snowpack is True:
        prune_deep_snowpack = 6
    return Options(
        abi_version=ABI_VERSION, device=device, mode=batch.mode, n_max_stream=int(n_max_stream),
        m_max=int(m_max) if batch.mode == MODE_ACTIVE else 0, max_layers=batch.L,
        max_batch=int(max_batch or batch.B), n_theta=len(batch.theta), n_inc=max(len(batch.theta_inc), 0),
        normalization=normalization_code(phase_normalization),
        rayleigh_jeans=1 if rayleigh_jeans_approximation else 0,
        prune_deep_snowpack=float(prune_deep_snowpack) if prune_deep_snowpack else 0.0, chunk=int(chunk), reserved=1 if serialize else 0)


class HostOutputs:
    """numpy output block of one solve."""

    def __init__(self, batch: ProblemBatch, n_max_stream: int):
        B, L = batch.B, batch.L
        if batch.mode == MODE_PASSIVE:
            self.values = np.zeros((B, 2, len(batch.theta)))
        else:
            self.values = np.zeros((B, 3, 3, len(batch.theta_inc)))
        self.ks = np.zeros((B, L))
        self.ka = np.zeros((B, L))
        self.eps_eff = np.zeros((B, L), dtype=np.complex128)
        self.n_streams = np.zeros(B, dtype=np.int32)
        self.stream_angles = np.full((B, n_max_stream), np.nan)
        self.optical_depth = np.zeros(B)
        self.status = np.zeros(B, dtype=np.int32)


def host_batch_struct(batch: ProblemBatch, out: HostOutputs, keep: list) -> Batch:
    """Batch struct with HOST pointers (numpy arrays are made contiguous and kept alive in `keep`)."""
    bt = Batch()
    bt.B = batch.B
    for fld, attr, dt in INPUT_FIELDS:
        val = getattr(batch, attr)
        if val is None:  # optional input the batch does not carry (interface_params without rough interfaces): NULL
            continue
        arr = np.ascontiguousarray(val, dtype=dt)
        keep.append(arr)
        setattr(bt, fld, arr.ctypes.data)
    theta = np.ascontiguousarray(batch.theta, dtype=np.float64)
    theta_inc = np.ascontiguousarray(batch.theta_inc if len(batch.theta_inc) else np.zeros(1), dtype=np.float64)
    keep += [theta, theta_inc]
    bt.theta, bt.theta_inc, bt.phi = theta.ctypes.data, theta_inc.ctypes.data, float(batch.phi)
    bt.values, bt.ks, bt.ka, bt.eps_eff = (out.values.ctypes.data, out.ks.ctypes.data, out.ka.ctypes.data,
                                          out.eps_eff.ctypes.data)
    bt.n_streams_out, bt.stream_angles = out.n_streams.ctypes.data, out.stream_angles.ctypes.data
    bt.optical_depth, bt.status = out.optical_depth.ctypes.data, out.status.ctypes.data
    return bt


class Plan:
    """Owner of one ``smrtb200_plan`` (device workspace + streams)."""

    def __init__(self, options: Options):
        self.lib = load_library()
        self.options = options
        self._h = C.c_void_p()
        _check(self.lib, self.lib.smrtb200_plan_create(C.byref(options), C.byref(self._h)), "smrtb200_plan_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.smrtb200_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        v = C.c_ulonglong()
        _check(self.lib, self.lib.smrtb200_plan_workspace_bytes(self._h, C.byref(v)), "workspace_bytes")
        return v.value

    @property
    def launch_count(self) -> int:
        v = C.c_ulonglong()
        _check(self.lib, self.lib.smrtb200_plan_launch_count(self._h, C.byref(v)), "launch_count")
        return v.value

    def sync_timing(self, stream: Optional[int] = None):
        _check(self.lib, self.lib.smrtb200_plan_sync_timing(self._h, C.c_void_p(stream or 0)), "sync_timing")

    def last_timing(self):
        a, b, c, n = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        _check(self.lib, self.lib.smrtb200_plan_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c),
                                                            C.byref(n)), "timing")
        return dict(total_ms=a.value, eigen_ms=b.value, boundary_ms=c.value, chunks=n.value)

    def solve_host(self, batch: ProblemBatch) -> HostOutputs:
        """HOST buffers in, HOST buffers out (H2D + kernels + D2H inside the call)."""
        out = HostOutputs(batch, self.options.n_max_stream)
        keep: list = []
        bt = host_batch_struct(batch, out, keep)
        _check(self.lib, self.lib.smrtb200_solve_batch_host(self._h, C.byref(bt)), "smrtb200_solve_batch_host")
        return out

    def solve_device(self, bt: Batch, stream: Optional[int] = None):
        """DEVICE pointers (e.g. torch tensors' data_ptr()); asynchronous on `stream` (cudaStream_t as int)."""
        _check(self.lib, self.lib.smrtb200_solve_batch_device(self._h, C.byref(bt), C.c_void_p(stream or 0)),
               "smrtb200_solve_batch_device")


def measure_fp64_peak(device=0, ms=200.0) -> float:
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.smrtb200_measure_fp64_peak(device, C.c_float(ms), C.byref(v)), "measure_fp64_peak")
    return v.value
