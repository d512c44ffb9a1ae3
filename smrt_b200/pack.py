"""Struct-of-arrays packing of (sensor, snowpack) simulations for the batched B200 DORT solve.

The reference walks Python objects one simulation at a time (``smrt/core/model.py:584-619``: one emmodel instance per
layer, one ``DORT`` instance per simulation).  Here every simulation of a ``Model.run`` call becomes one row of a set
of flat fp64/int32 arrays — the exact arrays the C ABI takes (``include/smrt_dort_b200.h``).  Inputs are read-only
duck-typed objects: the reference's own ``Snowpack`` / ``Layer`` / ``Sensor`` (``smrt/core/snowpack.py:37-46``,
``smrt/core/layer.py:42-156``, ``smrt/core/sensor.py:274-339``) or the light stand-ins of ``smrt_b200.inputs``.

Anything the device path does not implement raises ``SMRTError`` naming the feature — there is no CPU fallback.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from .error import SMRTError

# enumerations shared with include/smrt_dort_b200.h
EM_IBA, EM_DMRT_QCA_SR, EM_NONSCATTERING, EM_DMRT_QCACP_SR = 0, 1, 2, 3
MS_EXPONENTIAL, MS_SHS, MS_HOMOGENEOUS = 0, 1, 2
IF_FLAT, IF_TRANSPARENT = 0, 1
SUB_NONE, SUB_FLAT = 0, 1
MODE_PASSIVE, MODE_ACTIVE = 0, 1

_EMMODEL_NAMES = {
    "iba": EM_IBA,
    "dmrt_qca_shortrange": EM_DMRT_QCA_SR,
    "nonscattering": EM_NONSCATTERING,
    "dmrt_qcacp_shortrange": EM_DMRT_QCACP_SR,
}
_EMMODEL_CLASSNAMES = {"IBA": EM_IBA, "DMRT_QCA_ShortRange": EM_DMRT_QCA_SR, "NonScattering": EM_NONSCATTERING,
                       "DMRT_QCACP_ShortRange": EM_DMRT_QCACP_SR}
_DMRT_CODES = (EM_DMRT_QCA_SR, EM_DMRT_QCACP_SR)


def emmodel_code(em) -> int:
    """Map an emmodel given as name, class or specialised class to the device enumeration."""
    if em is None:
        raise SMRTError("an emmodel is required (make_model(emmodel, ...) or layer.emmodel)")
    if isinstance(em, str):
        if em in _EMMODEL_NAMES:
            return _EMMODEL_NAMES[em]
        raise SMRTError(f"emmodel '{em}' is not implemented on the B200 path (available: {sorted(_EMMODEL_NAMES)})")
    for cls in getattr(em, "__mro__", [type(em)]):
        if cls.__name__ in _EMMODEL_CLASSNAMES:
            return _EMMODEL_CLASSNAMES[cls.__name__]
    raise SMRTError(f"emmodel {em!r} is not implemented on the B200 path")


@dataclass
class ProblemBatch:
    """B independent (snowpack x frequency) problems, padded to L_max layers.  All float arrays are fp64."""

    mode: int  # MODE_PASSIVE / MODE_ACTIVE (uniform over the batch)
    frequency: np.ndarray  # (B,)
    nlayer: np.ndarray  # (B,) int32
    thickness: np.ndarray  # (B, L)
    temperature: np.ndarray  # (B, L)
    frac_volume: np.ndarray  # (B, L)
    eps_bg: np.ndarray  # (B, L) complex128 — background permittivity, layer.permittivity(0, f)
    eps_sc: np.ndarray  # (B, L) complex128 — scatterer permittivity,  layer.permittivity(1, f)
    emmodel: np.ndarray  # (B, L) int32
    ms_kind: np.ndarray  # (B, L) int32
    ms_p0: np.ndarray  # (B, L) corr_length | radius
    ms_p1: np.ndarray  # (B, L) stickiness
    interface: np.ndarray  # (B, L) int32 — interface ABOVE layer l
    substrate_kind: np.ndarray  # (B,) int32
    substrate_eps: np.ndarray  # (B,) complex128
    substrate_temperature: np.ndarray  # (B,)
    theta: np.ndarray  # (n_theta,) rad — viewing angles (shared by the batch)
    theta_inc: np.ndarray  # (n_inc,) rad — incidence angles (active); empty for passive
    phi: float = np.pi
    dense_snow_correction: np.ndarray = None  # (B, L) int32: 1 = invert the medium when frac_volume > 0.5

    def __post_init__(self):
        if self.dense_snow_correction is None:
            self.dense_snow_correction = np.zeros(self.thickness.shape, dtype=np.int32)

    @property
    def B(self) -> int:
        return len(self.frequency)

    @property
    def L(self) -> int:
        return self.thickness.shape[1]

    def subset(self, sl) -> "ProblemBatch":
        kw = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray) and k not in ("theta", "theta_inc"):
                kw[k] = v[sl]
            else:
                kw[k] = v
        return ProblemBatch(**kw)

    def to_problem(self, i: int, options: Optional[dict] = None) -> dict:
        """Row i as the plain dict the CPU oracle takes (tests / bench cpu_baseline only)."""
        n = int(self.nlayer[i])
        opts = dict(options or {})
        return dict(
            frequency=float(self.frequency[i]),
            mode="P" if self.mode == MODE_PASSIVE else "A",
            thickness=self.thickness[i, :n].copy(),
            temperature=self.temperature[i, :n].copy(),
            frac_volume=self.frac_volume[i, :n].copy(),
            eps_bg=self.eps_bg[i, :n].copy(),
            eps_sc=self.eps_sc[i, :n].copy(),
            emmodel=self.emmodel[i, :n].copy(),
            ms_kind=self.ms_kind[i, :n].copy(),
            ms_p0=self.ms_p0[i, :n].copy(),
            ms_p1=self.ms_p1[i, :n].copy(),
            interface=self.interface[i, :n].copy(),
            dense_snow_correction=self.dense_snow_correction[i, :n].copy(),
            substrate_kind=int(self.substrate_kind[i]),
            substrate_eps=complex(self.substrate_eps[i]),
            substrate_temperature=float(self.substrate_temperature[i]),
            theta=self.theta.copy() if self.mode == MODE_PASSIVE else self.theta_inc.copy(),
            phi=float(self.phi),
            options=opts,
        )

    def save_fields(self) -> dict:
        return {k: (v if isinstance(v, np.ndarray) else np.asarray(v)) for k, v in self.__dict__.items()}

    @staticmethod
    def from_fields(d) -> "ProblemBatch":
        kw = {}
        for k in ProblemBatch.__dataclass_fields__:
            v = d[k]
            if k == "mode":
                kw[k] = int(v)
            elif k == "phi":
                kw[k] = float(v)
            else:
                kw[k] = np.asarray(v)
        return ProblemBatch(**kw)


def concat_batches(batches: Sequence[ProblemBatch]) -> ProblemBatch:
    L = max(b.L for b in batches)
    kw = {}
    for k in ProblemBatch.__dataclass_fields__:
        v0 = getattr(batches[0], k)
        if isinstance(v0, np.ndarray) and k not in ("theta", "theta_inc"):
            parts = []
            for b in batches:
                v = getattr(b, k)
                if v.ndim == 2 and v.shape[1] < L:
                    pad = np.zeros((v.shape[0], L - v.shape[1]), dtype=v.dtype)
                    v = np.concatenate([v, pad], axis=1)
                parts.append(v)
            kw[k] = np.concatenate(parts, axis=0)
        else:
            kw[k] = v0
    return ProblemBatch(**kw)


# ---------------------------------------------------------------------------------------------------------------------
# object packer
# ---------------------------------------------------------------------------------------------------------------------
def _microstructure_params(layer):
    ms = getattr(layer, "microstructure", None)
    if ms is None:
        raise SMRTError("layer without a microstructure model")
    name = type(ms).__name__
    if name == "Exponential":
        return MS_EXPONENTIAL, float(ms.corr_length), 0.0
    if name == "StickyHardSpheres":
        return MS_SHS, float(ms.radius), float(getattr(ms, "stickiness", 1000))
    if name == "Homogeneous":
        return MS_HOMOGENEOUS, 0.0, 0.0
    raise SMRTError(f"microstructure model '{name}' is not implemented on the B200 path "
                    "(available: Exponential, StickyHardSpheres, Homogeneous)")


def _interface_code(iface):
    name = type(iface).__name__ if not isinstance(iface, type) else iface.__name__
    if name == "Flat":
        return IF_FLAT
    if name == "Transparent":
        return IF_TRANSPARENT
    raise SMRTError(f"interface '{name}' is not implemented on the B200 path (only Flat and Transparent: "
                    "rough interfaces make the boundary blocks dense)")


def _substrate(substrate, frequency):
    if substrate is None:
        return SUB_NONE, 0j, 0.0
    name = type(substrate).__name__
    if name != "Flat":
        raise SMRTError(f"substrate '{name}' is not implemented on the B200 path (only a flat half-space)")
    perm = substrate.permittivity(frequency)
    temp = getattr(substrate, "temperature", None)
    return SUB_FLAT, complex(perm), (float(temp) if temp is not None else 0.0)


def pack_simulations(simulations, emmodel, emmodel_options=None) -> ProblemBatch:
    """Pack a flat list of (sensor, snowpack) pairs (reference ``Model.prepare_simulations`` order,
    ``smrt/core/model.py:485-502``) into one ProblemBatch.

    ``sensor`` must have a scalar frequency (DORT broadcasts every sensor axis but frequency,
    ``smrt/rtsolver/dort.py:140-146``); every sensor of the batch must share mode and angles.
    """
    simulations = list(simulations)
    if not simulations:
        raise SMRTError("nothing to simulate")
    emmodel_options = emmodel_options or {}
    sensor0 = simulations[0][0]
    mode = MODE_PASSIVE if sensor0.mode == "P" else MODE_ACTIVE

    B = len(simulations)
    L = max(max(len(sp.layers) for _, sp in simulations), 1)
    z = lambda dt=np.float64: np.zeros((B, L), dtype=dt)  # noqa: E731
    batch = ProblemBatch(
        mode=mode, frequency=np.zeros(B), nlayer=np.zeros(B, dtype=np.int32), thickness=z(), temperature=z(),
        frac_volume=z(), eps_bg=z(np.complex128), eps_sc=z(np.complex128), emmodel=z(np.int32), ms_kind=z(np.int32),
        ms_p0=z(), ms_p1=z(), interface=z(np.int32), substrate_kind=np.zeros(B, dtype=np.int32),
        substrate_eps=np.zeros(B, dtype=np.complex128), substrate_temperature=np.zeros(B),
        theta=np.atleast_1d(np.asarray(sensor0.theta, dtype=float)).copy(),
        theta_inc=(np.atleast_1d(np.asarray(sensor0.theta_inc, dtype=float)).copy() if mode == MODE_ACTIVE
                   else np.zeros(0)),
        phi=float(np.atleast_1d(getattr(sensor0, "phi", 0.0))[0]),
        dense_snow_correction=z(np.int32),
    )
    phi = np.atleast_1d(getattr(sensor0, "phi", 0.0))
    if len(phi) > 1:
        raise SMRTError("phi as an array must be implemented")  # same as reference dort.py:180-187

    for b, (sensor, sp) in enumerate(simulations):
        if sensor.mode != sensor0.mode:
            raise SMRTError("all the sensors of a batch must have the same mode")
        f = np.atleast_1d(sensor.frequency)
        if len(f) != 1:
            raise SMRTError("internal error: the frequency axis must be split before packing")
        f = float(f[0])
        batch.frequency[b] = f
        th = np.atleast_1d(np.asarray(sensor.theta, dtype=float))
        if th.shape != batch.theta.shape or not np.array_equal(th, batch.theta):
            raise SMRTError("all the sensors of a batch must have the same viewing angles")
        if getattr(sp, "atmosphere", None) is not None:
            raise SMRTError("atmosphere is not implemented on the B200 path")
        n = len(sp.layers)
        batch.nlayer[b] = n
        if len(sp.interfaces) != n:
            raise SMRTError("the snowpack must have one interface per layer")
        for l, layer in enumerate(sp.layers):
            batch.thickness[b, l] = layer.thickness
            batch.temperature[b, l] = layer.temperature
            em = getattr(layer, "emmodel", None) or emmodel
            if isinstance(emmodel, (list, tuple)):
                em = emmodel[l]
            code = emmodel_code(em)
            opts = dict(getattr(em, "_smrt_options", {}) or {})  # class_specializer stand-in
            opts.update(getattr(layer, "emmodel_options", None) or emmodel_options)
            unknown = set(opts) - {"dense_snow_correction"}
            if unknown:
                raise SMRTError(f"emmodel options {sorted(unknown)} are not implemented on the B200 path")
            dsc = opts.get("dense_snow_correction", "auto" if code in _DMRT_CODES else None)
            if dsc not in (None, "auto"):
                raise SMRTError(f"dense_snow_correction={dsc!r} is not implemented")
            batch.dense_snow_correction[b, l] = 1 if dsc == "auto" else 0
            batch.emmodel[b, l] = code
            batch.frac_volume[b, l] = layer.frac_volume
            kind, p0, p1 = _microstructure_params(layer)
            if code in _DMRT_CODES and kind != MS_SHS:
                raise SMRTError("DMRT short range models are only compatible with SHS microstructure model")
            if code == EM_IBA and kind == MS_HOMOGENEOUS:
                raise SMRTError("IBA needs a microstructure with a Fourier transform (exponential, sticky hard spheres)")
            batch.ms_kind[b, l], batch.ms_p0[b, l], batch.ms_p1[b, l] = kind, p0, p1
            if getattr(layer, "inclusion_shape", None) not in (None, "spheres"):
                raise SMRTError("only spherical inclusions are implemented on the B200 path")
            if getattr(layer, "depolarization_factors", None) is not None or \
                    getattr(layer, "length_ratio", None) not in (None, 1, 1.0):
                raise SMRTError("anisotropic depolarization factors are not implemented on the B200 path")
            batch.eps_bg[b, l] = complex(layer.permittivity(0, f))
            batch.eps_sc[b, l] = complex(layer.permittivity(1, f))
            batch.interface[b, l] = _interface_code(sp.interfaces[l])
        kind, eps, temp = _substrate(sp.substrate, f)
        batch.substrate_kind[b], batch.substrate_eps[b], batch.substrate_temperature[b] = kind, eps, temp
    return batch


# ---------------------------------------------------------------------------------------------------------------------
# array ("ensemble") entry point — no Python objects at all
# ---------------------------------------------------------------------------------------------------------------------
def ice_permittivity_maetzler06(frequency, temperature):
    """Host-vectorised pure-ice permittivity (Mätzler 2006) — reference ``smrt/permittivity/ice.py:24-73``.

    Input-side preparation for the ensemble entry point: one numpy expression over the whole (B, L)
    block instead of B*L Python calls.
    """
    frequency = np.asarray(frequency, dtype=float)
    temperature = np.asarray(temperature, dtype=float)
    freqGHz = frequency / 1e9
    tempC = temperature - 273.15
    if np.any(tempC > 0):
        raise SMRTError("The ice temperature must be lower or equal to 273.15K")
    Ereal = 3.1884 + 9.1e-4 * tempC
    theta = 300.0 / temperature - 1.0
    alpha = (0.00504 + 0.0062 * theta) * np.exp(-22.1 * theta)
    B1, B2, b = 0.0207, 1.16e-11, 335.0
    deltabeta = np.exp(-9.963 + 0.0372 * tempC)
    betam = (B1 / temperature) * (np.exp(b / temperature) / ((np.exp(b / temperature) - 1) ** 2)) + B2 * freqGHz**2
    beta = betam + deltabeta
    return Ereal + 1j * (alpha / freqGHz + beta * freqGHz)


def pack_snow_ensemble(frequency, thickness, density, temperature, *, microstructure="exponential",
                       corr_length=None, radius=None, stickiness=None, emmodel="iba", mode="P", theta_deg=55.0,
                       theta_inc_deg=None, phi_deg=180.0) -> ProblemBatch:
    """Dry-snow ensemble given directly as arrays: ``(S, L)`` profiles x ``(F,)`` frequencies -> ``F*S`` problems in the
    reference's simulation order (frequency outermost, snowpack innermost; ``smrt/core/model.py:485-502``).

    Equivalent to ``make_snowpack(thickness[s], microstructure, density=density[s], temperature=temperature[s], ...)``
    for every member (``smrt/inputs/make_medium.py:158-232``: frac_volume = density / 916.7, background air eps = 1,
    scatterers = pure ice, flat interfaces, no substrate) without building any Python object.
    """
    thickness = np.atleast_2d(np.asarray(thickness, dtype=float))
    S, L = thickness.shape
    density = np.broadcast_to(np.asarray(density, dtype=float), (S, L))
    temperature = np.broadcast_to(np.asarray(temperature, dtype=float), (S, L))
    freqs = np.atleast_1d(np.asarray(frequency, dtype=float))
    F = len(freqs)
    B = F * S
    code = emmodel_code(emmodel)

    def tile(a, dt=np.float64):
        return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=dt), (S, L))[None].repeat(F, axis=0)
                                    .reshape(B, L))

    if microstructure == "exponential":
        kind, p0, p1 = MS_EXPONENTIAL, corr_length, 0.0
    elif microstructure == "sticky_hard_spheres":
        kind, p0, p1 = MS_SHS, radius, (1000 if stickiness is None else stickiness)
    else:
        raise SMRTError(f"microstructure '{microstructure}' is not implemented on the B200 path")
    if p0 is None:
        raise SMRTError("the microstructure parameter (corr_length or radius) is required")
    freq_b = np.repeat(freqs, S)
    temp_b = tile(temperature)
    eps_sc = ice_permittivity_maetzler06(freq_b[:, None], temp_b)
    theta = np.radians(np.atleast_1d(np.asarray(theta_deg, dtype=float)))
    if mode == "A":
        theta_inc = np.radians(np.atleast_1d(np.asarray(theta_deg if theta_inc_deg is None else theta_inc_deg,
                                                        dtype=float)))
    else:
        theta_inc = np.zeros(0)
    return ProblemBatch(
        mode=MODE_PASSIVE if mode == "P" else MODE_ACTIVE,
        frequency=freq_b, nlayer=np.full(B, L, dtype=np.int32), thickness=tile(thickness), temperature=temp_b,
        frac_volume=tile(density / 916.7), eps_bg=np.ones((B, L), dtype=np.complex128), eps_sc=eps_sc,
        emmodel=np.full((B, L), code, dtype=np.int32), ms_kind=np.full((B, L), kind, dtype=np.int32),
        ms_p0=tile(p0), ms_p1=tile(p1), interface=np.zeros((B, L), dtype=np.int32),
        substrate_kind=np.zeros(B, dtype=np.int32), substrate_eps=np.zeros(B, dtype=np.complex128),
        substrate_temperature=np.zeros(B), theta=theta, theta_inc=theta_inc, phi=float(np.radians(phi_deg)),
        dense_snow_correction=np.full((B, L), 1 if code in _DMRT_CODES else 0, dtype=np.int32),
    )
