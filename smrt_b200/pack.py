"""Struct-of-arrays packing of (sensor, snowpack) simulations for the batched B200 DORT solve.

The reference walks Python objects one simulation at a time (``smrt/core/model.py:584-619``: one emmodel instance per
layer, one ``DORT`` instance per simulation).  Here every simulation of a ``Model.run`` call becomes one row of a set
of flat fp64/int32 arrays — the exact arrays the C ABI takes (``include/smrt_dort_b200.h``).  Inputs are read-only
duck-typed objects: the reference's own ``Snowpack`` / ``Layer`` / ``Sensor`` (``smrt/core/snowpack.py:37-46``,
``smrt/core/layer.py:42-156``, ``smrt/core/sensor.py:274-339``) or the light stand-ins of ``smrt_b200.inputs``.

Anything the device path does not implement raises ``SMRTError`` naming the feature — there is no CPU fallback.
"""

from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from .error import SMRTError

# enumerations shared with include/smrt_dort_b200.h
EM_IBA, EM_DMRT_QCA_SR, EM_NONSCATTERING, EM_DMRT_QCACP_SR, EM_RAYLEIGH, EM_PRESCRIBED_KSKAEPS = 0, 1, 2, 3, 4, 5
EM_IBA_ORIGINAL, EM_IBA_MAXWELL_GARNETT = 6, 7  # smrt/emmodel/iba_original.py, iba_maxwell_garnett.py
MS_EXPONENTIAL, MS_SHS, MS_HOMOGENEOUS = 0, 1, 2
MS_INDEPENDENT_SPHERE, MS_TEUBNER_STREY, MS_UNIFIED_TS_1, MS_UNIFIED_TS_2, MS_SHS_T = 3, 4, 5, 6, 7
IF_FLAT, IF_TRANSPARENT, IF_IEM_FUNG92, IF_IEM_FUNG92_BRIOGONI10 = 0, 1, 2, 3
SUB_NONE, SUB_FLAT, SUB_SOIL_WEGMULLER, SUB_SOIL_QNH, SUB_REFLECTOR, SUB_ROUGH_CHOUDHURY = 0, 1, 2, 3, 4, 5
SUB_REFLECTOR_BACKSCATTER = 6
SUB_IEM_FUNG92, SUB_IEM_FUNG92_BRIOGONI10 = 7, 8
MODE_PASSIVE, MODE_ACTIVE = 0, 1

_EMMODEL_NAMES = {
    "iba": EM_IBA,
    "dmrt_qca_shortrange": EM_DMRT_QCA_SR,
    "nonscattering": EM_NONSCATTERING,
    "dmrt_qcacp_shortrange": EM_DMRT_QCACP_SR,
    "rayleigh": EM_RAYLEIGH,
    "prescribed_kskaeps": EM_PRESCRIBED_KSKAEPS,
    "iba_original": EM_IBA_ORIGINAL,
    "iba_maxwell_garnett": EM_IBA_MAXWELL_GARNETT,
}
_EMMODEL_CLASSNAMES = {"IBA": EM_IBA, "DMRT_QCA_ShortRange": EM_DMRT_QCA_SR, "NonScattering": EM_NONSCATTERING,
                       "DMRT_QCACP_ShortRange": EM_DMRT_QCACP_SR, "Rayleigh": EM_RAYLEIGH,
                       "Prescribed_KsKaEps": EM_PRESCRIBED_KSKAEPS, "IBA_original": EM_IBA_ORIGINAL,
                       "IBA_MaxwellGarnett": EM_IBA_MAXWELL_GARNETT}
SPHERICAL_INCLUSIONS = (1.0, 0.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0)
_DMRT_CODES = (EM_DMRT_QCA_SR, EM_DMRT_QCACP_SR)
_IBA_CODES = (EM_IBA, EM_IBA_ORIGINAL, EM_IBA_MAXWELL_GARNETT)


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
    substrate_params: np.ndarray = None  # (B, 4) parameters of the rough / prescribed substrates (see SUB_*)
    atmosphere: np.ndarray = None  # (B, 3) isotropic atmosphere: tb_down, tb_up (K), transmittance; (0, 0, 1) = none
    inclusion: np.ndarray = None  # (B, L, 5) weights of the spheres / needles solutions, depolarisation factors x, y, z
    interface_params: np.ndarray = None  # (B, L, 4) parameters of the rough interfaces (IF_* >= 2), None without any

    def __post_init__(self):
        if self.dense_snow_correction is None:
            self.dense_snow_correction = np.zeros(self.thickness.shape, dtype=np.int32)
        if self.substrate_params is None:
            self.substrate_params = np.zeros((len(self.frequency), 4))
        if self.atmosphere is None:
            self.atmosphere = np.tile(np.array([0.0, 0.0, 1.0]), (len(self.frequency), 1))
        if self.inclusion is None:
            self.inclusion = np.tile(np.array(SPHERICAL_INCLUSIONS), self.thickness.shape + (1,))

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
            substrate_params=self.substrate_params[i].copy(),
            atmosphere=self.atmosphere[i].copy(),
            inclusion=self.inclusion[i, :n].copy(),
            interface_params=None if self.interface_params is None else self.interface_params[i, :n].copy(),
            theta=self.theta.copy() if self.mode == MODE_PASSIVE else self.theta_inc.copy(),
            phi=float(self.phi),
            options=opts,
        )

    def save_fields(self) -> dict:
        return {k: (v if isinstance(v, np.ndarray) else np.asarray(v)) for k, v in self.__dict__.items() if v is not None}

    @staticmethod
    def from_fields(d) -> "ProblemBatch":
        kw = {}
        for k in ProblemBatch.__dataclass_fields__:
            if k in ("substrate_params", "atmosphere", "inclusion", "interface_params") and k not in d:  # fixtures written before these fields existed
                continue
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
        if k == "interface_params":
            if all(b.interface_params is None for b in batches):
                kw[k] = None
            else:
                parts = []
                for b in batches:
                    v = np.zeros((b.B, L, 4))
                    if b.interface_params is not None:
                        v[:, :b.L] = b.interface_params
                    parts.append(v)
                kw[k] = np.concatenate(parts, axis=0)
            continue
        if isinstance(v0, np.ndarray) and k not in ("theta", "theta_inc"):
            parts = []
            for b in batches:
                v = getattr(b, k)
                if v.ndim == 2 and v.shape[1] < L and k not in ("substrate_params", "atmosphere"):
                    pad = np.zeros((v.shape[0], L - v.shape[1]), dtype=v.dtype)
                    v = np.concatenate([v, pad], axis=1)
                elif k == "inclusion" and v.shape[1] < L:
                    pad = np.tile(np.array(SPHERICAL_INCLUSIONS), (v.shape[0], L - v.shape[1], 1))
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
    if name == "IndependentSphere":
        return MS_INDEPENDENT_SPHERE, float(ms.radius), 0.0
    if name == "TeubnerStrey":
        return MS_TEUBNER_STREY, float(ms.corr_length), float(ms.repeat_distance)
    if name == "UnifiedScaledExponential":  # unified_scaled_exponential.py:20-36: an exponential with a scaled length
        return MS_EXPONENTIAL, float(ms.corr_length), 0.0
    if name == "UnifiedTeubnerStrey":  # unified_teubner_strey.py:25-36: the two cases of Ruland 2010
        kind = MS_UNIFIED_TS_1 if ms.polydispersity >= 1 else MS_UNIFIED_TS_2
        return kind, float(ms.zeta1), float(ms.zeta2)
    if name == "UnifiedStickyHardSpheres":
        return MS_SHS_T, float(ms.radius), float(ms.t)
    raise SMRTError(f"microstructure model '{name}' is not implemented on the B200 path (available: Exponential, "
                    "StickyHardSpheres, IndependentSphere, TeubnerStrey, UnifiedScaledExponential, UnifiedTeubnerStrey, "
                    "UnifiedStickyHardSpheres, Homogeneous; models without an analytical Fourier transform are not)")


def _interface_code(iface):
    name = type(iface).__name__ if not isinstance(iface, type) else iface.__name__
    if name == "Flat":
        return IF_FLAT
    if name == "Transparent":
        return IF_TRANSPARENT
    if name in ("IEM_Fung92", "IEM_Fung92_Briogoni10"):
        return IF_IEM_FUNG92 if name == "IEM_Fung92" else IF_IEM_FUNG92_BRIOGONI10
    raise SMRTError(f"interface '{name}' is not implemented on the B200 path (Flat, Transparent, IEM_Fung92, "
                    "IEM_Fung92_Briogoni10: interfaces with a dense diffuse matrix make the boundary blocks dense)")


def _iem_params(obj):
    """(roughness_rms, corr_length, autocorrelation, series_truncation) of an IEM_Fung92 interface / substrate —
    reference smrt/interface/iem_fung92.py:60-67"""
    acf = getattr(obj, "autocorrelation_function", "exponential")
    if acf not in ("exponential", "gaussian"):
        raise SMRTError("The autocorrelation function must be exponential or gaussian")  # iem_fung92.py:189
    if getattr(obj, "warning_handling", "print") != "print":
        raise SMRTError("IEM_Fung92 on the B200 path follows warning_handling='print' (outside the validity range "
                        "the reference warns and goes on; no message is printed here)")
    N = int(getattr(obj, "series_truncation", 10))
    if not 1 <= N <= 64:
        raise SMRTError("series_truncation must be in 1..64")
    return [float(obj.roughness_rms), float(obj.corr_length), 1.0 if acf == "gaussian" else 0.0, float(N)]


def _reflector_value(substrate, frequency, polarization):
    """``Reflector._get_refl`` (reference smrt/substrate/reflector.py:83-111) for scalar / dict specifications"""
    spec = substrate.specular_reflection
    if spec is None:
        spec = 1
    if isinstance(spec, dict):
        for key in [(frequency, polarization), (polarization, frequency), frequency, polarization]:
            if key in spec:
                spec = spec[key]
                break
    if isinstance(spec, dict):
        raise SMRTError("The specular_reflection argument must be a scalar or a dict with the frequency and/or "
                        "polarization as a key. If both, provide frequency and polarization as a tuple key")
    if callable(spec):
        raise SMRTError("a Reflector with a specular_reflection function of theta is not implemented on the B200 path "
                        "(the stream angles are computed on the device): give scalars")
    return float(spec)


def _substrate(substrate, frequency, mode=MODE_PASSIVE):
    """-> (kind, permittivity, temperature, params[4]); reference smrt/substrate/flat.py, soil_wegmuller.py,
    soil_qnh.py, reflector.py, rough_choudhury79.py (all diagonal: Fresnel coefficients with a per-stream adjustment)"""
    par = np.zeros(4)
    if substrate is None:
        return SUB_NONE, 0j, 0.0, par
    name = type(substrate).__name__
    temp = getattr(substrate, "temperature", None)
    temp = float(temp) if temp is not None else 0.0
    if name == "Reflector":
        if mode != MODE_PASSIVE:  # reflector.py:56-57
            raise NotImplementedError("active model is not yet implemented, need modification for the third component")
        par[0] = _reflector_value(substrate, frequency, "V")
        par[1] = _reflector_value(substrate, frequency, "H")
        return SUB_REFLECTOR, 0j, temp, par
    if name == "ReflectorBackscatter":  # reflector_backscatter.py:66-135: passive and active (third component zero)
        spec, back = substrate.specular_reflection, substrate.backscattering_coefficient
        if spec is None and back is None:
            spec = 1  # reflector_backscatter.py:72-73
        if back is not None and not (isinstance(back, dict) and "VV" in back and "HH" in back):
            raise SMRTError("backscattering_coefficient must be a dictionary with keys VV and HH")
        if spec is None:
            raise SMRTError("a ReflectorBackscatter needs its specular_reflection next to the backscattering_coefficient "
                            "(the reference turns a missing one into NaN reflectivities)")
        vals = [spec["V"], spec["H"]] if isinstance(spec, dict) else [spec, spec]
        vals += [back["VV"], back["HH"]] if back is not None else [0.0, 0.0]
        if any(callable(v) for v in vals):
            raise SMRTError("a reflector given as a function of theta is not implemented on the B200 path (the stream "
                            "angles are computed on the device): give scalars")
        par[:] = [float(v) for v in vals]
        return SUB_REFLECTOR_BACKSCATTER, 0j, temp, par
    if name == "Flat":
        kind = SUB_FLAT
    elif name == "SoilWegmuller":
        kind = SUB_SOIL_WEGMULLER
        par[0] = float(substrate.roughness_rms)
    elif name == "ChoudhuryReflectivity":
        kind = SUB_ROUGH_CHOUDHURY
        par[0] = float(substrate.roughness_rms)
    elif name in ("IEM_Fung92", "IEM_Fung92_Briogoni10"):  # substrate/iem_fung92.py, iem_fung92_brogioni10.py
        kind = SUB_IEM_FUNG92 if name == "IEM_Fung92" else SUB_IEM_FUNG92_BRIOGONI10
        par[:] = _iem_params(substrate)
    elif name == "SoilQNH":
        kind = SUB_SOIL_QNH
        N = float(getattr(substrate, "N", 0.0))
        Nv, Nh = float(getattr(substrate, "Nv", np.nan)), float(getattr(substrate, "Nh", np.nan))
        par[:] = [float(substrate.H), float(getattr(substrate, "Q", 0.0)), N if np.isnan(Nv) else Nv,
                  N if np.isnan(Nh) else Nh]
    else:
        raise SMRTError(f"substrate '{name}' is not implemented on the B200 path (available: Flat, SoilWegmuller, "
                        "SoilQNH, ChoudhuryReflectivity, Reflector, ReflectorBackscatter, IEM_Fung92, IEM_Fung92_Briogoni10; rough substrates with a dense "
                        "diffuse reflection matrix are not: they make the boundary blocks dense)")
    return kind, complex(substrate.permittivity(frequency)), temp, par


def _atmosphere(atmosphere, frequency, mode=MODE_PASSIVE):
    """-> (tb_down, tb_up, transmittance) of an isotropic atmosphere (reference
    smrt/atmosphere/simple_isotropic_atmosphere.py:49-77: constants or frequency-keyed dicts)"""
    if atmosphere is None:
        return 0.0, 0.0, 1.0
    name = type(atmosphere).__name__
    if name != "SimpleIsotropicAtmosphere":
        raise SMRTError(f"atmosphere '{name}' is not implemented on the B200 path (available: "
                        "SimpleIsotropicAtmosphere)")

    def value(x):
        if isinstance(x, dict):
            x = x[frequency]
        return float(x)

    return value(atmosphere.constant_tbdown), value(atmosphere.constant_tbup), value(atmosphere.constant_trans)


def _shape_weights(inclusion_shape, mixing_ratio=None):
    """-> (weight of the "spheres" solution, weight of the "random_needles" one) of the reference's polder_van_santen
    (smrt/permittivity/generic_mixing_formula.py:88-141: a string, or a dict / sequence of shapes with mixing ratios)"""
    if inclusion_shape is None or inclusion_shape == "spheres":
        return 1.0, 0.0
    if inclusion_shape == "random_needles":
        return 0.0, 1.0
    if isinstance(inclusion_shape, str):
        raise SMRTError("inclusion_shape must be one of (or a list of) the following: 'spheres' (default) or "
                        "'random_needles'.")
    if isinstance(inclusion_shape, dict):
        if mixing_ratio is not None:
            raise SMRTError("Setting mixing_ratio and using a dict for inclusion_shape is ambiguous.")
        mixing_ratio = list(inclusion_shape.values())
        inclusion_shape = list(inclusion_shape.keys())
    try:
        mixing_ratio = list(mixing_ratio)
    except TypeError:
        mixing_ratio = [float(mixing_ratio)]
    if len(mixing_ratio) == len(inclusion_shape) - 1:
        mixing_ratio = mixing_ratio + [1 - np.sum(mixing_ratio)]
    elif len(mixing_ratio) != len(inclusion_shape):
        raise SMRTError("The length of inclusion_shape and mixing_ratio are incompatible. See the documentation.")
    w = [0.0, 0.0]
    for shape, mixing in zip(inclusion_shape, mixing_ratio):
        ws, wn = _shape_weights(shape)
        if w[0 if ws else 1] != 0.0:
            raise SMRTError("a shape given twice in inclusion_shape is not implemented on the B200 path")
        w[0 if ws else 1] = float(mixing)
    return w[0], w[1]


def _inclusion_params(layer, code):
    """Row of ProblemBatch.inclusion for a layer: effective-permittivity shape weights (used by the Polder - van Santen
    models: iba, iba_original, nonscattering) and the depolarisation factors of the IBA family (iba.py:112-119)."""
    ws, wn = 1.0, 0.0
    shape = getattr(layer, "inclusion_shape", None)
    if code == EM_IBA_MAXWELL_GARNETT:
        if shape not in (None, "spheres"):  # generic_mixing_formula.py:343-344
            raise SMRTError("inclusion_shape must be set to 'spheres'")
    elif code in _IBA_CODES or code == EM_NONSCATTERING:
        ws, wn = _shape_weights(shape, getattr(layer, "mixing_ratio", None))
    depol = getattr(layer, "depolarization_factors", None)
    if depol is not None:
        if callable(depol):
            raise SMRTError("callable depolarization_factors are not implemented on the B200 path")
        depol = np.asarray(depol, dtype=float)
        if depol.shape != (3,):
            raise SMRTError("depolarization_factors must hold three numbers")
    else:
        from .inputs import depolarization_factors_spheroids
        depol = depolarization_factors_spheroids(getattr(layer, "length_ratio", None))
    return ws, wn, float(depol[0]), float(depol[1]), float(depol[2])


_PERM_CONST, _PERM_ICE, _PERM_WETICE, _PERM_CALL = 0, 1, 2, 3
_KNOWN_ICE_MODELS = {"ice_permittivity_maetzler06": _PERM_ICE, "wetice_permittivity_bohren83": _PERM_WETICE}


def _classify_permittivity(model):
    """How the permittivity of one medium of a layer can be evaluated for a whole block at once: a constant, one of the
    two default ice models of the reference (recognised by name and module: ``smrt/permittivity/ice.py:24``,
    ``wetice.py:13``, or this package's restatements), or an arbitrary callable that has to be called layer by layer."""
    if not callable(model):
        return _PERM_CONST, complex(model)
    kind = _KNOWN_ICE_MODELS.get(getattr(model, "__name__", ""))
    module = getattr(model, "__module__", "") or ""
    if kind is not None and (module.startswith("smrt.permittivity") or module.startswith("smrt_b200")):
        return kind, 0j
    return _PERM_CALL, 0j


_MS_FAST = {"Exponential": (MS_EXPONENTIAL, "corr_length"), "IndependentSphere": (MS_INDEPENDENT_SPHERE, "radius")}
_IF_FAST = {"Flat": IF_FLAT, "Transparent": IF_TRANSPARENT}
_INCLUSION_KEYS = ("inclusion_shape", "depolarization_factors", "length_ratio", "mixing_ratio")


def _layer_rows(sp, emmodel, emmodel_options, L, defaults=None):
    """Frequency-independent rows of one snowpack (every quantity the device needs but the permittivities), plus how to
    evaluate the permittivities: (kind, constant) per medium and layer.  One pass over the layers; the values are
    collected in Python lists and converted once (element-wise writes into NumPy arrays cost more than the reads)."""
    layers = sp.layers
    n = len(layers)
    if len(sp.interfaces) != n:
        raise SMRTError("the snowpack must have one interface per layer")
    list_em = isinstance(emmodel, (list, tuple))
    map_em = isinstance(emmodel, Mapping)
    if list_em and len(emmodel) != n:
        raise SMRTError("the list of emmodels must have one entry per layer of the snowpack")
    list_opts = isinstance(emmodel_options, (list, tuple))
    if list_opts and len(emmodel_options) != n:
        raise SMRTError("the list of emmodel options must have one entry per layer of the snowpack")
    map_opts = map_em and emmodel_options and all(isinstance(o, Mapping) for o in emmodel_options.values())
    plain = not (list_em or map_em or list_opts)  # one emmodel, one option dict: resolved once by the caller
    thick, temp, fvol, p0s, p1s, lws = [], [], [], [], [], []
    codes, kinds, ifaces, dscs, kbg, ksc = [], [], [], [], [], []
    cbg, csc = [], []
    incl = None
    ipar = None  # parameters of the rough interfaces, (L, 4), only when the snowpack has one
    for l, layer in enumerate(layers):
        d = layer.__dict__
        own_em = d.get("emmodel")
        own_opts = d.get("emmodel_options")
        if plain and own_em is None and not own_opts and defaults is not None:
            code, dsc_flag = defaults
        else:
            # smrt/core/model.py:536-571: a list gives one emmodel per layer, a dict one per medium (layer.medium), else
            # the layer's own emmodel attribute wins over the model's; the options follow the same three shapes
            if list_em:
                em = emmodel[l]
            elif map_em:
                medium = getattr(layer, "medium", None)
                if medium not in emmodel:
                    raise SMRTError(f"no emmodel is given for the medium {medium!r} of layer {l}")
                em = emmodel[medium]
            else:
                em = own_em or emmodel
            code = emmodel_code(em)
            opts = getattr(em, "_smrt_options", None)  # class_specializer stand-in
            opts = dict(opts) if opts else {}
            if list_opts:
                opts.update(emmodel_options[l] or {})
            elif map_opts:
                opts.update(emmodel_options[getattr(layer, "medium", None)])
            else:
                opts.update(own_opts or emmodel_options)
            dsc_flag = _dense_snow_flag(opts, code)
        codes.append(code)
        dscs.append(dsc_flag)
        iface = sp.interfaces[l]
        icode = _IF_FAST.get(type(iface).__name__)
        if icode is None:
            icode = _interface_code(iface)
            if icode >= IF_IEM_FUNG92:
                if ipar is None:
                    ipar = np.zeros((L, 4))
                ipar[l] = _iem_params(iface)
        ifaces.append(icode)
        thick.append(layer.thickness)
        temp.append(layer.temperature)
        fvol.append(layer.frac_volume)
        if code == EM_PRESCRIBED_KSKAEPS:  # emmodel/prescribed_kskaeps.py:20-27: everything is given on the layer
            kinds.append(MS_HOMOGENEOUS)
            p0s.append(float(layer.ks))
            p1s.append(float(layer.ka))
            eps = complex(layer.effective_permittivity)
            kbg.append(_PERM_CONST); ksc.append(_PERM_CONST); cbg.append(eps); csc.append(eps); lws.append(0.0)
        else:
            ms = d.get("microstructure")
            fast = _MS_FAST.get(type(ms).__name__) if ms is not None else None
            if fast is not None:
                kind, p0, p1 = fast[0], float(getattr(ms, fast[1])), 0.0
            else:
                kind, p0, p1 = _microstructure_params(layer)
            if code in _DMRT_CODES and kind != MS_SHS:
                raise SMRTError("DMRT short range models are only compatible with SHS microstructure model")
            if code in _IBA_CODES and kind == MS_HOMOGENEOUS:
                raise SMRTError("IBA needs a microstructure with a Fourier transform (exponential, sticky hard spheres)")
            if code == EM_RAYLEIGH:  # emmodel/rayleigh.py:41-47
                if not hasattr(layer.microstructure, "radius"):
                    raise SMRTError("Only microstructure_model which defined a `radius` can be used with Rayleigh "
                                    "scattering")
                kind, p0, p1 = MS_HOMOGENEOUS, float(layer.microstructure.radius), 0.0
            kinds.append(kind); p0s.append(p0); p1s.append(p1)
            models = d.get("permittivity_model")
            if models is None:  # stand-in dry-snow layer of smrt_b200.inputs: ice (Maetzler 2006) in air
                kbg.append(_PERM_CONST); ksc.append(_PERM_ICE); cbg.append(1.0); csc.append(0j); lws.append(0.0)
            else:
                k0, c0 = _classify_permittivity(models[0])
                k1, c1 = _classify_permittivity(models[1])
                kbg.append(k0); ksc.append(k1); cbg.append(c0); csc.append(c1)
                lws.append((getattr(layer, "liquid_water", 0.0) or 0.0) if _PERM_WETICE in (k0, k1) else 0.0)
        # shape of the inclusions / depolarisation factors: only layers that say something about them
        if code == EM_IBA_MAXWELL_GARNETT or any(d.get(k) is not None for k in _INCLUSION_KEYS):
            if code == EM_IBA_MAXWELL_GARNETT or _needs_inclusion_params(layer):
                if incl is None:
                    incl = {}
                incl[l] = _inclusion_params(layer, code)
    f8 = np.zeros((6, L))
    f8[0, :n], f8[1, :n], f8[2, :n], f8[3, :n], f8[4, :n], f8[5, :n] = thick, temp, fvol, p0s, p1s, lws
    i4 = np.zeros((6, L), dtype=np.int32)
    i4[0, :n], i4[1, :n], i4[2, :n], i4[3, :n], i4[4, :n], i4[5, :n] = codes, kinds, ifaces, dscs, kbg, ksc
    const = np.zeros((2, L), dtype=np.complex128)
    const[0, :n], const[1, :n] = cbg, csc
    inclusion = np.empty((L, 5))
    inclusion[:] = SPHERICAL_INCLUSIONS
    if incl:
        for l, v in incl.items():
            inclusion[l] = v
    return n, f8, i4, inclusion, const, ipar


def _dense_snow_flag(opts, code):
    if opts:
        unknown = set(opts) - {"dense_snow_correction"}
        if unknown:
            raise SMRTError(f"emmodel options {sorted(unknown)} are not implemented on the B200 path")
    dsc = opts.get("dense_snow_correction", "auto" if code in _DMRT_CODES else None)
    if dsc not in (None, "auto"):
        raise SMRTError(f"dense_snow_correction={dsc!r} is not implemented")
    return 1 if dsc == "auto" else 0


def _needs_inclusion_params(layer):
    """False for the default layer (spherical inclusions, depolarisation factors 1/3): the common case is not charged
    the shape / depolarisation logic"""
    d = layer.__dict__ if hasattr(layer, "__dict__") else {}
    return (d.get("inclusion_shape") not in (None, "spheres") or d.get("depolarization_factors") is not None
            or d.get("length_ratio") not in (None, 1, 1.0) or d.get("mixing_ratio") is not None)


def pack_simulations(simulations, emmodel, emmodel_options=None, atmospheres=None) -> ProblemBatch:
    """Pack a flat list of (sensor, snowpack) pairs (reference ``Model.prepare_simulations`` order,
    ``smrt/core/model.py:485-502``) into one ProblemBatch.

    ``sensor`` must have a scalar frequency (DORT broadcasts every sensor axis but frequency,
    ``smrt/rtsolver/dort.py:140-146``); every sensor of the batch must share mode and angles.

    A snowpack that appears in several simulations (one per frequency of the sensor) is walked ONCE; the permittivities
    are evaluated for whole (snowpack, layer) blocks per frequency when the layer uses a constant or one of the
    reference's default ice models, and through ``layer.permittivity(i, f)`` otherwise (any user-defined model works).
    """
    simulations = list(simulations)
    if not simulations:
        raise SMRTError("nothing to simulate")
    emmodel_options = emmodel_options or {}
    sensor0 = simulations[0][0]
    mode = MODE_PASSIVE if sensor0.mode == "P" else MODE_ACTIVE
    theta0 = np.atleast_1d(np.asarray(sensor0.theta, dtype=float)).copy()
    phi = np.atleast_1d(getattr(sensor0, "phi", 0.0))
    if len(phi) > 1:
        raise SMRTError("phi as an array must be implemented")  # same as reference dort.py:180-187

    B = len(simulations)
    freq = np.empty(B)
    sp_index = np.empty(B, dtype=np.int64)
    unique, position = [], {}
    checked_sensors = {}
    for b, (sensor, sp) in enumerate(simulations):
        sid = id(sensor)
        f = checked_sensors.get(sid)
        if f is None:
            if sensor.mode != sensor0.mode:
                raise SMRTError("all the sensors of a batch must have the same mode")
            fa = np.atleast_1d(sensor.frequency)
            if len(fa) != 1:
                raise SMRTError("internal error: the frequency axis must be split before packing")
            th = np.atleast_1d(np.asarray(sensor.theta, dtype=float))
            if th.shape != theta0.shape or not np.array_equal(th, theta0):
                raise SMRTError("all the sensors of a batch must have the same viewing angles")
            f = checked_sensors[sid] = float(fa[0])
        freq[b] = f
        key = id(sp)
        u = position.get(key)
        if u is None:
            u = position[key] = len(unique)
            unique.append(sp)
        sp_index[b] = u

    U = len(unique)
    L = max(max(len(sp.layers) for sp in unique), 1)
    nlayer_u = np.zeros(U, dtype=np.int32)
    F8 = np.zeros((U, 6, L))
    I4 = np.zeros((U, 6, L), dtype=np.int32)
    INCL = np.empty((U, L, 5))
    CONST = np.zeros((U, 2, L), dtype=np.complex128)
    defaults = None
    if not isinstance(emmodel, (list, tuple, Mapping)) and not isinstance(emmodel_options, (list, tuple)) \
            and emmodel is not None:
        # one emmodel and one option dict for every layer that has none of its own: resolved once
        code0 = emmodel_code(emmodel)
        opts0 = dict(getattr(emmodel, "_smrt_options", None) or {})
        opts0.update(emmodel_options)
        defaults = (code0, _dense_snow_flag(opts0, code0))
    IPAR = None
    for u, sp in enumerate(unique):
        nlayer_u[u], F8[u], I4[u], INCL[u], CONST[u], ipar = _layer_rows(sp, emmodel, emmodel_options, L, defaults)
        if ipar is not None:
            if IPAR is None:
                IPAR = np.zeros((U, L, 4))
            IPAR[u] = ipar

    take = lambda a: np.ascontiguousarray(a[sp_index])  # noqa: E731
    temperature = take(F8[:, 1])
    batch = ProblemBatch(
        mode=mode, frequency=freq, nlayer=take(nlayer_u), thickness=take(F8[:, 0]), temperature=temperature,
        frac_volume=take(F8[:, 2]), eps_bg=np.zeros((B, L), dtype=np.complex128),
        eps_sc=np.zeros((B, L), dtype=np.complex128), emmodel=take(I4[:, 0]), ms_kind=take(I4[:, 1]),
        ms_p0=take(F8[:, 3]), ms_p1=take(F8[:, 4]), interface=take(I4[:, 2]),
        substrate_kind=np.zeros(B, dtype=np.int32), substrate_eps=np.zeros(B, dtype=np.complex128),
        substrate_temperature=np.zeros(B), theta=theta0,
        theta_inc=(np.atleast_1d(np.asarray(sensor0.theta_inc, dtype=float)).copy() if mode == MODE_ACTIVE
                   else np.zeros(0)),
        phi=float(phi[0]), dense_snow_correction=take(I4[:, 3]), inclusion=take(INCL),
        interface_params=None if IPAR is None else take(IPAR),
    )

    # permittivities: whole blocks per medium for constants and the default ice models, layer by layer otherwise
    liquid_water = take(F8[:, 5])
    for medium, target in ((0, batch.eps_bg), (1, batch.eps_sc)):
        kind = take(I4[:, 4 + medium])
        target[...] = take(CONST[:, medium])
        ice = kind == _PERM_ICE
        wet = kind == _PERM_WETICE
        if ice.any() or wet.any():
            fb = np.broadcast_to(freq[:, None], kind.shape)
            if ice.any():
                target[ice] = ice_permittivity_maetzler06(fb[ice], temperature[ice])
            if wet.any():
                target[wet] = wetice_permittivity_bohren83(fb[wet], temperature[wet], liquid_water[wet])
        call = np.argwhere(kind == _PERM_CALL)
        for b, l in call:
            target[b, l] = complex(unique[sp_index[b]].layers[l].permittivity(medium, float(freq[b])))

    # substrate and atmosphere (frequency dependent): only the simulations that have one
    for b, (sensor, sp) in enumerate(simulations):
        substrate = sp.substrate
        if substrate is not None:
            kind, eps, temp, par = _substrate(substrate, float(freq[b]), mode)
            batch.substrate_kind[b], batch.substrate_eps[b], batch.substrate_temperature[b] = kind, eps, temp
            batch.substrate_params[b] = par
        # smrt/core/model.py:615: snowpack.atmosphere or the (deprecated) atmosphere argument of the run
        atmos = getattr(sp, "atmosphere", None) or (atmospheres[b] if atmospheres is not None else None)
        if atmos is not None and mode == MODE_PASSIVE:  # active mode ignores it (rtsolver_utils.py:109-147, 302-305)
            batch.atmosphere[b] = _atmosphere(atmos, float(freq[b]), mode)
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


def water_permittivity_maetzler87(frequency, temperature):
    """Pure water, Mätzler & Wegmüller 1987 — reference ``smrt/permittivity/water.py:12-47`` (vectorised)."""
    frequency = np.asarray(frequency, dtype=float)
    temperature = np.asarray(temperature, dtype=float)
    if np.any(temperature < 273.15):
        raise SMRTError("The water temperature must be higher or equal to 273.15K")
    fghz = frequency / 1e9
    theta = 1 - 300.0 / temperature
    e0 = 77.66 - 103.3 * theta
    e1 = 0.0671 * e0
    f1 = 20.2 + 146.4 * theta + 316 * theta**2
    e2 = 3.52 + 7.52 * theta
    f2 = 39.8 * f1
    return e2 + (e1 - e2) / (1 - 1j * fghz / f2) + (e0 - e1) / (1 - 1j * fghz / f1)


def wetice_permittivity_bohren83(frequency, temperature, liquid_water):
    """Wet ice particles: Maxwell Garnett mixing with water as the background and ice as the inclusions — reference
    ``smrt/permittivity/wetice.py:13-41`` with ``maxwell_garnett_for_spheres``
    (``generic_mixing_formula.py:360-380``); dry ice where liquid_water <= 0 (vectorised)."""
    frequency, temperature, liquid_water = np.broadcast_arrays(np.asarray(frequency, dtype=float),
                                                              np.asarray(temperature, dtype=float),
                                                              np.asarray(liquid_water, dtype=float))
    eps = np.asarray(ice_permittivity_maetzler06(frequency, temperature), dtype=np.complex128).copy()
    wet = liquid_water > 0.0
    if np.any(wet):
        e0 = water_permittivity_maetzler87(frequency[wet], temperature[wet])
        ei = eps[wet]
        cplus = ei + 2 * e0
        cminus = (ei - e0) * (1 - liquid_water[wet])
        eps[wet] = (cplus + 2 * cminus) / (cplus - cminus) * e0
    return eps


def snow_frac_volumes(density, volumetric_liquid_water=None, liquid_water=None):
    """(frac_volume of ice + water, liquid_water = water / (ice + water)) of snow layers — reference
    ``SnowLayer.compute_frac_volumes``, ``smrt/inputs/make_medium.py:390-434`` (vectorised)."""
    density = np.asarray(density, dtype=float)
    rho_ice, rho_water = 916.7, 1000.0
    if volumetric_liquid_water is not None:
        if liquid_water is not None:
            raise SMRTError("Setting both liquid_water and volumetric_liquid_water is ambiguous")
        vlw = np.asarray(volumetric_liquid_water, dtype=float)
        frac_volume = (density - (rho_water - rho_ice) * vlw) / rho_ice
        liquid_water = vlw / frac_volume
    else:
        liquid_water = np.zeros_like(density) if liquid_water is None else np.asarray(liquid_water, dtype=float)
        frac_volume = density / (rho_ice * (1 - liquid_water) + rho_water * liquid_water)
    if not (np.all(frac_volume >= 0) and np.all(frac_volume <= 1.01)):
        raise SMRTError("the frac_volume of ice+water in snow must be between 0 and 1")
    if not (np.all(liquid_water >= 0) and np.all(liquid_water <= 1)):
        raise SMRTError("liquid_water must be between 0 and 1")
    return np.minimum(frac_volume, 1.0), np.broadcast_to(liquid_water, frac_volume.shape)


def pack_snow_ensemble(frequency, thickness, density, temperature, *, microstructure="exponential",
                       corr_length=None, radius=None, stickiness=None, emmodel="iba", mode="P", theta_deg=55.0,
                       theta_inc_deg=None, phi_deg=180.0, volumetric_liquid_water=None,
                       liquid_water=None) -> ProblemBatch:
    """Dry-snow ensemble given directly as arrays: ``(S, L)`` profiles x ``(F,)`` frequencies -> ``F*S`` problems in the
    reference's simulation order (frequency outermost, snowpack innermost; ``smrt/core/model.py:485-502``).

    Equivalent to ``make_snowpack(thickness[s], microstructure, density=density[s], temperature=temperature[s], ...)``
    for every member (``smrt/inputs/make_medium.py:158-232``: frac_volume = density / 916.7, background air eps = 1,
    scatterers = pure ice, flat interfaces, no substrate) without building any Python object.

    Wet snow: ``volumetric_liquid_water`` (or ``liquid_water``) as ``(S, L)`` arrays, like the reference's
    ``make_snowpack(..., volumetric_liquid_water=...)``: the fractional volumes follow
    ``SnowLayer.compute_frac_volumes`` and the scatterers are wet ice (``wetice_permittivity_bohren83``).
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
    if volumetric_liquid_water is None and liquid_water is None:
        frac_volume = density / 916.7
        eps_sc = ice_permittivity_maetzler06(freq_b[:, None], temp_b)
    else:
        bc = lambda a: None if a is None else np.broadcast_to(np.asarray(a, dtype=float), (S, L))  # noqa: E731
        frac_volume, lw = snow_frac_volumes(density, bc(volumetric_liquid_water), bc(liquid_water))
        eps_sc = wetice_permittivity_bohren83(freq_b[:, None], temp_b, tile(lw))
    theta = np.radians(np.atleast_1d(np.asarray(theta_deg, dtype=float)))
    if mode == "A":
        theta_inc = np.radians(np.atleast_1d(np.asarray(theta_deg if theta_inc_deg is None else theta_inc_deg,
                                                        dtype=float)))
    else:
        theta_inc = np.zeros(0)
    return ProblemBatch(
        mode=MODE_PASSIVE if mode == "P" else MODE_ACTIVE,
        frequency=freq_b, nlayer=np.full(B, L, dtype=np.int32), thickness=tile(thickness), temperature=temp_b,
        frac_volume=tile(frac_volume), eps_bg=np.ones((B, L), dtype=np.complex128), eps_sc=eps_sc,
        emmodel=np.full((B, L), code, dtype=np.int32), ms_kind=np.full((B, L), kind, dtype=np.int32),
        ms_p0=tile(p0), ms_p1=tile(p1), interface=np.zeros((B, L), dtype=np.int32),
        substrate_kind=np.zeros(B, dtype=np.int32), substrate_eps=np.zeros(B, dtype=np.complex128),
        substrate_temperature=np.zeros(B), theta=theta, theta_inc=theta_inc, phi=float(np.radians(phi_deg)),
        dense_snow_correction=np.full((B, L), 1 if code in _DMRT_CODES else 0, dtype=np.int32),
    )


# ---------------------------------------------------------------------------------------------------------------------
# sea-ice ensembles (BASELINE config 5): the permittivity chain of make_ice_column("multiyear", ...) vectorised over
# the whole (S, L) block.  Every function restates the reference formula it cites; pinned against the packed inputs of
# the reference-generated fixture tests/golden/cfg5_first6.npz (tests/test_host_api.py).
# ---------------------------------------------------------------------------------------------------------------------
_FREEZING_POINT = 273.15
_PSU = 1e-3
_EPS0 = 1.0 / (4e-7 * np.pi * 299792458.0**2)  # smrt/core/globalconstants.py:32


def water_freezing_temperature(salinity):
    """TEOS-10 polynomial fit at sea-level pressure — reference ``smrt/permittivity/brine.py:176-229``."""
    c = (0.017947064327968736, -6.076099099929818, 4.883198653547851, -11.88081601230542, 13.34658511480257,
         -8.722761043208607, 2.082038908808201, -7.389420998107497, -2.110913185058476, 0.2295491578006229,
         -0.9891538123307282, -0.08987150128406496, 0.3831132432071728, 1.054318231187074, 1.065556599652796,
         -0.7997496801694032, 0.3850133554097069, -2.078616693017569, 0.8756340772729538, -2.079022768390933,
         1.596435439942262, 0.1338002171109174, 1.242891021876471)
    s_r = np.asarray(salinity, dtype=float) * 1e1
    x = np.sqrt(s_r)
    p_r = 10.1325 * 1e-4
    t = (c[0] + s_r * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * (c[5] + c[6] * x))))) +
         p_r * (c[7] + p_r * (c[8] + c[9] * p_r)) +
         s_r * p_r * (c[10] + p_r * (c[12] + p_r * (c[15] + c[21] * s_r)) + s_r * (c[13] + c[17] * p_r + c[19] * s_r) +
                      x * (c[11] + p_r * (c[14] + c[18] * p_r) + s_r * (c[16] + c[20] * p_r + c[22] * s_r))))
    return t + 273.15


def brine_volume_cox83_lepparanta88(temperature, salinity):
    """Brine volume fraction (Cox & Weeks 1983; Leppäranta & Manninen 1988 above -2 degC), porosity 0 —
    reference ``smrt/permittivity/brine.py:232-329``."""
    temperature = np.asarray(temperature, dtype=float)
    salinity = np.broadcast_to(np.asarray(salinity, dtype=float), temperature.shape)
    T = temperature - _FREEZING_POINT
    if np.any(T < -38.0):
        raise SMRTError("the brine-volume polynomials of Cox and Weeks (1983) are unphysical below -38 degC")
    rho_ice = 916.7 / 1e3 - 1.403e-4 * T
    cold = T < -22.9
    warm = T >= -2.0
    a = np.where(warm[..., None], (-4.1221e-2, -1.8407e1, 5.8402e-1, 2.1454e-1),
                 np.where(cold[..., None], (9.899e3, 1.309e3, 5.527e1, 7.160e-1), (-4.732, -2.245e1, -6.397e-1, -1.074e-2)))
    b = np.where(warm[..., None], (9.0312e-2, -1.6111e-2, 1.2291e-4, 1.3603e-4),
                 np.where(cold[..., None], (8.547, 1.089, 4.518e-2, 5.819e-4), (8.903e-2, -1.763e-2, -5.33e-4, -8.801e-6)))
    # np.polyval([a3, a2, a1, a0], T): Horner from the highest power
    F1 = ((a[..., 3] * T + a[..., 2]) * T + a[..., 1]) * T + a[..., 0]
    F2 = ((b[..., 3] * T + b[..., 2]) * T + b[..., 1]) * T + b[..., 0]
    bulk_density = rho_ice * F1 / (F1 - rho_ice * salinity * _PSU**-1 * F2) * 1e3
    Vb = salinity / _PSU * bulk_density * 1e-3 / F1
    tf = water_freezing_temperature(salinity)
    Vb = np.where((Vb > 1.0) & (np.abs(temperature - tf) < 0.1), 1.0, Vb)
    Vb = np.where(temperature > tf, 1.0, Vb)
    if np.any((Vb < 0) | (Vb > 1)):
        raise SMRTError("the brine-volume polynomials give a fraction outside [0, 1] for these temperatures / salinities")
    return Vb


def brine_permittivity_stogryn85(frequency, temperature):
    """Brine permittivity (Stogryn & Desargant 1985) — reference ``smrt/permittivity/saline_water.py:131-155`` with
    ``brine.py:13-46, 146-175`` (conductivity, relaxation time, static and high-frequency limits)."""
    frequency = np.asarray(frequency, dtype=float)
    tempC = np.asarray(temperature, dtype=float) - _FREEZING_POINT
    eps_static = (939.66 - 19.068 * tempC) / (10.737 - tempC)
    tau = 0.1099 + 0.13603e-2 * tempC + 0.20894e-3 * tempC**2 + 0.28167e-5 * tempC**3
    sigma = np.where(tempC >= -22.9, -tempC * np.exp(0.5193 + 0.08755 * tempC), -tempC * np.exp(1.0334 + 0.1100 * tempC))
    eps_inf = (82.79 + 8.19 * tempC**2) / (15.68 + tempC**2)
    return (eps_inf + (eps_static - eps_inf) / (1.0 - tau * frequency / 1e9 * 1j)
            + sigma / (2.0 * np.pi * _EPS0 * frequency) * 1j)


def polder_van_santen_spheres(frac_volume, e0, eps):
    """Polder - van Santen mixing for spherical inclusions — ``smrt/permittivity/generic_mixing_formula.py:118-141``."""
    b_quad = eps - 2 * e0 - 3.0 * frac_volume * (eps - e0)
    c_quad = -eps * e0
    return (-b_quad + np.sqrt(b_quad**2 - 8.0 * c_quad)) / 4.0


def saline_ice_permittivity_pvs_mixing(frequency, temperature, brine_volume_fraction):
    """Pure ice + spherical brine pockets — reference ``smrt/permittivity/saline_ice.py:76-127`` (default models)."""
    return polder_van_santen_spheres(brine_volume_fraction, ice_permittivity_maetzler06(frequency, temperature),
                                     brine_permittivity_stogryn85(frequency, temperature))


def seawater_permittivity_klein76(frequency, temperature, salinity):
    """Klein & Swift (1976) — reference ``smrt/permittivity/saline_water.py:24-84``."""
    frequency = np.asarray(frequency, dtype=float)
    tempC = np.asarray(temperature, dtype=float) - _FREEZING_POINT
    Sppt = np.asarray(salinity, dtype=float) / _PSU
    tempF = -(0.0575 * Sppt - 1.710523e-3 * Sppt**1.5 + 2.154996e-4 * Sppt**2)
    if np.any(tempC < tempF - 0.1):
        raise SMRTError("The water temperature must be higher than the freezing point at the given salinity")
    omega = 2 * np.pi * frequency
    eps_inf = 4.9
    eps_s_T = 87.134 - 1.949e-1 * tempC - 1.276e-2 * tempC**2 + 2.491e-4 * tempC**3
    a_ST = 1.0 + 1.613e-5 * Sppt * tempC - 3.656e-3 * Sppt + 3.210e-5 * Sppt**2 - 4.232e-7 * Sppt**3
    eps_static = eps_s_T * a_ST
    tau_T0 = 1.768e-11 - 6.086e-13 * tempC + 1.104e-14 * tempC**2 - 8.111e-17 * tempC**3
    b_ST = 1.0 + 2.282e-5 * Sppt * tempC - 7.638e-4 * Sppt - 7.760e-6 * Sppt**2 + 1.105e-8 * Sppt**3
    tau = tau_T0 * b_ST
    delta = 25 - tempC
    beta = (2.0333e-2 + 1.266e-4 * delta + 2.464e-6 * delta**2
            - Sppt * (1.849e-5 - 2.551e-7 * delta + 2.551e-8 * delta**2))
    sigma_25S = Sppt * (0.182521 - 1.46192e-3 * Sppt + 2.09324e-5 * Sppt**2 - 1.28205e-7 * Sppt**3)
    sigma = sigma_25S * np.exp(-delta * beta)
    return eps_inf + (eps_static - eps_inf) / (1 - 1j * omega * tau) + 1j * sigma / (omega * _EPS0)


def pack_sea_ice_ensemble(frequency, thickness, temperature, salinity, porosity, corr_length, *, theta_deg=40.0,
                          water_substrate=True, water_temperature=_FREEZING_POINT - 1.8,
                          water_salinity=0.032, ice_type="multiyear", brine_inclusion_shape="spheres") -> ProblemBatch:
    """Sea-ice ensemble given directly as arrays: ``(S, L)`` profiles x ``(F,)`` frequencies -> ``F*S`` passive
    problems in the reference's simulation order (frequency outermost).

    ``ice_type="firstyear"`` (``smrt/inputs/make_medium.py:660-681``): background = pure ice (Mätzler 2006),
    scatterers = Stogryn-85 brine pockets at the Cox-Weeks brine volume, shaped as ``brine_inclusion_shape``
    ("spheres", "random_needles", or a dict of mixing ratios); ``porosity`` must be 0 there.  ``"multiyear"``:

    Equivalent, member by member, to ``make_ice_column("multiyear", thickness=..., temperature=..., salinity=...,
    porosity=..., microstructure_model="exponential", corr_length=..., brine_inclusion_shape="spheres",
    add_water_substrate="ocean")`` (``smrt/inputs/make_medium.py:437-571, 573-753, 962-990``): background = saline ice
    (Polder - van Santen mix of pure ice and Stogryn-85 brine at the Cox-Weeks brine volume), scatterers = air bubbles
    (fractional volume = porosity), flat interfaces, semi-infinite sea water (Klein & Swift) below.  IBA + DORT.
    """
    thickness = np.atleast_2d(np.asarray(thickness, dtype=float))
    S, L = thickness.shape
    freqs = np.atleast_1d(np.asarray(frequency, dtype=float))
    F = len(freqs)
    B = F * S

    def full(a):
        return np.broadcast_to(np.asarray(a, dtype=float), (S, L))

    def tile(a, dt=np.float64):
        return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=dt), (S, L))[None].repeat(F, axis=0)
                                    .reshape(B, L))

    temperature, salinity, porosity = full(temperature), full(salinity), full(porosity)
    if np.any(salinity >= 1):
        raise SMRTError("salinity must be given in kg/kg (multiply PSU by 1e-3)")
    freq_b = np.repeat(freqs, S)
    temp_b = tile(temperature)
    vb = tile(brine_volume_cox83_lepparanta88(temperature, salinity))
    inclusion = None
    if ice_type == "firstyear":
        if np.any(porosity != 0):
            raise SMRTError("first-year ice has no air bubbles in the reference's make_ice_column: porosity must be 0")
        eps_bg = np.asarray(ice_permittivity_maetzler06(freq_b[:, None], temp_b), dtype=np.complex128) + 0 * vb
        eps_sc = np.asarray(brine_permittivity_stogryn85(freq_b[:, None], temp_b), dtype=np.complex128) + 0 * vb
        frac = vb
        ws, wn = _shape_weights(brine_inclusion_shape)
        inclusion = np.tile(np.array((ws, wn) + SPHERICAL_INCLUSIONS[2:]), (B, L, 1))
    elif ice_type == "multiyear":
        if brine_inclusion_shape not in (None, "spheres"):
            raise SMRTError("only spherical brine pockets are implemented for the multi-year ensemble")
        eps_bg = saline_ice_permittivity_pvs_mixing(freq_b[:, None], temp_b, vb)
        eps_sc = np.ones((B, L), dtype=np.complex128)
        frac = tile(porosity)
    else:
        raise SMRTError("ice_type must be 'firstyear' or 'multiyear'")
    if water_substrate:
        sub_kind = np.full(B, SUB_FLAT, dtype=np.int32)
        sub_eps = np.asarray(seawater_permittivity_klein76(freq_b, water_temperature, water_salinity), dtype=np.complex128)
        sub_T = np.full(B, float(water_temperature))
    else:
        sub_kind, sub_eps, sub_T = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.complex128), np.zeros(B)
    return ProblemBatch(
        mode=MODE_PASSIVE, frequency=freq_b, nlayer=np.full(B, L, dtype=np.int32), thickness=tile(thickness),
        temperature=temp_b, frac_volume=frac, eps_bg=np.asarray(eps_bg, dtype=np.complex128),
        eps_sc=eps_sc, emmodel=np.full((B, L), emmodel_code("iba"), dtype=np.int32),
        ms_kind=np.full((B, L), MS_EXPONENTIAL, dtype=np.int32), ms_p0=tile(corr_length), ms_p1=np.zeros((B, L)),
        interface=np.zeros((B, L), dtype=np.int32), substrate_kind=sub_kind, substrate_eps=sub_eps,
        substrate_temperature=sub_T, theta=np.radians(np.atleast_1d(np.asarray(theta_deg, dtype=float))),
        theta_inc=np.zeros(0), phi=0.0, dense_snow_correction=np.zeros((B, L), dtype=np.int32), inclusion=inclusion,
    )
