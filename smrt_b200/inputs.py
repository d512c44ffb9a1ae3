"""Light-weight input objects with the attribute surface the packer reads.

The reference's builders (``smrt/inputs/make_medium.py``, ``smrt/inputs/sensor_list.py``, ``smrt/core/sensor.py``) are
OUT OF SCOPE to re-implement: when SMRT is installed its own ``Snowpack`` / ``Sensor`` objects
are consumed as they are.  On a box without SMRT (the GPU test box) these small stand-ins give the parity tests and the
examples the same vocabulary — ``make_snowpack(thickness, "exponential", density=..., temperature=..., corr_length=...)``,
``sensor_list.amsre("37V")``, ``sensor_list.active(13e9, 40)`` — for the media the device path covers: dry snow with
exponential or sticky-hard-sphere microstructure, flat / transparent interfaces, an optional flat substrate.
"""

from __future__ import annotations

import copy
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np

from .error import SMRTError
from .pack import ice_permittivity_maetzler06

DENSITY_OF_ICE = 916.7
FREEZING_POINT = 273.15
C_SPEED = 299792458.0


# ------------------------------------------------------------------------------------------------------------- sensors
class Sensor:
    """Sensor description: same attributes as reference ``smrt/core/sensor.py:235-366``."""

    def __init__(self, frequency, theta_inc_deg=None, theta_deg=None, phi_deg=None, polarization_inc=None,
                 polarization=None, channel_map=None, name=None):
        if frequency is None:
            raise SMRTError("Either frequency or wavelength is required")
        self.frequency = np.asarray(frequency).squeeze() if isinstance(frequency, (list, tuple, np.ndarray)) else frequency
        if isinstance(self.frequency, np.ndarray) and self.frequency.ndim == 0:
            self.frequency = float(self.frequency)
        self.channel_map = channel_map or {}
        self.name = name
        self.polarization = list(polarization) if isinstance(polarization, str) else polarization
        self.polarization_inc = list(polarization_inc) if isinstance(polarization_inc, str) else polarization_inc
        if theta_deg is None:
            raise SMRTError("Sensor requires the argument 'theta_deg' to be set")
        self.theta_deg = np.atleast_1d(theta_deg).flatten().astype(float)
        if len(np.unique(self.theta_deg)) != len(self.theta_deg):
            raise SMRTError("Zenith angle theta has duplicated values which is invalid.")
        self.theta = np.radians(self.theta_deg)
        if phi_deg is not None:
            self.phi_deg = np.atleast_1d(phi_deg).flatten().astype(float)
            self.phi = np.radians(self.phi_deg)
        else:
            self.phi = 0.0
        if theta_inc_deg is None:
            self.theta_inc_deg = None
            self.theta_inc = None
        else:
            self.theta_inc_deg = np.atleast_1d(theta_inc_deg).flatten().astype(float)
            if len(np.unique(self.theta_inc_deg)) != len(self.theta_inc_deg):
                raise SMRTError("Zenith angle theta_inc has duplicated values which is invalid.")
            self.theta_inc = np.radians(self.theta_inc_deg)

    @property
    def mode(self):
        return "P" if self.theta_inc is None else "A"

    @property
    def wavelength(self):
        return C_SPEED / self.frequency

    def configurations(self):
        for axis in ("frequency", "theta_inc", "polarization_inc", "theta", "phi", "polarization"):
            values = np.atleast_1d(getattr(self, axis))
            if len(values) > 1:
                yield axis, values

    def iterate(self, axis):
        for v in getattr(self, axis):
            sub = copy.copy(self)
            setattr(sub, axis, v)
            yield sub


def passive(frequency, theta, polarization=None, channel_map=None, name=None):
    """Generic radiometer (reference ``smrt/core/sensor.py:24-75``)."""
    return Sensor(frequency, None, theta, None, None, ["V", "H"] if polarization is None else list(polarization),
                  channel_map=channel_map, name=name)


def active(frequency, theta_inc, theta=None, phi=None, polarization_inc=None, polarization=None, channel_map=None,
           name=None):
    """Generic radar, backscatter by default (reference ``smrt/core/sensor.py:119-199``)."""
    return Sensor(frequency, theta_inc_deg=theta_inc, theta_deg=theta_inc if theta is None else theta,
                  phi_deg=180.0 if phi is None else phi,
                  polarization_inc=["V", "H"] if polarization_inc is None else list(polarization_inc),
                  polarization=["V", "H"] if polarization is None else list(polarization),
                  channel_map=channel_map, name=name)


_AMSRE = {"06": 6.925e9, "10": 10.65e9, "19": 18.7e9, "23": 23.8e9, "37": 36.5e9, "89": 89e9}


def amsre(channel=None, theta=55):
    """AMSR-E channels (reference ``smrt/inputs/sensor_list.py:22-66, 149-203``)."""
    channel_map = {f + p: dict(frequency=_AMSRE[f], polarization=p, theta=theta) for f in _AMSRE for p in "HV"}
    if channel is not None:
        wanted = []
        for ch in [channel] if isinstance(channel, str) else list(channel):
            wanted += [ch + "H", ch + "V"] if ch[-1] not in "HV" else [ch]
        for ch in wanted:
            if "18" in ch:
                channel_map[ch] = channel_map.pop("19" + ch[-1])
            if "36" in ch:
                channel_map[ch] = channel_map.pop("37" + ch[-1])
        try:
            channel_map = {ch: channel_map[ch] for ch in wanted}
        except KeyError:
            raise SMRTError(f"AMSR-E channel not recognized. Expected one of: {', '.join(_AMSRE)}")
    conf = {}
    for k in ("frequency", "polarization", "theta"):
        x = np.unique([channel_map[ch][k] for ch in channel_map])
        conf[k] = x[0] if len(x) == 1 else x
    return passive(conf["frequency"], conf["theta"], conf["polarization"], channel_map=channel_map, name="amsre")


sensor_list = SimpleNamespace(passive=passive, active=active, amsre=amsre)


# ---------------------------------------------------------------------------------------------- media and interfaces
class Flat:
    """Flat (Fresnel) interface — reference ``smrt/interface/flat.py:11-75``."""


class Transparent:
    """Transparent interface — reference ``smrt/interface/transparent.py:7-49``."""


class Exponential:
    def __init__(self, frac_volume, corr_length):
        self.frac_volume, self.corr_length = float(frac_volume), float(corr_length)


class StickyHardSpheres:
    def __init__(self, frac_volume, radius, stickiness=1000):
        self.frac_volume, self.radius, self.stickiness = float(frac_volume), float(radius), float(stickiness)


class Homogeneous:
    def __init__(self, frac_volume):
        self.frac_volume = float(frac_volume)


class IndependentSphere:
    def __init__(self, frac_volume, radius):
        self.frac_volume, self.radius = float(frac_volume), float(radius)


class TeubnerStrey:
    def __init__(self, frac_volume, corr_length, repeat_distance):
        self.frac_volume, self.corr_length = float(frac_volume), float(corr_length)
        self.repeat_distance = float(repeat_distance)


class UnifiedScaledExponential:
    """reference ``smrt/microstructure_model/unified_scaled_exponential.py:19-26``"""

    def __init__(self, frac_volume, porod_length, polydispersity):
        self.frac_volume, self.porod_length = float(frac_volume), float(porod_length)
        self.polydispersity = float(polydispersity)
        self.corr_length = self.polydispersity * self.porod_length


class UnifiedTeubnerStrey:
    """reference ``smrt/microstructure_model/unified_teubner_strey.py:18-36``"""

    def __init__(self, frac_volume, porod_length, polydispersity):
        self.frac_volume, self.porod_length = float(frac_volume), float(porod_length)
        self.polydispersity = float(polydispersity)
        K32 = self.polydispersity ** (3 / 2)
        if self.polydispersity >= 1:
            b = self.porod_length * K32
            delta = np.sqrt(1 - 1 / K32)
            self.zeta1, self.zeta2 = b * (1 - delta), b * (1 + delta)
        else:
            self.zeta1 = self.porod_length
            self.zeta2 = self.porod_length * np.sqrt(1 / (1 / K32 - 1))


class UnifiedStickyHardSpheres:
    """reference ``smrt/microstructure_model/unified_sticky_hard_spheres.py:18-31``"""

    def __init__(self, frac_volume, porod_length, polydispersity):
        self.frac_volume, self.porod_length = float(frac_volume), float(porod_length)
        self.polydispersity = float(polydispersity)
        self.radius = 3 / 4 * self.porod_length / (1 - self.frac_volume)
        K_32 = self.polydispersity ** (-3 / 2)
        self.t = (1 + 2 * self.frac_volume - 3 / (8 * np.sqrt(2)) * K_32) / (self.frac_volume * (1.0 - self.frac_volume))


_UNIFIED = ("porod_length", "polydispersity")
_MICROSTRUCTURES = {"exponential": (Exponential, ("corr_length",)),
                    "sticky_hard_spheres": (StickyHardSpheres, ("radius", "stickiness")),
                    "independent_sphere": (IndependentSphere, ("radius",)),
                    "teubner_strey": (TeubnerStrey, ("corr_length", "repeat_distance")),
                    "unified_scaled_exponential": (UnifiedScaledExponential, _UNIFIED),
                    "unified_teubner_strey": (UnifiedTeubnerStrey, _UNIFIED),
                    "unified_sticky_hard_spheres": (UnifiedStickyHardSpheres, _UNIFIED),
                    "homogeneous": (Homogeneous, ())}


def depolarization_factors_spheroids(length_ratio=None):
    """Depolarisation factors (x, y, z) of spheroids of a given horizontal / vertical length ratio (Löwe et al. 2013
    eq. 4, Mätzler 1996 eq. 7) — reference ``smrt/permittivity/depolarization_factors.py:9-46``; 1/3 each for spheres."""
    if length_ratio is None:
        length_ratio = 1.0
    if length_ratio == 1:
        q = 1.0 / 3.0
    elif length_ratio > 1:
        chi_b = np.sqrt(1.0 - 1.0 / (length_ratio**2.0))
        ln_term = np.log((1.0 + chi_b) / (1.0 - chi_b))
        q = 0.5 * (1.0 + (1.0 / (length_ratio**2.0 - 1.0)) * (1.0 - (1.0 / (2.0 * chi_b)) * ln_term))
    else:
        chi_a = np.sqrt(1.0 / length_ratio**2.0 - 1.0)
        q = 0.5 * (1.0 + (1.0 / (length_ratio**2.0 - 1.0)) * (1.0 - (1.0 / chi_a) * np.arctan(chi_a)))
    return np.array([q, q, (1.0 - 2.0 * q)])


class Layer:
    """Dry-snow layer: ice scatterers (Mätzler 2006 permittivity) in air — the subset of reference
    ``smrt/core/layer.py:35-156`` + ``smrt/inputs/make_medium.py:235-315`` read by the packer."""

    def __init__(self, thickness, microstructure, temperature=FREEZING_POINT, density=None, permittivity_model=None,
                 emmodel=None, emmodel_options=None):
        if temperature < 0:
            raise SMRTError("Layer temperature is negative. Temperature must be in Kelvin")
        self.thickness = float(thickness)
        self.temperature = float(temperature)
        self.density = density
        self.microstructure = microstructure
        self.permittivity_model = permittivity_model  # (background, scatterers): numbers or f(frequency, temperature)
        self.inclusion_shape = None
        self.medium = "snow"  # make_medium.py:361-363 (SnowLayer): the key of a dict of emmodels
        self.emmodel = emmodel
        self.emmodel_options = emmodel_options

    @property
    def frac_volume(self):
        return self.microstructure.frac_volume

    def permittivity(self, i, frequency):
        if self.permittivity_model is None:
            return 1.0 if i == 0 else complex(ice_permittivity_maetzler06(frequency, self.temperature))
        model = self.permittivity_model[i]
        return model(frequency, self.temperature) if callable(model) else model


class FlatSubstrate:
    """Flat half-space under the snowpack — reference ``smrt/substrate/flat.py:15-17``."""

    def __init__(self, temperature=None, permittivity_model=None):
        self.temperature = temperature
        self.permittivity_model = permittivity_model

    def permittivity(self, frequency):
        if self.permittivity_model is None:
            raise SMRTError("No permittivity_model have been given to the substrate.")
        pm = self.permittivity_model
        return pm(frequency, self.temperature) if callable(pm) else pm


FlatSubstrate.__name__ = "Flat"  # the packer recognises substrates and interfaces by class name, like the reference's


class SoilWegmuller(FlatSubstrate):
    """Data of the rough soil of Wegmüller & Mätzler 1999 — reference ``smrt/substrate/soil_wegmuller.py:20-22``
    (the reflectivity model itself runs on the device)."""

    def __init__(self, temperature=None, permittivity_model=None, roughness_rms=None):
        super().__init__(temperature, permittivity_model)
        if roughness_rms is None:
            raise SMRTError("Parameter roughness_rms must be specified")
        self.roughness_rms = roughness_rms


class ChoudhuryReflectivity(SoilWegmuller):
    """Rough reflectivity of Choudhury et al. 1979 — reference ``smrt/substrate/rough_choudhury79.py:19-21``."""


class SoilQNH(FlatSubstrate):
    """QNH soil parameters — reference ``smrt/substrate/soil_qnh.py:22-24``."""

    def __init__(self, temperature=None, permittivity_model=None, H=None, Q=0.0, N=0.0, Nv=np.nan, Nh=np.nan):
        super().__init__(temperature, permittivity_model)
        if H is None:
            raise SMRTError("Parameter H must be specified")
        self.H, self.Q, self.N, self.Nv, self.Nh = H, Q, N, Nv, Nh


class IEM_Fung92(FlatSubstrate):
    """Moderately rough surface, backscatter by the IEM of Fung et al. 1992 — reference ``smrt/substrate/iem_fung92.py``
    (``smrt/interface/iem_fung92.py:44-67``)."""

    def __init__(self, temperature=None, permittivity_model=None, roughness_rms=None, corr_length=None,
                 autocorrelation_function="exponential", warning_handling="print", series_truncation=10):
        super().__init__(temperature, permittivity_model)
        if roughness_rms is None or corr_length is None:
            raise SMRTError("Parameters roughness_rms and corr_length must be specified")
        self.roughness_rms, self.corr_length = roughness_rms, corr_length
        self.autocorrelation_function = autocorrelation_function
        self.warning_handling, self.series_truncation = warning_handling, series_truncation


class IEM_Fung92_Briogoni10(IEM_Fung92):
    """The same with the Fresnel coefficients taken at normal incidence for ks kl > sqrt(eps_r) — reference
    ``smrt/interface/iem_fung92_brogioni10.py:31-54``."""


class Reflector:
    """Prescribed specular reflection (scalar, or dict keyed by polarisation and / or frequency) — reference
    ``smrt/substrate/reflector.py:51-111``."""

    def __init__(self, temperature=None, specular_reflection=None):
        self.temperature = temperature
        self.specular_reflection = specular_reflection


class ReflectorBackscatter:
    """Prescribed specular reflection (scalar or {"V", "H"} dict) and backscattering coefficient ({"VV", "HH"} dict,
    linear) — reference ``smrt/substrate/reflector_backscatter.py:54-135``; usable in active mode."""

    def __init__(self, temperature=None, specular_reflection=None, backscattering_coefficient=None):
        self.temperature = temperature
        self.specular_reflection = specular_reflection
        self.backscattering_coefficient = backscattering_coefficient


def make_reflector(temperature=None, specular_reflection=None, backscattering_coefficient=None):
    """``smrt.substrate.reflector.make_reflector``; with a backscattering_coefficient, the one of
    ``smrt.substrate.reflector_backscatter``"""
    if backscattering_coefficient is not None:
        return ReflectorBackscatter(temperature=temperature, specular_reflection=specular_reflection,
                                    backscattering_coefficient=backscattering_coefficient)
    return Reflector(temperature=temperature, specular_reflection=specular_reflection)


_SOILS = {"flat": FlatSubstrate, "soil_wegmuller": SoilWegmuller, "soil_qnh": SoilQNH,
          "rough_choudhury79": ChoudhuryReflectivity, "iem_fung92": IEM_Fung92,
          "iem_fung92_brogioni10": IEM_Fung92_Briogoni10}


def make_soil(substrate_model, permittivity_model=None, temperature=FREEZING_POINT, **kwargs):
    """``make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.001, temperature=265)`` as in
    reference ``smrt/inputs/make_soil.py:50-111`` with a complex number or a function of (frequency, temperature) as
    permittivity model (the named soil permittivity models of the reference are not re-implemented here)."""
    if substrate_model not in _SOILS:
        raise SMRTError(f"substrate model '{substrate_model}' is not implemented on the B200 path "
                        f"(available: {sorted(_SOILS)})")
    if permittivity_model is None or isinstance(permittivity_model, str):
        raise SMRTError("give the soil permittivity as a complex number or a function of (frequency, temperature)")
    return _SOILS[substrate_model](temperature=temperature, permittivity_model=permittivity_model, **kwargs)


class SimpleIsotropicAtmosphere:
    """Isotropic atmosphere given by constants or frequency-keyed dicts — reference
    ``smrt/atmosphere/simple_isotropic_atmosphere.py:49-53``."""

    def __init__(self, tb_down=0.0, tb_up=0.0, transmittance=1.0):
        self.constant_tbdown = tb_down
        self.constant_tbup = tb_up
        self.constant_trans = transmittance

    def __add__(self, other):  # atmosphere + snowpack, reference smrt/core/atmosphere.py / snowpack.py:263-283
        if isinstance(other, Snowpack):
            if other.atmosphere is not None:
                raise SMRTError("The snowpack has already an atmosphere")
            return Snowpack(other.layers, other.interfaces, other.substrate, self)
        raise SMRTError("Attempt to add an incorrect object to an atmosphere. Only adding an atmosphere and a "
                        "snowpack (in that order) is a valid operation.")


def make_atmosphere(atmosphere_model="simple_isotropic_atmosphere", **kwargs):
    if atmosphere_model not in ("simple_isotropic_atmosphere", "simple_isotopic_atmosphere", SimpleIsotropicAtmosphere):
        raise SMRTError(f"atmosphere model '{atmosphere_model}' is not implemented on the B200 path")
    return SimpleIsotropicAtmosphere(**kwargs)


class Snowpack:
    """Container with the attributes of reference ``smrt/core/snowpack.py:37-46``."""

    def __init__(self, layers=None, interfaces=None, substrate=None, atmosphere=None):
        self.layers = layers or []
        self.interfaces = interfaces or []
        self.substrate = substrate
        self.atmosphere = atmosphere

    @property
    def nlayer(self):
        return len(self.layers)

    @property
    def layer_thicknesses(self):
        return [lay.thickness for lay in self.layers]

    def __add__(self, other):
        if isinstance(other, (FlatSubstrate, Reflector, ReflectorBackscatter)):
            return Snowpack(self.layers, self.interfaces, other, self.atmosphere)
        raise SMRTError("only `snowpack + substrate` and `atmosphere + snowpack` are supported")


def _get(x, i):
    if isinstance(x, (str, type)) or x is None or np.isscalar(x):
        return x
    return x[i]


def make_interface(interface, **kwargs):
    """``smrt.core.interface.make_interface``: a name, a class or an instance; the IEM interfaces take their parameters
    as keywords (``make_interface("iem_fung92", roughness_rms=0.004, corr_length=0.05)``; the same stand-in classes serve
    as interfaces and as substrates, like the reference's ``substrate_from_interface``)"""
    if isinstance(interface, str) and interface in ("iem_fung92", "iem_fung92_brogioni10"):
        return _SOILS[interface](**kwargs)
    if interface is None or interface in ("flat", Flat):
        return Flat()
    if interface in ("transparent", Transparent):
        return Transparent()
    if isinstance(interface, (Flat, Transparent)):
        return interface
    if type(interface).__name__ in ("Flat", "Transparent", "IEM_Fung92", "IEM_Fung92_Briogoni10"):
        return interface
    raise SMRTError(f"interface {interface!r} is not implemented on the B200 path")


def make_snowpack(thickness, microstructure_model, density, interface=None, substrate=None, temperature=FREEZING_POINT,
                  atmosphere=None, **kwargs):
    """Multi-layer dry snowpack, same call as reference ``smrt/inputs/make_medium.py:158-232``:
    ``make_snowpack([1, 10], "exponential", density=[200, 300], temperature=[240, 250], corr_length=[2e-4, 3e-4])``."""
    if not isinstance(thickness, (Sequence, np.ndarray)):
        raise SMRTError("The thickness argument must be iterable, that is, a list of numbers, numpy array or pandas "
                        "Series or DataFrame.")
    sp = Snowpack(substrate=substrate, atmosphere=atmosphere)
    for i, dz in enumerate(thickness):
        if dz <= 0:
            continue
        name = _get(microstructure_model, i)
        if name not in _MICROSTRUCTURES:
            raise SMRTError(f"microstructure model '{name}' is not implemented on the B200 path")
        cls, params = _MICROSTRUCTURES[name]
        rho = float(_get(density, i))
        frac_volume = min(rho / DENSITY_OF_ICE, 1.0)  # SnowLayer.compute_frac_volumes, make_medium.py:390-434
        try:
            ms = cls(frac_volume, **{p: _get(kwargs[p], i) for p in params if p in kwargs})
        except TypeError:
            raise SMRTError(f"microstructure '{name}' requires the parameters {params}")
        extra = {k: _get(kwargs[k], i) for k in ("emmodel", "emmodel_options", "permittivity_model") if k in kwargs}
        if "ice_permittivity_model" in kwargs:  # make_medium.py:317-358: (air background, given ice permittivity)
            ipm = kwargs["ice_permittivity_model"]
            extra["permittivity_model"] = (1.0, ipm if (callable(ipm) or np.isscalar(ipm)) else _get(ipm, i))
        sp.layers.append(Layer(dz, ms, temperature=_get(temperature, i), density=rho, **extra))
        sp.interfaces.append(make_interface(interface[i] if isinstance(interface, (list, tuple)) else interface))
    return sp
