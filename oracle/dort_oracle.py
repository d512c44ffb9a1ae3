"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy/SciPy) of the reference's IBA / DMRT-QCA-SR + DORT hot path.

This module is the *oracle* of SURVEY.md §8(c): a flat, object-free restatement of what
``make_model(em, "dort").run(sensor, snowpack)`` computes in the reference (smrt-model/smrt @ cb8de48), one
(snowpack x frequency) problem at a time, written from the reference's algorithm with file:line citations on every
function.  It follows the reference's *numerical choices* (SURVEY.md appendix "parity traps"): 2^6+1-point Romberg
for the IBA ks, 16/64-sample DFT in azimuth for the phase Fourier modes, mu-difference stream weights, rigorous
Maezawa Fresnel, LAPACK real Schur with forced upper-triangular form for the layer eigenproblem (the reference's
default ``diagonalization_method="schur_forcedtriu"``) and LAPACK band LU (``scipy.linalg.solve_banded``) for the
boundary system.

Parity pinning: ``tests/test_oracle_golden.py`` checks this module against (i) the golden literals held by the
reference's own tests and (ii) fixtures in ``tests/golden/*.npz`` produced by running the
unmodified reference in the authoring container (``oracle/gen_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  The shipped product (``smrt_b200``) never does; it fails loudly when the CUDA library is missing.

Problem description (a plain dict, all SI units; the same fields the C ABI takes, see include/smrt_dort_b200.h):

    frequency      float   Hz
    mode           "P" | "A"
    thickness      (L,)    m
    temperature    (L,)    K
    frac_volume    (L,)
    eps_bg         (L,) complex   background permittivity   (layer.permittivity(0, f))
    eps_sc         (L,) complex   scatterer permittivity    (layer.permittivity(1, f))
    emmodel        (L,) int       0 = IBA, 1 = DMRT-QCA short range, 2 = non-scattering, 3 = DMRT-QCACP short range,
                                  4 = Rayleigh (ms_p0 = radius), 5 = prescribed ks / ka / eps (eps_bg = effective
                                  permittivity, ms_p0 = ks, ms_p1 = ka), 6 = IBA original (Maetzler 1998 absorption),
                                  7 = IBA with the Maxwell-Garnett effective permittivity
    ms_kind        (L,) int       0 = exponential (ms_p0 = corr_length), 1 = sticky hard spheres (ms_p0 = radius,
                                  ms_p1 = stickiness), 2 = homogeneous, 3 = independent sphere (radius), 4 = Teubner-
                                  Strey (corr_length, repeat_distance), 5 / 6 = unified Teubner-Strey, polydispersity
                                  >= 1 / < 1 (zeta1, zeta2), 7 = sticky hard spheres with t given (radius, t)
    ms_p0, ms_p1   (L,)
    interface      (L,) int       interface ABOVE layer l: 0 = flat (Fresnel), 1 = transparent
    substrate_kind int            0 = none, 1 = flat half-space (substrate_eps, substrate_temperature), 2 = rough soil
                                  of Wegmueller & Maetzler 1999 (params[0] = roughness_rms), 3 = QNH soil (params = H, Q,
                                  Nv, Nh), 4 = reflector (params = specular reflection V, H; passive only), 5 = rough
                                  reflectivity of Choudhury 1979 (params[0] = roughness_rms)
    substrate_params (4,)         see substrate_kind (optional)
    atmosphere     (3,)           isotropic atmosphere: tb_down, tb_up (K), transmittance (optional; passive only)
    theta          (n_theta,) rad          viewing angles (passive) / incidence = viewing angles (active)
    phi            float rad      relative azimuth (active; pi for backscatter)
    inclusion      (L, 5)         optional: weights of the spheres / needles Polder - van Santen solutions, then the
                                  three depolarisation factors (default 1, 0, 1/3, 1/3, 1/3)
    dense_snow_correction (L,) int  1 = model the layer as the inverted medium when frac_volume > 0.5
    options        dict: n_max_stream, m_max, phase_normalization, prune_deep_snowpack, rayleigh_jeans_approximation,
                         error_handling
"""

from __future__ import annotations

import math

import numpy as np
import scipy.linalg
from scipy.special import roots_legendre

# reference smrt/core/globalconstants.py:25-33
DENSITY_OF_ICE = 916.7
FREEZING_POINT = 273.15
C_SPEED = 299792458.0
PLANCK_CONSTANT = 6.62607015e-34
BOLTZMANN_CONSTANT = 1.380649e-23

EM_IBA, EM_DMRT_QCA_SR, EM_NONSCATTERING, EM_DMRT_QCACP_SR, EM_RAYLEIGH, EM_PRESCRIBED_KSKAEPS = 0, 1, 2, 3, 4, 5
EM_IBA_ORIGINAL, EM_IBA_MAXWELL_GARNETT = 6, 7
EM_IBA_FAMILY = (EM_IBA, EM_IBA_ORIGINAL, EM_IBA_MAXWELL_GARNETT)
MS_EXPONENTIAL, MS_SHS, MS_HOMOGENEOUS = 0, 1, 2
MS_INDEPENDENT_SPHERE, MS_TEUBNER_STREY, MS_UNIFIED_TS_1, MS_UNIFIED_TS_2, MS_SHS_T = 3, 4, 5, 6, 7
IF_FLAT, IF_TRANSPARENT, IF_IEM_FUNG92, IF_IEM_FUNG92_BRIOGONI10 = 0, 1, 2, 3
SUB_NONE, SUB_FLAT, SUB_SOIL_WEGMULLER, SUB_SOIL_QNH, SUB_REFLECTOR, SUB_ROUGH_CHOUDHURY = 0, 1, 2, 3, 4, 5
SUB_REFLECTOR_BACKSCATTER = 6
SUB_IEM_FUNG92, SUB_IEM_FUNG92_BRIOGONI10 = 7, 8

# status codes shared with the C ABI (include/smrt_dort_b200.h)
ST_OK = 0
ST_NORMALIZATION = 1  # phase re-normalisation exceeds 30 % (dort.py:792-801)
ST_EIGEN = 2  # diagonalisation failed / not real (dort.py:1068-1085)
ST_SINGULAR = 3  # boundary system singular
ST_INPUT = 4  # invalid per-problem input
ST_SUBSTRATE = 5  # substrate model outside its validity range (substrate/rough_choudhury79.py:29-31 raises Warning)
ST_SHALLOW_WARNING = 16  # flag bit: optically shallow without substrate (dort.py:460-467)


class OracleError(Exception):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


# --------------------------------------------------------------------------------------------------------------------
# a1  permittivity of pure ice  — reference smrt/permittivity/ice.py:24-73
# --------------------------------------------------------------------------------------------------------------------
def ice_permittivity_maetzler06(frequency, temperature):
    freqGHz = frequency / 1e9
    tempC = temperature - FREEZING_POINT
    Ereal = 3.1884 + 9.1e-4 * tempC
    theta = 300.0 / temperature - 1.0
    alpha = (0.00504 + 0.0062 * theta) * np.exp(-22.1 * theta)
    B1, B2, b = 0.0207, 1.16e-11, 335.0
    deltabeta = np.exp(-9.963 + 0.0372 * tempC)
    betam = (B1 / temperature) * (np.exp(b / temperature) / ((np.exp(b / temperature) - 1) ** 2)) + B2 * freqGHz**2
    beta = betam + deltabeta
    Eimag = alpha / freqGHz + beta * freqGHz
    return Ereal + 1j * Eimag


# --------------------------------------------------------------------------------------------------------------------
# a4  FT of the autocorrelation function
# --------------------------------------------------------------------------------------------------------------------
def ft_autocorr_exponential(k, frac_volume, corr_length):
    """reference smrt/microstructure_model/exponential.py:53-58"""
    X = (k * corr_length) ** 2
    return frac_volume * (1.0 - frac_volume) * 8 * np.pi * corr_length**3 / (1.0 + X) ** 2


def ft_autocorr_shs(k, frac_volume, radius, stickiness, t=None):
    """Percus-Yevick sticky hard spheres — reference smrt/microstructure_model/sticky_hard_spheres.py:63-130; with `t`
    given: unified_sticky_hard_spheres.py:45-106 (the same structure factor, t prescribed by the polydispersity)"""
    d = 2 * radius
    phi_2 = frac_volume
    tau = stickiness
    k = np.asarray(k, dtype=float)
    shape = k.shape
    X = np.atleast_1d(k).ravel() * d / 2.0
    if t is not None:
        pass
    elif np.isfinite(tau) and phi_2 > 0.0:
        t = (6 * tau * phi_2 - 6 * phi_2 - 6 * tau
             + (36 * tau**2 * phi_2**2 - 72 * tau * phi_2**2 - 72 * tau**2 * phi_2 + 30 * phi_2**2
                + 72 * tau * phi_2 + 36 * tau**2 - 12 * phi_2) ** 0.5) / (phi_2 * (-1 + phi_2))
    else:
        t = 0
    vd = 4.0 / 3 * np.pi * (d / 2.0) ** 3
    sqrt_vint__vd = np.empty_like(X)
    zerok = np.isclose(X, 0, atol=1e-03)
    nzerok = ~zerok
    sqrt_vint__vd[nzerok] = 3 * (np.sinc(X[nzerok] / np.pi) - np.cos(X[nzerok])) / X[nzerok] ** 2
    sqrt_vint__vd[zerok] = 1
    Psi = np.sinc(X / np.pi) / sqrt_vint__vd
    Phi = 1.0
    A = phi_2 / (1 - phi_2) * ((1 - t * phi_2 + 3 * phi_2 / (1 - phi_2)) * Phi + (3 - t * (1 - phi_2)) * Psi) \
        + np.cos(X) / sqrt_vint__vd
    B = phi_2 / (1 - phi_2) * X * Phi + np.sin(X) / sqrt_vint__vd
    S = 1 / (A**2 + B**2)
    Ctilde = phi_2 * vd * S
    Ctilde[zerok] = phi_2 * vd / (phi_2 / (1 - phi_2) * ((1 - t * phi_2 + 3 * phi_2 / (1 - phi_2))
                                                         + (3 - t * (1 - phi_2))) + 1) ** 2
    return Ctilde.reshape(shape)


def shs_compute_t(frac_volume, stickiness):
    """reference smrt/microstructure_model/sticky_hard_spheres.py:132-167"""
    if stickiness == np.inf:
        return 0.0
    f = frac_volume
    a = f / 12.0
    b = -(stickiness + f / (1 - f))
    c = (1 + f / 2) / (1 - f) ** 2
    discr2 = b**2 - 4 * a * c
    if discr2 < 0:
        raise OracleError(ST_EIGEN, "negative discriminant")
    discr = math.sqrt(discr2)
    t = (-b - discr) / (2 * a)
    mhu = t * f * (1 - f)
    mhulim = 1 + 2 * f
    if mhu > mhulim:
        t = (-b + discr) / (2 * a)
        mhu = t * f * (1 - f)
    if mhu > mhulim:
        raise OracleError(ST_EIGEN, "no solution for the t parameter")
    return t


def ft_autocorr_independent_sphere(k, frac_volume, radius):
    """reference smrt/microstructure_model/independent_sphere.py:62-80"""
    k = np.asarray(k, dtype=float)
    X = radius * k
    volume_sphere = 4.0 / 3 * np.pi * radius**3
    bessel_term = np.empty_like(X)
    zero_X = np.isclose(X, 0)
    nz = np.logical_not(zero_X)
    Xn = X[nz]
    bessel_term[nz] = 9 * ((np.sin(Xn) - Xn * np.cos(Xn)) / Xn**3) ** 2
    bessel_term[zero_X] = 1.0
    return frac_volume * (1.0 - frac_volume) * volume_sphere * bessel_term


def ft_autocorr_teubner_strey(k, frac_volume, corr_length, repeat_distance):
    """reference smrt/microstructure_model/teubner_strey.py:53-62"""
    X = (k * corr_length) ** 2
    Y = (2 * np.pi * corr_length / repeat_distance) ** 2
    return frac_volume * (1.0 - frac_volume) * (8 * np.pi * corr_length**3 / ((1 + Y) ** 2 + 2 * (1 - Y) * X + X**2))


def ft_autocorr_unified_teubner_strey(k, frac_volume, zeta1, zeta2, case1):
    """reference smrt/microstructure_model/unified_teubner_strey.py:64-80 (zeta1, zeta2 of :25-36)"""
    if case1:
        ft = (4 * np.pi * zeta1 * zeta2 * (zeta1 + zeta2)) / ((1 + (zeta1 * k) ** 2) * (1 + (zeta2 * k) ** 2))
    else:
        x1 = k * zeta1
        r12 = zeta1 / zeta2
        ft = 8 * np.pi * zeta1**3 / ((1 + (x1 - r12) ** 2) * (1 + (x1 + r12) ** 2))
    return frac_volume * (1.0 - frac_volume) * ft


def ft_autocorr(k, ms_kind, f, p0, p1):
    if ms_kind == MS_EXPONENTIAL:
        return ft_autocorr_exponential(k, f, p0)
    elif ms_kind == MS_SHS:
        return ft_autocorr_shs(k, f, p0, p1)
    elif ms_kind == MS_INDEPENDENT_SPHERE:
        return ft_autocorr_independent_sphere(k, f, p0)
    elif ms_kind == MS_TEUBNER_STREY:
        return ft_autocorr_teubner_strey(k, f, p0, p1)
    elif ms_kind in (MS_UNIFIED_TS_1, MS_UNIFIED_TS_2):
        return ft_autocorr_unified_teubner_strey(k, f, p0, p1, ms_kind == MS_UNIFIED_TS_1)
    elif ms_kind == MS_SHS_T:
        return ft_autocorr_shs(k, f, p0, None, t=p1)
    raise ValueError("unknown microstructure kind")


# --------------------------------------------------------------------------------------------------------------------
# a2/a3  per-layer electromagnetic quantities
# --------------------------------------------------------------------------------------------------------------------
def polder_van_santen_spheres(frac_volume, e0, eps):
    """reference smrt/permittivity/generic_mixing_formula.py:118-141 (spherical inclusions)"""
    a_quad = 2.0
    b_quad = eps - 2 * e0 - 3.0 * frac_volume * (eps - e0)
    c_quad = -eps * e0
    return (-b_quad + np.sqrt(b_quad**2 - 4.0 * a_quad * c_quad + 0j)) / (2.0 * a_quad)


def polder_van_santen_needles(frac_volume, e0, eps):
    """reference smrt/permittivity/generic_mixing_formula.py:131-141 (randomly oriented needles)"""
    a_quad = 1.0
    b_quad = eps - e0 - 5.0 / 3.0 * frac_volume * (eps - e0)
    c_quad = -eps * (e0 + 1.0 / 3.0 * frac_volume * (eps - e0))
    return (-b_quad + np.sqrt(b_quad**2 - 4.0 * a_quad * c_quad + 0j)) / (2.0 * a_quad)


def polder_van_santen_shapes(frac_volume, e0, eps, inclusion=None):
    """reference smrt/permittivity/generic_mixing_formula.py:88-141: spheres, random needles or a weighted mixture;
    inclusion = (weight of spheres, weight of needles, depolarisation factors x, y, z) or None = spheres"""
    if inclusion is None or (inclusion[0] == 1.0 and inclusion[1] == 0.0):
        return polder_van_santen_spheres(frac_volume, e0, eps)
    if inclusion[0] == 0.0 and inclusion[1] == 1.0:
        return polder_van_santen_needles(frac_volume, e0, eps)
    return sum((float(inclusion[0]) * polder_van_santen_spheres(frac_volume, e0, eps),
                float(inclusion[1]) * polder_van_santen_needles(frac_volume, e0, eps)))


def romb65(y, dx):
    """scipy.integrate.romb for 2**6+1 samples, restated (the reference calls it at smrt/emmodel/iba.py:179)."""
    n_interv = 64
    R = {}
    h = n_interv * dx
    R[(0, 0)] = (y[0] + y[-1]) / 2.0 * h
    start = stop = step = n_interv
    for i in range(1, 7):
        start >>= 1
        R[(i, 0)] = 0.5 * (R[(i - 1, 0)] + h * y[start:stop:step].sum())
        step >>= 1
        for j in range(1, i + 1):
            prev = R[(i, j - 1)]
            R[(i, j)] = prev + (prev - R[(i - 1, j - 1)]) / ((1 << (2 * j)) - 1)
        h /= 2.0
    return R[(6, 6)]


def layer_optics(frequency, f, e0, eps, emmodel, ms_kind, p0, p1, invert_dense=False, inclusion=None):
    """Return dict(eps_eff, ks, ka, iba_coeff, f, k0, ms_kind, p0, p1) for one layer.

    IBA: reference smrt/emmodel/iba.py:85-137, 139-162, 168-226, 246-265.
    DMRT-QCA short range: reference smrt/emmodel/dmrt_qca_shortrange.py:65-112.
    Non-scattering: ks = 0, ka = 2 k0 Im sqrt(eps_bg) (reference smrt/emmodel/nonscattering.py).
    """
    e0 = complex(e0)
    eps = complex(eps)
    out = dict(ms_kind=ms_kind, p0=p0, p1=p1)
    if emmodel in EM_IBA_FAMILY:
        if f > 0.5 and invert_dense:  # dense_snow_correction="auto": iba.py:95-96, core/layer.py:186-201
            f, e0, eps = 1.0 - f, eps, e0
        k0 = 2 * np.pi * frequency / C_SPEED
        # depolarization_factors.py:9-46 (1/3 each for length_ratio 1), or the layer's own (iba.py:112-119)
        depol = np.array([1.0 / 3, 1.0 / 3, 1.0 / 3]) if inclusion is None else np.asarray(inclusion[2:5], dtype=float)
        if emmodel == EM_IBA_MAXWELL_GARNETT:
            # reference smrt/permittivity/generic_mixing_formula.py:346-358, smrt/emmodel/iba_maxwell_garnett.py:47-51
            eeff = complex(np.mean(e0 * (1 + f * (eps - e0) / (e0 + (1.0 - f) * depol * (eps - e0))),
                                   dtype=np.complex128))
            eapp = e0
        else:
            eeff = polder_van_santen_shapes(f, e0, eps, inclusion)
            eapp = eeff * (1 - depol) + e0 * depol
        y2 = (1.0 / 3.0) * np.sum(np.absolute(eapp / (eapp + (eps - e0) * depol)) ** 2.0)
        iba_coeff = (1.0 / (4.0 * np.pi)) * np.absolute(eps - e0) ** 2.0 * y2 * k0**4
        if emmodel == EM_IBA_ORIGINAL:  # reference smrt/emmodel/iba_original.py:43-44 (Maetzler 1998)
            ka = k0 * f * eps.imag * abs(y2)
        else:
            ka = 2 * k0 * np.sqrt(eeff).imag
        mu = np.linspace(1, -1, 65)
        sintheta_2 = np.sqrt((1.0 - mu) / 2.0)
        k_diff = 2.0 * k0 * sintheta_2 * abs(np.sqrt(eeff))
        ft = ft_autocorr(k_diff, ms_kind, f, p0, p1)
        y = (iba_coeff * ft).real * mu**2 + (iba_coeff * ft).real
        ks = romb65(y, mu[0] - mu[1]) / 4.0
        out.update(eps_eff=eeff, ks=float(ks), ka=float(ka), iba_coeff=float(iba_coeff), f=f, k0=k0)
    elif emmodel == EM_DMRT_QCA_SR:
        if f > 0.5 and invert_dense:  # the packer sets the flag: "auto" is DMRT's default (dmrt_qca_shortrange.py:65)
            f, e0, eps = 1.0 - f, eps, e0
        lmda = C_SPEED / frequency
        radius = p0
        t = shs_compute_t(f, p1)
        y = (eps - e0) / (eps + 2 * e0)
        fy = f * y
        k0 = (2 * math.pi / lmda) * np.sqrt(e0).real
        Eeff = e0 + 3 * fy * e0 / (1 - fy) * (1 + 2j / 3 * (k0 * radius) ** 3 * y
                                              * (1 - f) ** 4 / ((1 - fy) * (1 + 2 * f - t * f * (1 - f)) ** 2))
        Ks = 2 / (9 * f) * k0 * (k0 * radius) ** 3 * (
            np.abs(Eeff / e0 - 1) ** 2 * (1 - f) ** 4 / (1 + 2 * f - t * f * (1 - f)) ** 2)
        beta = 2 * k0 * np.sqrt(Eeff).imag
        out.update(eps_eff=complex(Eeff), ks=float(Ks), ka=float(beta - Ks), iba_coeff=0.0, f=f, k0=k0)
    elif emmodel == EM_DMRT_QCACP_SR:
        # reference smrt/emmodel/dmrt_qcacp_shortrange.py:63-125
        if f > 0.5 and invert_dense:
            f, e0, eps = 1.0 - f, eps, e0
        lmda = C_SPEED / frequency
        radius = p0
        t = shs_compute_t(f, p1)
        b = (eps - e0) * (1.0 - 4.0 * f) / 3.0 - e0
        c = -e0 * (eps - e0) * (1.0 - f) / 3.0
        discriminant = b**2 - 4 * c
        Eeff0 = 0.5 * (-b + np.sqrt(discriminant + 0j))
        if Eeff0.real < 1:
            Eeff0 = 0.5 * (-b - np.sqrt(discriminant + 0j))
        Eeff = e0 + (Eeff0 - e0) * (1 + 2.0j / 9.0 * (2 * math.pi * radius / lmda) ** 3
                                    * np.sqrt(Eeff0) * (eps - e0) / (1.0 + (eps - e0) / (3 * Eeff0) * (1.0 - f))
                                    * (1.0 - f) ** 4 / (1.0 + 2 * f - t * f * (1.0 - f)) ** 2)
        albedo = 2.0 / 9.0 * (2 * np.pi * radius / lmda) ** 3 * f / (2 * np.sqrt(Eeff).imag) * \
            abs((eps - e0) / (1 + (eps - e0) / (3 * Eeff0) * (1.0 - f))) ** 2 * \
            (1.0 - f) ** 4 / (1.0 + 2 * f - t * f * (1.0 - f)) ** 2
        beta = 2 * math.pi / lmda * 2 * np.sqrt(Eeff).imag
        ks = albedo * beta
        out.update(eps_eff=complex(Eeff), ks=float(ks), ka=float(beta - ks), iba_coeff=0.0, f=f,
                   k0=2 * math.pi / lmda)
    elif emmodel == EM_NONSCATTERING:
        # reference smrt/emmodel/nonscattering.py:19-34
        k0 = 2 * np.pi * frequency / C_SPEED
        eeff = polder_van_santen_shapes(f, e0, eps, inclusion)
        out.update(eps_eff=eeff, ks=0.0, ka=float(2 * k0 * np.sqrt(eeff).imag), iba_coeff=0.0, f=f, k0=k0)
    elif emmodel == EM_RAYLEIGH:
        # reference smrt/emmodel/rayleigh.py:21-39 (sparse medium: the effective permittivity is the background's)
        lmda = C_SPEED / frequency
        radius = p0
        k0 = 2 * np.pi / lmda
        ks = f * 2 * abs((eps - e0) / (eps + 2 * e0)) ** 2 * radius**3 * abs(e0) ** 2 * k0**4
        ka = f * k0 * eps.imag * abs(3 * e0 / (eps + 2 * e0)) ** 2 + (1 - f) * 2 * k0 * np.sqrt(e0).imag
        out.update(eps_eff=e0, ks=float(ks), ka=float(ka), iba_coeff=0.0, f=f, k0=k0)
    elif emmodel == EM_PRESCRIBED_KSKAEPS:
        # reference smrt/emmodel/prescribed_kskaeps.py:20-27: layer.ks, layer.ka, layer.effective_permittivity
        out.update(eps_eff=e0, ks=float(p0), ka=float(p1), iba_coeff=0.0, f=f, k0=2 * np.pi * frequency / C_SPEED)
    else:
        raise ValueError("unknown emmodel")
    out["emmodel"] = emmodel
    return out


# --------------------------------------------------------------------------------------------------------------------
# a7  streams — reference smrt/rtsolver/streams.py:136-223, 300-330
# --------------------------------------------------------------------------------------------------------------------
_gl_cache = {}


def gauss_legendre_quadrature(n):
    """positive Gauss-Legendre nodes of order 2n in descending order (streams.py:300-313, core/lib.py:669-684)"""
    if n not in _gl_cache:
        mu, weight = roots_legendre(2 * n)
        _gl_cache[n] = mu[-1:n - 1:-1].copy()
    return _gl_cache[n]


def compute_weight(mu):
    """streams.py:316-330 — weights from node differences, NOT the Gauss weights"""
    w = np.empty_like(mu)
    w[0] = 1 - 0.5 * (mu[0] + mu[1])
    w[-1] = abs(0.5 * (mu[-2] + mu[-1]))
    w[1:-1] = np.abs(0.5 * (mu[0:-2] - mu[2:]))
    return w


def compute_streams(n_max_stream, eps_eff):
    eps_eff = np.asarray(eps_eff, dtype=complex)
    k_most = int(np.argmax(eps_eff))  # complex argmax: lexicographic (streams.py:155)
    real_index_air = np.real(np.sqrt(eps_eff[k_most] / 1.0))
    mu_most = gauss_legendre_quadrature(n_max_stream)
    real_index = np.real(np.sqrt(eps_eff[k_most] / eps_eff))
    relsin = real_index[:, None] * np.sqrt(1 - mu_most[None, :] ** 2)
    real_reflection = relsin < 1
    mus, ws = [], []
    for layer in range(len(eps_eff)):
        m = np.sqrt(1 - relsin[layer, real_reflection[layer]] ** 2)
        if len(m) < 2:
            raise OracleError(ST_EIGEN, "fewer than 2 streams in a layer")
        mus.append(m)
        ws.append(compute_weight(m))
    relsin_air = real_index_air * np.sqrt(1 - mu_most**2)
    outmu = np.sqrt(1 - relsin_air[relsin_air < 1] ** 2)
    outweight = compute_weight(outmu)  # compute_outweight has no abs(); identical for descending mu
    return dict(mu=mus, weight=ws, n=np.array([len(m) for m in mus]), outmu=outmu, outweight=outweight,
                n_air=len(outmu))


# --------------------------------------------------------------------------------------------------------------------
# a8  Fresnel — reference smrt/core/fresnel.py:99-146, 417-474; smrt/interface/flat.py, transparent.py
# --------------------------------------------------------------------------------------------------------------------
def fresnel_coefficients(eps_1, eps_2, mu):
    eps_1 = complex(eps_1)
    eps_2 = complex(eps_2)
    n1 = np.sqrt(eps_1)
    kiz2 = n1.real**2 * (1 - mu**2)
    kyi = -np.sqrt(eps_1 - kiz2, dtype=np.complex128)
    kyt = -np.sqrt(eps_2 - kiz2, dtype=np.complex128)
    rh = (kyi - kyt) / (kyi.conjugate() + kyt)
    rv = n1.conjugate() * (eps_2 * kyi - eps_1 * kyt) / (n1 * (eps_2 * kyi.conjugate() + eps_1.conjugate() * kyt))
    mu2 = -kyt.real / np.sqrt(eps_2).real
    return rv, rh, mu2


def abs2(z):
    return z.real**2 + z.imag**2


def kirchhoff_factors(rms, freq, eps_1, eps_2, mu):
    """factors of the coherent reflection / transmission of a rough surface under the Kirchhoff approximation: reference
    smrt/interface/interface_utils.py:21-64 (the reference's k2 carries |eps_1|^2; kept)"""
    eps_1, eps_2 = complex(eps_1), complex(eps_2)
    k0 = 2 * np.pi * freq / C_SPEED
    k2 = k0**2 * abs2(eps_1)
    k_iz = k0 * np.sqrt(eps_1).real * mu
    k_sz = k0 * np.sqrt(eps_2 - (1 - mu**2) * eps_1).real
    return np.exp(-4 * k2 * rms**2 * mu**2), np.exp(-((k_sz - k_iz) ** 2) * rms**2)


def interface_R_T(kind, eps_1, eps_2, mu, npol, par=None, freq=None):
    """Diagonal COHERENT power reflection / transmission (npol, len(mu)) for medium 1 above/below medium 2; par, freq:
    parameters of a rough interface (IEM_Fung92: interface/iem_fung92.py:44-67) and the frequency."""
    mu = np.atleast_1d(mu)
    if kind == IF_TRANSPARENT:
        return np.zeros((npol, len(mu))), np.ones((npol, len(mu)))
    if kind in (IF_IEM_FUNG92, IF_IEM_FUNG92_BRIOGONI10):
        R, T = interface_R_T(IF_FLAT, eps_1, eps_2, mu, npol)
        fr, ft = kirchhoff_factors(float(par[0]), freq, eps_1, eps_2, mu)
        return R * fr, T * ft
    rv, rh, mu2 = fresnel_coefficients(eps_1, eps_2, mu)
    R = np.ones((npol, len(mu)))
    T = np.zeros((npol, len(mu)))
    R[0] = abs2(rv)
    R[1] = abs2(rh)
    T[0] = 1 - abs2(rv)
    T[1] = 1 - abs2(rh)
    if npol >= 3:
        R[2] = (rv * np.conj(rh)).real
        T[2] = mu2 / mu * ((1 + rv) * np.conj(1 + rh)).real
    return R, T


def compress_diag(mat_pol_mu, mode):
    """smrt_matrix 'diagonal4' compress with auto_reduce_npol (core/lib.py:314-366): (pol, mu) -> mu*npol + pol"""
    if mat_pol_mu.shape[0] == 3 and mode == 0:
        mat_pol_mu = mat_pol_mu[0:2]
    return np.transpose(mat_pol_mu).reshape(-1)


def substrate_R_T(problem, eps_1, mu, npol):
    """specular_reflection_matrix / emissivity_matrix of the substrate under the last layer (permittivity eps_1) on that
    layer's streams: reference smrt/substrate/flat.py:15-17 (+ core/interface.py:94-154), soil_wegmuller.py:20-81,
    soil_qnh.py:22-89, reflector.py:51-111, rough_choudhury79.py:19-79.  Only the V and H components are modified by
    the rough models, the third component keeps its Fresnel value (as in the reference)."""
    kind = int(problem["substrate_kind"])
    mu = np.atleast_1d(np.asarray(mu, dtype=float))
    par = np.asarray(problem.get("substrate_params", np.zeros(4)), dtype=float)
    if kind == SUB_REFLECTOR:
        if npol > 2:
            raise NotImplementedError("active model is not yet implemented, need modification for the third component")
        R = np.zeros((npol, len(mu)))
        R[0] = par[0]
        R[1] = par[1]
        return R, 1 - R
    if kind == SUB_REFLECTOR_BACKSCATTER:  # reflector_backscatter.py:72-88, 118-135: the third component stays zero
        R = np.zeros((npol, len(mu)))
        T = np.zeros((npol, len(mu)))
        R[0] = par[0]
        R[1] = par[1]
        T[0] = 1 - par[0]
        T[1] = 1 - par[1]
        return R, T
    R, T = interface_R_T(IF_FLAT, eps_1, problem["substrate_eps"], mu, npol)
    if kind == SUB_FLAT:
        return R, T
    freq = float(problem["frequency"])
    if kind in (SUB_IEM_FUNG92, SUB_IEM_FUNG92_BRIOGONI10):
        # coherent part under the Kirchhoff approximation, all components: interface/interface_utils.py:21-64 (the
        # reference's k2 carries |eps_1|^2; kept); the emissivity of the substrate is the coherent transmission
        # (core/interface.py:191-195)
        fr, ft = kirchhoff_factors(par[0], freq, eps_1, problem["substrate_eps"], mu)
        return R * fr, T * ft

    def adjust(rh, rv):  # in place, like the reference
        if kind in (SUB_SOIL_WEGMULLER, SUB_ROUGH_CHOUDHURY):
            ksigma = (2 * np.pi * freq * np.sqrt((1 / 2.9979e8) ** 2 * complex(eps_1)) * par[0]).real
        if kind == SUB_SOIL_WEGMULLER:
            rh *= np.exp(-(ksigma ** (np.sqrt(0.1 * mu))))
            mask = mu < np.cos(60 * np.pi / 180)
            rv[~mask] = rh[~mask] * mu[~mask] ** 0.655
            rv[mask] = rh[mask] * (0.635 - 0.0014 * (np.arccos(mu[mask]) * 180 / np.pi - 60))
        elif kind == SUB_ROUGH_CHOUDHURY:
            if ksigma > 0.1:
                raise OracleError(ST_SUBSTRATE, "Reflectivity may be outside validity range. ksigma should be << 1")
            rh *= np.exp(-4 * ksigma**2 * mu**2)
            rv *= np.exp(-4 * ksigma**2 * mu**2)
        elif kind == SUB_SOIL_QNH:
            H, Q, Nv, Nh = par
            coef_h = np.exp(-H * (mu**Nh))
            coef_v = np.exp(-H * (mu**Nv))
            trv = ((1 - Q) * rv + Q * rh) * coef_v
            rh[:] = ((1 - Q) * rh + Q * rv) * coef_h
            rv[:] = trv
        else:
            raise ValueError(f"unknown substrate kind {kind}")

    adjust(R[1], R[0])
    rh = 1 - T[1]
    rv = 1 - T[0]
    adjust(rh, rv)
    T[1] = 1 - rh
    T[0] = 1 - rv
    return R, T


def iem_fung92_backscatter(freq, eps_1, eps_2, mu, par, brogioni):
    """sigma0_vv, sigma0_hh of the IEM of Fung et al. 1992 for backscatter on the streams mu of medium 1: reference
    smrt/interface/iem_fung92.py:88-189 (Kirchhoff + complementary terms, series of `series_truncation` terms,
    exponential or Gaussian surface spectrum); iem_fung92_brogioni10.py:45-54 switches the Fresnel coefficients to
    normal incidence when ks kl > sqrt(eps_r).  par = roughness_rms, corr_length, autocorrelation (0 exponential,
    1 gaussian), series_truncation.  The validity checks only warn in the reference (warning_handling='print')."""
    rms, lc, acf, N = float(par[0]), float(par[1]), int(par[2]), int(par[3])
    eps_1, eps_2 = complex(eps_1), complex(eps_2)
    mu = np.asarray(mu, dtype=float)[None, :]
    knorm = 2 * np.pi * freq / C_SPEED * np.sqrt(eps_1).real
    kz, kx = knorm * mu, knorm * np.sqrt(1 - mu**2)
    eps_r = eps_2 / eps_1
    ks, kl = abs(knorm * rms), abs(knorm * lc)
    sq = np.sqrt(eps_r)
    at_nadir = brogioni and ((ks * kl, 0.0) > (sq.real, sq.imag))  # numpy orders complex numbers lexicographically
    Rv, Rh, _ = fresnel_coefficients(eps_1, eps_2, np.ones(1) if at_nadir else mu[0])
    fvv = 2 * Rv / mu
    fhh = -2 * Rh / mu
    n = np.arange(1, N + 1, dtype=np.float64)[:, None]
    rms2 = rms**2
    Iscalar_n = (2 * kz) ** n * np.exp(-rms2 * kz**2)
    Ivv_n = Iscalar_n * fvv
    Ihh_n = Iscalar_n * fhh
    mu2 = mu**2
    sin2 = 1 - mu2
    tan2 = sin2 / mu2
    Ivv_n = Ivv_n + kz**n * (sin2 / mu * (1 + Rv) ** 2 * (1 - 1 / eps_r) * (1 + tan2 / eps_r))
    Ihh_n = Ihh_n - (kz**n) * (sin2 / mu * (1 + Rh) ** 2 * (eps_r - 1) / mu2)
    rms2_over_factorial = np.cumprod(rms2 / n)[:, None]
    kq = -2 * kx
    if acf == 1:
        W_n = (lc**2 / (2 * n)) * np.exp(-((kq * lc) ** 2) / (4 * n))
    else:
        W_n = (lc / n) ** 2 * (1 + (kq * lc / n) ** 2) ** (-1.5)
    coef = knorm**2 / 2 * np.exp(-2 * rms2 * kz**2)
    coef_n = rms2_over_factorial * W_n
    sigma_vv = coef * np.sum(coef_n * abs2(Ivv_n), axis=0)
    sigma_hh = coef * np.sum(coef_n * abs2(Ihh_n), axis=0)
    return sigma_vv.reshape(-1), sigma_hh.reshape(-1)


def backscatter_diffuse_diagonal(s_vv, s_hh, mu, w, mode, m_max):
    """backscattering coefficients -> diagonal diffuse reflection of azimuth mode `mode`, multiplied by the mode's
    integration coefficient and compressed (mu * npol + pol): iem_fung92.py:174-176, 191-214 / reflector_backscatter.py:
    90-116 (spread over 1 + 2 m_max modes), rtsolver_utils.py:728-740 (x weights), 690-709 (2 pi | pi)"""
    npol = 2 if mode == 0 else 3
    coef = 1.0 if mode == 0 else (-2.0 if mode % 2 == 1 else 2.0)
    coef = coef / (1 + 2 * m_max) / (4 * np.pi * mu)
    diff = np.zeros((npol, len(mu)))
    diff[0] = coef * s_vv
    diff[1] = coef * s_hh
    diff *= w
    return (2 * np.pi if mode == 0 else np.pi) * np.transpose(diff).reshape(-1)


def interface_diffuse_reflection(problem, l, eps_1, eps_2, mu, w, mode, m_max):
    """diagonal diffuse reflection of the (rough) interface ABOVE layer l seen from the medium eps_1 on the streams
    (mu, w), or None for flat / transparent interfaces: rtsolver_utils.py:489-501 (top of a layer), 570-582 (bottom),
    631-642 (from the air)"""
    kind = int(problem["interface"][l])
    if kind not in (IF_IEM_FUNG92, IF_IEM_FUNG92_BRIOGONI10):
        return None
    par = np.asarray(problem["interface_params"][l], dtype=float)
    mu = np.asarray(mu, dtype=float)
    s_vv, s_hh = iem_fung92_backscatter(float(problem["frequency"]), eps_1, eps_2, mu, par,
                                        kind == IF_IEM_FUNG92_BRIOGONI10)
    return backscatter_diffuse_diagonal(s_vv, s_hh, mu, np.asarray(w, dtype=float), mode, m_max)


def with_diffuse_reflection(R, problem, streams, mode, coherent_only, where, l=None):
    """R (compressed diagonal reflection) + the diagonal diffuse reflection of the rough surface it belongs to, unless
    this is the coherent pass: where = "top" (interface above layer l seen from layer l), "bottom" (interface or
    substrate below layer l seen from layer l), "air" (top interface seen from the air, on the air streams) —
    combine_coherent_diffuse_matrix, rtsolver_utils.py:646-709."""
    if coherent_only:
        return R
    eps = problem.get("_eps_eff")
    m_max = int(problem.get("_m_max", 0))
    L = len(problem["thickness"])
    d = None
    if where == "top":
        d = interface_diffuse_reflection(problem, l, eps[l], eps[l - 1] if l > 0 else 1, streams["mu"][l],
                                         streams["weight"][l], mode, m_max)
    elif where == "bottom" and l < L - 1:
        d = interface_diffuse_reflection(problem, l + 1, eps[l], eps[l + 1], streams["mu"][l], streams["weight"][l],
                                         mode, m_max)
    elif where == "bottom":
        d = substrate_diffuse_reflection(problem, streams, mode, m_max, eps[-1])
    elif where == "air":
        d = interface_diffuse_reflection(problem, 0, 1, eps[0], streams["outmu"], streams["outweight"], mode, m_max)
    return R if d is None else R + d


def substrate_diffuse_reflection(problem, streams, mode, m_max, eps_1=None):
    """Diagonal diffuse (backscatter) reflection of the substrate for azimuth mode `mode`, already multiplied by the
    mode's integration coefficient and compressed like the coherent part (mu * npol + pol), or None: reference
    smrt/substrate/reflector_backscatter.py:90-116 (the prescribed backscatter spread over the 1 + 2 m_max modes),
    rtsolver_utils.py:728-740 (normalize_diffuse_matrix, 'diagonal5' with mu_i is mu_st: x weights) and 690-709
    (combine_coherent_diffuse_matrix: 2 pi for mode 0, pi above)."""
    kind = int(problem.get("substrate_kind", SUB_NONE))
    if kind not in (SUB_REFLECTOR_BACKSCATTER, SUB_IEM_FUNG92, SUB_IEM_FUNG92_BRIOGONI10):
        return None
    par = np.asarray(problem.get("substrate_params", np.zeros(4)), dtype=float)
    mu = np.asarray(streams["mu"][-1], dtype=float)
    w = np.asarray(streams["weight"][-1], dtype=float)
    if kind == SUB_REFLECTOR_BACKSCATTER:
        s_vv, s_hh = par[2], par[3]
    else:  # iem_fung92.py:174-176, 191-214: the same spreading of the backscatter over the modes
        s_vv, s_hh = iem_fung92_backscatter(float(problem["frequency"]), eps_1, problem["substrate_eps"], mu, par,
                                            kind == SUB_IEM_FUNG92_BRIOGONI10)
    return backscatter_diffuse_diagonal(s_vv, s_hh, mu, w, mode, m_max)


def compute_interfaces(problem, eps_eff, streams, npol):
    """reference smrt/rtsolver/rtsolver_utils.py:473-644 for Flat / Transparent interfaces and a flat substrate."""
    L = len(eps_eff)
    kinds = problem["interface"]
    ipar = problem.get("interface_params")
    freq = float(problem["frequency"])

    def par(l):
        return None if ipar is None else ipar[l]

    Rtop, Ttop, Rbot, Tbot = {}, {}, {}, {}
    for l in range(L):
        eps_lm1 = eps_eff[l - 1] if l > 0 else 1
        eps_l = eps_eff[l]
        Rtop[l], Ttop[l] = interface_R_T(kinds[l], eps_l, eps_lm1, streams["mu"][l], npol, par(l), freq)
        if l < L - 1:
            Rbot[l], Tbot[l] = interface_R_T(kinds[l + 1], eps_l, eps_eff[l + 1], streams["mu"][l], npol, par(l + 1),
                                             freq)
        elif problem.get("substrate_kind", SUB_NONE) != SUB_NONE:
            Rbot[l], Tbot[l] = substrate_R_T(problem, eps_l, streams["mu"][l], npol)
        else:
            Rbot[l] = None
            Tbot[l] = None
    Rbot[-1], Tbot[-1] = interface_R_T(kinds[0], 1, eps_eff[0], streams["outmu"], npol, par(0), freq)
    return Rtop, Ttop, Rbot, Tbot


# --------------------------------------------------------------------------------------------------------------------
# a5/a6  Fourier modes of the phase matrix
# --------------------------------------------------------------------------------------------------------------------
def rayleigh_phase_and_angle(mu_s, mu_i, dphi, npol):
    """reference smrt/emmodel/common.py:9-53 (+ core/lib.py:623-652)"""
    dphi = dphi[:, None, None]
    mu_s = mu_s[None, :, None]
    mu_i = mu_i[None, None, :]
    sin_i = np.sqrt(1.0 - mu_i**2)
    sin_s = np.sqrt(1.0 - mu_s**2)
    sinphi = np.sin(dphi)
    cosphi = np.cos(dphi)
    fvv = cosphi * mu_s * mu_i + sin_s * sin_i
    fhv = -sinphi * mu_i
    fhh = cosphi
    fvh = sinphi * mu_s
    fvv, fvh, fhv, fhh = np.broadcast_arrays(fvv, fvh, fhv, fhh)
    if npol == 2:
        p = [[fvv**2, fvh**2], [fhv**2, fhh**2]]
    else:
        p = [[fvv**2, fvh**2, fvh * fvv],
             [fhv**2, fhh**2, fhh * fhv],
             [2 * (fvv * fhv), 2 * (fvh * fhh), fvv * fhh + fvh * fhv]]
    p = np.array(p)
    cosT = np.clip(mu_s * mu_i + sin_s * sin_i * cosphi, -1.0, 1.0)
    sin_half_scatt = np.sqrt(0.5 * (1 - cosT))
    return p, sin_half_scatt


def estimate_ft_number_samples(m_max):
    """reference smrt/emmodel/common.py:401-414"""
    return int(2 ** np.ceil(4 + np.log(m_max + 1) / np.log(2)))


def generic_ft_even_matrix(phase_function, m_max, nsamples):
    """reference smrt/emmodel/common.py:56-131 — DFT over azimuth of the mirrored phase samples"""
    dphi = np.linspace(0, np.pi, int(nsamples // 2 + 1))
    p = phase_function(dphi)
    npol = p.shape[0]
    p_mirror = p[:, :, -2:0:-1, :, :].copy()
    if npol >= 3:
        p_mirror[0:2, 2] = -p_mirror[0:2, 2]
        p_mirror[2, 0:2] = -p_mirror[2, 0:2]
    p = np.concatenate((p, p_mirror), axis=2)
    ft_p = np.fft.fft(p, axis=2)
    ft_even_p = np.empty((npol, npol, m_max + 1, p.shape[-2], p.shape[-1]))
    ft_even_p[:, :, 0] = ft_p[:, :, 0].real * (1.0 / nsamples)
    if npol == 2:
        ft_even_p[:, :, 1:] = ft_p[:, :, 1:m_max + 1].real * (2.0 / nsamples)
    else:
        delta = 2.0 / nsamples
        ft_even_p[0:2, 0:2, 1:] = ft_p[0:2, 0:2, 1:m_max + 1].real * delta
        ft_even_p[0:2, 2, 1:] = ft_p[0:2, 2, 1:m_max + 1].imag * delta
        ft_even_p[2, 0:2, 1:] = -ft_p[2, 0:2, 1:m_max + 1].imag * delta
        ft_even_p[2, 2, 1:] = ft_p[2, 2, 1:m_max + 1].real * delta
    return ft_even_p


def iba_ft_even_phase(opt, mu_s, mu_i, m_max, npol):
    """reference smrt/emmodel/common.py:349-399 + smrt/emmodel/iba.py:228-244"""
    if np.any(mu_i == 1) and npol > 2:
        raise OracleError(ST_EIGEN, "Phase matrix signs for sine elements of mode m = 2 incorrect")

    def phase_function(dphi):
        p, sin_half_scatt = rayleigh_phase_and_angle(mu_s, mu_i, dphi, npol)
        k_diff = 2.0 * opt["k0"] * np.sqrt(opt["eps_eff"]).real * sin_half_scatt
        ft_corr_fn = ft_autocorr(k_diff, opt["ms_kind"], opt["f"], opt["p0"], opt["p1"])
        return ft_corr_fn * opt["iba_coeff"] * p

    return generic_ft_even_matrix(phase_function, m_max, estimate_ft_number_samples(m_max))


def rayleigh_ft_even_phase(ks, mu_s, mu_i, m_max):
    """reference smrt/emmodel/rayleigh.py:52-127 (ft_even_phase_baseonUlaby)"""
    npol = 2 if m_max == 0 else 3
    P = np.empty((npol, npol, m_max + 1, len(mu_s), len(mu_i)))
    mu_i2 = mu_i**2
    mu_s2 = mu_s**2
    v, h, u = 0, 1, 2
    P[v, v, 0] = 0.5 * np.outer(mu_s2, mu_i2) + np.outer(1 - mu_s2, 1 - mu_i2)
    P[v, h, 0] = 0.5 * mu_s2[:, None]
    if npol >= 3:
        P[v, u] = 0
    P[h, v, 0] = 0.5 * mu_i2[None, :]
    P[h, h, 0] = 0.5
    if npol >= 3:
        P[h, u, 0] = 0
        P[u, v, 0] = 0
        P[u, h, 0] = 0
        P[u, u, 0] = 0
    if m_max >= 1:
        sint_s = np.sqrt(1.0 - mu_s2)
        sint_i = np.sqrt(1.0 - mu_i2)
        cossint_s = mu_s * sint_s
        cossint_i = mu_i * sint_i
        P[v, v, 1] = 2 * np.outer(cossint_s, cossint_i)
        P[v, h, 1] = 0
        P[v, u, 1] = np.outer(cossint_s, sint_i)
        P[h, v, 1] = 0
        P[h, h, 1] = 0
        P[h, u, 1] = 0
        P[u, v, 1] = -2 * np.outer(sint_s, cossint_i)
        P[u, h, 1] = 0
        P[u, u, 1] = np.outer(sint_s, sint_i)
    if m_max >= 2:
        P[v, v, 2] = 0.5 * np.outer(mu_s2, mu_i2)
        P[v, h, 2] = -0.5 * mu_s2[:, None]
        P[v, u, 2] = 0.5 * np.outer(mu_s2, mu_i)
        P[h, v, 2] = -0.5 * mu_i2[None, :]
        P[h, h, 2] = 0.5
        P[h, u, 2] = -0.5 * mu_i[None, :]
        P[u, v, 2] = -np.outer(mu_s, mu_i2)
        P[u, h, 2] = mu_s[:, None]
        P[u, u, 2] = np.outer(mu_s, mu_i)
    if m_max >= 3:
        P[:, :, 3:, :, :] = 0
    if npol == 3:
        P[v, u, :] = -P[v, u, :]
        P[h, u, :] = -P[h, u, :]
    return P * (3 * ks / 2)


def compress_dense(P5, mode):
    """smrt_matrix 'dense5' .compress(mode, auto_reduce_npol=True): (pol_s,pol_i,m,mu_s,mu_i) -> (mu_s*pol_s, mu_i*pol_i)
    reference smrt/core/lib.py:314-347, 443-446"""
    pola = slice(0, 2) if (P5.shape[0] == 3 and mode == 0) else slice(None)
    mat = P5[pola, pola, mode, :, :]
    mat = np.moveaxis(mat, (0, 1), (1, 3))
    return np.reshape(mat, (mat.shape[0] * mat.shape[1], mat.shape[2] * mat.shape[3]))


# --------------------------------------------------------------------------------------------------------------------
# a9/a10  layer eigenproblem — reference smrt/rtsolver/dort.py:617-889, 1068-1103
# --------------------------------------------------------------------------------------------------------------------
class LayerEigen:
    def __init__(self, opt, mu, weight, m_max, npol_em, normalization=True, method="schur_forcedtriu"):
        self.opt, self.mu, self.weight, self.m_max = opt, mu, weight, m_max
        self.npol_em = npol_em
        self.normalization = normalization
        self.method = method
        self.norm_0 = None
        self.norm_m = None
        self._phase = None

    def phase(self):
        if self._phase is None:
            fullmu = np.concatenate((self.mu, -self.mu))
            if self.opt["ks"] == 0:
                self._phase = 0
            elif self.opt["emmodel"] in EM_IBA_FAMILY:
                self._phase = iba_ft_even_phase(self.opt, fullmu, fullmu, self.m_max, self.npol_em)
            elif self.opt["emmodel"] in (EM_DMRT_QCA_SR, EM_DMRT_QCACP_SR, EM_RAYLEIGH, EM_PRESCRIBED_KSKAEPS):
                self._phase = rayleigh_ft_even_phase(self.opt["ks"], fullmu, fullmu, self.m_max)
            else:
                self._phase = 0
        return self._phase

    def no_scattering(self, m):
        npol = 2 if m == 0 else 3
        n = npol * len(self.mu)
        invmu = np.repeat(1.0 / self.mu, npol)
        invmu = np.concatenate((invmu, -invmu))
        ke = self.opt["ks"] + self.opt["ka"]
        beta = invmu * ke
        E = np.eye(2 * n)
        return beta, E[0:n, :], E[n:, :]

    def build_A(self, m):
        """dort.py:714-749"""
        P = self.phase()
        if isinstance(P, int) or not np.any(P):
            return None
        A = compress_dense(P, m).copy()
        npol = 2 if m == 0 else 3
        invmu = np.repeat(1.0 / self.mu, npol)
        invmu = np.concatenate((invmu, -invmu))
        coef = 0.5 if m == 0 else 0.25
        coef_weight = np.tile(np.repeat(-coef * self.weight, npol), 2)
        A *= coef_weight[None, :]
        k = A.shape[0]
        if self.normalization:
            A = self.normalize(m, A, self.opt["ks"])
        A[np.diag_indices(k)] += self.opt["ks"] + self.opt["ka"]
        A = invmu[0:k, None] * A
        return A

    def normalize(self, m, A, ks):
        """dort.py:782-819"""
        if m == 0:
            if ks == 0:
                return A
            self.norm_0 = -ks / np.sum(A, axis=1)
            norm = self.norm_0
            if self.normalization != "forced" and np.any(np.abs(self.norm_0 - 1.0) > 0.3):
                raise OracleError(ST_NORMALIZATION, "The re-normalization of the phase function exceeds the "
                                  "predefined threshold of 30%.")
        else:
            if self.norm_m is None:
                if self.norm_0 is None:
                    raise RuntimeError("mode 0 must be normalised first")
                npol = 3
                self.norm_m = np.empty(len(self.norm_0) // 2 * npol)
                self.norm_m[0::npol] = self.norm_0[0::2]
                self.norm_m[1::npol] = self.norm_0[1::2]
                self.norm_m[2::npol] = np.sqrt(self.norm_0[0::2] * self.norm_0[1::2])
            norm = self.norm_m
        A *= norm[:, None]
        return A

    def solve(self, m, coherent_only=False):
        if coherent_only:
            return self.no_scattering(m)
        A = self.build_A(m)
        if A is None:
            return self.no_scattering(m)
        npol = 2 if m == 0 else 3
        n = npol * len(self.mu)
        if self.method == "eig":
            beta, E = scipy.linalg.eig(A)
        elif self.method == "half_rank_eig":
            return self.half_rank(m, A)
        else:
            T, Z = scipy.linalg.schur(A)  # dort.py:842
            if self.method == "schur_forcedtriu":
                T[np.tril_indices(T.shape[0], k=-1)] = 0  # dort.py:848
            beta, E = scipy.linalg.eig(T)  # dort.py:850
            E = Z @ E
        return self.validate(beta, E[0:n, :], E[n:, :])

    def half_rank(self, m, A):
        """dort.py:891-962"""
        n = A.shape[1] // 2
        alpha_mat = -A[0:n, 0:n]
        beta_mat = -A[0:n, n:].copy()
        if m > 0:
            beta_mat[:, 2::3] = -beta_mat[:, 2::3]
        half_rank_A = (alpha_mat - beta_mat) @ (alpha_mat + beta_mat)
        beta, Ep = scipy.linalg.eig(half_rank_A)
        beta = np.sqrt(beta.real)
        Em = (alpha_mat + beta_mat) @ (Ep * (1 / beta)[None, :])
        Eu = np.hstack((0.5 * (Ep - Em), 0.5 * (Ep + Em)))
        Ed = np.hstack((Eu[:, n:], Eu[:, 0:n]))
        if m > 0:
            Ed[2::3, :] = -Ed[2::3, :]
        beta = np.concatenate((beta, -beta))
        return self.validate(beta, Eu, Ed)

    @staticmethod
    def validate(beta, Eu, Ed):
        """dort.py:1068-1103"""
        iscomplex_beta = not np.allclose(beta.imag, 0, atol=np.max(beta.real) * 1e-07)
        iscomplex_Eu = not np.allclose(Eu.imag, 0, atol=1e-6)
        iscomplex_Ed = not np.allclose(Ed.imag, 0, atol=1e-6)
        if iscomplex_beta or iscomplex_Eu or iscomplex_Ed:
            raise OracleError(ST_EIGEN, "diagonalization failed: complex eigenvalues / eigenvectors")
        return beta.real, Eu.real, Ed.real


# --------------------------------------------------------------------------------------------------------------------
# a13  Planck — reference smrt/core/lib.py:594-620
# --------------------------------------------------------------------------------------------------------------------
def planck_function(frequency, temperature):
    if temperature <= 1e-10:
        return 0.0
    b = (PLANCK_CONSTANT / BOLTZMANN_CONSTANT) * frequency / temperature
    return (2.0 * PLANCK_CONSTANT / C_SPEED**2) * frequency**3 / (np.exp(b) - 1.0)


def inverse_planck_function(frequency, radiance):
    radiance = np.asarray(radiance, dtype=float)
    out = np.zeros_like(radiance)
    pos = radiance > 1e-40
    x = (2.0 * PLANCK_CONSTANT / C_SPEED**2) * frequency**3 / radiance[pos]
    out[pos] = (PLANCK_CONSTANT / BOLTZMANN_CONSTANT) * frequency / np.log(1 + x)
    return out


# --------------------------------------------------------------------------------------------------------------------
# a11  boundary system for one azimuth mode — reference smrt/rtsolver/dort.py:263-488
# --------------------------------------------------------------------------------------------------------------------
def _todiag(bmat, oi, oj, dmat):
    """dort.py:556-580 (scatter of a dense block into LAPACK band storage), vectorised"""
    u = (bmat.shape[0] - 1) // 2
    n, m = dmat.shape
    I = np.arange(n)[:, None]
    J = np.arange(m)[None, :]
    bmat[u + (oi + I) - (oj + J), (oj + J) + 0 * I] = dmat


def dort_modem_banded(problem, mode, streams, eigs, iface, intensity_down, planck, coherent_only=False,
                      prune_deep_snowpack=None, info=None):
    Rtop_, Ttop_, Rbot_, Tbot_ = iface
    npol = 2 if mode == 0 else 3
    ns = streams["n"]
    L = len(ns)
    thickness = problem["thickness"]
    temperature = problem["temperature"] if problem["mode"] == "P" else None

    jl = 2 * (np.cumsum(ns) - ns) * npol
    il_top = jl.copy()
    il_bottom = il_top + ns * npol
    nboundary = int(sum(ns) * 2 * npol)
    if L >= 2:
        nband = int(npol * max(np.max(2 * ns[1:] + ns[:-1]), np.max(ns[1:] + 2 * ns[:-1])))
    else:
        nband = int(3 * npol * np.max(ns))
    bBC = np.zeros((2 * nband + 1, nboundary))
    nvector = intensity_down.shape[1]
    b = np.zeros((nboundary, nvector))
    optical_depth = 0.0

    def cdiag(mat, l):
        return None if mat[l] is None else compress_diag(mat[l], mode)

    for l in range(L):
        nsl_npol = ns[l] * npol
        nslm1_npol = ns[l - 1] * npol if l > 0 else streams["n_air"] * npol
        nslp1_npol = ns[l + 1] * npol if l < L - 1 else None
        beta, Eu, Ed = eigs[l].solve(mode, coherent_only)
        transt = np.exp(-np.maximum(beta, 0) * thickness[l])
        transb = np.exp(np.minimum(beta, 0) * thickness[l])
        if l == 0:
            Eu_0, transt_0 = Eu, transt
        Rtop_l = with_diffuse_reflection(cdiag(Rtop_, l), problem, streams, mode, coherent_only, "top", l)
        _todiag(bBC, il_top[l], jl[l], (Ed - Rtop_l[:, None] * Eu) * transt[None, :])
        Tbottom_lp1 = None
        if l < L - 1:
            Tbottom_lp1 = cdiag(Tbot_, l)
            if np.any(Tbottom_lp1):
                nc_b = min(len(Tbottom_lp1), nslp1_npol)
                _todiag(bBC, il_top[l + 1], jl[l], -(Tbottom_lp1[:, None] * Ed * transb[None, :])[:nc_b, :])
            else:
                Tbottom_lp1 = None
        Tl = temperature[l] if temperature is not None else None
        if mode == 0 and Tl is not None and Tl > 0:
            b[il_top[l]:il_top[l] + nsl_npol, :] -= ((1.0 - Rtop_l) * planck(Tl))[:, None]
            if l < L - 1 and Tbottom_lp1 is not None:
                b[il_top[l + 1]:il_top[l + 1] + nc_b, :] += (Tbottom_lp1 * planck(Tl))[:nc_b, None]
        if l == 0:
            Tbottom_air_down = compress_diag(Tbot_[-1], mode)
            if np.any(Tbottom_air_down):
                nc = min(len(Tbottom_air_down), nsl_npol)
                b[il_top[l]:il_top[l] + nc, :] += Tbottom_air_down[:, None] * intensity_down

        Rbottom_l = cdiag(Rbot_, l)
        if Rbottom_l is None:
            Rbottom_l = np.zeros(nsl_npol)
        Rbottom_l = with_diffuse_reflection(Rbottom_l, problem, streams, mode, coherent_only, "bottom", l)
        _todiag(bBC, il_bottom[l], jl[l], (Eu - Rbottom_l[:, None] * Ed) * transb[None, :])
        Ttop_lm1 = None
        if l > 0:
            Ttop_lm1 = cdiag(Ttop_, l)
            if np.any(Ttop_lm1):
                nc_t = min(len(Ttop_lm1), nslm1_npol)
                _todiag(bBC, il_bottom[l - 1], jl[l], -(Ttop_lm1[:, None] * Eu * transt[None, :])[:nc_t, :])
            else:
                Ttop_lm1 = None
        if mode == 0 and Tl is not None and Tl > 0:
            b[il_bottom[l]:il_bottom[l] + nsl_npol, :] -= ((1.0 - Rbottom_l) * planck(Tl))[:, None]
            if l > 0 and Ttop_lm1 is not None:
                b[il_bottom[l - 1]:il_bottom[l - 1] + nc_t, :] += (Ttop_lm1 * planck(Tl))[:nc_t, None]
        if (mode == 0 and l == L - 1 and problem.get("substrate_kind", SUB_NONE) != SUB_NONE
                and problem.get("substrate_temperature") is not None and temperature is not None):
            Tbottom_sub = cdiag(Tbot_, l)
            nc = min(len(Tbottom_sub), nsl_npol)
            if np.any(Tbottom_sub):
                b[il_bottom[l]:il_bottom[l] + nc, :] += (Tbottom_sub * planck(problem["substrate_temperature"]))[:nc, None]

        optical_depth += np.min(np.abs(beta)) * thickness[l]
        if prune_deep_snowpack is not None and optical_depth > prune_deep_snowpack:
            nboundary = int(sum(ns[0:l + 1]) * 2 * npol)
            bBC = bBC[:, 0:nboundary]
            b = b[0:nboundary, :]
            break

    if info is not None:
        info["optical_depth"] = optical_depth
        info["shallow"] = bool(problem.get("substrate_kind", SUB_NONE) == SUB_NONE and optical_depth < 5)

    try:
        x = scipy.linalg.solve_banded((nband, nband), bBC, b)
    except (scipy.linalg.LinAlgError, ValueError) as e:
        raise OracleError(ST_SINGULAR, f"boundary system: {e}")

    nsl2_npol = 2 * ns[0] * npol
    I1up_m = (Eu_0 * transt_0[None, :]) @ x[0:nsl2_npol, :]
    if mode == 0 and temperature is not None and temperature[0] > 0:
        I1up_m += planck(temperature[0])
    Rbottom_air_down = with_diffuse_reflection(compress_diag(Rbot_[-1], mode), problem, streams, mode, coherent_only,
                                               "air")
    Ttop_0 = compress_diag(Ttop_[0], mode)
    I0up_m = Rbottom_air_down[:, None] * intensity_down + (Ttop_0[:, None] * I1up_m)[0:streams["n_air"] * npol, :]
    I0up_m = np.array(I0up_m).squeeze()
    if np.ndim(I0up_m) == 1:  # dort.py:491-511
        return I0up_m.reshape((I0up_m.shape[0] // npol, npol)).transpose()
    return I0up_m.reshape((I0up_m.shape[0] // npol, npol, I0up_m.shape[1] // npol, npol)).transpose(1, 0, 3, 2)


# --------------------------------------------------------------------------------------------------------------------
# a12/a14  mode summation, interpolation — reference smrt/rtsolver/rtsolver_utils.py:91-320
# --------------------------------------------------------------------------------------------------------------------
def prepare_incident_streams(outmu, theta_inc):
    inc = set()
    for mu_inc in np.cos(theta_inc):
        i0 = int(np.searchsorted(-outmu, -mu_inc))
        if i0 == 0:
            inc.add(i0)
        elif i0 == len(outmu):
            inc.add(i0 - 1)
        else:
            inc.add(i0)
            inc.add(i0 - 1)
    return sorted(inc)


def interpolate_intensity(mode, outmu, intensity, theta):
    """rtsolver_utils.py:179-239 — linear in mu with extrapolation (scipy interp1d fill_value='extrapolate')"""
    user_mu = np.cos(theta)
    mu_axis = 1 if mode == "P" else 2
    if np.max(user_mu) > np.max(outmu):
        imumax = int(np.argmax(outmu))
        if mode == "P":
            outmu = np.insert(outmu, 0, 1.0)
            mean_H_V = np.mean(intensity.take(imumax, axis=mu_axis), axis=0)
            intensity = np.insert(intensity, 0, mean_H_V, axis=mu_axis)
        else:
            copol = (intensity[0, 0, imumax] + intensity[1, 1, imumax]) / 2
            crosspol = (intensity[1, 0, imumax] + intensity[0, 1, imumax]) / 2
            intensity = np.insert(intensity, 0,
                                  [[copol, crosspol, intensity[0, 2, imumax]],
                                   [crosspol, copol, intensity[1, 2, imumax]],
                                   intensity[2, :, imumax]], axis=mu_axis)
            outmu = np.insert(outmu, 0, 1.0)
    import scipy.interpolate
    if len(outmu) == 1:
        return np.repeat(intensity, len(user_mu), axis=mu_axis)
    intfct = scipy.interpolate.interp1d(outmu, intensity, axis=mu_axis, fill_value="extrapolate", bounds_error=False,
                                        assume_sorted=False)
    return intfct(user_mu)


def solve_problem(problem, method="schur_forcedtriu", return_details=False):
    """One (snowpack x frequency) DORT solve — reference smrt/rtsolver/dort.py:189-261.

    Returns dict(values=..., status=int, ks, ka, ke, eps_eff, stream_angles) with
      passive: values (2, n_theta) brightness temperature [V, H]
      active : values (3, 3, n_theta) intensity with axis order (polarization_inc-label, polarization-label, theta_inc)
               exactly as the reference lays it out; sigma0 = 4 pi cos(theta) * values.
    """
    opts = dict(n_max_stream=32, m_max=2, phase_normalization="auto", prune_deep_snowpack=None,
                rayleigh_jeans_approximation=False, error_handling="exception")
    opts.update(problem.get("options", {}))
    freq = float(problem["frequency"])
    mode = problem["mode"]
    L = len(problem["thickness"])
    theta = np.atleast_1d(np.asarray(problem["theta"], dtype=float))
    out = dict(status=ST_OK)

    if L == 0:  # empty snowpack: Tb = 0 (reference test/test_model.py:36-43)
        raise NotImplementedError("empty snowpack is handled by the host layer")

    dsc = problem.get("dense_snow_correction")
    if dsc is None:
        dsc = np.isin(np.asarray(problem["emmodel"]), (EM_DMRT_QCA_SR, EM_DMRT_QCACP_SR)).astype(int)
    optics = [layer_optics(freq, problem["frac_volume"][l], problem["eps_bg"][l], problem["eps_sc"][l],
                           int(problem["emmodel"][l]), int(problem["ms_kind"][l]), problem["ms_p0"][l],
                           problem["ms_p1"][l], bool(dsc[l]),
                           None if problem.get("inclusion") is None else problem["inclusion"][l]) for l in range(L)]
    eps_eff = np.array([o["eps_eff"] for o in optics])
    out.update(eps_eff=eps_eff, ks=np.array([o["ks"] for o in optics]), ka=np.array([o["ka"] for o in optics]))
    out["ke"] = out["ks"] + out["ka"]

    try:
        streams = compute_streams(int(opts["n_max_stream"]), eps_eff)
        out["stream_angles"] = np.rad2deg(np.arccos(streams["outmu"]))
        m_max = int(opts["m_max"]) if mode == "A" else 0
        npol = 2 if mode == "P" else 3
        iface = compute_interfaces(problem, eps_eff, streams, npol)
        problem = dict(problem, _m_max=m_max, _eps_eff=eps_eff)  # for the diffuse reflection of rough surfaces
        norm = opts["phase_normalization"]
        if norm == "auto":
            norm = True  # IBA, DMRT: _respect_reciprocity_principle defaults to True (dort.py:240-242)
        eigs = [LayerEigen(optics[l], streams["mu"][l], streams["weight"][l], m_max, npol, norm, method)
                for l in range(L)]

        if opts["rayleigh_jeans_approximation"]:
            planck = lambda T: T  # noqa: E731
            inv_planck = lambda I: I  # noqa: E731
        else:
            planck = lambda T: planck_function(freq, T)  # noqa: E731
            inv_planck = lambda I: inverse_planck_function(freq, I)  # noqa: E731

        n_air = streams["n_air"]
        info = {}
        kw = dict(problem=problem, streams=streams, eigs=eigs, iface=iface, planck=planck,
                  prune_deep_snowpack=opts["prune_deep_snowpack"], info=info)
        if mode == "P":
            atmos = problem.get("atmosphere")
            intensity_0 = np.zeros((2 * n_air, 1))
            if atmos is not None:  # isotropic atmosphere (atmosphere/simple_isotropic_atmosphere.py:55-77,
                intensity_0[:] = planck(float(atmos[0]))  # core/atmosphere.py:134-162, rtsolver_utils.py:141-147)
            I = dort_modem_banded(mode=0, intensity_down=intensity_0, **kw)
            intensity_up = np.zeros((2, n_air))
            intensity_up[0:2] += I[0:2]
            if atmos is not None:  # rtsolver_utils.py:302-304
                intensity_up = planck(float(atmos[1])) + float(atmos[2]) * intensity_up
            intensity_up = inv_planck(intensity_up)
            outmu = streams["outmu"]
        else:
            inc = prepare_incident_streams(streams["outmu"], theta)
            intensity_0 = np.zeros((2 * n_air, 2 * len(inc)))
            intensity_higher = np.zeros((3 * n_air, 3 * len(inc)))
            j0 = jh = 0
            for i in inc:
                power = 1.0 / (2 * np.pi * streams["outweight"][i])
                for ipol in (0, 1):
                    intensity_0[2 * i + ipol, j0] = power
                    j0 += 1
                for ipol in (0, 1, 2):
                    intensity_higher[3 * i + ipol, jh] = 2 * power
                    jh += 1
            intensity_up = np.zeros((3, n_air, 3, len(inc)))
            coh = dort_modem_banded(mode=0, intensity_down=intensity_0, coherent_only=True, **kw)
            phi = float(problem.get("phi", np.pi))
            for m in range(m_max + 1):
                I = dort_modem_banded(mode=m, intensity_down=intensity_0 if m == 0 else intensity_higher, **kw)
                I[0:2, :, 0:2, :] -= coh * (1 + float(m > 0))
                if m == 0:
                    intensity_up[0:2, :, 0:2] += I[0:2, :, 0:2]
                else:
                    intensity_up[0:2] += I[0:2] * np.cos(m * phi)
                    intensity_up[2:] += I[2:] * np.sin(m * phi)
            back = np.empty((3, 3, len(inc)))
            for j, i in enumerate(inc):
                back[:, :, j] = intensity_up[:, i, :, j]
            outmu = streams["outmu"][inc]
            intensity_up = back
        if info.get("shallow"):
            out["status"] |= ST_SHALLOW_WARNING
        out["optical_depth"] = info.get("optical_depth")
        # make_result (rtsolver_utils.py:338-342): stream_angles are those of the RETURNED outmu (active: incident only)
        out["stream_angles"] = np.rad2deg(np.arccos(outmu))
        out["values"] = interpolate_intensity(mode, outmu, intensity_up, theta)
    except OracleError as e:
        if opts["error_handling"] == "nan":
            shape = (2, len(theta)) if mode == "P" else (3, 3, len(theta))
            out["values"] = np.full(shape, np.nan)
            out["status"] = e.status
        else:
            raise
    if return_details:
        out["streams"] = streams
    return out
