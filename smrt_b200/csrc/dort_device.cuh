// dort_device.cuh — per-thread device functions of the DORT hot path: complex arithmetic, layer optics
// (IBA / DMRT-QCA(-CP) short range / non-scattering), microstructure FTs, streams, Fresnel coefficients, Planck and the
// Fourier modes of the phase matrix.  Every function cites the reference lines whose arithmetic it reproduces
// (reference = smrt-model/smrt, paths relative to smrt/).
#pragma once
#include "simt.h"
#include <math.h>

#define SMRT_PI 3.141592653589793238462643383279502884
#define SMRT_C_SPEED 299792458.0           // core/globalconstants.py:30
#define SMRT_PLANCK 6.62607015e-34         // core/globalconstants.py:31
#define SMRT_BOLTZMANN 1.380649e-23        // core/globalconstants.py:32

// enumerations: keep in sync with include/smrt_dort_b200.h
enum {
  EM_IBA = 0, EM_DMRT_QCA_SR = 1, EM_NONSCATTERING = 2, EM_DMRT_QCACP_SR = 3, EM_RAYLEIGH = 4, EM_PRESCRIBED_KSKAEPS = 5,
  EM_IBA_ORIGINAL = 6, EM_IBA_MAXWELL_GARNETT = 7
};
// the IBA family shares the scattering coefficient, the phase matrix and the dense-snow inversion (iba.py:85-265)
SMRT_DEV bool em_is_iba(int emmodel) {
  return emmodel == EM_IBA || emmodel == EM_IBA_ORIGINAL || emmodel == EM_IBA_MAXWELL_GARNETT;
}
enum {
  MS_EXPONENTIAL = 0, MS_SHS = 1, MS_HOMOGENEOUS = 2, MS_INDEPENDENT_SPHERE = 3, MS_TEUBNER_STREY = 4,
  MS_UNIFIED_TS_1 = 5, MS_UNIFIED_TS_2 = 6, MS_SHS_T = 7
};
enum { IF_FLAT = 0, IF_TRANSPARENT = 1, IF_IEM_FUNG92 = 2, IF_IEM_FUNG92_BRIOGONI10 = 3 };
enum {
  SUB_NONE = 0,
  SUB_FLAT = 1,
  SUB_SOIL_WEGMULLER = 2,
  SUB_SOIL_QNH = 3,
  SUB_REFLECTOR = 4,
  SUB_ROUGH_CHOUDHURY = 5,
  SUB_REFLECTOR_BACKSCATTER = 6,
  SUB_IEM_FUNG92 = 7,
  SUB_IEM_FUNG92_BRIOGONI10 = 8
};
enum { ST_OK = 0, ST_NORMALIZATION = 1, ST_EIGEN = 2, ST_SINGULAR = 3, ST_INPUT = 4, ST_SUBSTRATE = 5, ST_WARN_SHALLOW = 16 };

// ---------------------------------------------------------------------------------------------------- complex numbers
struct cplx {
  double re, im;
};
SMRT_DEV cplx c_make(double re, double im) {
  cplx z;
  z.re = re;
  z.im = im;
  return z;
}
SMRT_DEV cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
SMRT_DEV cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
SMRT_DEV cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SMRT_DEV cplx c_scale(cplx a, double s) { return c_make(a.re * s, a.im * s); }
SMRT_DEV cplx c_conj(cplx a) { return c_make(a.re, -a.im); }
SMRT_DEV double c_abs2(cplx a) { return a.re * a.re + a.im * a.im; }
SMRT_DEV double c_abs(cplx a) { return hypot(a.re, a.im); }
// Smith's algorithm (what numpy uses for complex128 division)
SMRT_DEV cplx c_div(cplx a, cplx b) {
  if (fabs(b.re) >= fabs(b.im)) {
    double rat = b.im / b.re;
    double scl = 1.0 / (b.re + b.im * rat);
    return c_make((a.re + a.im * rat) * scl, (a.im - a.re * rat) * scl);
  } else {
    double rat = b.re / b.im;
    double scl = 1.0 / (b.re * rat + b.im);
    return c_make((a.re * rat + a.im) * scl, (a.im * rat - a.re) * scl);
  }
}
// principal square root (C99 csqrt algorithm, finite inputs)
SMRT_DEV cplx c_sqrt(cplx z) {
  if (z.re == 0.0 && z.im == 0.0) return c_make(0.0, z.im);
  double t;
  if (z.re >= 0.0) {
    t = sqrt((z.re + hypot(z.re, z.im)) * 0.5);
    return c_make(t, z.im / (2.0 * t));
  } else {
    t = sqrt((-z.re + hypot(z.re, z.im)) * 0.5);
    return c_make(fabs(z.im) / (2.0 * t), copysign(t, z.im));
  }
}

// ---------------------------------------------------------------------------------------------------- Planck
// core/lib.py:594-620
SMRT_DEV double planck_function(double frequency, double temperature, int rayleigh_jeans) {
  if (rayleigh_jeans) return temperature;
  if (!(temperature > 1e-10)) return 0.0;
  double b = (SMRT_PLANCK / SMRT_BOLTZMANN) * frequency / temperature;
  return (2.0 * SMRT_PLANCK / (SMRT_C_SPEED * SMRT_C_SPEED)) * frequency * frequency * frequency / (exp(b) - 1.0);
}
SMRT_DEV double inverse_planck_function(double frequency, double radiance, int rayleigh_jeans) {
  if (rayleigh_jeans) return radiance;
  if (!(radiance > 1e-40)) return 0.0;
  double x = (2.0 * SMRT_PLANCK / (SMRT_C_SPEED * SMRT_C_SPEED)) * frequency * frequency * frequency / radiance;
  return (SMRT_PLANCK / SMRT_BOLTZMANN) * frequency / log(1.0 + x);
}

// ---------------------------------------------------------------------------------------------------- microstructure
struct MicroParams {
  int kind;      // MS_*
  double f;      // fractional volume (after a possible medium inversion)
  double p0;     // corr_length | radius
  double shs_t;  // Percus-Yevick t of sticky_hard_spheres.py:84-91 (NOT compute_t)
  double c0;     // exponential: f(1-f) 8 pi l^3 ; SHS: f * vd
  double a1, a2; // SHS: the two constants of A(X)
  double ct0;    // SHS: value at X ~ 0
};

// microstructure_model/exponential.py:53-58 and sticky_hard_spheres.py:63-130, prepared once per layer
SMRT_DEV MicroParams micro_prepare(int kind, double f, double p0, double p1) {
  MicroParams mp;
  mp.kind = kind;
  mp.f = f;
  mp.p0 = p0;
  mp.shs_t = 0.0;
  mp.c0 = mp.a1 = mp.a2 = mp.ct0 = 0.0;
  if (kind == MS_EXPONENTIAL) {
    mp.c0 = f * (1.0 - f) * 8.0 * SMRT_PI * p0 * p0 * p0;
  } else if (kind == MS_INDEPENDENT_SPHERE) {  // independent_sphere.py:62-80: f (1 - f) * sphere volume
    mp.c0 = f * (1.0 - f) * (4.0 / 3.0 * SMRT_PI * (p0 * p0 * p0));
  } else if (kind == MS_TEUBNER_STREY) {  // teubner_strey.py:53-62; a1 = Y = (2 pi l / d)^2
    const double y = 2.0 * SMRT_PI * p0 / p1;
    mp.a1 = y * y;
    mp.c0 = 8.0 * SMRT_PI * (p0 * p0 * p0);
    mp.ct0 = f * (1.0 - f);
  } else if (kind == MS_UNIFIED_TS_1) {  // unified_teubner_strey.py:70-73 (p0 = zeta1, p1 = zeta2)
    mp.c0 = 4.0 * SMRT_PI * p0 * p1 * (p0 + p1);
    mp.a1 = p1;
    mp.ct0 = f * (1.0 - f);
  } else if (kind == MS_UNIFIED_TS_2) {  // unified_teubner_strey.py:75-78; a1 = zeta1 / zeta2
    mp.c0 = 8.0 * SMRT_PI * (p0 * p0 * p0);
    mp.a1 = p0 / p1;
    mp.ct0 = f * (1.0 - f);
  } else if (kind == MS_SHS || kind == MS_SHS_T) {
    double tau = p1, phi2 = f;
    double t = 0.0;
    if (kind == MS_SHS_T) {  // unified_sticky_hard_spheres.py:27-31: t prescribed
      t = p1;
    } else if (isfinite(tau) && phi2 > 0.0) {
      double disc = 36.0 * tau * tau * phi2 * phi2 - 72.0 * tau * phi2 * phi2 - 72.0 * tau * tau * phi2 +
                    30.0 * phi2 * phi2 + 72.0 * tau * phi2 + 36.0 * tau * tau - 12.0 * phi2;
      t = (6.0 * tau * phi2 - 6.0 * phi2 - 6.0 * tau + sqrt(disc)) / (phi2 * (-1.0 + phi2));
    }
    mp.shs_t = t;
    double vd = 4.0 / 3.0 * SMRT_PI * p0 * p0 * p0;
    mp.c0 = phi2 * vd;
    double r = phi2 / (1.0 - phi2);
    mp.a1 = r * (1.0 - t * phi2 + 3.0 * phi2 / (1.0 - phi2));
    mp.a2 = r * (3.0 - t * (1.0 - phi2));
    double den = mp.a1 + mp.a2 + 1.0;
    mp.ct0 = phi2 * vd / (den * den);
  }
  return mp;
}

// FT of the autocorrelation function at wavenumber k given as k^2 (saves a sqrt for the exponential model)
SMRT_DEV double micro_ft(const MicroParams& mp, double k2) {
  if (mp.kind == MS_EXPONENTIAL) {
    double d = 1.0 + k2 * mp.p0 * mp.p0;
    return mp.c0 / (d * d);
  } else if (mp.kind == MS_INDEPENDENT_SPHERE) {
    const double X = sqrt(k2) * mp.p0;
    if (fabs(X) <= 1e-8) return mp.c0;  // np.isclose(X, 0)
    double s, c;
    sincos(X, &s, &c);
    const double b = (s - X * c) / (X * X * X);
    return mp.c0 * (9.0 * (b * b));
  } else if (mp.kind == MS_TEUBNER_STREY) {
    const double X = k2 * mp.p0 * mp.p0, Y = mp.a1;
    return mp.ct0 * (mp.c0 / ((1.0 + Y) * (1.0 + Y) + 2.0 * (1.0 - Y) * X + X * X));
  } else if (mp.kind == MS_UNIFIED_TS_1) {
    return mp.ct0 * (mp.c0 / ((1.0 + mp.p0 * mp.p0 * k2) * (1.0 + mp.a1 * mp.a1 * k2)));
  } else if (mp.kind == MS_UNIFIED_TS_2) {
    const double x1 = sqrt(k2) * mp.p0, r12 = mp.a1;
    return mp.ct0 * (mp.c0 / ((1.0 + (x1 - r12) * (x1 - r12)) * (1.0 + (x1 + r12) * (x1 + r12))));
  } else if (mp.kind == MS_SHS || mp.kind == MS_SHS_T) {
    double X = sqrt(k2) * mp.p0;  // k * d / 2
    if (fabs(X) <= 1e-3) return mp.ct0;  // np.isclose(X, 0, atol=1e-3): |X| <= atol (rtol * 0 = 0)
    double s, c;
    sincos(X, &s, &c);
    double sinc = s / X;
    double v = 3.0 * (sinc - c) / (X * X);  // sqrt(intersection volume) / vd
    double Psi = sinc / v;
    double A = mp.a1 + mp.a2 * Psi + c / v;
    double Bq = mp.f / (1.0 - mp.f) * X + s / v;
    return mp.c0 / (A * A + Bq * Bq);
  }
  return 0.0;
}

// sticky_hard_spheres.py:132-167 (compute_t used by the DMRT models). Returns false when there is no solution.
SMRT_DEV bool shs_compute_t(double f, double stickiness, double* t_out) {
  if (isinf(stickiness)) {
    *t_out = 0.0;
    return true;
  }
  double a = f / 12.0;
  double b = -(stickiness + f / (1.0 - f));
  double c = (1.0 + f / 2.0) / ((1.0 - f) * (1.0 - f));
  double discr2 = b * b - 4.0 * a * c;
  if (discr2 < 0.0) return false;
  double discr = sqrt(discr2);
  double t = (-b - discr) / (2.0 * a);
  double mhu = t * f * (1.0 - f);
  double mhulim = 1.0 + 2.0 * f;
  if (mhu > mhulim) {
    t = (-b + discr) / (2.0 * a);
    mhu = t * f * (1.0 - f);
  }
  if (mhu > mhulim) return false;
  *t_out = t;
  return true;
}

// ---------------------------------------------------------------------------------------------------- layer optics
struct LayerOptics {
  cplx eps_eff;
  double ks, ka;
  double iba_coeff;  // IBA only
  double kk;         // (2 k0 Re sqrt(eps_eff))^2, scale of the squared wavevector difference in the IBA phase
  double f;          // fractional volume after inversion
  int status;
};

// generic_mixing_formula.py:118-141 (spheres)
SMRT_DEV cplx polder_van_santen_spheres(double f, cplx e0, cplx eps) {
  cplx d = c_sub(eps, e0);
  cplx b = c_sub(c_sub(eps, c_scale(e0, 2.0)), c_scale(d, 3.0 * f));
  cplx cq = c_scale(c_mul(eps, e0), -1.0);
  cplx disc = c_sub(c_mul(b, b), c_scale(cq, 8.0));  // b^2 - 4 a c, a = 2
  cplx root = c_sqrt(disc);
  return c_scale(c_sub(root, b), 0.25);  // (-b + sqrt) / (2 a)
}

// generic_mixing_formula.py:131-141 (randomly oriented needles, Shokr 1998 eq. 18)
SMRT_DEV cplx polder_van_santen_needles(double f, cplx e0, cplx eps) {
  cplx d = c_sub(eps, e0);
  cplx b = c_sub(d, c_scale(d, 5.0 / 3.0 * f));
  cplx cq = c_scale(c_mul(eps, c_add(e0, c_scale(d, 1.0 / 3.0 * f))), -1.0);
  cplx disc = c_sub(c_mul(b, b), c_scale(cq, 4.0));  // b^2 - 4 a c, a = 1
  return c_scale(c_sub(c_sqrt(disc), b), 0.5);
}
// polder_van_santen for layer.inclusion_shape = spheres, random_needles or a mixture (generic_mixing_formula.py:88-141):
// incl = (weight of the spheres solution, weight of the needles solution, ...) or NULL = spheres
SMRT_DEV cplx polder_van_santen_shapes(double f, cplx e0, cplx eps, const double* incl) {
  if (!incl || (incl[0] == 1.0 && incl[1] == 0.0)) return polder_van_santen_spheres(f, e0, eps);
  if (incl[0] == 0.0 && incl[1] == 1.0) return polder_van_santen_needles(f, e0, eps);
  return c_add(c_scale(polder_van_santen_spheres(f, e0, eps), incl[0]), c_scale(polder_van_santen_needles(f, e0, eps), incl[1]));
}

// scipy.integrate.romb for 2^6 + 1 samples (emmodel/iba.py:176-180); y[65], dx = sample spacing
SMRT_DEV double romb65(const double* y, double dx) {
  double R[7][7];
  double h = 64.0 * dx;
  R[0][0] = (y[0] + y[64]) / 2.0 * h;
  int start = 64, step = 64;
  for (int i = 1; i <= 6; ++i) {
    start >>= 1;
    double s = 0.0;
    for (int j = start; j < 64; j += step) s += y[j];
    R[i][0] = 0.5 * (R[i - 1][0] + h * s);
    step >>= 1;
    for (int j = 1; j <= i; ++j) {
      double prev = R[i][j - 1];
      R[i][j] = prev + (prev - R[i - 1][j - 1]) / (double)((1 << (2 * j)) - 1);
    }
    h /= 2.0;
  }
  return R[6][6];
}

// emmodel/iba.py:85-137,139-162,168-226,246-265 ; emmodel/dmrt_qca_shortrange.py:65-112 ;
// emmodel/dmrt_qcacp_shortrange.py:63-125 ; emmodel/nonscattering.py:19-34
SMRT_DEV LayerOptics layer_optics(double frequency, double f, cplx e0, cplx eps, int emmodel, int ms_kind, double p0,
                                  double p1, int invert_dense, MicroParams* mp_out, const double* incl = nullptr) {
  LayerOptics o;
  o.status = ST_OK;
  o.iba_coeff = 0.0;
  o.kk = 0.0;
  if (emmodel == EM_RAYLEIGH) {  // emmodel/rayleigh.py:21-39: sparse medium, eps_eff = background; p0 = radius
    const double k0 = 2.0 * SMRT_PI / (SMRT_C_SPEED / frequency);
    const cplx e2 = c_add(eps, c_scale(e0, 2.0));
    const double a1 = c_abs(c_div(c_sub(eps, e0), e2)), a0 = c_abs(e0), a3 = c_abs(c_div(c_scale(e0, 3.0), e2));
    const double k02 = k0 * k0;
    o.f = f;
    o.eps_eff = e0;
    o.ks = f * 2.0 * (a1 * a1) * (p0 * p0 * p0) * (a0 * a0) * (k02 * k02);
    o.ka = f * k0 * eps.im * (a3 * a3) + (1.0 - f) * 2.0 * k0 * c_sqrt(e0).im;
    if (mp_out) *mp_out = micro_prepare(MS_HOMOGENEOUS, f, 0.0, 0.0);
    return o;
  }
  if (emmodel == EM_PRESCRIBED_KSKAEPS) {  // emmodel/prescribed_kskaeps.py:20-27: eps_bg = eps_eff, p0 = ks, p1 = ka
    o.f = f;
    o.eps_eff = e0;
    o.ks = p0;
    o.ka = p1;
    if (mp_out) *mp_out = micro_prepare(MS_HOMOGENEOUS, f, 0.0, 0.0);
    return o;
  }
  if (f > 0.5 && invert_dense && emmodel != EM_NONSCATTERING) {  // core/layer.py:186-201
    f = 1.0 - f;
    cplx tmp = e0;
    e0 = eps;
    eps = tmp;
  }
  o.f = f;
  MicroParams mp = micro_prepare(ms_kind, f, p0, p1);
  if (em_is_iba(emmodel)) {
    double k0 = 2.0 * SMRT_PI * frequency / SMRT_C_SPEED;
    // depolarisation factors of the three axes (iba.py:112-119): (1/3, 1/3, 1/3) for spheres
    const double Ax[3] = {incl ? incl[2] : 1.0 / 3.0, incl ? incl[3] : 1.0 / 3.0, incl ? incl[4] : 1.0 / 3.0};
    cplx de = c_sub(eps, e0);
    cplx eeff;
    if (emmodel == EM_IBA_MAXWELL_GARNETT) {
      // generic_mixing_formula.py:346-358: mean of the three components; iba_maxwell_garnett.py:47-51
      cplx acc = c_make(0.0, 0.0);
      for (int i = 0; i < 3; ++i) {
        cplx den = c_add(e0, c_scale(de, (1.0 - f) * Ax[i]));
        acc = c_add(acc, c_mul(e0, c_add(c_make(1.0, 0.0), c_div(c_scale(de, f), den))));
      }
      eeff = c_make(acc.re / 3.0, acc.im / 3.0);  // (x0 + x1 + x2) / 3, as numpy's mean evaluates it
    } else {
      eeff = polder_van_santen_shapes(f, e0, eps, incl);
    }
    // mean_sq_field_ratio (iba.py:150-162; apparent permittivity = background for Maxwell-Garnett)
    double ysum = 0.0;
    for (int i = 0; i < 3; ++i) {
      const cplx eapp = (emmodel == EM_IBA_MAXWELL_GARNETT) ? e0 : c_add(c_scale(eeff, 1.0 - Ax[i]), c_scale(e0, Ax[i]));
      const double ar = c_abs(c_div(eapp, c_add(eapp, c_scale(de, Ax[i]))));
      ysum += ar * ar;
    }
    double y2 = (1.0 / 3.0) * ysum;
    double ade = c_abs(de);
    double k02 = k0 * k0;
    o.iba_coeff = (1.0 / (4.0 * SMRT_PI)) * (ade * ade) * y2 * (k02 * k02);
    cplx n = c_sqrt(eeff);
    o.ka = (emmodel == EM_IBA_ORIGINAL) ? k0 * f * eps.im * fabs(y2)  // iba_original.py:43-44
                                        : 2.0 * k0 * n.im;
    // ks: Romberg on mu = linspace(1, -1, 65) of (iba_coeff*ft).real * mu^2 + (iba_coeff*ft).real
    double absn = c_abs(n);
    double y[65];
    for (int i = 0; i <= 64; ++i) {
      double mu = 1.0 - i * (2.0 / 64.0);
      if (i == 64) mu = -1.0;
      double sh = sqrt((1.0 - mu) / 2.0);
      double kd = 2.0 * k0 * sh * absn;
      double ft = micro_ft(mp, kd * kd);
      double v = o.iba_coeff * ft;
      y[i] = v * (mu * mu) + v * 1.0;
    }
    o.ks = romb65(y, 2.0 / 64.0) / 4.0;
    o.eps_eff = eeff;
    double kr = 2.0 * k0 * n.re;
    o.kk = kr * kr;
  } else if (emmodel == EM_DMRT_QCA_SR || emmodel == EM_DMRT_QCACP_SR) {
    double t;
    if (ms_kind != MS_SHS || !shs_compute_t(f, p1, &t)) {
      o.status = ST_INPUT;
      o.eps_eff = e0;
      o.ks = 0.0;
      o.ka = 0.0;
      if (mp_out) *mp_out = mp;
      return o;
    }
    double lmda = SMRT_C_SPEED / frequency;
    double radius = p0;
    double omf = 1.0 - f;
    double omf4 = (omf * omf) * (omf * omf);
    double q = 1.0 + 2.0 * f - t * f * omf;
    if (emmodel == EM_DMRT_QCA_SR) {
      cplx y = c_div(c_sub(eps, e0), c_add(eps, c_scale(e0, 2.0)));
      cplx fy = c_scale(y, f);
      cplx one = c_make(1.0, 0.0);
      double k0 = (2.0 * SMRT_PI / lmda) * c_sqrt(e0).re;
      double kr3 = (k0 * radius) * (k0 * radius) * (k0 * radius);
      // Eeff = e0 + 3 fy e0/(1-fy) * (1 + 2j/3 kr3 y (1-f)^4 / ((1-fy) q^2))
      cplx omfy = c_sub(one, fy);
      cplx inner = c_div(c_scale(y, kr3 * omf4), c_scale(omfy, q * q));  // kr3*y*(1-f)^4/((1-fy) q^2)
      cplx corr = c_add(one, c_mul(c_make(0.0, 2.0 / 3.0), inner));
      cplx lead = c_div(c_scale(c_mul(fy, e0), 3.0), omfy);
      cplx Eeff = c_add(e0, c_mul(lead, corr));
      cplx rel = c_sub(c_div(Eeff, e0), one);
      double ar = c_abs(rel);
      double Ks = 2.0 / (9.0 * f) * k0 * kr3 * ((ar * ar) * omf4 / (q * q));
      double beta = 2.0 * k0 * c_sqrt(Eeff).im;
      o.eps_eff = Eeff;
      o.ks = Ks;
      o.ka = beta - Ks;
    } else {
      cplx de = c_sub(eps, e0);
      cplx one = c_make(1.0, 0.0);
      cplx b = c_sub(c_scale(de, (1.0 - 4.0 * f) / 3.0), e0);
      cplx c = c_scale(c_mul(e0, de), -(1.0 - f) / 3.0);
      cplx disc = c_sub(c_mul(b, b), c_scale(c, 4.0));
      cplx sq = c_sqrt(disc);
      cplx Eeff0 = c_scale(c_sub(sq, b), 0.5);
      if (Eeff0.re < 1.0) Eeff0 = c_scale(c_add(sq, b), -0.5);
      double x = 2.0 * SMRT_PI * radius / lmda;
      double x3 = x * x * x;
      // g = (es - e0) / (1 + (es - e0) / (3 Eeff0) (1 - f))
      cplx g = c_div(de, c_add(one, c_scale(c_div(de, c_scale(Eeff0, 3.0)), omf)));
      cplx corr = c_mul(c_mul(c_make(0.0, 2.0 / 9.0 * x3), c_sqrt(Eeff0)), c_scale(g, omf4 / (q * q)));
      cplx Eeff = c_add(e0, c_mul(c_sub(Eeff0, e0), c_add(one, corr)));
      double imn = c_sqrt(Eeff).im;
      double ag = c_abs(g);
      double albedo = 2.0 / 9.0 * x3 * f / (2.0 * imn) * (ag * ag) * omf4 / (q * q);
      double beta = 2.0 * SMRT_PI / lmda * 2.0 * imn;
      o.eps_eff = Eeff;
      o.ks = albedo * beta;
      o.ka = beta - o.ks;
    }
  } else {  // EM_NONSCATTERING
    double k0 = 2.0 * SMRT_PI * frequency / SMRT_C_SPEED;
    cplx eeff = polder_van_santen_shapes(f, e0, eps, incl);
    o.eps_eff = eeff;
    o.ks = 0.0;
    o.ka = 2.0 * k0 * c_sqrt(eeff).im;
  }
  if (mp_out) *mp_out = mp;
  return o;
}

// ---------------------------------------------------------------------------------------------------- streams
// rtsolver/streams.py:136-223.  eps_eff: [nlayer][2].  Returns the index of the most refringent layer
// (np.argmax over complex = lexicographic (re, im), first occurrence).
SMRT_DEV int most_refringent_layer(const double* eps_eff, int nlayer) {
  int k = 0;
  double br = eps_eff[0], bi = eps_eff[1];
  for (int l = 1; l < nlayer; ++l) {
    double r = eps_eff[2 * l], i = eps_eff[2 * l + 1];
    if (r > br || (r == br && i > bi)) {
      k = l;
      br = r;
      bi = i;
    }
  }
  return k;
}

// number of streams of a medium with Re sqrt(eps*/eps_medium) = real_index: the kept streams are a prefix because
// relsin_j = real_index * sqrt(1 - mu*_j^2) increases with j (mu* descending)  — streams.py:182-194
SMRT_DEV int stream_count(double real_index, const double* gl_mu, int n) {
  int lo = 0, hi = n;  // first j with relsin >= 1
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    double relsin = real_index * sqrt(1.0 - gl_mu[mid] * gl_mu[mid]);
    if (relsin < 1.0)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
SMRT_DEV double stream_mu(double real_index, const double* gl_mu, int j) {
  double relsin = real_index * sqrt(1.0 - gl_mu[j] * gl_mu[j]);
  return sqrt(1.0 - relsin * relsin);
}
// streams.py:316-330: weights from node differences (NOT the Gauss weights). mu: the layer's nodes, n >= 2.
SMRT_DEV double stream_weight(const double* mu, int n, int j) {
  if (j == 0) return 1.0 - 0.5 * (mu[0] + mu[1]);
  if (j == n - 1) return fabs(0.5 * (mu[n - 2] + mu[n - 1]));
  return fabs(0.5 * (mu[j - 1] - mu[j + 1]));
}

// ---------------------------------------------------------------------------------------------------- Fresnel
// core/fresnel.py:99-146 (rigorous Maezawa & Miyauchi 2009), 417-474 (power matrices). Medium 1 holds the stream mu.
struct FresnelRT {
  double R[3], T[3];
};
// field reflection coefficients (rv, rh) and the cosine in medium 2: core/fresnel.py:99-146
SMRT_DEV void fresnel_field(cplx eps_1, cplx eps_2, double mu, cplx& rv, cplx& rh, double& mu2) {
  cplx n1 = c_sqrt(eps_1);
  double kiz2 = n1.re * n1.re * (1.0 - mu * mu);
  cplx kyi = c_scale(c_sqrt(c_make(eps_1.re - kiz2, eps_1.im)), -1.0);
  cplx kyt = c_scale(c_sqrt(c_make(eps_2.re - kiz2, eps_2.im)), -1.0);
  rh = c_div(c_sub(kyi, kyt), c_add(c_conj(kyi), kyt));
  cplx num = c_mul(c_conj(n1), c_sub(c_mul(eps_2, kyi), c_mul(eps_1, kyt)));
  cplx den = c_mul(n1, c_add(c_mul(eps_2, c_conj(kyi)), c_mul(c_conj(eps_1), kyt)));
  rv = c_div(num, den);
  mu2 = -kyt.re / c_sqrt(eps_2).re;
}
SMRT_DEV_NOINLINE FresnelRT fresnel_power(int kind, cplx eps_1, cplx eps_2, double mu) {
  FresnelRT o;
  if (kind == IF_TRANSPARENT) {  // interface/transparent.py:12-46
    o.R[0] = o.R[1] = o.R[2] = 0.0;
    o.T[0] = o.T[1] = o.T[2] = 1.0;
    return o;
  }
  cplx rv, rh;
  double mu2;
  fresnel_field(eps_1, eps_2, mu, rv, rh, mu2);
  o.R[0] = c_abs2(rv);
  o.R[1] = c_abs2(rh);
  o.R[2] = c_mul(rv, c_conj(rh)).re;
  o.T[0] = 1.0 - o.R[0];
  o.T[1] = 1.0 - o.R[1];
  cplx one = c_make(1.0, 0.0);
  o.T[2] = mu2 / mu * c_mul(c_add(one, rv), c_conj(c_add(one, rh))).re;
  return o;
}

// ------------------------------------------------------------------------------------------------------- substrates
// k sigma of the rough-surface models: Re(2 pi f sqrt((1 / 2.9979e8)^2 eps_1)) * roughness_rms
// (substrate/soil_wegmuller.py:29-30, rough_choudhury79.py:26-27; the reference's own rounded speed of light)
SMRT_DEV double substrate_ksigma(double freq, cplx eps_1, double roughness_rms) {
  const double ic = 1.0 / 2.9979e8, c2 = ic * ic;
  const cplx sq = c_sqrt(c_make(c2 * eps_1.re, c2 * eps_1.im));
  return (2.0 * SMRT_PI * freq) * sq.re * roughness_rms;
}
// in-place adjustment of the H and V power reflectivities of a rough substrate
SMRT_DEV void substrate_adjust(int kind, const double* par, double ksigma, double mu, double& rh, double& rv) {
  if (kind == SUB_SOIL_WEGMULLER) {  // soil_wegmuller.py:24-43
    rh *= exp(-pow(ksigma, sqrt(0.1 * mu)));
    if (mu < 0.5000000000000001)  // np.cos(60 * np.pi / 180)
      rv = rh * (0.635 - 0.0014 * (acos(mu) * 180.0 / SMRT_PI - 60.0));
    else
      rv = rh * pow(mu, 0.655);
  } else if (kind == SUB_ROUGH_CHOUDHURY) {  // rough_choudhury79.py:23-37 (validity checked by the caller)
    const double f = exp(-4.0 * (ksigma * ksigma) * (mu * mu));
    rh *= f;
    rv *= f;
  } else if (kind == SUB_SOIL_QNH) {  // soil_qnh.py:26-42; par = H, Q, Nv, Nh
    const double H = par[0], Q = par[1];
    const double coef_h = exp(-H * pow(mu, par[3])), coef_v = exp(-H * pow(mu, par[2]));
    const double trv = ((1.0 - Q) * rv + Q * rh) * coef_v;
    rh = ((1.0 - Q) * rh + Q * rv) * coef_h;
    rv = trv;
  }
}
// Diagonal diffuse reflection of a substrate with a prescribed backscattering coefficient sigma0 (linear), azimuth mode
// m of m_max, stream (mu, weight w) -- substrate/reflector_backscatter.py:90-116: the backscatter is spread over the
// 1 + 2 m_max modes with signs (+1, -2, +2, ...) and converted to scattering by 1 / (4 pi mu); rtsolver_utils.py:735-737
// multiplies by the weight and 690-709 by the mode integral (2 pi | pi): +- sigma0 w / (2 mu (1 + 2 m_max)).
SMRT_DEV double substrate_backscatter(double sigma0, int m, int m_max, double mu, double w) {
  const double sgn = (m & 1) ? -1.0 : 1.0;
  return sgn * 0.5 * sigma0 * w / ((double)(1 + 2 * m_max) * mu);
}

// Backscattering coefficients (VV, HH) of a moderately rough surface by the IEM of Fung et al. 1992 on the stream mu of
// medium 1 -- interface/iem_fung92.py:88-189: Kirchhoff and complementary terms, `series_truncation` terms of the series,
// exponential or Gaussian surface spectrum; brogioni: Fresnel coefficients at normal incidence when
// ks kl > sqrt(eps_r) (iem_fung92_brogioni10.py:45-54; complex numbers compare lexicographically in NumPy).
// par = roughness_rms, corr_length, autocorrelation (0 exponential, 1 gaussian), series_truncation.
SMRT_DEV_NOINLINE void iem_fung92_sigma0(double freq, cplx eps_1, cplx eps_2, double mu, const double* par,
                                         bool brogioni, double& svv, double& shh) {
  const double rms = par[0], lc = par[1];
  const bool gauss = par[2] != 0.0;
  const int N = (int)par[3];
  const double knorm = 2.0 * SMRT_PI * freq / SMRT_C_SPEED * c_sqrt(eps_1).re;
  const double mu2 = mu * mu, sin2 = 1.0 - mu2, tan2 = sin2 / mu2;
  const double kz = knorm * mu, kx = knorm * sqrt(sin2);
  const cplx eps_r = c_div(eps_2, eps_1);
  const cplx sq = c_sqrt(eps_r);
  const double kskl = fabs(knorm * rms) * fabs(knorm * lc);
  const bool nadir = brogioni && (kskl > sq.re || (kskl == sq.re && 0.0 > sq.im));
  cplx rv, rh;
  double mut;
  fresnel_field(eps_1, eps_2, nadir ? 1.0 : mu, rv, rh, mut);
  const cplx one = c_make(1.0, 0.0);
  const cplx fvv = c_scale(rv, 2.0 / mu), fhh = c_scale(rh, -2.0 / mu);
  const cplx inv_er = c_div(one, eps_r);
  const cplx opv = c_add(one, rv), oph = c_add(one, rh);
  const cplx cv = c_scale(c_mul(c_mul(c_mul(opv, opv), c_sub(one, inv_er)), c_add(one, c_scale(inv_er, tan2))), sin2 / mu);
  const cplx ch = c_scale(c_mul(c_mul(oph, oph), c_sub(eps_r, one)), sin2 / (mu * mu2));
  const double rms2 = rms * rms, e1 = exp(-rms2 * kz * kz);
  const double kql = -2.0 * kx * lc;
  double p1 = 1.0, p2 = 1.0, fact = 1.0, sv = 0.0, sh = 0.0;
  for (int n = 1; n <= N; ++n) {
    p1 *= kz;
    p2 *= 2.0 * kz;
    fact *= rms2 / (double)n;
    const cplx ivv = c_add(c_scale(fvv, p2 * e1), c_scale(cv, p1));
    const cplx ihh = c_sub(c_scale(fhh, p2 * e1), c_scale(ch, p1));
    const double ln = lc / (double)n;
    const double W = gauss ? (lc * lc / (2.0 * n)) * exp(-(kql * kql) / (4.0 * n))
                           : ln * ln * pow(1.0 + (kql / n) * (kql / n), -1.5);
    sv = fma(fact * W, c_abs2(ivv), sv);
    sh = fma(fact * W, c_abs2(ihh), sh);
  }
  const double coef = 0.5 * knorm * knorm * exp(-2.0 * rms2 * kz * kz);
  svv = coef * sv;
  shh = coef * sh;
}

// specular reflection and emissivity of the substrate under a layer of permittivity eps_1, on the stream mu:
// substrate/flat.py:15-17, soil_wegmuller.py:45-81, soil_qnh.py:44-89, reflector.py:51-81, rough_choudhury79.py:39-79.
// The third Stokes component keeps its Fresnel value (as in the reference).
// coherent reflection / transmission of a rough surface under the Kirchhoff approximation, every component:
// interface/interface_utils.py:21-64 (k2 carries |eps_1|^2 as in the reference)
SMRT_DEV void kirchhoff_coherent(FresnelRT& o, double rms, double freq, cplx eps_1, cplx eps_2, double mu) {
  const double k0 = 2.0 * SMRT_PI * freq / SMRT_C_SPEED, rms2 = rms * rms;
  const double fr = exp(-4.0 * (k0 * k0 * c_abs2(eps_1)) * rms2 * (mu * mu));
  const double k_iz = k0 * c_sqrt(eps_1).re * mu;
  const double s2 = 1.0 - mu * mu;
  const double k_sz = k0 * c_sqrt(c_make(eps_2.re - s2 * eps_1.re, eps_2.im - s2 * eps_1.im)).re;
  const double ft = exp(-((k_sz - k_iz) * (k_sz - k_iz)) * rms2);
  for (int p = 0; p < 3; ++p) {
    o.R[p] *= fr;
    o.T[p] *= ft;
  }
}
// Rough INTERFACE between two media (kinds IF_IEM_FUNG92 / _BRIOGONI10; par = roughness_rms, corr_length,
// autocorrelation, series_truncation): Kirchhoff coherent part and, when m_diff >= 0, the IEM backscatter as a diagonal
// diffuse reflection of azimuth mode m_diff on the stream (mu, w) -- interface/iem_fung92.py, interface_utils.py:16-64;
// the IEM has no diffuse transmission (rtsolver_utils.py:506-522: the attribute is missing, the matrix is zero).
SMRT_DEV_NOINLINE FresnelRT rough_interface_power(int kind, const double* par, double freq, cplx eps_1, cplx eps_2,
                                                  double mu, double w, int m_diff, int m_max) {
  FresnelRT o = fresnel_power(IF_FLAT, eps_1, eps_2, mu);
  kirchhoff_coherent(o, par[0], freq, eps_1, eps_2, mu);
  if (m_diff >= 0) {
    double svv, shh;
    iem_fung92_sigma0(freq, eps_1, eps_2, mu, par, kind == IF_IEM_FUNG92_BRIOGONI10, svv, shh);
    o.R[0] += substrate_backscatter(svv, m_diff, m_max, mu, w);
    o.R[1] += substrate_backscatter(shh, m_diff, m_max, mu, w);
  }
  return o;
}
// w, m_diff, m_max: weight of the stream, azimuth mode whose diffuse (backscatter) reflection is added to R for the
// substrates that have one (kinds 6 - 8; m_diff < 0: none, the coherent pass), number of modes it is spread over.
SMRT_DEV FresnelRT substrate_specular(int kind, const double* par, double freq, cplx eps_1, cplx eps_2, double mu);
SMRT_DEV_NOINLINE FresnelRT substrate_power(int kind, const double* par, double freq, cplx eps_1, cplx eps_2, double mu,
                                            double w = 0.0, int m_diff = -1, int m_max = 0) {
  FresnelRT o = substrate_specular(kind, par, freq, eps_1, eps_2, mu);
  if (m_diff >= 0 && par && kind >= SUB_REFLECTOR_BACKSCATTER && kind <= SUB_IEM_FUNG92_BRIOGONI10) {
    double svv = par[2], shh = par[3];
    if (kind != SUB_REFLECTOR_BACKSCATTER) iem_fung92_sigma0(freq, eps_1, eps_2, mu, par, kind == SUB_IEM_FUNG92_BRIOGONI10, svv, shh);
    o.R[0] += substrate_backscatter(svv, m_diff, m_max, mu, w);
    o.R[1] += substrate_backscatter(shh, m_diff, m_max, mu, w);
  }
  return o;
}
SMRT_DEV FresnelRT substrate_specular(int kind, const double* par, double freq, cplx eps_1, cplx eps_2, double mu) {
  FresnelRT o;
  const double zero4[4] = {0.0, 0.0, 0.0, 0.0};
  if (!par) par = zero4;
  if (kind == SUB_REFLECTOR || kind == SUB_REFLECTOR_BACKSCATTER) {  // (the diffuse part of kind 6: substrate_backscatter)
    o.R[0] = par[0];
    o.R[1] = par[1];
    o.R[2] = 0.0;
    o.T[0] = 1.0 - par[0];
    o.T[1] = 1.0 - par[1];
    o.T[2] = 0.0;
    return o;
  }
  o = fresnel_power(IF_FLAT, eps_1, eps_2, mu);
  if (kind == SUB_FLAT) return o;
  if (kind == SUB_IEM_FUNG92 || kind == SUB_IEM_FUNG92_BRIOGONI10) {  // the emissivity is the coherent transmission
    kirchhoff_coherent(o, par[0], freq, eps_1, eps_2, mu);
    return o;
  }
  const double ksigma =
      (kind == SUB_SOIL_WEGMULLER || kind == SUB_ROUGH_CHOUDHURY) ? substrate_ksigma(freq, eps_1, par[0]) : 0.0;
  substrate_adjust(kind, par, ksigma, mu, o.R[1], o.R[0]);
  double rh = 1.0 - o.T[1], rv = 1.0 - o.T[0];
  substrate_adjust(kind, par, ksigma, mu, rh, rv);
  o.T[1] = 1.0 - rh;
  o.T[0] = 1.0 - rv;
  return o;
}

// ---------------------------------------------------------------------------------------------------- phase matrix
// Fourier mode m (in azimuth) of the IBA phase matrix for ONE pair of streams — emmodel/common.py:56-131, 349-414 with
// emmodel/iba.py:228-244 and common.py:9-53 restated as cosine / sine sums over the same nsamples/2+1 azimuth samples
// (SURVEY.md §8 a5).  ctab/stab hold cos(pi j / K), sin(pi j / K) for j in [0, 2K).
//   out[ps*npol + pi], ps = scattered polarisation, pi = incident polarisation; npol = 2 (m == 0) or 3 (m > 0).
SMRT_DEV void iba_phase_mode(int m, int K, const double* ctab, const double* stab, double mu_s, double mu_i,
                             double iba_coeff, double kk, const MicroParams& mp, double* out) {
  const int npol = (m == 0) ? 2 : 3;
  double sin_s = sqrt(1.0 - mu_s * mu_s);
  double sin_i = sqrt(1.0 - mu_i * mu_i);
  double mm = mu_s * mu_i, ss = sin_s * sin_i;
  double acc[9];
  for (int e = 0; e < 9; ++e) acc[e] = 0.0;
  const int twoK = 2 * K;
  for (int k = 0; k <= K; ++k) {
    double cphi = ctab[k], sphi = stab[k];
    double cosT = mm + ss * cphi;
    cosT = fmin(1.0, fmax(-1.0, cosT));
    double sh2 = 0.5 * (1.0 - cosT);  // sin^2(Theta / 2)
    double ft = micro_ft(mp, kk * sh2) * iba_coeff;
    double fvv = cphi * mm + ss;
    double fhh = cphi;
    double fvh = sphi * mu_s;
    double fhv = -sphi * mu_i;
    int mk = (m * k) % twoK;
    double wend = (k == 0 || k == K) ? 1.0 : 2.0;
    double wc = wend * ctab[mk] * ft;
    acc[0 * npol + 0] += wc * (fvv * fvv);
    acc[0 * npol + 1] += wc * (fvh * fvh);
    acc[1 * npol + 0] += wc * (fhv * fhv);
    acc[1 * npol + 1] += wc * (fhh * fhh);
    if (npol == 3) {
      acc[8] += wc * (fvv * fhh + fvh * fhv);
      double ws = 2.0 * stab[mk] * ft;  // mirrored samples double the sine sums; zero at k = 0 and k = K
      acc[0 * 3 + 2] -= ws * (fvh * fvv);
      acc[1 * 3 + 2] -= ws * (fhh * fhv);
      acc[2 * 3 + 0] += ws * (2.0 * (fvv * fhv));
      acc[2 * 3 + 1] += ws * (2.0 * (fvh * fhh));
    }
  }
  double scale = ((m == 0) ? 1.0 : 2.0) / (double)twoK;
  for (int e = 0; e < npol * npol; ++e) out[e] = acc[e] * scale;
}

// Analytic Fourier modes m = 0, 1, 2 of the Rayleigh phase matrix times 1.5 ks — emmodel/rayleigh.py:52-127
SMRT_DEV void rayleigh_phase_mode(int m, double mu_s, double mu_i, double ks, double* out) {
  const int npol = (m == 0) ? 2 : 3;
  double mus2 = mu_s * mu_s, mui2 = mu_i * mu_i;
  double coef = 3.0 * ks / 2.0;
  if (m == 0) {
    out[0] = coef * (0.5 * mus2 * mui2 + (1.0 - mus2) * (1.0 - mui2));
    out[1] = coef * (0.5 * mus2);
    out[2] = coef * (0.5 * mui2);
    out[3] = coef * 0.5;
    return;
  }
  for (int e = 0; e < 9; ++e) out[e] = 0.0;
  if (m == 1) {
    double sint_s = sqrt(1.0 - mus2), sint_i = sqrt(1.0 - mui2);
    double cs_s = mu_s * sint_s, cs_i = mu_i * sint_i;
    out[0] = 2.0 * cs_s * cs_i;
    out[2] = -(cs_s * sint_i);  // (v,u) with the sign flip of rayleigh.py:120-122
    out[6] = -2.0 * sint_s * cs_i;
    out[8] = sint_s * sint_i;
  } else if (m == 2) {
    out[0] = 0.5 * mus2 * mui2;
    out[1] = -0.5 * mus2;
    out[2] = -(0.5 * mus2 * mu_i);
    out[3] = -0.5 * mui2;
    out[4] = 0.5;
    out[5] = -(-0.5 * mu_i);
    out[6] = -mu_s * mui2;
    out[7] = mu_s;
    out[8] = mu_s * mu_i;
  }
  for (int e = 0; e < npol * npol; ++e) out[e] *= coef;
}
