"""TEST INFRASTRUCTURE ONLY — NumPy model of the algorithm the CUDA kernels implement.

The reference solves, per layer and azimuth mode, a non-symmetric N x N eigenproblem with LAPACK dgees/dgeev and then a
banded LU of the whole multi-layer boundary system.  The device path computes the same quantities differently:

  (1) half-rank reduction (the reference's own ``diagonalize_half_rank_eig``, smrt/rtsolver/dort.py:891-962) followed
      by an exact *symmetrising similarity*: with row scale d_a = norm_a/mu_a, column scale c_a = coef*w_a and the
      reciprocity weights q = (1, 1, 2) of the (V, H, U) Stokes components, S = diag(sqrt(d q / c)) turns
      alpha and beta*D into symmetric matrices, so (alpha - beta D)(alpha + beta D) ~ X_minus X_plus with X+- symmetric
      positive definite.  Cholesky X- = L L^T, X+ = C C^T gives k = singular values of M = C^T L; with M V = U Sigma:
      E+ = S C^-T (U Sigma),  E- = -S C U.   The device obtains U Sigma by one-sided Jacobi on M.
  (2) bottom-up elimination of the block-tridiagonal boundary system carrying only the h x h reflection operator
      R_l and the source vector s_l of the stack below each layer (only x_0 is needed, dort.py:472-476); each layer
      step is two h x h Gauss-Jordan eliminations and three h x h products (see solve_mode).

This file exists so the algebra is checked against the oracle on the CPU (tests/test_b200_algorithm_model.py) before
and independently of the CUDA implementation.  It is never imported by the product.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg

from . import dort_oracle as O


def layer_eigen_symmetric(P_half, mu, weight, ks, ke, m, norm_rows=None, normalization=True):
    """Half-rank symmetric eigen-solve of one layer and one mode.

    P_half: (h, 2h) compressed Fourier mode m of the phase matrix for scattered mu_s > 0 rows, all incident columns.
    Returns k (h,), F, G (h, h) with Eu = [F | G], Ed = D [G | F], beta = [k, -k]; and the row norms used.
    """
    npol = 2 if m == 0 else 3
    h = npol * len(mu)
    mu_a = np.repeat(mu, npol)
    w_a = np.repeat(weight, npol)
    coef = 0.5 if m == 0 else 0.25
    q = np.tile(np.array([1.0, 1.0, 2.0])[:npol], len(mu))
    Dsign = np.tile(np.array([1.0, 1.0, -1.0])[:npol], len(mu))

    Ppp = P_half[:, :h]
    Ppm = P_half[:, h:] * Dsign[None, :]  # beta * D

    if normalization and ks != 0:
        if m == 0:
            rowsum = -(coef * (P_half * np.tile(w_a, 2)[None, :]).sum(axis=1))
            norm = -ks / rowsum
            if normalization != "forced" and np.any(np.abs(norm - 1.0) > 0.3):
                raise O.OracleError(O.ST_NORMALIZATION, "normalization > 30 %")
        else:
            norm = norm_rows
    else:
        norm = np.ones(h)

    g = np.sqrt(norm * q * coef * w_a / mu_a)
    s = np.sqrt(norm * q / (mu_a * coef * w_a))
    Gpp = g[:, None] * (Ppp / q[:, None]) * g[None, :]
    Gpm = g[:, None] * (Ppm / q[:, None]) * g[None, :]
    asym = max(np.abs(Gpp - Gpp.T).max(), np.abs(Gpm - Gpm.T).max()) / max(np.abs(Gpp).max(), 1e-300)
    Gpp = 0.5 * (Gpp + Gpp.T)
    Gpm = 0.5 * (Gpm + Gpm.T)
    Xm = np.diag(ke / mu_a) - Gpp + Gpm
    Xp = np.diag(ke / mu_a) - Gpp - Gpm
    try:
        Lc = np.linalg.cholesky(Xm)
        Cc = np.linalg.cholesky(Xp)
    except np.linalg.LinAlgError:
        raise O.OracleError(O.ST_EIGEN, "X+- not positive definite")
    M = Cc.T @ Lc
    U, sig, Vt = np.linalg.svd(M)
    W = U * sig[None, :]
    Ep = s[:, None] * scipy.linalg.solve_triangular(Cc.T, W, lower=False)
    Em = -s[:, None] * (Cc @ U)
    F = 0.5 * (Ep - Em)
    G = 0.5 * (Ep + Em)
    return sig, F, G, norm, asym


def mode_norm_rows(norm0, npol_m):
    """dort.py:803-816: reuse the mode-0 row norms for m > 0 (U = geometric mean of V and H)"""
    n = len(norm0) // 2
    out = np.empty(n * npol_m)
    out[0::npol_m] = norm0[0::2]
    out[1::npol_m] = norm0[1::2]
    if npol_m == 3:
        out[2::3] = np.sqrt(norm0[0::2] * norm0[1::2])
    return out


def solve_mode(problem, mode, streams, layers, iface, intensity_down, planck, coherent_only=False,
               prune_deep_snowpack=None, info=None):
    """Boundary system of one azimuth mode by bottom-up reflection-operator recursion.

    layers[l] = dict(k, F, G) in the half-rank representation (or no-scattering: F = I, G = 0, k = ke/mu).
    """
    Rtop_, Ttop_, Rbot_, Tbot_ = iface
    npol = 2 if mode == 0 else 3
    ns = streams["n"]
    L = len(ns)
    thickness = problem["thickness"]
    temperature = problem["temperature"] if problem["mode"] == "P" else None
    nrhs = intensity_down.shape[1]

    def cdiag(mat, l):
        return None if mat[l] is None else O.compress_diag(mat[l], mode)

    def lay(l):
        if coherent_only or layers[l][mode] is None:
            h = npol * ns[l]
            ke = layers[l]["ke"]
            return ke / np.repeat(streams["mu"][l], npol), np.eye(h), np.zeros((h, h))
        d = layers[l][mode]
        return d["k"], d["F"], d["G"]

    # optical depth, top-down, to find the last layer kept (dort.py:444-452)
    optical_depth = 0.0
    l_end = L - 1
    for l in range(L):
        k, _, _ = lay(l)
        optical_depth += np.min(np.abs(k)) * thickness[l]
        if prune_deep_snowpack is not None and optical_depth > prune_deep_snowpack:
            l_end = l
            break
    if info is not None:
        info["optical_depth"] = optical_depth
        info["shallow"] = bool(problem.get("substrate_kind", 0) == 0 and optical_depth < 5)

    Dsign = lambda h: np.tile(np.array([1.0, 1.0, -1.0])[:npol], h // npol)  # noqa: E731
    Rop = None
    svec = None
    for l in range(l_end, -1, -1):
        h = npol * ns[l]
        k, F, G = lay(l)
        t = np.exp(-k * thickness[l])
        D = Dsign(h)
        Rt = cdiag(Rtop_, l)
        Rb = cdiag(Rbot_, l)
        if Rb is None:
            Rb = np.zeros(h)
        # diagonal diffuse (backscatter) reflection of rough interfaces / substrates
        Rb = O.with_diffuse_reflection(Rb, problem, streams, mode, coherent_only, "bottom", l)
        Rt = O.with_diffuse_reflection(Rt, problem, streams, mode, coherent_only, "top", l)
        Rbm = np.diag(Rb)  # effective bottom reflection: interface + (T R T) of the stack below
        b_top = np.zeros((h, nrhs))
        b_bot = np.zeros((h, nrhs))
        Tl = temperature[l] if temperature is not None else None
        if mode == 0 and Tl is not None and Tl > 0:
            b_top -= ((1.0 - Rt) * planck(Tl))[:, None]
            b_bot -= ((1.0 - Rb) * planck(Tl))[:, None]
        # contribution of the layer above (l-1) into the top rows of l
        if l > 0 and mode == 0 and temperature is not None and temperature[l - 1] > 0:
            Tb_lm1 = cdiag(Tbot_, l - 1)
            r = min(len(Tb_lm1), h)
            b_top[:r] += (Tb_lm1 * planck(temperature[l - 1]))[:r, None]
        if l == 0:
            Tair = O.compress_diag(Tbot_[-1], mode)
            r = min(len(Tair), h)
            b_top[:r] += (Tair[:, None] * intensity_down)[:r]
        # contribution of the layer below (l+1) into the bottom rows of l
        if l < l_end:
            Tt_lp1 = cdiag(Ttop_, l + 1)
            r = min(len(Tt_lp1), h)
            if mode == 0 and temperature is not None and temperature[l + 1] > 0:
                b_bot[:r] += (Tt_lp1 * planck(temperature[l + 1]))[:r, None]
            Tb_l = cdiag(Tbot_, l)
            Rbm[:r, :r] += Tt_lp1[:r, None] * Rop[:r, :r] * Tb_l[None, :r]
            b_bot[:r] += Tt_lp1[:r, None] * svec[:r]
        elif (l == L - 1 and mode == 0 and problem.get("substrate_kind", 0) != 0 and temperature is not None):
            Tsub = cdiag(Tbot_, l)
            b_bot += (Tsub * planck(problem["substrate_temperature"]))[:, None]
        # Unknowns x+ (modes referenced at the bottom), x- (referenced at the top); with Eu = [F | G], Ed = D [G | F]:
        #   bottom rows: A21 x+ + A22 x- = b_bot,  A21 = F - Rb' D G,      A22 = (G - Rb' D F) t
        #   top rows   : A11 x+ + A12 x- = b_top,  A11 = (D G - Rt F) t,   A12 = D F - Rt G
        # Gauss-Jordan on [A21 | A22 | b_bot] gives Y22 = A21^-1 A22, Yr = A21^-1 b_bot; with Y~ = t Y22, y~r = t Yr:
        #   P = F - G Y~,  K = G - F Y~,  Schur S = D P - Rt K,  b' = b_top - D G y~r + Rt F y~r
        #   reflection operator of the stack seen from above  R = K S^-1,  source  s = F y~r + R b'
        DG = D[:, None] * G
        DF = D[:, None] * F
        Tm = np.hstack((F - Rbm @ DG, (G - Rbm @ DF) * t[None, :], b_bot))
        try:
            sol = np.linalg.solve(Tm[:, :h], Tm[:, h:])
            Yt = t[:, None] * sol[:, :h]
            yr = t[:, None] * sol[:, h:]
            P = F - G @ Yt
            K = G - F @ Yt
            S = D[:, None] * P - Rt[:, None] * K
            v = F @ yr
            bp = b_top - D[:, None] * (G @ yr) + Rt[:, None] * v
            if l > 0:
                Rop = np.linalg.solve(S.T, K.T).T
                svec = v + Rop @ bp
            else:
                svec = v + K @ np.linalg.solve(S, bp)
        except np.linalg.LinAlgError:
            raise O.OracleError(O.ST_SINGULAR, "singular boundary block")

    I1up = svec
    if mode == 0 and temperature is not None and temperature[0] > 0:
        I1up = I1up + planck(temperature[0])
    Rair = O.with_diffuse_reflection(O.compress_diag(Rbot_[-1], mode), problem, streams, mode, coherent_only, "air")
    Ttop0 = O.compress_diag(Ttop_[0], mode)
    n_air = streams["n_air"]
    I0up = Rair[:, None] * intensity_down + (Ttop0[:, None] * I1up)[0:n_air * npol, :]
    I0up = np.array(I0up).squeeze()
    if np.ndim(I0up) == 1:
        return I0up.reshape((I0up.shape[0] // npol, npol)).transpose()
    return I0up.reshape((I0up.shape[0] // npol, npol, I0up.shape[1] // npol, npol)).transpose(1, 0, 3, 2)


def solve_problem(problem, collect=None):
    """Same contract as ``dort_oracle.solve_problem`` but through the device algorithm."""
    opts = dict(n_max_stream=32, m_max=2, phase_normalization="auto", prune_deep_snowpack=None,
                rayleigh_jeans_approximation=False, error_handling="exception")
    opts.update(problem.get("options", {}))
    freq = float(problem["frequency"])
    mode = problem["mode"]
    L = len(problem["thickness"])
    theta = np.atleast_1d(np.asarray(problem["theta"], dtype=float))
    dsc = problem.get("dense_snow_correction")
    if dsc is None:
        dsc = np.isin(np.asarray(problem["emmodel"]), (O.EM_DMRT_QCA_SR, O.EM_DMRT_QCACP_SR)).astype(int)
    optics = [O.layer_optics(freq, problem["frac_volume"][l], problem["eps_bg"][l], problem["eps_sc"][l],
                             int(problem["emmodel"][l]), int(problem["ms_kind"][l]), problem["ms_p0"][l],
                             problem["ms_p1"][l], bool(dsc[l]),
                           None if problem.get("inclusion") is None else problem["inclusion"][l]) for l in range(L)]
    eps_eff = np.array([o["eps_eff"] for o in optics])
    out = dict(status=O.ST_OK, eps_eff=eps_eff, ks=np.array([o["ks"] for o in optics]),
               ka=np.array([o["ka"] for o in optics]))
    streams = O.compute_streams(int(opts["n_max_stream"]), eps_eff)
    m_max = int(opts["m_max"]) if mode == "A" else 0
    npol = 2 if mode == "P" else 3
    iface = O.compute_interfaces(problem, eps_eff, streams, npol)
    problem = dict(problem, _m_max=m_max, _eps_eff=eps_eff)
    norm = opts["phase_normalization"]
    if norm == "auto":
        norm = True

    layers = []
    max_asym = 0.0
    for l in range(L):
        o = optics[l]
        mu, w = streams["mu"][l], streams["weight"][l]
        entry = {"ke": o["ks"] + o["ka"]}
        for m in range(m_max + 1):
            entry[m] = None
        if o["ks"] != 0:
            fullmu = np.concatenate((mu, -mu))
            if o["emmodel"] in O.EM_IBA_FAMILY:
                P5 = O.iba_ft_even_phase(o, mu, fullmu, m_max, npol)
            else:
                P5 = O.rayleigh_ft_even_phase(o["ks"], mu, fullmu, m_max)
            norm0 = None
            for m in range(m_max + 1):
                P_half = O.compress_dense(P5, m)
                if not np.any(P_half):
                    continue
                nr = None if m == 0 else mode_norm_rows(norm0, 3) if norm0 is not None else None
                k, F, G, nrm, asym = layer_eigen_symmetric(P_half, mu, w, o["ks"], entry["ke"], m, nr, norm)
                max_asym = max(max_asym, asym)
                if m == 0:
                    norm0 = nrm if (norm and o["ks"] != 0) else None
                entry[m] = dict(k=k, F=F, G=G)
        layers.append(entry)
    if collect is not None:
        collect["max_asym"] = max_asym
        collect["layers"] = layers
        collect["streams"] = streams

    if opts["rayleigh_jeans_approximation"]:
        planck = lambda T: T  # noqa: E731
        inv_planck = lambda I: I  # noqa: E731
    else:
        planck = lambda T: O.planck_function(freq, T)  # noqa: E731
        inv_planck = lambda I: O.inverse_planck_function(freq, I)  # noqa: E731
    n_air = streams["n_air"]
    info = {}
    kw = dict(problem=problem, streams=streams, layers=layers, iface=iface, planck=planck,
              prune_deep_snowpack=opts["prune_deep_snowpack"], info=info)
    if mode == "P":
        atmos = problem.get("atmosphere")  # isotropic atmosphere, as in dort_oracle.solve_problem
        idown = np.zeros((2 * n_air, 1))
        if atmos is not None:
            idown[:] = planck(float(atmos[0]))
        I = solve_mode(mode=0, intensity_down=idown, **kw)
        intensity_up = I[0:2].copy()
        if atmos is not None:
            intensity_up = planck(float(atmos[1])) + float(atmos[2]) * intensity_up
        intensity_up = inv_planck(intensity_up)
        outmu = streams["outmu"]
    else:
        inc = O.prepare_incident_streams(streams["outmu"], theta)
        i0 = np.zeros((2 * n_air, 2 * len(inc)))
        ih = np.zeros((3 * n_air, 3 * len(inc)))
        j0 = jh = 0
        for i in inc:
            power = 1.0 / (2 * np.pi * streams["outweight"][i])
            for ipol in (0, 1):
                i0[2 * i + ipol, j0] = power
                j0 += 1
            for ipol in (0, 1, 2):
                ih[3 * i + ipol, jh] = 2 * power
                jh += 1
        intensity_up = np.zeros((3, n_air, 3, len(inc)))
        coh = solve_mode(mode=0, intensity_down=i0, coherent_only=True, **kw)
        phi = float(problem.get("phi", np.pi))
        for m in range(m_max + 1):
            I = solve_mode(mode=m, intensity_down=i0 if m == 0 else ih, **kw)
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
        out["status"] |= O.ST_SHALLOW_WARNING
    out["optical_depth"] = info.get("optical_depth")
    out["stream_angles"] = np.rad2deg(np.arccos(outmu))
    out["values"] = O.interpolate_intensity(mode, outmu, intensity_up, theta)
    return out
