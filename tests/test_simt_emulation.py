"""The CUDA device code (smrt_b200/csrc/*.cuh) compiled for the HOST through the SIMT emulator (tests/simt_emu) and
checked against the reference fixtures and the oracle.  This is what keeps the kernels honest in the authoring
container, where nvcc cross-compiles but no GPU exists; the `-m gpu` tests repeat the comparison on the B200."""
import ctypes as C

import numpy as np
import pytest

from emu_util import emu_lib, emu_solve, load_golden, rel_err
from oracle import dort_oracle as O

# a representative subset of the reference-generated fixtures (every one of them runs on the GPU: test_gpu_parity.py);
# one problem each: the emulator spends its time in pthread barriers
SMALL = ["cfg1_iba_onelayer", "ref_iba_2layer_passive", "ref_dmrt_qcacp_2layer_passive", "nonscattering_transparent",
         "iba_options_prune_rj", "iba_exp_substrate_passive", "soil_wegmuller_passive", "reflector_passive",
         "atmosphere_passive", "ref_physics_law", "iba_microstructures_passive", "prescribed_kskaeps_passive",
         "ref_iba_original_2layer_passive", "iba_maxwell_garnett_passive", "emmodel_per_medium_passive",
         "inclusion_shapes_passive", "iem_fung92_interface_passive"]
SMALL_ACTIVE = ["ref_dmrt_less_refringent_active", "rayleigh_active",
                "depolarization_active", "ref_rayleigh_mmax6_active", "iem_fung92_active",
                "iem_fung92_interface_active"]


@pytest.mark.parametrize("name", SMALL + SMALL_ACTIVE)
def test_emulated_kernels_match_reference_fixture(name):
    d, batch, opts = load_golden(name)
    batch = batch.subset(slice(0, 1))
    out = emu_solve(batch, opts, threads=64)
    ref = d["ref_values"][:batch.B]
    tol = 1e-10 if batch.mode == 0 else 1e-6
    assert np.all((out.status & 15) == 0)
    assert rel_err(out.values, ref, batch.mode) <= tol
    for b in range(batch.B):
        n = batch.nlayer[b]
        np.testing.assert_allclose(out.ks[b, :n], d["ref_ks"][b, :n], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(out.ka[b, :n], d["ref_ka"][b, :n], rtol=1e-12)
        np.testing.assert_allclose(out.eps_eff[b, :n], d["ref_eps_eff"][b, :n], rtol=1e-13)
        assert out.n_streams[b] == d["ref_n_air"][b]


@pytest.mark.parametrize("name", ["ref_iba_2layer_passive", "iba_exp_substrate_passive", "nonscattering_transparent",
                                  "atmosphere_passive", "ref_dmrt_less_refringent_active"])
def test_emulated_boundary_kernel_with_staged_operands(name, monkeypatch):
    """the boundary instantiation that keeps only [T | R] resident and stages F / G into them (two CTAs per SM on the
    device): TMA copies for even block sizes, thread copies for odd ones, generated operands for non-scattering layers"""
    monkeypatch.setenv("SMRT_EMU_STREAM_FG", "1")
    d, batch, opts = load_golden(name)
    batch = batch.subset(slice(0, 2))
    for threads in (64, 128):
        out = emu_solve(batch, opts, threads=threads)
        assert np.all((out.status & 15) == 0)
        assert rel_err(out.values, d["ref_values"][:batch.B], batch.mode) <= (1e-10 if batch.mode == 0 else 1e-6)


def test_emulated_kernels_full_warp_block():
    """same code with 256 threads per block (the launch configuration used on the GPU)"""
    d, batch, opts = load_golden("ref_iba_2layer_passive")
    out = emu_solve(batch, opts, threads=256)
    assert rel_err(out.values, d["ref_values"], 0) <= 1e-10


def test_gauss_legendre_nodes_match_scipy():
    lib = emu_lib()
    for n in (2, 16, 32, 64, 128):
        mu = np.zeros(n)
        lib.emu_gauss_legendre(n, mu.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(mu, O.gauss_legendre_quadrature(n), rtol=0, atol=3e-16)


def test_layer_optics_device_functions():
    lib = emu_lib()
    lib.emu_layer_optics.argtypes = [C.c_double] * 6 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                                        C.POINTER(C.c_double)]
    rng = np.random.default_rng(3)
    for em, ms in ((0, 0), (0, 1), (1, 1), (3, 1), (2, 2)):
        for _ in range(5):
            f = rng.uniform(0.1, 0.45)
            freq = rng.choice([6.925e9, 36.5e9, 89e9])
            es = O.ice_permittivity_maetzler06(freq, rng.uniform(240, 272))
            p0 = rng.uniform(5e-5, 3e-4)
            out = np.zeros(5)
            st = lib.emu_layer_optics(freq, f, 1.0, 0.0, es.real, es.imag, em, ms, p0, 0.2, 1,
                                      out.ctypes.data_as(C.POINTER(C.c_double)))
            assert st == 0
            ref = O.layer_optics(freq, f, 1.0, es, em, ms, p0, 0.2, True)
            np.testing.assert_allclose(out[0] + 1j * out[1], ref["eps_eff"], rtol=1e-14)
            np.testing.assert_allclose(out[2], ref["ks"], rtol=1e-13)
            np.testing.assert_allclose(out[3], ref["ka"], rtol=1e-13)


def test_fresnel_device_function():
    lib = emu_lib()
    lib.emu_fresnel.argtypes = [C.c_int] + [C.c_double] * 5 + [C.POINTER(C.c_double)]
    rng = np.random.default_rng(4)
    for _ in range(20):
        e1 = complex(rng.uniform(1, 3.2), rng.uniform(0, 0.01))
        e2 = complex(rng.uniform(1, 80), rng.uniform(0, 40))
        mu = rng.uniform(0.05, 1.0)
        out = np.zeros(6)
        lib.emu_fresnel(0, e1.real, e1.imag, e2.real, e2.imag, mu, out.ctypes.data_as(C.POINTER(C.c_double)))
        R, T = O.interface_R_T(O.IF_FLAT, e1, e2, np.array([mu]), 3)
        np.testing.assert_allclose(out[:3], R[:, 0], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(out[3:], T[:, 0], rtol=1e-12, atol=1e-15)


def test_iba_phase_fourier_modes_device_function():
    """cosine / sine sums over the azimuth samples == the reference's FFT-based decomposition, m = 0, 1, 2"""
    lib = emu_lib()
    lib.emu_iba_phase_mode.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int] + [C.c_double] * 3 + [
        C.POINTER(C.c_double)]
    opt = O.layer_optics(36.5e9, 0.3, 1.0, O.ice_permittivity_maetzler06(36.5e9, 260.0), O.EM_IBA, O.MS_EXPONENTIAL,
                         2e-4, 0.0)
    kk = (2 * opt["k0"] * np.sqrt(opt["eps_eff"]).real) ** 2
    mus = np.array([0.93, 0.41])
    mui = np.array([0.77, -0.35, 0.12])
    for m_max, npol in ((0, 2), (2, 3)):
        P5 = O.iba_ft_even_phase(opt, mus, mui, m_max, npol)
        for m in range(m_max + 1):
            np_m = 2 if m == 0 else 3
            for a, ms_ in enumerate(mus):
                for b, mi_ in enumerate(mui):
                    out = np.zeros(9)
                    lib.emu_iba_phase_mode(m, m_max, ms_, mi_, opt["iba_coeff"], kk, 0, 0.3, 2e-4, 0.0,
                                           out.ctypes.data_as(C.POINTER(C.c_double)))
                    ref = P5[:np_m, :np_m, m, a, b]
                    np.testing.assert_allclose(out[:np_m * np_m].reshape(np_m, np_m), ref, rtol=1e-11,
                                               atol=1e-14 * np.abs(P5).max())


@pytest.mark.parametrize("h,threads", [(8, 64), (33, 64), (64, 256)])
def test_one_sided_jacobi_device_function(h, threads):
    lib = emu_lib()
    rng = np.random.default_rng(h)
    M = rng.normal(size=(h, h)) + np.diag(rng.uniform(1, 5, h))
    W = np.asfortranarray(M.copy())
    sweeps = lib.emu_jacobi(W.ctypes.data_as(C.POINTER(C.c_double)), h, threads)
    assert 0 < sweeps < 20
    sig = np.sort(np.linalg.norm(W, axis=0))[::-1]
    np.testing.assert_allclose(sig, np.linalg.svd(M, compute_uv=False), rtol=1e-13)
    U = W / np.linalg.norm(W, axis=0)
    assert np.abs(U.T @ U - np.eye(h)).max() < 1e-13


@pytest.mark.parametrize("h,threads,kind", [(6, 64, "dense"), (15, 64, "dense"), (33, 64, "dense"), (47, 128, "dense"),
                                            (64, 128, "dense"), (64, 128, "degenerate"), (44, 128, "neardiag"),
                                            (64, 256, "neardiag"), (97, 512, "neardiag")])
def test_register_blocked_jacobi_device_function(h, threads, kind):
    """block_jacobi_svd_fast (2-column blocks in registers, tracked norms, MUFU-seeded tangent): singular values and
    orthogonality to rounding, including odd sizes (zero pad row), clusters of equal singular values and the nearly
    diagonal matrices of weakly scattering layers"""
    lib = emu_lib()
    rng = np.random.default_rng(100 + h)
    if kind == "dense":
        M = rng.normal(size=(h, h)) + np.diag(rng.uniform(1, 5, h))
    elif kind == "degenerate":
        Q1, _ = np.linalg.qr(rng.normal(size=(h, h)))
        Q2, _ = np.linalg.qr(rng.normal(size=(h, h)))
        sv = np.repeat(rng.uniform(1, 10, h // 4), 4) * (1 + 1e-13 * rng.normal(size=h))
        M = (Q1 * sv) @ Q2.T
    else:
        M = np.diag(rng.uniform(1, 11, h)) + 1e-3 * rng.normal(size=(h, h))
    ld = lib.emu_jacobi_ld(h)
    assert ld % 2 == 0 and ld >= h
    W = np.zeros((ld, h), order="F")
    W[:h, :] = M
    sweeps = lib.emu_jacobi_fast(W.ctypes.data_as(C.POINTER(C.c_double)), h, ld, threads)
    assert 0 < sweeps < 20
    assert np.all(W[h:, :] == 0.0)
    Wh = W[:h, :]
    sig = np.sort(np.linalg.norm(Wh, axis=0))[::-1]
    # (the smallest singular value of the random 33 x 33 case sits at 1e-13 relative whatever the order of the
    # rotations and reductions: 5e-13 leaves room for that case, every other one is below 5e-14)
    np.testing.assert_allclose(sig, np.linalg.svd(M, compute_uv=False), rtol=5e-13)
    U = Wh / np.linalg.norm(Wh, axis=0)
    assert np.abs(U.T @ U - np.eye(h)).max() < 1e-13
    # W = M V with V orthogonal: M^-1 W must be orthogonal
    V = np.linalg.solve(M, Wh)
    assert np.abs(V.T @ V - np.eye(h)).max() < 1e-11


@pytest.mark.parametrize("h,nR,threads", [(5, 3, 64), (16, 17, 64), (33, 34, 128), (44, 45, 256), (64, 65, 512),
                                          (64, 64, 128), (47, 1, 64)])
def test_blocked_gauss_jordan_device_function(h, nR, threads):
    """block_gj_rows_blocked (panel of 8 columns factorised in registers by one warp, rank-8 updates by register
    tiles): A^-1 R against LAPACK, on matrices that need row interchanges"""
    lib = emu_lib()
    rng = np.random.default_rng(7 * h + nR)
    A = rng.normal(size=(h, h))
    A[rng.permutation(h), np.arange(h)] += 3.0   # large entries off the diagonal: pivoting is exercised
    R = rng.normal(size=(h, nR))
    Ac, Rc = np.asfortranarray(A.copy()), np.asfortranarray(R.copy())
    X = np.zeros((h, nR), order="F")
    P = C.POINTER(C.c_double)
    rc = lib.emu_gj_blocked(Ac.ctypes.data_as(P), h, Rc.ctypes.data_as(P), nR, threads, X.ctypes.data_as(P))
    assert rc == 0
    ref = np.linalg.solve(A, R)
    assert np.abs(X - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()) * np.linalg.cond(A)


def test_blocked_gauss_jordan_reports_singular_matrix():
    lib = emu_lib()
    h, nR = 24, 4
    A = np.asfortranarray(np.ones((h, h)))
    R = np.asfortranarray(np.ones((h, nR)))
    X = np.zeros((h, nR), order="F")
    P = C.POINTER(C.c_double)
    assert lib.emu_gj_blocked(A.ctypes.data_as(P), h, R.ctypes.data_as(P), nR, 64, X.ctypes.data_as(P)) == 1


@pytest.mark.parametrize("h,threads", [(1, 64), (2, 64), (7, 64), (33, 64), (64, 128), (63, 128), (100, 64)])
def test_left_looking_cholesky_device_function(h, threads):
    """team_cholesky_fast: two concurrent teams, two columns per step, one barrier per pair of columns; only the lower
    triangle is read or written; odd sizes and sizes larger than the team (several rows per thread)"""
    lib = emu_lib()
    rng = np.random.default_rng(h)
    P = C.POINTER(C.c_double)
    mats = []
    for _ in range(2):
        Q = rng.normal(size=(h, h))
        mats.append(Q @ Q.T + h * np.eye(h))
    A = [np.asfortranarray(np.tril(m) + np.triu(np.full((h, h), 7e77), 1)) for m in mats]   # poison above the diagonal
    rc = lib.emu_cholesky_pair(A[0].ctypes.data_as(P), A[1].ctypes.data_as(P), h, threads)
    assert rc == 0
    for a, m in zip(A, mats):
        np.testing.assert_allclose(np.tril(a), np.linalg.cholesky(m), rtol=1e-12, atol=1e-13)
        assert np.all(a[np.triu_indices(h, 1)] == 7e77)


def test_left_looking_cholesky_reports_indefinite_matrix():
    lib = emu_lib()
    P = C.POINTER(C.c_double)
    h = 12
    good = np.asfortranarray(np.eye(h) * 2.0)
    bad = np.asfortranarray(np.eye(h))
    bad[5, 5] = -1.0
    assert lib.emu_cholesky_pair(good.ctypes.data_as(P), bad.ctypes.data_as(P), h, 64) == 2


@pytest.mark.parametrize("h,threads", [(3, 64), (8, 64), (13, 64), (44, 128), (64, 128), (64, 32)])
def test_paired_lane_back_substitution_device_function(h, threads):
    lib = emu_lib()
    lib.emu_jacobi_ld.restype = C.c_int
    rng = np.random.default_rng(31 * h)
    Cm = np.asfortranarray(np.tril(rng.normal(size=(h, h))) + 3 * np.eye(h))
    ld = lib.emu_jacobi_ld(h)
    W = np.zeros((ld, h), order="F")
    W0 = rng.normal(size=(h, h))
    W[:h] = W0
    P = C.POINTER(C.c_double)
    lib.emu_backsolve_lt(Cm.ctypes.data_as(P), W.ctypes.data_as(P), h, ld, threads)
    ref = np.linalg.solve(Cm.T, W0)
    np.testing.assert_allclose(W[:h], ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
    assert np.all(W[h:] == 0.0)


@pytest.mark.parametrize("M,K,threads,dual", [(64, 64, 256, True), (44, 44, 256, True), (20, 20, 64, False),
                                              (64, 64, 64, True), (7, 5, 64, True)])
def test_block_matvec_device_function(M, K, threads, dual):
    lib = emu_lib()
    rng = np.random.default_rng(M + K)
    A1 = np.asfortranarray(rng.normal(size=(M, K)))
    A2 = np.asfortranarray(rng.normal(size=(M, K)))
    x = rng.normal(size=K)
    y1, y2 = np.zeros(M), np.zeros(M)
    P = C.POINTER(C.c_double)
    lib.emu_matvec_dual(A1.ctypes.data_as(P), A2.ctypes.data_as(P) if dual else None, M, K, x.ctypes.data_as(P),
                        y1.ctypes.data_as(P), y2.ctypes.data_as(P), threads)
    np.testing.assert_allclose(y1, A1 @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(y2, A2 @ x if dual else 0.0, rtol=1e-13, atol=1e-13)


def test_conservative_layer_is_reported_not_solved():
    """iba_original on an inverted medium (dense_snow_correction="auto") has ka = k0 f Im(eps_air) |y2| = 0: scattering
    albedo 1, the symmetric factors of the layer matrix are singular.  The reference returns whatever LAPACK's rounding
    gives for the double eigenvalue 0; the device path reports ERR_EIGEN and NaN instead of a number."""
    d, batch, opts = load_golden("iba_original_dense_active")
    batch.dense_snow_correction[:] = 1
    out = emu_solve(batch, opts, threads=64)
    assert (out.status[0] & 15) == 2 and np.all(np.isnan(out.values[0]))
    assert out.ka[0, 1] == 0.0 and out.ks[0, 1] > 0.0


@pytest.mark.parametrize("n_max_stream,mode,layers", [(36, "P", 2)])
def test_emulated_kernels_for_blocks_of_65_to_128_unknowns(n_max_stream, mode, layers, monkeypatch):
    """The 64 < h <= 128 instantiations (eigen_kernel<2>: packed lower-triangular C, 16-lane Jacobi groups;
    boundary_kernel<.., kMid>: one resident matrix or resident right block, product-form elimination, staged GEMMs)
    against the oracle: 36 streams passive (blocks of 72; 512 emulated threads).  The GPU suite runs them at full size,
    passive and active (tests/test_gpu_parity.py::test_blocks_of_65_to_128_unknowns_against_oracle)."""
    from oracle import dort_oracle as O
    from smrt_b200.pack import pack_snow_ensemble

    monkeypatch.setenv("SMRT_EMU_EIGEN_MID", "1")
    monkeypatch.setenv("SMRT_EMU_BOUNDARY_MID", "1")
    rng = np.random.default_rng(11)
    th = np.concatenate((rng.uniform(0.05, 0.5, (1, layers - 1)), np.full((1, 1), 1000.0)), axis=1)
    rho = rng.uniform(150, 450, (1, layers)); T = rng.uniform(240, 272, (1, layers))
    pc = rng.uniform(5e-5, 3e-4, (1, layers))
    kw = dict(mode="A", theta_inc_deg=40.0, theta_deg=40.0) if mode == "A" else dict(theta_deg=55.0)
    batch = pack_snow_ensemble([36.5e9], th, rho, T, corr_length=pc, **kw)
    opts = dict(n_max_stream=n_max_stream, m_max=2)
    ref = np.asarray(O.solve_problem(batch.to_problem(0, opts))["values"])
    out = emu_solve(batch, opts, threads=64)
    assert (out.status[0] & 15) == 0
    assert rel_err(out.values[0], ref, batch.mode) <= (1e-9 if mode == "P" else 1e-6)
