"""Pin the CPU oracle (oracle/dort_oracle.py) to the reference: fixtures produced by running the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.npz) and the golden literals of the reference's own tests."""
import numpy as np
import pytest

from emu_util import load_golden, rel_err
from oracle import dort_oracle as O

FAST = ["cfg1_iba_onelayer", "ref_iba_2layer_passive", "ref_iba_2layer_active", "ref_dmrt_qcacp_2layer_passive",
        "ref_dmrt_less_refringent_active", "nonscattering_transparent", "nonscattering_active",
        "iba_multiangle_passive", "iba_options_prune_rj", "iba_shs_active_multiangle", "iba_exp_substrate_passive",
        "cfg3_first4", "cfg5_first6", "ref_sea_ice_128streams", "soil_wegmuller_passive", "soil_qnh_passive",
        "reflector_passive", "choudhury_passive", "atmosphere_passive", "ref_physics_law", "soil_active",
        "iba_microstructures_passive", "iba_microstructures_active", "rayleigh_passive", "rayleigh_active",
        "prescribed_kskaeps_passive", "ref_iba_original_2layer_passive", "ref_mixed_emmodel_passive", "iba_original_passive",
        "iba_original_dense_active", "iba_maxwell_garnett_passive", "iba_maxwell_garnett_dense_active",
        "emmodel_per_medium_passive", "inclusion_shapes_passive", "depolarization_active", "iba_original_depolarization_passive",
        "iba_maxwell_garnett_depolarization_passive", "ref_rayleigh_mmax6_active", "iba_mmax5_active",
        "reflector_backscatter_active", "reflector_backscatter_active_mmax4", "reflector_backscatter_passive",
        "iem_fung92_active", "iem_fung92_brogioni10_active", "iem_fung92_passive",
        "iem_fung92_interface_active", "iem_fung92_interface_passive"]


def solve_all(batch, opts, limit=None):
    n = batch.B if limit is None else min(limit, batch.B)
    return [O.solve_problem(batch.to_problem(i, opts)) for i in range(n)]


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_fixture(name):
    d, batch, opts = load_golden(name)
    outs = solve_all(batch, opts)
    vals = np.stack([o["values"] for o in outs])
    tol = 1e-11 if batch.mode == 0 else 1e-6  # active: the coherent subtraction amplifies LAPACK-level rounding to ~1e-8
    assert rel_err(vals, d["ref_values"], batch.mode) <= tol
    for b, o in enumerate(outs):
        n = batch.nlayer[b]
        np.testing.assert_allclose(o["ka"], d["ref_ka"][b, :n], rtol=1e-13)
        np.testing.assert_allclose(o["ks"], d["ref_ks"][b, :n], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(o["eps_eff"], d["ref_eps_eff"][b, :n], rtol=1e-14)
        assert len(o["stream_angles"]) == d["ref_n_air"][b]
        np.testing.assert_allclose(o["stream_angles"], d["ref_stream_angles"][b, :len(o["stream_angles"])], atol=1e-11)


def test_oracle_cfg2_member0_all_frequencies():
    d, batch, opts = load_golden("cfg2_first4")
    idx = [f * 4 for f in range(6)]  # member 0 at the 6 AMSR-E frequencies
    vals = np.stack([O.solve_problem(batch.to_problem(i, opts))["values"] for i in idx])
    assert rel_err(vals, d["ref_values"][idx], 0) <= 1e-11
    # SURVEY.md §8(d) member-0 table
    np.testing.assert_allclose(vals[0, :, 0], [248.695458986944, 226.167355602153], rtol=1e-12)
    np.testing.assert_allclose(vals[5, :, 0], [204.878081370754, 190.650942815642], rtol=1e-12)


def test_reference_test_literals():
    """Literals hard-coded in the reference's tests."""
    d, batch, opts = load_golden("cfg1_iba_onelayer")  # examples/iba_onelayer_example.py, SURVEY §6
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    np.testing.assert_allclose(v[:, 0], [268.2217269495297, 251.75293752732134], rtol=1e-13)
    d, batch, opts = load_golden("ref_iba_2layer_passive")  # test/test_integration_iba.py:48-49 (atol 1e-4)
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    np.testing.assert_allclose(v[:, 0], [248.09044325849692, 237.3487270223389], atol=1e-4)
    d, batch, opts = load_golden("ref_iba_2layer_active")  # test/test_integration_iba.py:67-69
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    sig = 4 * np.pi * np.cos(np.deg2rad(55.0)) * v[:, :, 0]
    np.testing.assert_allclose(10 * np.log10([sig[0, 0], sig[1, 1], sig[1, 0]]),
                               [-24.044882546524693, -24.416295329469907, -51.544272924876886], atol=1e-3)
    d, batch, opts = load_golden("ref_iba_original_2layer_passive")  # test/test_integration_iba_original.py:44-45
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    np.testing.assert_allclose(v[:, 0], [247.92662874568973, 237.1283359660738], atol=1e-4)
    d, batch, opts = load_golden("ref_mixed_emmodel_passive")  # test/test_mixed_emmodel.py:39-40
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    np.testing.assert_allclose(v[:, 0], [204.510189893163, 190.53692754287889], atol=1e-4)
    d, batch, opts = load_golden("ref_sea_ice_128streams")  # test/test_iba_sea_ice.py:31-32
    v = np.stack([O.solve_problem(batch.to_problem(i, opts))["values"] for i in range(2)])
    np.testing.assert_allclose(v[0, :, 0], [256.0184487450634, 228.46148449852473], atol=1e-4)
    np.testing.assert_allclose(v[1, :, 0], [257.5733413408494, 232.02001231655734], atol=1e-4)
    d, batch, opts = load_golden("ref_dmrt_less_refringent_active")  # test/test_dmrtdort.py:107-108
    v = O.solve_problem(batch.to_problem(0, opts))["values"]
    sig = 4 * np.pi * np.cos(np.deg2rad(45.0)) * v[:, :, 0]
    assert abs(sig[0, 0] - 7.54253344e-05) < 1e-7 and abs(sig[1, 1] - 7.09606407e-05) < 1e-7


def test_all_diagonalisation_methods_agree():
    """eig / schur / schur_forcedtriu / half_rank_eig give the same answer (reference test_integration_iba.py:33-49)"""
    d, batch, opts = load_golden("ref_iba_2layer_passive")
    vals = [O.solve_problem(batch.to_problem(0, opts), method=m)["values"]
            for m in ("eig", "schur", "schur_forcedtriu", "half_rank_eig")]
    for v in vals[1:]:
        np.testing.assert_allclose(v, vals[0], rtol=1e-12)


def test_romb65_is_scipy_romb():
    import scipy.integrate

    y = np.random.default_rng(0).normal(size=65)
    assert O.romb65(y, 1 / 32) == pytest.approx(scipy.integrate.romb(y, 1 / 32), rel=1e-15)


def test_error_handling_nan_and_exception():
    d, batch, opts = load_golden("cfg1_iba_onelayer")
    p = batch.to_problem(0, opts)
    p["frequency"] = 89e9
    p["ms_p0"] = np.array([5e-3])  # 5 mm grains at 89 GHz: normalisation far beyond 30 %
    with pytest.raises(O.OracleError):
        O.solve_problem(p)
    p["options"] = dict(opts, error_handling="nan")
    out = O.solve_problem(p)
    assert np.all(np.isnan(out["values"])) and out["status"] in (O.ST_NORMALIZATION, O.ST_EIGEN)
