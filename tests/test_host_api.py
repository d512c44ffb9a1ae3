"""Host-side logic (packer, Model.run batching/ordering, Result API, error mapping) on CPU.

The batched solve is stood in by the SIMT-emulated kernels (tests/simt_emu) so that the host code is exercised end to
end without a GPU; the `-m gpu` twin (tests/test_gpu_api.py) runs the same calls through the CUDA library."""
import warnings

import numpy as np
import pandas as pd
import pytest

import smrt_b200
from emu_util import emu_solve, load_golden
from smrt_b200 import capi, make_model, make_snowpack, model as model_mod, sensor_list
from smrt_b200.error import SMRTError, SMRTWarning


class EmuPlan:
    def __init__(self, opts):
        self.opts = opts
        self.options = type("O", (), dict(max_batch=10 ** 9, n_max_stream=opts["n_max_stream"]))()

    def solve_host(self, batch):
        return emu_solve(batch, self.opts, threads=64)


@pytest.fixture(autouse=True)
def emulated_plans(monkeypatch):
    monkeypatch.setattr(model_mod._PLANS, "get", lambda batch, opts, device: EmuPlan(opts))


def two_layer():
    return make_snowpack([0.1, 100], "exponential", density=[200, 400], temperature=[250.0, 250.0],
                         corr_length=[5e-5, 5e-5])


def test_reads_like_the_reference_test_passive():
    # reference smrt/test/test_integration_iba.py:33-49
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16))
    res = m.run(sensor_list.amsre("37V"), two_layer())
    d, batch, opts = load_golden("ref_iba_2layer_passive")
    assert isinstance(res.TbV(), float)
    # 16 streams instead of 32: close to, not equal to, the 32-stream literal
    assert abs(res.TbV() - 248.09044325849692) < 1.0 and abs(res.TbH() - 237.3487270223389) < 1.0
    assert res.Tb(channel="37V") == res.TbV()
    assert set(res.other_data) == {"stream_angles", "effective_permittivity", "ks", "ke", "ka", "thickness"}
    np.testing.assert_allclose(res.optical_depth().values, (res.ks().values + res.ka().values) * [0.1, 100])


def test_dims_and_order_of_a_multi_frequency_multi_snowpack_run():
    sps = [make_snowpack([0.3, 10], "exponential", density=[250, 350], temperature=[260, 265], corr_length=c)
           for c in (1e-4, 2e-4, 3e-4)]
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8))
    res = m.run(sensor_list.passive([18.7e9, 36.5e9], [40, 55]), sps)
    assert res.data.dims == ("frequency", "snowpack", "polarization", "theta")
    assert res.data.shape == (2, 3, 2, 2)
    assert res.other_data["ks"].dims == ("frequency", "snowpack", "layer")
    # same numbers as running each simulation alone (ordering: frequency outermost, snowpack innermost)
    single = m.run(sensor_list.passive(36.5e9, [40, 55]), sps[1])
    np.testing.assert_allclose(res.data.sel(frequency=36.5e9, snowpack=1).values, single.data.values, rtol=1e-12)
    assert res.TbV(frequency=18.7e9, snowpack=2, theta=55) == res.data.values[0, 2, 0, 1]
    df = res.to_dataframe(channel_axis=None)
    assert len(df) == 24


def test_snowpack_containers():
    sps = [two_layer(), two_layer()]
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8))
    s = sensor_list.passive(37e9, 55)
    r = m.run(s, {"a": sps[0], "b": sps[1]})
    assert list(r.coords["snowpack"].values) == ["a", "b"]
    r = m.run(s, sps, snowpack_dimension=("time", [10.0, 20.0]))
    np.testing.assert_allclose(r.time, [10.0, 20.0])
    with pytest.raises(SMRTError):
        m.run(s, sps, snowpack_dimension=([1, 2], "time"))
    df = pd.DataFrame(dict(snowpack=sps, site=["x", "y"]), index=pd.Index([5, 6], name="id"))
    r = m.run(s, df)
    out = r.to_dataframe(channel_axis=None)
    assert "site" in out.columns and list(r.coords["id"].values) == [5, 6]


def test_active_result_accessors():
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8))
    res = m.run(sensor_list.active(13e9, 45), two_layer())
    assert res.data.dims == ("polarization_inc", "polarization", "theta_inc")
    vv = res.sigmaVV()
    assert vv == pytest.approx(4 * np.pi * np.cos(np.deg2rad(45)) * res.data.values[0, 0, 0])
    assert res.sigmaHV() == pytest.approx(4 * np.pi * np.cos(np.deg2rad(45)) * res.data.values[1, 0, 0])
    assert res.sigmaVV_dB() == pytest.approx(10 * np.log10(vv))


def test_error_mapping_exception_and_nan():
    big = make_snowpack([1, 10], "exponential", density=[300, 300], temperature=[260, 260], corr_length=[5e-3, 5e-3])
    s = sensor_list.passive(89e9, 55)
    with pytest.raises(SMRTError):
        make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(s, big)
    r = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8, error_handling="nan")).run(s, [big, two_layer()])
    assert np.all(np.isnan(r.data.values[0])) and np.all(np.isfinite(r.data.values[1]))


def test_shallow_warning():
    sp = make_snowpack([0.2, 0.3], "sticky_hard_spheres", density=[290.0, 250.0], temperature=[260, 260],
                       radius=[1e-4, 1e-4], stickiness=[0.2, 0.2])
    m = make_model("dmrt_qcacp_shortrange", "dort", rtsolver_options=dict(n_max_stream=8))
    with pytest.warns(SMRTWarning, match="optically shallow"):
        m.run(sensor_list.active(10e9, 45), sp)


def test_unsupported_features_are_rejected_loudly():
    with pytest.raises(SMRTError):
        make_model("sft_rayleigh", "dort")
    with pytest.raises(SMRTError):
        make_model("iba", "iterative_first_order")
    with pytest.raises(SMRTError):
        make_model("iba", "dort", rtsolver_options=dict(process_coherent_layers=True))
    with pytest.raises(TypeError):
        make_model("iba", "dort", rtsolver_options=dict(not_an_option=1))
    with pytest.raises(SMRTError):
        make_model("iba", "dort").run("not a sensor", two_layer())


def test_rtsolver_plugin_seam():
    """smrt_b200.DORT has the reference's plugin contract: C(**options).solve(snowpack, emmodels, sensor, atmosphere)"""
    class IBA:  # stands for the reference's emmodel instance: only its class name matters to the packer
        pass

    sp = two_layer()
    solver = smrt_b200.DORT(n_max_stream=8)
    res = solver.solve(sp, [IBA(), IBA()], sensor_list.passive(37e9, 55), None, parallel_computation="none")
    ref = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(sensor_list.passive(37e9, 55), sp)
    assert res.TbV() == ref.TbV()


def test_sea_ice_ensemble_packer_matches_the_reference_inputs():
    """pack_sea_ice_ensemble (vectorised brine volume / brine + saline-ice + sea-water permittivities) against the inputs
    the reference itself produced for the first six members of BASELINE config 5 (make_ice_column("multiyear", ...,
    add_water_substrate="ocean") -> layer.permittivity / substrate.permittivity), tests/golden/cfg5_first6.npz"""
    import bench
    from smrt_b200 import pack_sea_ice_ensemble

    d, ref_batch, _ = load_golden("cfg5_first6")
    th, T, sal, por, pc = bench.sea_ice_members(6)
    b = pack_sea_ice_ensemble(1.4e9, th, T, sal, por, pc, theta_deg=40.0)
    assert b.B == ref_batch.B and b.L == ref_batch.L and b.mode == ref_batch.mode
    for name in ("thickness", "temperature", "frac_volume", "eps_sc", "ms_p0", "ms_p1", "substrate_temperature", "theta"):
        np.testing.assert_allclose(getattr(b, name), getattr(ref_batch, name), rtol=1e-15, atol=0, err_msg=name)
    np.testing.assert_allclose(b.eps_bg, ref_batch.eps_bg, rtol=1e-13)
    np.testing.assert_allclose(b.substrate_eps, ref_batch.substrate_eps, rtol=1e-14)
    for name in ("nlayer", "emmodel", "ms_kind", "interface", "substrate_kind", "dense_snow_correction"):
        assert np.array_equal(getattr(b, name), getattr(ref_batch, name)), name


def test_sea_ice_permittivity_building_blocks():
    """spot values of the restated formulas (reference smrt/permittivity/test_saline_water.py, test_saline_ice.py style)"""
    from smrt_b200 import pack as P

    # Cox & Weeks regime boundaries are continuous enough and inside [0, 1]
    T = np.array([250.0, 255.0, 265.0, 271.0, 271.3])
    vb = P.brine_volume_cox83_lepparanta88(T, 0.005)
    assert np.all((vb > 0) & (vb < 1)) and np.all(np.diff(vb) > 0)
    # brine is lossy and strongly dispersive at L band; sea water around (77, 44) at 1.4 GHz, 271.35 K, 32 PSU
    eb = P.brine_permittivity_stogryn85(1.4e9, 265.0)
    assert eb.real > 40 and eb.imag > 20
    ew = P.seawater_permittivity_klein76(1.4e9, 271.35, 0.032)
    assert abs(ew - (76.94890621 + 44.08814627j)) < 1e-6
    with pytest.raises(SMRTError):
        P.seawater_permittivity_klein76(1.4e9, 260.0, 0.032)
    assert abs(P.water_freezing_temperature(0.0) - 273.15) < 0.05


def test_substrate_and_atmosphere_packing():
    """soil / reflector substrates and the isotropic atmosphere become plain per-problem parameters
    (reference smrt/substrate/soil_wegmuller.py, soil_qnh.py, reflector.py, atmosphere/simple_isotropic_atmosphere.py)"""
    import smrt_b200 as S
    from smrt_b200 import pack

    sensor = S.sensor_list.passive(37e9, 55)

    def sp(**kw):
        return S.make_snowpack([0.3], "exponential", density=[300], temperature=265, corr_length=1e-4, **kw)

    soil = S.make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.001, temperature=265)
    qnh = S.make_soil("soil_qnh", permittivity_model=lambda f, t: complex(8, 2), H=0.5, Q=0.1, N=1.0, temperature=260)
    refl = S.make_reflector(temperature=250, specular_reflection={(37e9, "H"): 0.5, (37e9, "V"): 0.6})
    atm = S.SimpleIsotropicAtmosphere(tb_down={37e9: 20.0}, tb_up=5.0, transmittance=0.9)
    b = pack.pack_simulations([(sensor, sp(substrate=soil)), (sensor, sp(substrate=qnh)), (sensor, sp(substrate=refl)),
                               (sensor, atm + sp())], "iba")
    assert list(b.substrate_kind) == [pack.SUB_SOIL_WEGMULLER, pack.SUB_SOIL_QNH, pack.SUB_REFLECTOR, pack.SUB_NONE]
    np.testing.assert_array_equal(b.substrate_params, [[0.001, 0, 0, 0], [0.5, 0.1, 1.0, 1.0], [0.6, 0.5, 0, 0],
                                                       [0, 0, 0, 0]])
    np.testing.assert_array_equal(b.substrate_eps[:2], [10 + 1j, 8 + 2j])
    np.testing.assert_array_equal(b.atmosphere, [[0, 0, 1], [0, 0, 1], [0, 0, 1], [20, 5, 0.9]])
    # the (deprecated) atmosphere argument of the solver seam, and concatenation of batches
    b2 = pack.pack_simulations([(sensor, sp())], "iba", atmospheres=[atm])
    np.testing.assert_array_equal(b2.atmosphere, [[20, 5, 0.9]])
    cat = pack.concat_batches([b, b2])
    assert cat.atmosphere.shape == (5, 3) and cat.substrate_params.shape == (5, 4)
    with pytest.raises(NotImplementedError):  # reflector.py:56-57
        pack.pack_simulations([(S.sensor_list.active(13e9, 40), sp(substrate=refl))], "iba")
    with pytest.raises(S.SMRTError):
        pack.pack_simulations([(sensor, sp(substrate=S.make_reflector(specular_reflection=np.cos)))], "iba")


def test_physics_laws_through_the_public_api():
    """reference smrt/test/test_physics_law.py ("Shallow" case, 16 streams) through make_model(...).run() with a soil
    substrate and an isotropic atmosphere: isothermal universe and Kirchhoff's law"""
    import smrt_b200 as S

    T = 265.0

    def snowpack(atmosphere=None):
        substrate = S.make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.001, temperature=T)
        return make_snowpack([0.3, 0.1], "exponential", density=[200, 300], temperature=T, corr_length=0.8e-3,
                             ice_permittivity_model=complex(1.7, 0.00001), substrate=substrate, atmosphere=atmosphere)

    radiometer = sensor_list.passive(37e9, [10, 40, 70])
    m = make_model("iba", "dort", rtsolver_options=dict(rayleigh_jeans_approximation=True, n_max_stream=16))
    iso = m.run(radiometer, snowpack(S.SimpleIsotropicAtmosphere(tb_down=T, tb_up=0, transmittance=1)))
    np.testing.assert_allclose(iso.TbV(), T, atol=0.01)
    np.testing.assert_allclose(iso.TbH(), T, atol=0.01)
    r0 = m.run(radiometer, snowpack())
    r1 = m.run(radiometer, S.SimpleIsotropicAtmosphere(tb_down=1, tb_up=0, transmittance=1) + snowpack())
    for tb0, tb1 in ((r0.TbV(), r1.TbV()), (r0.TbH(), r1.TbH())):
        np.testing.assert_allclose((tb0 + tb1) / 2 / T, 1 - (tb1 - tb0), atol=0.002)


def test_choudhury_outside_validity_raises_like_the_reference():
    from smrt_b200.inputs import ChoudhuryReflectivity

    sub = ChoudhuryReflectivity(temperature=265.0, permittivity_model=complex(10, 1), roughness_rms=5e-3)
    sp = make_snowpack([0.3], "exponential", density=[300], temperature=265, corr_length=1e-4, substrate=sub)
    with pytest.raises(Warning, match="outside validity range"):
        make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(sensor_list.passive(37e9, 55), sp)


def test_iba_variants_through_the_public_api():
    """reference test/test_integration_iba_original.py:12-45 and test/test_mixed_emmodel.py:9-40 read the same here; a
    subclass of the reference's IBA is mapped by its own name, not by its base class"""
    from smrt_b200.pack import EM_IBA_MAXWELL_GARNETT, EM_IBA_ORIGINAL, emmodel_code

    res = make_model("iba_original", "dort").run(sensor_list.amsre("37V"), two_layer())
    assert abs(res.TbV() - 247.92662874568973) < 1e-4 and abs(res.TbH() - 237.1283359660738) < 1e-4
    sp = make_snowpack([0.1, 100], "sticky_hard_spheres", density=[200, 400], temperature=[250.0, 250.0],
                       radius=[2e-4, 2e-4], stickiness=[0.1, 0.1])
    res = make_model(["dmrt_qcacp_shortrange", "iba"], "dort").run(sensor_list.amsre("37V"), sp)
    assert abs(res.TbV() - 204.510189893163) < 1e-4 and abs(res.TbH() - 190.53692754287889) < 1e-4

    class IBA:
        pass

    class IBA_original(IBA):
        pass

    class IBA_MaxwellGarnett(IBA):
        pass

    assert emmodel_code(IBA_original) == EM_IBA_ORIGINAL and emmodel_code(IBA_original()) == EM_IBA_ORIGINAL
    assert emmodel_code(IBA_MaxwellGarnett) == EM_IBA_MAXWELL_GARNETT and emmodel_code("iba_maxwell_garnett") == 7


def test_emmodels_and_options_per_medium_or_per_layer():
    """reference core/model.py:536-571: dict keyed by layer.medium, list per layer, for emmodels and their options"""
    sp = make_snowpack([0.1, 100], "exponential", density=[200, 600], temperature=[250.0, 250.0],
                       corr_length=[5e-5, 5e-5])
    sensor, opts = sensor_list.passive(37e9, 55), dict(n_max_stream=8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SMRTWarning)
        ref = make_model("iba", "dort", rtsolver_options=opts,
                         emmodel_options=dict(dense_snow_correction="auto")).run(sensor, sp)
        by_medium = make_model({"snow": "iba"}, "dort", rtsolver_options=opts,
                               emmodel_options={"snow": dict(dense_snow_correction="auto")}).run(sensor, sp)
        by_layer = make_model(["iba", "iba"], "dort", rtsolver_options=opts,
                              emmodel_options=[{}, dict(dense_snow_correction="auto")]).run(sensor, sp)
        plain = make_model("iba", "dort", rtsolver_options=opts).run(sensor, sp)
    assert by_medium.TbV() == ref.TbV() and by_layer.TbV() == ref.TbV() and plain.TbV() != ref.TbV()
    with pytest.raises(SMRTError):
        make_model({"ice": "iba"}, "dort", rtsolver_options=opts).run(sensor, sp)
    with pytest.raises(SMRTError):
        make_model(["iba"], "dort", rtsolver_options=opts).run(sensor, sp)


def test_inclusion_shapes_and_depolarization_packing():
    """reference permittivity/generic_mixing_formula.py:88-116 (shape mixtures), depolarization_factors.py:9-46"""
    from smrt_b200.inputs import depolarization_factors_spheroids
    from smrt_b200.pack import _shape_weights, pack_simulations

    assert _shape_weights(None) == (1.0, 0.0) and _shape_weights("random_needles") == (0.0, 1.0)
    assert _shape_weights({"spheres": 0.3, "random_needles": 0.7}) == (0.3, 0.7)
    assert _shape_weights(("random_needles", "spheres"), 0.25) == (0.75, 0.25)
    with pytest.raises(SMRTError):
        _shape_weights("cubes")
    with pytest.raises(SMRTError):
        _shape_weights({"spheres": 0.5, "random_needles": 0.5}, 0.5)
    np.testing.assert_allclose(depolarization_factors_spheroids(None), [1 / 3] * 3)
    for lr in (0.6, 1.5):  # oblate / prolate: the three factors sum to one
        d = depolarization_factors_spheroids(lr)
        assert abs(d.sum() - 1.0) < 1e-15 and (d[2] > d[0]) == (lr < 1)
    sp = two_layer()
    sp.layers[0].length_ratio = 1.5
    sp.layers[1].inclusion_shape = "random_needles"
    batch = pack_simulations([(sensor_list.passive(37e9, 55), sp)], "iba")
    np.testing.assert_allclose(batch.inclusion[0, 0], (1.0, 0.0) + tuple(depolarization_factors_spheroids(1.5)))
    np.testing.assert_allclose(batch.inclusion[0, 1], (0.0, 1.0, 1 / 3, 1 / 3, 1 / 3))
    res = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(sensor_list.passive(37e9, 55), sp)
    ref = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(sensor_list.passive(37e9, 55), two_layer())
    assert res.TbV() != ref.TbV() and abs(res.TbV() - ref.TbV()) < 5.0


def test_high_azimuthal_mode_counts_run_like_the_reference_schur_test():
    """reference rtsolver/test_dort.py:13-42: m_max = 16 with IBA / 32 streams runs (there: only with the Schur-based
    diagonalisation; here the symmetrised eigenproblem has no complex-eigenvalue failure mode at all)"""
    sp = make_snowpack(thickness=[1000], microstructure_model="exponential", density=280, temperature=265,
                       corr_length=0.05e-3)
    scatt = sensor_list.active(10e9, 50)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SMRTWarning)
        s16 = make_model("iba", "dort", rtsolver_options=dict(m_max=16, n_max_stream=32,
                                                              diagonalization_method="schur_forcedtriu")).run(scatt, sp)
        s2 = make_model("iba", "dort", rtsolver_options=dict(m_max=2, n_max_stream=32)).run(scatt, sp)
    assert np.isfinite(s16.sigmaVV()) and s16.sigmaVV() > 0
    assert abs(s16.sigmaVV_dB() - s2.sigmaVV_dB()) < 0.5  # small grains: the modes beyond 2 carry almost nothing
    with pytest.raises(SMRTError):
        make_model("iba", "dort", rtsolver_options=dict(m_max=17))


def test_first_year_sea_ice_ensemble_packer_matches_the_reference_inputs():
    """pack_sea_ice_ensemble(ice_type="firstyear") against what the reference's make_ice_column produced: the 9-layer
    column of test/test_iba_sea_ice.py (spheres) and the needle / mixed brine columns of
    tests/golden/inclusion_shapes_passive.npz"""
    from smrt_b200 import pack_sea_ice_ensemble

    d, ref, _ = load_golden("ref_sea_ice_128streams")  # member 0 = first-year ice, 1.4 GHz, 40 degrees
    L = 9
    b = pack_sea_ice_ensemble(1.4e9, [[1.5 / L] * L], np.linspace(273.15 - 20.0, 273.15 - 1.8, L)[None],
                              np.linspace(2.0, 10.0, L)[None] * 1e-3, 0.0, 500e-6, theta_deg=40.0, ice_type="firstyear")
    for name in ("thickness", "temperature", "ms_p0", "substrate_temperature", "theta"):
        np.testing.assert_allclose(getattr(b, name)[0], getattr(ref, name)[0], rtol=1e-15, err_msg=name)
    for name in ("frac_volume", "eps_bg", "eps_sc", "substrate_eps"):
        np.testing.assert_allclose(getattr(b, name)[0], getattr(ref, name)[0], rtol=1e-13, err_msg=name)
    d, ref, _ = load_golden("inclusion_shapes_passive")  # members 0 / 1: needles / 30 % spheres + 70 % needles; 3 frequencies
    for member, shape in ((0, "random_needles"), (1, {"spheres": 0.3, "random_needles": 0.7})):
        b = pack_sea_ice_ensemble([1.4e9, 6.925e9, 18.7e9], [[0.25, 0.35, 0.4]], [[258.0, 264.0, 269.0]],
                                  np.array([[5.0, 7.0, 9.0]]) * 1e-3, 0.0, [[2e-4, 3e-4, 4e-4]], theta_deg=[40, 55],
                                  ice_type="firstyear", brine_inclusion_shape=shape)
        rows = [f * 3 + member for f in range(3)]  # frequency-major order, three snowpacks per frequency
        for name in ("frac_volume", "eps_bg", "eps_sc", "substrate_eps"):
            got, want = getattr(b, name), getattr(ref, name)[rows]
            np.testing.assert_allclose(got, want[:, :3] if want.ndim == 2 else want, rtol=1e-13, err_msg=name)
        np.testing.assert_allclose(b.inclusion, ref.inclusion[rows][:, :3], rtol=1e-15)
    with pytest.raises(SMRTError):
        pack_sea_ice_ensemble(1.4e9, [[1.0]], [[260.0]], [[5e-3]], 0.05, 5e-4, ice_type="firstyear")
