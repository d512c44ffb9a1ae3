"""Host-side logic (packer, Model.run batching/ordering, Result API, error mapping) on CPU.

The batched solve is stood in by the SIMT-emulated kernels (tests/simt_emu) so that the host code is exercised end to
end without a GPU; the `-m gpu` twin (tests/test_gpu_api.py) runs the same calls through the CUDA library."""
import threading
import warnings

import numpy as np
import pandas as pd
import pytest

import smrt_b200
from emu_util import emu_solve, load_golden
from smrt_b200 import capi, make_model, make_snowpack, model as model_mod, sensor_list
from smrt_b200.error import SMRTError, SMRTWarning


_EMU_LOCK = threading.Lock()  # the SIMT emulator runs one launch at a time (the real plans are one per device)


class EmuPlan:
    def __init__(self, opts):
        self.opts = opts
        self.options = type("O", (), dict(max_batch=10 ** 9, n_max_stream=opts["n_max_stream"]))()

    def solve_host(self, batch):
        with _EMU_LOCK:
            return emu_solve(batch, self.opts, threads=64)


@pytest.fixture(autouse=True)
def emulated_plans(monkeypatch):
    monkeypatch.setattr(model_mod._PLANS, "get", lambda batch, opts, device, max_batch=0: EmuPlan(opts))


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
    # reflector with a prescribed backscattering coefficient (reference substrate/reflector_backscatter.py): passive and
    # active, the four numbers travel as substrate parameters
    rb = S.make_reflector(temperature=250, specular_reflection={"V": 0.3, "H": 0.4},
                          backscattering_coefficient={"VV": 0.1, "HH": 0.05})
    rb1 = S.make_reflector(specular_reflection=0.2, backscattering_coefficient={"VV": 0.03, "HH": 0.02})
    radar = S.sensor_list.active(13e9, 40)
    b3 = pack.pack_simulations([(radar, sp(substrate=rb)), (radar, sp(substrate=rb1))], "iba")
    assert list(b3.substrate_kind) == [pack.SUB_REFLECTOR_BACKSCATTER] * 2
    np.testing.assert_array_equal(b3.substrate_params, [[0.3, 0.4, 0.1, 0.05], [0.2, 0.2, 0.03, 0.02]])
    assert pack.pack_simulations([(sensor, sp(substrate=rb))], "iba").substrate_kind[0] == pack.SUB_REFLECTOR_BACKSCATTER
    # rough soil with the IEM backscatter of Fung et al. 1992 (reference substrate/iem_fung92.py, iem_fung92_brogioni10.py)
    iem = S.make_soil("iem_fung92", permittivity_model=complex(12, 2), roughness_rms=0.005, corr_length=0.05, temperature=265)
    iemb = S.make_soil("iem_fung92_brogioni10", permittivity_model=complex(8, 1), roughness_rms=0.01, corr_length=0.1,
                       autocorrelation_function="gaussian", series_truncation=6, temperature=265)
    b4 = pack.pack_simulations([(radar, sp(substrate=iem)), (radar, sp(substrate=iemb))], "iba")
    assert list(b4.substrate_kind) == [pack.SUB_IEM_FUNG92, pack.SUB_IEM_FUNG92_BRIOGONI10]
    np.testing.assert_array_equal(b4.substrate_params, [[0.005, 0.05, 0, 10], [0.01, 0.1, 1, 6]])
    np.testing.assert_array_equal(b4.substrate_eps, [12 + 2j, 8 + 1j])
    with pytest.raises(S.SMRTError):
        pack.pack_simulations([(radar, sp(substrate=S.make_soil(
            "iem_fung92", permittivity_model=complex(12, 2), roughness_rms=0.005, corr_length=0.05,
            autocorrelation_function="power")))], "iba")
    # the same IEM model as an INTERFACE (rough snow surface / internal interface): kinds per layer + a [B, L, 4] block of
    # parameters that exists only when a snowpack has such an interface
    rough = S.make_interface("iem_fung92", roughness_rms=0.004, corr_length=0.05)
    rough_b = S.make_interface("iem_fung92_brogioni10", roughness_rms=0.006, corr_length=0.1,
                               autocorrelation_function="gaussian")
    sp2 = S.make_snowpack([0.2, 0.3], "exponential", density=[250, 350], temperature=265, corr_length=1e-4,
                          interface=[rough, "flat"])
    sp3 = S.make_snowpack([0.2, 0.3], "exponential", density=[250, 350], temperature=265, corr_length=1e-4,
                          interface=["transparent", rough_b])
    b5 = pack.pack_simulations([(radar, sp2), (radar, sp3), (radar, sp())], "iba")
    np.testing.assert_array_equal(b5.interface, [[pack.IF_IEM_FUNG92, pack.IF_FLAT],
                                                 [pack.IF_TRANSPARENT, pack.IF_IEM_FUNG92_BRIOGONI10], [pack.IF_FLAT, 0]])
    np.testing.assert_array_equal(b5.interface_params[0], [[0.004, 0.05, 0, 10], [0, 0, 0, 0]])
    np.testing.assert_array_equal(b5.interface_params[1], [[0, 0, 0, 0], [0.006, 0.1, 1, 10]])
    np.testing.assert_array_equal(b5.interface_params[2], np.zeros((2, 4)))
    assert b4.interface_params is None and pack.concat_batches([b4, b4]).interface_params is None
    cat2 = pack.concat_batches([b4, b5])
    assert cat2.interface_params.shape == (5, 2, 4) and not cat2.interface_params[:2].any()
    assert b5.subset(slice(0, 2)).interface_params.shape == (2, 2, 4)
    assert b5.to_problem(1)["interface_params"].shape == (2, 4) and b4.to_problem(0)["interface_params"] is None
    for bad in (dict(specular_reflection=0.2, backscattering_coefficient=0.1),
                dict(backscattering_coefficient={"VV": 0.1, "HH": 0.1}),
                dict(specular_reflection=0.1, backscattering_coefficient={"VV": np.cos, "HH": 0.1})):
        with pytest.raises(S.SMRTError):
            pack.pack_simulations([(sensor, sp(substrate=S.make_reflector(**bad)))], "iba")


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


def _ragged_snowpacks():
    rng = np.random.default_rng(3)
    sps = []
    for n in (1, 3, 2):
        sps.append(make_snowpack(list(rng.uniform(0.1, 0.5, n - 1)) + [20.0], "exponential",
                                 density=rng.uniform(200, 420, n), temperature=rng.uniform(245, 270, n),
                                 corr_length=rng.uniform(5e-5, 3e-4, n)))
    return sps


def _assert_same_result(a, b):
    assert a.data.dims == b.data.dims and type(a) is type(b)
    np.testing.assert_array_equal(np.asarray(a.data.values), np.asarray(b.data.values))
    for d in a.data.dims:
        assert list(a.data.coords[d].values) == list(b.data.coords[d].values)
    assert a.channel_map == b.channel_map
    assert set(a.other_data) == set(b.other_data)
    for k in a.other_data:
        assert a.other_data[k].dims == b.other_data[k].dims, k
        np.testing.assert_array_equal(np.asarray(a.other_data[k].values), np.asarray(b.other_data[k].values))


def test_one_shot_result_equals_the_stacked_per_simulation_results():
    """Model.run builds the N-d Result directly from the output block; the reference stacks one Result per simulation
    (core/model.py:401-404).  Ragged layer counts and ragged numbers of air streams are NaN padded the same way."""
    from smrt_b200.result import concat_results

    sps = _ragged_snowpacks()
    sensor = sensor_list.amsre(["19", "37"])
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8))
    fast = m.run(sensor, sps)
    assert fast.data.dims == ("frequency", "snowpack", "polarization", "theta") and fast.data.shape == (2, 3, 2, 1)
    assert fast.other_data["ks"].shape == (2, 3, 3) and np.isnan(fast.other_data["ks"].values[0, 0, 1:]).all()
    sims, dimensions = m.prepare_simulations(sensor, sps, None, "snowpack")
    from smrt_b200.model import check_dort_options
    results = m._run_simulations(sims, check_dort_options(m.rtsolver_options))
    for dimension in reversed(dimensions):
        n = len(dimension[1])
        results = [concat_results(results[i:i + n], dimension) for i in range(0, len(results), n)]
    _assert_same_result(fast, results[0])
    assert fast.Tb(channel="37V", snowpack=2) == fast.data.values[1, 2, 0, 0]
    # chunked pack / solve pipeline (and the same through two "devices"): identical block
    m2 = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8), devices=[0, 1])
    m2.CHUNK_SIMULATIONS = 2
    _assert_same_result(fast, m2.run(sensor, sps))
    # active mode
    m4 = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=4, m_max=1))
    fa = m4.run(sensor_list.active([13e9, 17e9], 40), sps[:2])
    assert fa.data.dims == ("frequency", "snowpack", "polarization_inc", "polarization", "theta_inc")
    one = m4.run(sensor_list.active(17e9, 40), sps[1])
    np.testing.assert_array_equal(fa.data.values[1, 1], one.data.values)


def test_wet_snow_packers_agree_and_follow_the_reference_formulas():
    """Wet snow: the array packer, the object packer (reference wetice model recognised by name) and the scalar
    formulas of smrt/permittivity/wetice.py:13-41, water.py:12-47, make_medium.py:390-434."""
    from smrt_b200 import pack

    f, T, lw = 18.7e9, 273.15, 0.06
    theta = 1 - 300.0 / T
    e0 = 77.66 - 103.3 * theta
    e1 = 0.0671 * e0
    f1 = 20.2 + 146.4 * theta + 316 * theta**2
    e2 = 3.52 + 7.52 * theta
    ew = e2 + (e1 - e2) / complex(1, -f / 1e9 / (39.8 * f1)) + (e0 - e1) / complex(1, -f / 1e9 / f1)
    assert complex(pack.water_permittivity_maetzler87(f, T)) == pytest.approx(ew, rel=1e-15)
    ei = complex(pack.ice_permittivity_maetzler06(f, T))
    cplus, cminus = ei + 2 * ew, (ei - ew) * (1 - lw)
    assert complex(pack.wetice_permittivity_bohren83(f, T, lw)) == pytest.approx((cplus + 2 * cminus) / (cplus - cminus) * ew,
                                                                               rel=1e-15)
    assert complex(pack.wetice_permittivity_bohren83(f, 260.0, 0.0)) == complex(pack.ice_permittivity_maetzler06(f, 260.0))
    fv, lwf = pack.snow_frac_volumes(np.array([300.0, 400.0]), volumetric_liquid_water=np.array([0.0, 0.03]))
    np.testing.assert_allclose(fv, [300 / 916.7, (400 - 83.3 * 0.03) / 916.7], rtol=1e-15)
    np.testing.assert_allclose(lwf, [0.0, 0.03 / fv[1]], rtol=1e-15)

    # object packer on layers that carry the reference's default permittivity model (by name) and a liquid_water attribute
    def wetice_permittivity_bohren83(frequency, temperature=None, liquid_water=None, _properties_to_inject=None):
        p = _properties_to_inject
        return complex(pack.wetice_permittivity_bohren83(frequency, p.temperature, p.liquid_water))

    wetice_permittivity_bohren83.__module__ = "smrt.permittivity.wetice"
    th = np.array([[0.2, 0.4, 10.0]]); rho = np.array([[250.0, 320.0, 400.0]]); T3 = np.array([[273.15] * 3])
    vlw = np.array([[0.02, 0.0, 0.05]]); pc = np.array([[1e-4, 2e-4, 3e-4]])
    arr = pack.pack_snow_ensemble([18.7e9, 36.5e9], th, rho, T3, corr_length=pc, volumetric_liquid_water=vlw)
    fvol, lwat = pack.snow_frac_volumes(rho[0], vlw[0])
    sp = make_snowpack(th[0], "exponential", density=rho[0], temperature=T3[0], corr_length=pc[0])
    for l, layer in enumerate(sp.layers):
        layer.microstructure.frac_volume = float(fvol[l])
        layer.liquid_water = float(lwat[l])
        layer.permittivity_model = (1.0, wetice_permittivity_bohren83)
        layer.permittivity = (lambda lay: lambda i, fr: lay.permittivity_model[i](fr, _properties_to_inject=lay)
                              if callable(lay.permittivity_model[i]) else lay.permittivity_model[i])(layer)
    sims = [(s, sp) for s in sensor_list.passive([18.7e9, 36.5e9], 55).iterate("frequency")]
    obj = pack.pack_simulations(sims, "iba")
    np.testing.assert_allclose(obj.eps_sc, arr.eps_sc, rtol=1e-15)
    np.testing.assert_allclose(obj.frac_volume, arr.frac_volume, rtol=1e-15)
    assert abs(arr.eps_sc[0, 0].imag) > 10 * abs(arr.eps_sc[0, 1].imag)  # the wet layers absorb far more
