"""`-m gpu` twin of tests/test_host_api.py: the public API (make_model(...).run(), Result accessors) through the CUDA
library on a real B200, written the way the reference's own tests are (smrt/test/test_integration_iba.py,
smrt/test/test_dmrtdort.py, smrt/rtsolver/test_rtsolver.py)."""
import warnings

import numpy as np
import pytest

from smrt_b200 import make_model, make_snowpack, sensor_list
from smrt_b200.error import SMRTError, SMRTWarning
from smrt_b200.inputs import FlatSubstrate, Transparent

pytestmark = pytest.mark.gpu


@pytest.fixture
def setup_snowpack_2():
    return make_snowpack(thickness=[0.1, 100], microstructure_model="exponential", density=[200, 400],
                         temperature=[250.0, 250.0], corr_length=[5e-5, 5e-5])


def test_iba_dort_oneconfig_passive(setup_snowpack_2):
    m = make_model("iba", "dort")
    res = m.run(sensor_list.amsre("37V"), setup_snowpack_2)
    np.testing.assert_allclose(res.TbV(), 248.09044325849692, atol=1e-4)
    np.testing.assert_allclose(res.TbH(), 237.3487270223389, atol=1e-4)


@pytest.mark.parametrize("method", ["eig", "schur", "half_rank_eig"])
def test_diagonalization_method_option_is_accepted(setup_snowpack_2, method):
    m = make_model("iba", "dort", rtsolver_options=dict(diagonalization_method=method))
    res = m.run(sensor_list.amsre("37V"), setup_snowpack_2)
    np.testing.assert_allclose(res.TbV(), 248.09044325849692, atol=1e-4)


def test_iba_dort_oneconfig_active(setup_snowpack_2):
    m = make_model("iba", "dort")
    res = m.run(sensor_list.active(frequency=19e9, theta_inc=55), setup_snowpack_2)
    np.testing.assert_allclose(res.sigmaVV_dB(), -24.044882546524693, atol=1e-3)
    np.testing.assert_allclose(res.sigmaHH_dB(), -24.416295329469907, atol=1e-3)
    np.testing.assert_allclose(res.sigmaHV_dB(), -51.544272924876886, atol=1e-3)


def test_dmrt_twoconfig():
    sp = make_snowpack([0.1, 1000], "sticky_hard_spheres", density=[200, 400], temperature=[250.0, 250.0],
                       radius=[2e-4, 2e-4], stickiness=[0.1, 0.1])
    res = make_model("dmrt_qcacp_shortrange", "dort").run(sensor_list.amsre(["19", "37"]), sp)
    assert (res.Tb(channel="37V") - 202.1726891947754) < 1e-4
    assert (res.Tb(channel="37H") - 187.45835882462404) < 1e-4
    assert abs(res.Tb(channel="19V") - 242.53348671) < 1e-4  # the reference's current output (fixture), not its stale literal
    assert abs(res.Tb(channel="19H") - 230.13167636301958) < 1e-4


def test_less_refringent_bottom_layer_VH():
    sp = make_snowpack([0.2, 0.3], "sticky_hard_spheres", density=[290.0, 250.0], radius=[1e-4, 1e-4],
                       stickiness=[0.2, 0.2])
    m = make_model("dmrt_qcacp_shortrange", "dort", rtsolver_options=dict(diagonalization_method="schur_forcedtriu"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=SMRTWarning)
        res = m.run(sensor_list.active(10e9, 45), sp)
    assert abs(res.sigmaVV() - 7.54253344e-05) < 1e-7
    assert abs(res.sigmaHH() - 7.09606407e-05) < 1e-7


def test_noabsorption_and_nadir():
    # reference smrt/rtsolver/test_rtsolver.py:36-47, 104-113
    sp = make_snowpack([100], "homogeneous", density=[300], temperature=[250], interface=[Transparent])
    res = make_model("nonscattering", "dort").run(sensor_list.passive(37e9, [30, 40]), sp)
    np.testing.assert_allclose(res.TbV(), 250, atol=0.01)
    np.testing.assert_allclose(res.coords["theta"], [30, 40])
    res = make_model("nonscattering", "dort").run(sensor_list.passive(37e9, [0, 5]), sp)
    np.testing.assert_allclose(res.TbV(), 250)


def test_output_stream_angles_active():
    # reference smrt/rtsolver/test_rtsolver.py:64-74
    sp = make_snowpack([0.5, 1000], "homogeneous", density=[250, 300], temperature=2 * [250],
                       interface=2 * [Transparent])
    res = make_model("nonscattering", "dort").run(sensor_list.active(13e9, 45), sp)
    np.testing.assert_allclose(res.other_data["stream_angles"], np.array([41.91460595, 45.86542465]))
    assert res.sigmaVV() == 0


def test_shallow_snowpack_warns():
    # reference smrt/rtsolver/test_rtsolver.py:115-127
    sp = make_snowpack([0.5, 0.5], "homogeneous", density=[300, 250], temperature=2 * [250],
                       interface=2 * [Transparent])
    with pytest.warns(SMRTWarning, match="optically shallow"):
        make_model("nonscattering", "dort").run(sensor_list.active(13e9, 45), sp).sigmaVV()


def test_rayleigh_jeans_approximation():
    sp = make_snowpack([100], "homogeneous", density=[300], temperature=[250], interface=[Transparent])
    s = sensor_list.passive(300e9, [30, 40])
    rj = make_model("nonscattering", "dort", rtsolver_options=dict(rayleigh_jeans_approximation=True)).run(s, sp)
    full = make_model("nonscattering", "dort", rtsolver_options=dict(rayleigh_jeans_approximation=False)).run(s, sp)
    np.testing.assert_allclose(rj.data.values, full.data.values, rtol=0.01)


def test_multi_frequency_snowpack_list_and_substrate():
    rng = np.random.default_rng(0)
    sps = [make_snowpack(rng.uniform(0.05, 0.5, 4), "exponential", density=rng.uniform(150, 450, 4),
                         temperature=rng.uniform(240, 272, 4), corr_length=rng.uniform(5e-5, 3e-4, 4),
                         substrate=FlatSubstrate(temperature=265.0, permittivity_model=complex(10, 1)))
           for _ in range(5)]
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16))
    res = m.run(sensor_list.amsre(), sps)
    assert res.data.dims == ("frequency", "snowpack", "polarization", "theta")
    assert res.data.shape == (6, 5, 2, 1)
    single = m.run(sensor_list.amsre("37"), sps[3])
    np.testing.assert_allclose(res.TbV(frequency=36.5e9, snowpack=3), single.TbV(), rtol=1e-12)
    assert np.all(res.data.values > 100) and np.all(res.data.values < 273.15)


def test_error_handling():
    big = make_snowpack([1, 10], "exponential", density=[300, 300], temperature=[260, 260], corr_length=[5e-3, 5e-3])
    with pytest.raises(SMRTError):
        make_model("iba", "dort").run(sensor_list.passive(89e9, 55), big)
    r = make_model("iba", "dort", rtsolver_options=dict(error_handling="nan")).run(sensor_list.passive(89e9, 55), big)
    assert np.all(np.isnan(r.data.values))


# ------------------------------------------------------------------ reference smrt/test/test_physics_law.py on the GPU
PHYSICS_CASES = [("High scattering", 0.8e-3, 10), ("Low scattering", 0.05e-3, 10), ("Shallow", 0.8e-3, 0.1)]


def _physics_snowpack(pc, thickness, T, atmosphere=None):
    from smrt_b200 import make_soil

    substrate = make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.001, temperature=T)
    return make_snowpack([0.3, thickness], "exponential", density=[200, 300], temperature=T, corr_length=pc,
                         ice_permittivity_model=complex(1.7, 0.00001), substrate=substrate, atmosphere=atmosphere)


@pytest.mark.parametrize("test,pc,thickness", PHYSICS_CASES)
def test_isothermal_universe(test, pc, thickness):
    """test/test_physics_law.py:9-43: soil, snow and sky at the same temperature radiate that temperature"""
    from smrt_b200 import SimpleIsotropicAtmosphere

    T = 265
    snowpack = _physics_snowpack(pc, thickness, T, SimpleIsotropicAtmosphere(tb_down=T, tb_up=0, transmittance=1))
    m = make_model("iba", "dort", rtsolver_options=dict(rayleigh_jeans_approximation=True))
    sresult = m.run(sensor_list.passive(37e9, range(10, 80, 5)), snowpack)
    np.testing.assert_allclose(sresult.TbV(), T, atol=0.01)
    np.testing.assert_allclose(sresult.TbH(), T, atol=0.01)


@pytest.mark.parametrize("test,pc,thickness", PHYSICS_CASES)
def test_kirchoff_law(test, pc, thickness):
    """test/test_physics_law.py:46-95: emissivity = 1 - reflectivity"""
    from smrt_b200 import SimpleIsotropicAtmosphere

    T = 265.0
    snowpack = _physics_snowpack(pc, thickness, T)
    radiometer = sensor_list.passive(37e9, range(10, 80, 5))
    m = make_model("iba", "dort", rtsolver_options=dict(rayleigh_jeans_approximation=True))
    sresult_0 = m.run(radiometer, snowpack)
    sresult_1 = m.run(radiometer, SimpleIsotropicAtmosphere(tb_down=1, tb_up=0, transmittance=1) + snowpack)
    for tb0, tb1 in ((sresult_0.TbV(), sresult_1.TbV()), (sresult_0.TbH(), sresult_1.TbH())):
        emissivity = (tb0 + tb1) / 2 / T
        reflectivity = tb1 - tb0
        np.testing.assert_allclose(emissivity, 1 - reflectivity, atol=0.002)


def test_choudhury_outside_validity_raises_like_the_reference():
    """substrate/rough_choudhury79.py:29-31: `raise Warning(...)` when k sigma > 0.1"""
    from smrt_b200.inputs import ChoudhuryReflectivity

    sub = ChoudhuryReflectivity(temperature=265.0, permittivity_model=complex(10, 1), roughness_rms=5e-3)
    sp = make_snowpack([0.3], "exponential", density=[300], temperature=265, corr_length=1e-4, substrate=sub)
    with pytest.raises(Warning, match="outside validity range"):
        make_model("iba", "dort").run(sensor_list.passive(37e9, 55), sp)


@pytest.mark.parametrize("microstructure_model,m_max,emmodel", [("independent_sphere", 6, "rayleigh"),
                                                                ("exponential", 16, "iba")])
def test_schur_based_diagonalisation_settings_run(microstructure_model, m_max, emmodel):
    """reference rtsolver/test_dort.py:13-42 (settings where scipy.linalg.eig fails in the reference)"""
    sp = make_snowpack(thickness=[1000], microstructure_model=microstructure_model, density=280, temperature=265,
                       radius=0.05e-3, corr_length=0.05e-3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SMRTWarning)
        m = make_model(emmodel, "dort", rtsolver_options=dict(m_max=m_max, n_max_stream=32,
                                                              diagonalization_method="schur"))
        s = m.run(sensor_list.active(10e9, 50), sp).sigmaVV()
    assert np.isfinite(s) and s > 0


def test_iba_variants_and_per_medium_emmodels(setup_snowpack_2):
    """reference test/test_integration_iba_original.py:44-45, test/test_mixed_emmodel.py:39-40, core/model.py:547-548"""
    res = make_model("iba_original", "dort").run(sensor_list.amsre("37V"), setup_snowpack_2)
    np.testing.assert_allclose([res.TbV(), res.TbH()], [247.92662874568973, 237.1283359660738], atol=1e-4)
    sp = make_snowpack([0.1, 100], "sticky_hard_spheres", density=[200, 400], temperature=[250.0, 250.0],
                       radius=[2e-4, 2e-4], stickiness=[0.1, 0.1])
    res = make_model(["dmrt_qcacp_shortrange", "iba"], "dort").run(sensor_list.amsre("37V"), sp)
    np.testing.assert_allclose([res.TbV(), res.TbH()], [204.510189893163, 190.53692754287889], atol=1e-4)
    by_medium = make_model({"snow": "iba_original"}, "dort").run(sensor_list.amsre("37V"), setup_snowpack_2)
    np.testing.assert_allclose(by_medium.TbV(), 247.92662874568973, atol=1e-4)


def test_one_shot_result_pipeline_and_devices():
    """Model.run at a size where the pack / solve pipeline is used: identical to the single-call path, dims and
    NaN-padded ragged layers like the reference's stacked result."""
    rng = np.random.default_rng(5)
    sps = []
    for k in range(40):
        n = 2 + k % 4
        sps.append(make_snowpack(list(rng.uniform(0.05, 0.5, n - 1)) + [30.0], "exponential",
                                 density=rng.uniform(150, 450, n), temperature=rng.uniform(240, 272, n),
                                 corr_length=rng.uniform(5e-5, 3e-4, n)))
    sensor = sensor_list.amsre(["19", "37", "89"])
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16))
    whole = m.run(sensor, sps)
    assert whole.data.dims == ("frequency", "snowpack", "polarization", "theta") and whole.data.shape == (3, 40, 2, 1)
    assert whole.other_data["ks"].shape == (3, 40, 5) and np.isnan(whole.other_data["ks"].values[0, 0, 2:]).all()
    piped = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16), devices=[0])
    piped.CHUNK_SIMULATIONS = 21
    res = piped.run(sensor, sps)
    np.testing.assert_array_equal(res.data.values, whole.data.values)
    np.testing.assert_array_equal(res.other_data["ke"].values, whole.other_data["ke"].values)
    one = m.run(sensor_list.amsre("37"), sps[17])
    np.testing.assert_allclose(whole.Tb(channel="37V", snowpack=17), one.TbV(), rtol=1e-12)


def test_two_devices_in_one_process():
    """make_model(..., devices=[0, 1]): the chunks of one run() go round-robin to one worker thread per GPU (needs a
    box with two GPUs: `gpurun --gpus 2`); bit-identical to the single-device run."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(9)
    sps = [make_snowpack(list(rng.uniform(0.05, 0.5, 5)) + [30.0], "exponential", density=rng.uniform(150, 450, 6),
                         temperature=rng.uniform(240, 272, 6), corr_length=rng.uniform(5e-5, 3e-4, 6)) for _ in range(64)]
    sensor = sensor_list.amsre()
    one = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16)).run(sensor, sps)
    m2 = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16), devices=[0, 1])
    m2.CHUNK_SIMULATIONS = 48
    two = m2.run(sensor, sps)
    np.testing.assert_array_equal(two.data.values, one.data.values)
    np.testing.assert_array_equal(two.other_data["ks"].values, one.other_data["ks"].values)


def test_rough_surfaces_with_diagonal_backscatter_through_the_public_api():
    """reflector with a prescribed backscattering coefficient (reference substrate/reflector_backscatter.py) and the IEM
    of Fung et al. 1992 as the snow surface and as the soil (interface/iem_fung92.py, substrate/iem_fung92.py) through
    make_model(...).run() with this package's builders; the literals are the unmodified reference's outputs for the
    same calls (16 streams, m_max = 2, 13.5 GHz, 40 degrees)"""
    from smrt_b200 import make_interface, make_reflector, make_soil

    kw = dict(density=[250, 350], temperature=[260, 265], radius=[3e-4, 5e-4], stickiness=0.2)
    m = make_model("iba", "dort", rtsolver_options=dict(n_max_stream=16, m_max=2))
    radar, radiometer = sensor_list.active(13.5e9, 40), sensor_list.passive(13.5e9, 40)
    sub = make_reflector(temperature=265, specular_reflection={"V": 0.3, "H": 0.4},
                         backscattering_coefficient={"VV": 0.1, "HH": 0.05})
    sp = make_snowpack([0.3, 0.7], "sticky_hard_spheres", substrate=sub, **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=SMRTWarning)
        res = m.run(radar, sp)
        np.testing.assert_allclose([res.sigmaVV_dB(), res.sigmaHH_dB(), res.sigmaVH_dB()],
                                   [-11.573163275586909, -11.857597269280516, -30.616256833957248], atol=1e-4)
        res = m.run(radiometer, sp)
        np.testing.assert_allclose([res.TbV(), res.TbH()], [195.98956243324483, 171.1949806679902], atol=1e-4)
        soil = make_soil("iem_fung92", permittivity_model=complex(10, 1), roughness_rms=0.005, corr_length=0.06,
                         temperature=268.0)
        sp = make_snowpack([0.3, 0.7], "sticky_hard_spheres", substrate=soil,
                           interface=[make_interface("iem_fung92", roughness_rms=0.004, corr_length=0.05), "flat"], **kw)
        res = m.run(radar, sp)
        np.testing.assert_allclose([res.sigmaVV_dB(), res.sigmaHH_dB(), res.sigmaVH_dB()],
                                   [-13.149321808779032, -12.468096175478424, -35.39501167583748], atol=1e-4)
        res = m.run(radiometer, sp)
        np.testing.assert_allclose([res.TbV(), res.TbH()], [15.115552515198775, 14.736331861270559], atol=1e-4)
