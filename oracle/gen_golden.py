"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/smrt).

Run in the authoring container only (the reference does not exist on the GPU box):

    python oracle/gen_golden.py [case ...]

For every case it (1) builds snowpacks and a sensor with the reference's own builders, (2) runs the reference
``make_model(em, "dort").run(sensor, snowpacks, parallel_computation="none")``, (3) packs the very same objects with
``smrt_b200.pack.pack_simulations`` and (4) writes the packed inputs + the reference outputs to one ``.npz``.
The synthetic ensembles are SURVEY.md §8(d)'s generators (same seeds).  xarray is replaced by oracle/xarray_shim.
"""

import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "xarray_shim"))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np  # noqa: E402

from smrt import PSU, make_ice_column, make_model, make_snowpack, sensor_list  # noqa: E402
from smrt.core.error import SMRTWarning  # noqa: E402
from smrt.interface.transparent import Transparent  # noqa: E402

from smrt_b200.pack import pack_simulations  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
AMSRE_FREQS = [6.925e9, 10.65e9, 18.7e9, 23.8e9, 36.5e9, 89e9]


def snow(rng, L, kind):
    """SURVEY.md §8(d) generators"""
    th = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
    if kind == "exponential":
        rho = rng.uniform(150, 450, L)
        T = rng.uniform(240, 272, L)
        pc = rng.uniform(5e-5, 3e-4, L)
        return make_snowpack(th, "exponential", density=rho, temperature=T, corr_length=pc)
    rho = rng.uniform(200, 400, L)
    T = rng.uniform(240, 270, L)
    a = rng.uniform(1e-4, 3e-4, L)
    return make_snowpack(th, "sticky_hard_spheres", density=rho, temperature=T, radius=a, stickiness=0.2)


def seaice(rng, L=30):
    H = rng.uniform(1, 3)
    dT = rng.uniform(10, 30)
    ss = rng.uniform(0.5, 1.5)
    por = rng.uniform(0.02, 0.12)
    pc = rng.uniform(0.5e-3, 1.5e-3)
    return make_ice_column("multiyear", thickness=np.full(L, H / L), temperature=np.linspace(273.15 - dT, 273.15 - 1.8, L),
                           microstructure_model="exponential", brine_inclusion_shape="spheres",
                           salinity=np.linspace(2, 10, L) * PSU * ss, porosity=por, corr_length=pc,
                           add_water_substrate="ocean")


def simulations_of(sensor, snowpacks):
    """frequency outermost, snowpack innermost — reference smrt/core/model.py:485-502"""
    freqs = np.atleast_1d(sensor.frequency)
    sims = []
    if len(freqs) > 1:
        for sub in sensor.iterate("frequency"):
            sims += [(sub, sp) for sp in snowpacks]
    else:
        sims = [(sensor, sp) for sp in snowpacks]
    return sims


def run_case(name, emmodel, sensor, snowpacks, rtsolver_options=None, emmodel_options=None):
    rtsolver_options = rtsolver_options or {}
    t0 = time.time()
    m = make_model(emmodel, "dort", rtsolver_options=rtsolver_options, emmodel_options=emmodel_options)
    sims = simulations_of(sensor, snowpacks)
    batch = pack_simulations(sims, emmodel, emmodel_options)
    values, ks, ka, eps, angles = [], [], [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SMRTWarning)
        for sen, sp in sims:
            res = m.run(sen, sp, parallel_computation="none")
            values.append(np.asarray(res.data.values, dtype=float))
            ks.append(np.asarray(res.other_data["ks"].values, dtype=float))
            ka.append(np.asarray(res.other_data["ka"].values, dtype=float))
            eps.append(np.asarray(res.other_data["effective_permittivity"].values, dtype=complex))
            angles.append(np.asarray(res.other_data["stream_angles"].values, dtype=float))
    L = batch.L

    def pad(rows, n, dt):
        out = np.full((len(rows), n), np.nan, dtype=dt)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return out

    nmax = max(len(a) for a in angles)
    out = {f"in_{k}": v for k, v in batch.save_fields().items()}
    out.update(ref_values=np.stack(values), ref_ks=pad(ks, L, float), ref_ka=pad(ka, L, float),
               ref_eps_eff=pad(eps, L, complex), ref_stream_angles=pad(angles, nmax, float),
               ref_n_air=np.array([len(a) for a in angles]),
               rtsolver_options=np.array(json.dumps(rtsolver_options)), emmodel=np.array(str(emmodel)))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: B={batch.B} L={L} values{np.stack(values).shape} in {time.time() - t0:.1f}s  first={np.stack(values)[0].ravel()[:4]}")


def two_layer_iba():
    return make_snowpack(thickness=[0.1, 100], microstructure_model="exponential", density=[200, 400],
                         temperature=[250.0, 250.0], corr_length=[5e-5, 5e-5])


def two_layer_shs():
    return make_snowpack([0.1, 1000], "sticky_hard_spheres", density=[200, 400], temperature=[250.0, 250.0],
                         radius=[2e-4, 2e-4], stickiness=[0.1, 0.1])


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


@case
def cfg1_iba_onelayer():
    sp = make_snowpack([100], "exponential", density=[320], temperature=[270], corr_length=[5e-5])
    run_case("cfg1_iba_onelayer", "iba", sensor_list.amsre("37V"), [sp])


@case
def ref_iba_2layer_passive():  # reference test/test_integration_iba.py:33-49
    run_case("ref_iba_2layer_passive", "iba", sensor_list.amsre("37V"), [two_layer_iba()])


@case
def ref_iba_2layer_active():  # reference test/test_integration_iba.py:52-69
    run_case("ref_iba_2layer_active", "iba", sensor_list.active(frequency=19e9, theta_inc=55), [two_layer_iba()])


@case
def ref_dmrt_qcacp_2layer_passive():  # reference test/test_dmrtdort.py:41-76
    run_case("ref_dmrt_qcacp_2layer_passive", "dmrt_qcacp_shortrange", sensor_list.amsre(["19", "37"]),
             [two_layer_shs()])


@case
def ref_dmrt_less_refringent_active():  # reference test/test_dmrtdort.py:99-108
    sp = make_snowpack([0.2, 0.3], "sticky_hard_spheres", density=[290.0, 250.0], radius=1e-4, stickiness=0.2)
    run_case("ref_dmrt_less_refringent_active", "dmrt_qcacp_shortrange", sensor_list.active(10e9, 45), [sp])


@case
def ref_sea_ice_128streams():  # reference test/test_iba_sea_ice.py:30-68
    layer = 9
    thickness = np.array([1.5 / layer] * layer)
    temperature = np.linspace(273.15 - 20.0, 273.15 - 1.8, layer)
    salinity = np.linspace(2.0, 10.0, layer) * PSU
    sps = []
    for ice_type, porosity, pex in (("firstyear", 0, 500e-6), ("multiyear", 0.08, 1000e-6)):
        sps.append(make_ice_column(ice_type=ice_type, thickness=thickness, temperature=temperature,
                                   microstructure_model="exponential", brine_inclusion_shape="spheres",
                                   salinity=salinity, porosity=porosity, corr_length=np.array([pex] * layer),
                                   add_water_substrate="ocean"))
    run_case("ref_sea_ice_128streams", "iba", sensor_list.passive(1.4e9, 40.0), sps, dict(n_max_stream=128))


@case
def nonscattering_transparent():  # reference rtsolver/test_rtsolver.py:16-51, 104-113
    sps = [make_snowpack([100], "homogeneous", density=[300], temperature=[250], interface=[Transparent]),
           make_snowpack([0.5, 1000], "homogeneous", density=[300, 250], temperature=2 * [250],
                         interface=2 * [Transparent])]
    run_case("nonscattering_transparent", "nonscattering", sensor_list.passive(37e9, [0, 5, 30, 40]), sps)


@case
def nonscattering_active():  # reference rtsolver/test_rtsolver.py:64-101
    sps = [make_snowpack([0.5, 1000], "homogeneous", density=[250, 300], temperature=2 * [250],
                         interface=2 * [Transparent]),
           make_snowpack([0.5, 1000], "homogeneous", density=[300, 250], temperature=2 * [250])]
    run_case("nonscattering_active", "nonscattering", sensor_list.active(13e9, 45), sps)


@case
def iba_multiangle_passive():
    rng = np.random.default_rng(11)
    sps = [snow(rng, 5, "exponential") for _ in range(3)]
    run_case("iba_multiangle_passive", "iba", sensor_list.passive([10.65e9, 36.5e9], [0, 10, 25, 40, 55, 70, 85]),
             sps, dict(n_max_stream=16))


@case
def iba_options_passive():
    rng = np.random.default_rng(12)
    sps = [snow(rng, 8, "exponential") for _ in range(2)]
    run_case("iba_options_prune_rj", "iba", sensor_list.passive(36.5e9, [40, 55]), sps,
             dict(n_max_stream=16, prune_deep_snowpack=6, rayleigh_jeans_approximation=True))


@case
def iba_shs_active_multiangle():
    rng = np.random.default_rng(13)
    sps = []
    for _ in range(2):
        th = np.concatenate((rng.uniform(0.05, 0.5, 3), [1000.0]))
        sps.append(make_snowpack(th, "sticky_hard_spheres", density=rng.uniform(200, 400, 4),
                                 temperature=rng.uniform(240, 270, 4), radius=rng.uniform(1e-4, 3e-4, 4),
                                 stickiness=0.3))
    run_case("iba_shs_active_multiangle", "iba", sensor_list.active(13.5e9, [20, 35, 50]), sps, dict(n_max_stream=16))


@case
def iba_exp_substrate_passive():
    from smrt.substrate.flat import Flat as FlatSub

    rng = np.random.default_rng(14)
    sps = []
    for _ in range(2):
        L = 4
        sub = FlatSub(temperature=265.0, permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)))
        sps.append(make_snowpack(rng.uniform(0.05, 0.5, L), "exponential", density=rng.uniform(150, 450, L),
                                 temperature=rng.uniform(240, 272, L), corr_length=rng.uniform(5e-5, 3e-4, L),
                                 substrate=sub))
    run_case("iba_exp_substrate_passive", "iba", sensor_list.passive([18.7e9, 36.5e9], 55), sps, dict(n_max_stream=16))


@case
def cfg2_first4():
    rng = np.random.default_rng(2)
    sps = [snow(rng, 20, "exponential") for _ in range(4)]
    run_case("cfg2_first4", "iba", sensor_list.amsre(), sps, dict(n_max_stream=32))


@case
def cfg3_first4():
    rng = np.random.default_rng(3)
    sps = [snow(rng, 10, "shs") for _ in range(4)]
    run_case("cfg3_first4", "dmrt_qca_shortrange", sensor_list.active([5.4e9, 9.6e9, 13.5e9], 40), sps,
             dict(n_max_stream=16, m_max=2))


@case
def cfg4_first1():
    rng = np.random.default_rng(4)
    sps = [snow(rng, 50, "exponential")]
    freqs = [1.4135e9, 5.4e9, 6.925e9, 7.3e9, 9.6e9, 10.65e9, 13.5e9, 18.7e9, 23.8e9, 31.4e9, 36.5e9, 89.0e9]
    run_case("cfg4_first1", "iba", sensor_list.passive(freqs, 55), sps, dict(n_max_stream=64))


@case
def cfg5_first6():
    rng = np.random.default_rng(5)
    sps = [seaice(rng) for _ in range(6)]
    run_case("cfg5_first6", "iba", sensor_list.passive(1.4e9, 40), sps, dict(n_max_stream=32))


def _thin_snowpacks(seed, n, substrate_of, atmosphere_of=None, L=4):
    """thin snowpacks (the substrate matters) with one substrate / atmosphere object each"""
    rng = np.random.default_rng(seed)
    sps = []
    for i in range(n):
        sps.append(make_snowpack(rng.uniform(0.05, 0.3, L), "exponential", density=rng.uniform(150, 450, L),
                                 temperature=rng.uniform(240, 272, L), corr_length=rng.uniform(5e-5, 3e-4, L),
                                 substrate=substrate_of(i, rng),
                                 atmosphere=atmosphere_of(i, rng) if atmosphere_of else None))
    return sps


@case
def soil_wegmuller_passive():  # reference substrate/soil_wegmuller.py; both branches of the V reflectivity (theta >< 60)
    from smrt import make_soil

    sps = _thin_snowpacks(21, 3, lambda i, rng: make_soil(
        "soil_wegmuller", permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        roughness_rms=[0.001, 0.01, 0.03][i], temperature=rng.uniform(255, 272)))
    run_case("soil_wegmuller_passive", "iba", sensor_list.passive([10.65e9, 36.5e9], [30, 55, 70]), sps,
             dict(n_max_stream=16))


@case
def soil_qnh_passive():  # reference substrate/soil_qnh.py (N, Nv / Nh, Q)
    from smrt import make_soil

    args = [dict(H=0.5, Q=0.1, N=1.0), dict(H=1.2, Q=0.0, Nv=0.5, Nh=1.5), dict(H=0.2, Q=0.3)]
    sps = _thin_snowpacks(22, 3, lambda i, rng: make_soil(
        "soil_qnh", permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        temperature=rng.uniform(255, 272), **args[i]))
    run_case("soil_qnh_passive", "iba", sensor_list.passive([1.4e9, 18.7e9], [40, 55]), sps, dict(n_max_stream=16))


@case
def reflector_passive():  # reference substrate/reflector.py: scalar, polarisation dict, (frequency, polarisation) dict
    from smrt.substrate.reflector import make_reflector

    specs = [0.7, {"V": 0.6, "H": 0.8}, {(18.7e9, "H"): 0.5, (18.7e9, "V"): 0.6, (36.5e9, "H"): 0.7, (36.5e9, "V"): 0.8}]
    sps = _thin_snowpacks(23, 3, lambda i, rng: make_reflector(temperature=rng.uniform(255, 272),
                                                               specular_reflection=specs[i]))
    run_case("reflector_passive", "iba", sensor_list.passive([18.7e9, 36.5e9], 55), sps, dict(n_max_stream=16))


@case
def reflector_backscatter_active():  # reference substrate/reflector_backscatter.py: prescribed sigma0 as a DIAGONAL
    from smrt.substrate.reflector_backscatter import make_reflector  # diffuse reflection, active mode, m_max = 2 / 4

    specs = [({"V": 0.3, "H": 0.4}, {"VV": 0.1, "HH": 0.05}), (0.2, {"VV": 0.03, "HH": 0.03}),
             ({"V": 0.0, "H": 0.0}, {"VV": 0.3, "HH": 0.2})]
    sps = _thin_snowpacks(27, 3, lambda i, rng: make_reflector(temperature=rng.uniform(255, 272),
                                                               specular_reflection=specs[i][0],
                                                               backscattering_coefficient=specs[i][1]))
    run_case("reflector_backscatter_active", "iba", sensor_list.active([5.4e9, 13.5e9], [30, 40]), sps,
             dict(n_max_stream=16, m_max=2))
    run_case("reflector_backscatter_active_mmax4", "iba", sensor_list.active(13.5e9, 35), sps[:2],
             dict(n_max_stream=12, m_max=4))


@case
def reflector_backscatter_passive():  # the same substrate under a radiometer: one mode, the backscatter adds to R
    from smrt.substrate.reflector_backscatter import make_reflector

    specs = [({"V": 0.3, "H": 0.4}, {"VV": 0.1, "HH": 0.05}), (0.6, None), (None, None)]
    sps = _thin_snowpacks(28, 3, lambda i, rng: make_reflector(temperature=rng.uniform(255, 272),
                                                               specular_reflection=specs[i][0],
                                                               backscattering_coefficient=specs[i][1]))
    run_case("reflector_backscatter_passive", "iba", sensor_list.passive([18.7e9, 36.5e9], 55), sps,
             dict(n_max_stream=16))


@case
def iem_fung92_active():  # reference substrate/iem_fung92.py + iem_fung92_brogioni10.py: IEM backscatter of a rough soil
    from smrt import make_soil  # as a DIAGONAL diffuse reflection, Kirchhoff coherent part; exponential / Gaussian spectra

    args = [dict(roughness_rms=0.005, corr_length=0.05), dict(roughness_rms=0.012, corr_length=0.08,
            autocorrelation_function="gaussian"), dict(roughness_rms=0.003, corr_length=0.02, series_truncation=6)]
    sps = _thin_snowpacks(29, 3, lambda i, rng: make_soil(
        "iem_fung92", permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        temperature=rng.uniform(255, 272), **args[i]))
    run_case("iem_fung92_active", "iba", sensor_list.active([5.4e9, 13.5e9], [25, 40]), sps,
             dict(n_max_stream=16, m_max=2))
    sps = _thin_snowpacks(30, 2, lambda i, rng: make_soil(
        "iem_fung92_brogioni10", permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        temperature=rng.uniform(255, 272), roughness_rms=[0.004, 0.02][i], corr_length=[0.03, 0.15][i]))
    run_case("iem_fung92_brogioni10_active", "iba", sensor_list.active([5.4e9, 13.5e9], 35), sps,
             dict(n_max_stream=16, m_max=2))


@case
def iem_fung92_passive():  # the same soil under a radiometer ("not suitable for emissivity calculations", but it runs)
    from smrt import make_soil

    sps = _thin_snowpacks(31, 2, lambda i, rng: make_soil(
        "iem_fung92", permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        temperature=rng.uniform(255, 272), roughness_rms=[0.002, 0.006][i], corr_length=[0.03, 0.06][i]))
    run_case("iem_fung92_passive", "iba", sensor_list.passive([10.65e9, 18.7e9], 55), sps, dict(n_max_stream=16))


def _rough_surface_snowpacks(seed, n, interfaces_of, substrate_of=None, L=3):
    rng = np.random.default_rng(seed)
    sps = []
    for i in range(n):
        sps.append(make_snowpack(rng.uniform(0.1, 0.5, L), "sticky_hard_spheres", density=rng.uniform(200, 400, L),
                                 temperature=rng.uniform(245, 270, L), radius=rng.uniform(2e-4, 6e-4, L), stickiness=0.3,
                                 interface=interfaces_of(i), substrate=substrate_of(i, rng) if substrate_of else None))
    return sps


@case
def iem_fung92_interface_active():  # reference interface/iem_fung92.py as the snow SURFACE and as an internal interface
    from smrt import make_soil
    from smrt.interface.flat import Flat
    from smrt.interface.iem_fung92 import IEM_Fung92
    from smrt.interface.iem_fung92_brogioni10 import IEM_Fung92_Briogoni10

    ifaces = [[IEM_Fung92(roughness_rms=0.004, corr_length=0.05), Flat(), Flat()],
              [Flat(), IEM_Fung92(roughness_rms=0.002, corr_length=0.03, autocorrelation_function="gaussian"), Flat()],
              [IEM_Fung92_Briogoni10(roughness_rms=0.006, corr_length=0.1), Flat(),
               IEM_Fung92(roughness_rms=0.003, corr_length=0.04, series_truncation=5)]]
    soil = lambda i, rng: make_soil("iem_fung92", permittivity_model=complex(8, 1), roughness_rms=0.005,  # noqa: E731
                                    corr_length=0.06, temperature=268.0) if i == 2 else None
    sps = _rough_surface_snowpacks(33, 3, lambda i: ifaces[i], soil)
    run_case("iem_fung92_interface_active", "iba", sensor_list.active([5.4e9, 13.5e9], [30, 45]), sps,
             dict(n_max_stream=16, m_max=2))


@case
def iem_fung92_interface_passive():  # the rough surface / interface under a radiometer, with an isotropic atmosphere
    from smrt import make_soil
    from smrt.atmosphere.simple_isotropic_atmosphere import SimpleIsotropicAtmosphere
    from smrt.interface.flat import Flat
    from smrt.interface.iem_fung92 import IEM_Fung92

    ifaces = [[IEM_Fung92(roughness_rms=0.003, corr_length=0.04), Flat(), Flat()],
              [IEM_Fung92(roughness_rms=0.001, corr_length=0.02), IEM_Fung92(roughness_rms=0.002, corr_length=0.03), Flat()]]
    soil = lambda i, rng: make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.005,  # noqa: E731
                                    temperature=268.0)
    sps = _rough_surface_snowpacks(34, 2, lambda i: ifaces[i], soil)
    sps[1].atmosphere = SimpleIsotropicAtmosphere(tb_down=25.0, tb_up=8.0, transmittance=0.9)
    run_case("iem_fung92_interface_passive", "iba", sensor_list.passive([10.65e9, 18.7e9], [40, 55]), sps,
             dict(n_max_stream=16))


@case
def choudhury_passive():  # reference substrate/rough_choudhury79.py (k sigma << 1)
    from smrt.substrate.rough_choudhury79 import ChoudhuryReflectivity

    sps = _thin_snowpacks(24, 2, lambda i, rng: ChoudhuryReflectivity(
        temperature=rng.uniform(255, 272), permittivity_model=complex(rng.uniform(4, 20), rng.uniform(0.5, 5)),
        roughness_rms=[1e-4, 2e-4][i]))
    run_case("choudhury_passive", "iba", sensor_list.passive([6.925e9, 10.65e9], 55), sps, dict(n_max_stream=16))


@case
def atmosphere_passive():  # reference atmosphere/simple_isotropic_atmosphere.py: constants and frequency dicts
    from smrt import make_soil
    from smrt.atmosphere.simple_isotropic_atmosphere import SimpleIsotropicAtmosphere

    atm = [SimpleIsotropicAtmosphere(tb_down=25.0, tb_up=8.0, transmittance=0.9),
           SimpleIsotropicAtmosphere(tb_down={18.7e9: 15.2, 36.5e9: 23.5}, tb_up={18.7e9: 5.0, 36.5e9: 9.0},
                                     transmittance={18.7e9: 0.95, 36.5e9: 0.85}),
           SimpleIsotropicAtmosphere(tb_down=30.0)]
    sps = _thin_snowpacks(25, 3, lambda i, rng: make_soil("soil_wegmuller", permittivity_model=complex(10, 1),
                                                          roughness_rms=0.005, temperature=268.0) if i < 2 else None,
                          lambda i, rng: atm[i])
    run_case("atmosphere_passive", "iba", sensor_list.passive([18.7e9, 36.5e9], [35, 55]), sps, dict(n_max_stream=16))


@case
def ref_physics_law():  # reference test/test_physics_law.py:9-43, 46-95 (isothermal universe, Kirchhoff's law)
    from smrt import make_soil
    from smrt.atmosphere.simple_isotropic_atmosphere import SimpleIsotropicAtmosphere

    T = 265.0
    sps = []
    for pc, thickness in [(0.8e-3, 10), (0.05e-3, 10), (0.8e-3, 0.1)]:
        for atmosphere in (SimpleIsotropicAtmosphere(tb_down=T, tb_up=0, transmittance=1), None,
                           SimpleIsotropicAtmosphere(tb_down=1, tb_up=0, transmittance=1)):
            substrate = make_soil("soil_wegmuller", permittivity_model=complex(10, 1), roughness_rms=0.001,
                                  temperature=T)
            sps.append(make_snowpack([0.3, thickness], "exponential", density=[200, 300], temperature=T,
                                     corr_length=pc, ice_permittivity_model=complex(1.7, 0.00001),
                                     substrate=substrate, atmosphere=atmosphere))
    run_case("ref_physics_law", "iba", sensor_list.passive(37e9, list(range(10, 80, 5))), sps,
             dict(rayleigh_jeans_approximation=True))


@case
def soil_active():  # third Stokes component untouched by the rough soil models (soil_wegmuller.py:54-58)
    from smrt import make_soil

    rng = np.random.default_rng(26)
    sps = []
    for i in range(2):
        sub = make_soil(["soil_wegmuller", "soil_qnh"][i], permittivity_model=complex(8, 1.5), temperature=268.0,
                        **([dict(roughness_rms=0.01), dict(H=0.6, Q=0.1, N=1.0)][i]))
        sps.append(make_snowpack(rng.uniform(0.05, 0.3, 3), "sticky_hard_spheres", density=rng.uniform(200, 400, 3),
                                 temperature=rng.uniform(240, 270, 3), radius=rng.uniform(1e-4, 3e-4, 3),
                                 stickiness=0.3, substrate=sub))
    run_case("soil_active", "iba", sensor_list.active(13.5e9, 40), sps, dict(n_max_stream=16))


@case
def iba_microstructures_passive():  # every microstructure model with an analytical Fourier transform, mixed in one snowpack
    rng = np.random.default_rng(31)
    models = ["independent_sphere", "teubner_strey", "unified_scaled_exponential", "unified_teubner_strey",
              "unified_teubner_strey", "unified_sticky_hard_spheres", "exponential", "sticky_hard_spheres"]
    L = len(models)
    sps = []
    for _ in range(2):
        th = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
        sps.append(make_snowpack(th, models, density=rng.uniform(150, 450, L), temperature=rng.uniform(240, 272, L),
                                 radius=rng.uniform(1e-4, 3e-4, L), stickiness=0.3,
                                 corr_length=rng.uniform(5e-5, 3e-4, L), repeat_distance=rng.uniform(5e-4, 3e-3, L),
                                 porod_length=rng.uniform(5e-5, 2e-4, L),
                                 polydispersity=np.array([1.0, 1.0, 1.3, 1.5, 0.7, 1.2, 1.0, 1.0])))
    run_case("iba_microstructures_passive", "iba", sensor_list.passive([18.7e9, 36.5e9, 89e9], 55), sps,
             dict(n_max_stream=16))


@case
def iba_microstructures_active():
    rng = np.random.default_rng(32)
    models = ["teubner_strey", "unified_teubner_strey", "independent_sphere", "unified_sticky_hard_spheres"]
    L = len(models)
    th = np.concatenate((rng.uniform(0.05, 0.5, L - 1), [1000.0]))
    sp = make_snowpack(th, models, density=rng.uniform(150, 450, L), temperature=rng.uniform(240, 272, L),
                       radius=rng.uniform(1e-4, 3e-4, L), corr_length=rng.uniform(5e-5, 3e-4, L),
                       repeat_distance=rng.uniform(5e-4, 3e-3, L), porod_length=rng.uniform(5e-5, 2e-4, L),
                       polydispersity=np.array([1.0, 0.8, 1.0, 1.4]))
    run_case("iba_microstructures_active", "iba", sensor_list.active(13.5e9, 40), [sp], dict(n_max_stream=16))


@case
def rayleigh_passive_active():  # reference emmodel/rayleigh.py (sparse media), passive and active in two fixtures
    rng = np.random.default_rng(41)
    sps = []
    for _ in range(2):
        L = 3
        th = np.concatenate((rng.uniform(0.1, 1.0, L - 1), [1000.0]))
        sps.append(make_snowpack(th, "independent_sphere", density=rng.uniform(50, 150, L),
                                 temperature=rng.uniform(240, 270, L), radius=rng.uniform(1e-4, 4e-4, L)))
    run_case("rayleigh_passive", "rayleigh", sensor_list.passive([18.7e9, 36.5e9], 55), sps, dict(n_max_stream=16))
    run_case("rayleigh_active", "rayleigh", sensor_list.active(13.5e9, 40), sps[:1], dict(n_max_stream=16))


@case
def prescribed_kskaeps_passive():  # reference emmodel/prescribed_kskaeps.py: ks, ka, eps_eff set on the layers
    rng = np.random.default_rng(42)
    sps = []
    for _ in range(2):
        L = 3
        th = np.concatenate((rng.uniform(0.1, 1.0, L - 1), [1000.0]))
        sp = make_snowpack(th, "homogeneous", density=rng.uniform(200, 400, L), temperature=rng.uniform(240, 270, L))
        for lay in sp.layers:
            lay.ks = float(rng.uniform(0.05, 2.0))
            lay.ka = float(rng.uniform(0.05, 0.5))
            lay.effective_permittivity = complex(rng.uniform(1.3, 1.9), rng.uniform(1e-4, 1e-3))
        sps.append(sp)
    run_case("prescribed_kskaeps_passive", "prescribed_kskaeps", sensor_list.passive(36.5e9, [30, 55]), sps,
             dict(n_max_stream=16))


@case
def iba_variants():  # reference emmodel/iba_original.py, iba_maxwell_garnett.py, test/test_mixed_emmodel.py
    # test/test_integration_iba_original.py:12-45 (literals TbV 247.92662874568973, TbH 237.1283359660738)
    run_case("ref_iba_original_2layer_passive", "iba_original", sensor_list.amsre("37V"), [two_layer_iba()])
    # test/test_mixed_emmodel.py:9-40 (one emmodel per layer)
    run_case("ref_mixed_emmodel_passive", ["dmrt_qcacp_shortrange", "iba"], sensor_list.amsre("37V"), [two_layer_shs()])
    rng = np.random.default_rng(43)
    sps = []
    for k in range(3):
        L = 4
        th = np.concatenate((rng.uniform(0.05, 0.6, L - 1), [1000.0]))
        dens = rng.uniform(150, 450, L)
        if k == 2:
            dens[1] = 650.0  # denser than half the ice density: the medium is inverted with dense_snow_correction="auto"
        sps.append(make_snowpack(th, "exponential", density=dens, temperature=rng.uniform(235, 270, L),
                                 corr_length=rng.uniform(5e-5, 3e-4, L)))
    for em in ("iba_original", "iba_maxwell_garnett"):
        run_case(em + "_passive", em, sensor_list.passive([18.7e9, 36.5e9], [40, 55]), sps[:2], dict(n_max_stream=16))
        # iba_original on an inverted medium has ka = k0 f Im(eps_air) |y2| = 0 exactly (a conservative layer whose
        # eigenproblem is singular; the B200 path reports ERR_EIGEN for it): the dense layer is kept as it is there
        run_case(em + "_dense_active", em, sensor_list.active(13.5e9, 40), sps[2:], dict(n_max_stream=16),
                 dict(dense_snow_correction="auto") if em == "iba_maxwell_garnett" else None)


@case
def emmodel_per_medium():  # reference core/model.py:547-548, 561-566: a dict of emmodels (and options) keyed by medium
    ice = make_ice_column(ice_type="firstyear", thickness=[0.3, 0.7], temperature=[262.0, 268.0],
                          microstructure_model="exponential", brine_inclusion_shape="spheres",
                          salinity=np.array([6.0, 8.0]) * PSU, corr_length=[3e-4, 3e-4], add_water_substrate="ocean")
    snow = make_snowpack([0.1, 0.15], "exponential", density=[250, 620], temperature=[255.0, 258.0],
                         corr_length=[1e-4, 2e-4])
    run_case("emmodel_per_medium_passive", {"snow": "iba", "ice": "iba_original"},
             sensor_list.passive([6.925e9, 18.7e9], 55), [snow + ice], dict(n_max_stream=16),
             {"snow": dict(dense_snow_correction="auto"), "ice": {}})


@case
def inclusion_shapes():  # reference permittivity/generic_mixing_formula.py:88-141, depolarization_factors.py:9-46
    # first-year sea ice with brine needles / a mixture of spheres and needles (examples/iba_sea_ice.py), snow on top
    # with oblate / prolate grains (length_ratio) and explicit depolarisation factors
    thickness = np.array([0.25, 0.35, 0.4])
    temperature = np.array([258.0, 264.0, 269.0])
    salinity = np.array([5.0, 7.0, 9.0]) * PSU
    sps = []
    for shape in ("random_needles", {"spheres": 0.3, "random_needles": 0.7}):
        sps.append(make_ice_column(ice_type="firstyear", thickness=thickness, temperature=temperature,
                                   microstructure_model="exponential", brine_inclusion_shape=shape, salinity=salinity,
                                   corr_length=np.array([2e-4, 3e-4, 4e-4]), add_water_substrate="ocean"))
    snow = make_snowpack([0.1, 0.2], "exponential", density=[220, 340], temperature=[252.0, 255.0],
                         corr_length=[1.2e-4, 2.5e-4])
    snow.layers[0].length_ratio = 1.4
    snow.layers[1].length_ratio = 0.7
    sps.append(snow + sps[0])
    run_case("inclusion_shapes_passive", "iba", sensor_list.passive([1.4e9, 6.925e9, 18.7e9], [40, 55]), sps,
             dict(n_max_stream=16))
    snow2 = make_snowpack([0.3, 1000.0], "exponential", density=[250, 380], temperature=[255.0, 262.0],
                          corr_length=[1.5e-4, 3e-4])
    snow2.layers[0].depolarization_factors = np.array([0.25, 0.3, 0.45])
    snow2.layers[1].length_ratio = 1.2
    run_case("depolarization_active", "iba", sensor_list.active(13.5e9, 40), [snow2], dict(n_max_stream=16))
    for em in ("iba_original", "iba_maxwell_garnett"):
        run_case(em + "_depolarization_passive", em, sensor_list.passive(36.5e9, 55), [snow2], dict(n_max_stream=16))


@case
def high_azimuthal_modes():  # reference rtsolver/test_dort.py:13-42: m_max = 6 / 16, where scipy.linalg.eig fails
    sp1 = make_snowpack(thickness=[1000], microstructure_model="independent_sphere", density=280, temperature=265,
                        radius=0.05e-3)
    run_case("ref_rayleigh_mmax6_active", "rayleigh", sensor_list.active(10e9, 50), [sp1],
             dict(m_max=6, n_max_stream=32, diagonalization_method="schur"))
    sp2 = make_snowpack(thickness=[0.5, 1000], microstructure_model="exponential", density=[250, 330],
                        temperature=265, corr_length=[0.2e-3, 0.3e-3])
    run_case("iba_mmax5_active", "iba", sensor_list.active(13e9, [30, 50]), [sp2],
             dict(m_max=5, n_max_stream=16, diagonalization_method="schur_forcedtriu"))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
