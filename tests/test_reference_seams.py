"""Drop-in seams inside the UNMODIFIED reference: `Model.run(..., runner=B200Runner())` and
`make_model("iba", smrt_b200.DORT)`.  Needs the reference tree (authoring container only: /root/reference does not exist
on the GPU box, where these tests skip); the batched solve is stood in by the SIMT-emulated kernels."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "smrt")), reason="reference tree not available")


@pytest.fixture(scope="module")
def smrt_ref():
    here = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(os.path.dirname(here), "oracle", "xarray_shim")
    added = [p for p in (shim, REF) if p not in sys.path]
    sys.path[:0] = added
    import smrt

    yield smrt
    for p in added:
        sys.path.remove(p)


@pytest.fixture(autouse=True)
def emulated_plans(monkeypatch):
    from emu_util import emu_solve
    from smrt_b200 import model as model_mod

    class EmuPlan:
        def __init__(self, opts):
            self.opts = opts

        def solve_host(self, batch):
            return emu_solve(batch, self.opts, threads=64)

    monkeypatch.setattr(model_mod._PLANS, "get", lambda batch, opts, device, max_batch=0: EmuPlan(opts))


def test_runner_seam_matches_the_reference(smrt_ref):
    import smrt_b200

    smrt = smrt_ref
    sps = [smrt.make_snowpack([0.2, 0.4, 10], "exponential", density=[220, 300, 380], temperature=[255, 260, 265],
                              corr_length=[8e-5, 1.5e-4, 2.5e-4]),
           smrt.make_snowpack([0.5, 10], "exponential", density=[250, 350], temperature=[250, 262],
                              corr_length=[1e-4, 2e-4])]
    sensor = smrt.sensor_list.passive([18.7e9, 36.5e9], 55)
    m = smrt.make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8))
    ref = m.run(sensor, sps, parallel_computation="none")
    ours = m.run(sensor, sps, runner=smrt_b200.B200Runner())
    assert type(ours) is type(ref) and ours.data.dims == ref.data.dims
    np.testing.assert_allclose(np.asarray(ours.data.values), np.asarray(ref.data.values), rtol=1e-9)
    np.testing.assert_allclose(ours.TbV(frequency=36.5e9, snowpack=1), ref.TbV(frequency=36.5e9, snowpack=1), rtol=1e-9)
    for k in ("ks", "ka", "ke", "thickness", "effective_permittivity"):
        np.testing.assert_allclose(np.asarray(ours.other_data[k].values), np.asarray(ref.other_data[k].values),
                                   rtol=1e-12, equal_nan=True)


def test_rtsolver_plugin_seam_matches_the_reference(smrt_ref):
    import smrt_b200

    smrt = smrt_ref
    sp = smrt.make_snowpack([0.1, 100], "exponential", density=[200, 400], temperature=[250.0, 250.0],
                            corr_length=[5e-5, 5e-5])
    radar = smrt.sensor_list.active(frequency=19e9, theta_inc=55)
    opts = dict(n_max_stream=8)
    ref = smrt.make_model("iba", "dort", rtsolver_options=opts).run(radar, sp, parallel_computation="none")
    ours = smrt.make_model("iba", smrt_b200.DORT, rtsolver_options=opts).run(radar, sp, parallel_computation="none")
    np.testing.assert_allclose(ours.sigmaVV(), ref.sigmaVV(), rtol=1e-6)
    np.testing.assert_allclose(ours.sigmaHH(), ref.sigmaHH(), rtol=1e-6)
    np.testing.assert_allclose(ours.sigmaHV(), ref.sigmaHV(), rtol=1e-6)


def test_rtsolver_plugin_seam_keeps_the_dense_snow_correction(smrt_ref):
    """Model-level emmodel options are applied by the reference to its emmodel INSTANCES before the rtsolver sees
    them (smrt/core/model.py:529-582): with dense_snow_correction="auto" a layer denser than 458 kg m-3 is solved as
    the inverted medium (iba.py:95-106); the plugin must recover that from the instances."""
    import smrt_b200

    smrt = smrt_ref
    sp = smrt.make_snowpack([0.3, 0.5, 100], "exponential", density=[300, 700, 850], temperature=[255.0, 260.0, 265.0],
                            corr_length=[1e-4, 2e-4, 3e-4])
    sensor = smrt.sensor_list.passive(36.5e9, 55)
    for em_opts in (dict(dense_snow_correction="auto"), {}):
        kw = dict(emmodel_options=em_opts, rtsolver_options=dict(n_max_stream=8))
        ref = smrt.make_model("iba", "dort", **kw).run(sensor, sp, parallel_computation="none")
        ours = smrt.make_model("iba", smrt_b200.DORT, **kw).run(sensor, sp, parallel_computation="none")
        np.testing.assert_allclose(ours.TbV(), ref.TbV(), rtol=1e-9)
        np.testing.assert_allclose(ours.TbH(), ref.TbH(), rtol=1e-9)
    # the two settings differ by kelvins: the test would catch a dropped option
    a = smrt.make_model("iba", "dort", emmodel_options=dict(dense_snow_correction="auto"),
                        rtsolver_options=dict(n_max_stream=8)).run(sensor, sp, parallel_computation="none").TbV()
    b = smrt.make_model("iba", "dort", rtsolver_options=dict(n_max_stream=8)).run(sensor, sp,
                                                                                 parallel_computation="none").TbV()
    assert abs(a - b) > 0.1
