"""Result accessors of smrt_b200.result against the behaviour the reference pins in smrt/core/test_result.py
(lines 11-70: the example results; 107-158: dataframes, series, concatenation; 161-180: diagnostics), and — where the
reference is importable (authoring container) — frame-by-frame against the reference's own classes."""
import copy
import os
import sys

import numpy as np
import pandas as pd
import pytest

from smrt_b200 import result
from smrt_b200.labelled import DataArray

LAYER = ("layer", [0, 1, 2])
CHANNELS = {"VV": dict(polarization="V", polarization_inc="V"), "VH": dict(polarization="H", polarization_inc="V")}
DATA1 = [[[[4.01445680e-03, 3.77746658e-03, 0.0]], [[3.83889082e-03, 3.85904771e-03, 0.0]],
          [[2.76453599e-20, -2.73266027e-20, 0.0]]]]
BLOCK2 = [[[4e-03, 3e-03, 0], [8e-03, 6e-03, 0]], [[3e-03, 3.85904771e-03, 0], [6e-03, 6.85904771e-03, 0]],
          [[0, 0, 0], [0, 0, 0]]]


def _coords(angles):
    pol = ["V", "H", "U"]
    return [("theta", angles), ("polarization", pol), ("theta_inc", angles), ("polarization_inc", pol)]


def _other(mod, ks, ke):
    mk = lambda v: mod([float(x) for x in v], coords=[LAYER])  # noqa: E731
    return {"ks": mk(ks), "ka": mk([3, 2, 1]), "ke": mk(ke), "thickness": mk([0.1, 0.1, 0.1])}


def examples(cls, array):
    r1 = cls(DATA1, coords=_coords([35]), channel_map=copy.deepcopy(CHANNELS), other_data=_other(array, [1, 2, 3], [4, 4, 4]))
    r2 = cls([BLOCK2, BLOCK2], coords=_coords([45, 50]), channel_map=copy.deepcopy(CHANNELS),
             other_data=_other(array, [2, 4, 6], [5, 6, 7]))
    return r1, r2


RES1, RES2 = examples(result.ActiveResult, DataArray)


def test_sigma_accessors():
    assert RES1.sigmaVV() > 0 and RES1.sigmaVH() > 0 and RES1.sigmaHV() > 0 and RES1.sigmaHH() > 0
    np.testing.assert_allclose(RES1.sigmaVV_dB(), -13.8379882755357)
    np.testing.assert_allclose(RES1.sigmaVH_dB(), -14.0321985560285)
    assert RES2.sigmaVV_dB().name == "sigmaVV_dB" and RES2.sigmaHV().name == "sigmaHV"
    with pytest.raises(Exception):
        RES2.sigma(theta=45, theta_inc=50)


def test_dataframes_and_series():
    for df in (RES1.sigma_dB_as_dataframe(channel_axis="column"), RES1.to_dataframe(channel_axis="column")):
        np.testing.assert_allclose(df["VV"], -13.8379882755357)
        np.testing.assert_allclose(df["VH"], -14.0321985560285)
    long = RES1.to_dataframe(channel_axis=None)
    np.testing.assert_allclose(long.loc[(35, "V", "V"), :], (35, -13.8379882755357))  # (theta, sigma), test_result.py:130
    np.testing.assert_allclose(long.loc[(35, "H", "V"), :], (35, -14.0321985560285))
    series = RES1.to_series()
    np.testing.assert_allclose(series.loc["VV"], -13.8379882755357)
    np.testing.assert_allclose(series.loc["VH"], -14.0321985560285)
    passive = result.PassiveResult([[250.0], [240.0]], coords=[("polarization", ["V", "H"]), ("theta", [55.0])],
                                   channel_map={"37V": dict(polarization="V", theta=55.0),
                                                "37H": dict(polarization="H", theta=55.0)})
    stacked = passive.to_dataframe(channel_axis="index")
    assert list(stacked.index) == ["37V", "37H"] and list(stacked.columns) == ["Tb"]
    np.testing.assert_allclose(stacked["Tb"], [250.0, 240.0])
    with pytest.raises(Exception):
        RES1.to_dataframe(channel_axis="rows")
    with pytest.raises(Exception):
        result.ActiveResult(DATA1, coords=_coords([35])).to_dataframe(channel_axis="column")  # no channel map


def test_concat_and_diagnostics():
    both = result.concat_results((RES1, RES2), coord=("dim0", [0, 1]))
    assert "dim0" in both.data.dims and len(both.data["dim0"]) == 2
    assert both.other_data["ks"].dims == ("dim0", "layer")
    assert both.data.shape == (2, 3, 3, 3, 3) and np.isnan(np.asarray(both.data.sel(dim0=0, theta=45))).all()
    unnamed = result.concat_results([RES1, RES1], pd.Index([7, 8]))
    assert unnamed.data.dims[0] == "snowpack_index"
    with pytest.raises(Exception):
        result.concat_results([RES1, RES2], [0, 1])
    passive = result.PassiveResult([[250.0], [240.0]], coords=[("polarization", ["V", "H"]), ("theta", [55.0])])
    with pytest.raises(Exception):
        result.concat_results([RES1, passive], ("x", [0, 1]))
    np.testing.assert_allclose(RES1.ks(), [1, 2, 3])
    np.testing.assert_allclose(RES1.ka(), [3, 2, 1])
    np.testing.assert_allclose(RES1.single_scattering_albedo(), [0.25, 0.5, 0.75])
    np.testing.assert_allclose(RES1.optical_depth(), [0.4, 0.4, 0.4])


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "smrt")), reason="reference sources not on this box")
def test_frames_identical_to_reference_classes():
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(os.path.dirname(here), "oracle", "xarray_shim"), REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import xarray as xr
    from smrt.core import result as ref

    mother = pd.DataFrame({"site": ["a", "b"]}, index=pd.Index([10, 20], name="snowpack"))
    for ours_pair, ref_pair in [(examples(result.ActiveResult, DataArray), examples(ref.ActiveResult, xr.DataArray))]:
        for ours, theirs in zip(ours_pair, ref_pair):
            for axis in (None, "column", "index"):
                for natural in (False, True):
                    fn = "sigma_as_dataframe" if natural else "to_dataframe"
                    try:
                        expected = getattr(theirs, fn)(channel_axis=axis)
                    except ValueError:  # pandas refuses to stack the duplicated theta columns of a bistatic result
                        with pytest.raises(ValueError):
                            getattr(ours, fn)(channel_axis=axis)
                        continue
                    pd.testing.assert_frame_equal(getattr(ours, fn)(channel_axis=axis), expected)
            pd.testing.assert_series_equal(ours.to_series(), theirs.to_series())
            np.testing.assert_allclose(np.asarray(ours.sigma_dB(theta=ours.data.theta.values[0])),
                                       np.asarray(theirs.sigma_dB(theta=theirs.data.theta.values[0])))
        # along a snowpack dimension with the DataFrame the snowpacks came from
        o = result.concat_results([ours_pair[0], ours_pair[0]], ("snowpack", [10, 20]))
        t = ref.concat_results([ref_pair[0], ref_pair[0]], ("snowpack", [10, 20]))
        o.mother_df, t.mother_df = mother, mother
        for axis in (None, "column"):
            pd.testing.assert_frame_equal(o.to_dataframe(channel_axis=axis), t.to_dataframe(channel_axis=axis))
        unnamed = mother.rename_axis(None)
        o.mother_df, t.mother_df = unnamed, unnamed
        pd.testing.assert_frame_equal(o.to_dataframe(channel_axis=None), t.to_dataframe(channel_axis=None))
