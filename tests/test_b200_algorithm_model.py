"""The NumPy model of the device algorithm (symmetrised half-rank eigen-solve + bottom-up reflection-operator
elimination, oracle/b200_algorithm.py) against the reference fixtures: checks the ALGEBRA the CUDA kernels implement,
independently of CUDA."""
import numpy as np
import pytest

from emu_util import load_golden, rel_err
from oracle import b200_algorithm as A

CASES = ["cfg1_iba_onelayer", "ref_iba_2layer_passive", "ref_iba_2layer_active", "ref_dmrt_qcacp_2layer_passive",
         "ref_dmrt_less_refringent_active", "iba_multiangle_passive", "iba_options_prune_rj",
         "iba_shs_active_multiangle", "iba_exp_substrate_passive", "nonscattering_active", "cfg3_first4", "cfg5_first6",
         "soil_wegmuller_passive", "atmosphere_passive", "ref_physics_law", "reflector_backscatter_active",
         "reflector_backscatter_passive", "iem_fung92_active",
         "iem_fung92_interface_active", "iem_fung92_interface_passive"]


@pytest.mark.parametrize("name", CASES)
def test_device_algorithm_model(name):
    d, batch, opts = load_golden(name)
    n = min(batch.B, 4)
    collect = {}
    vals = np.stack([A.solve_problem(batch.to_problem(i, opts), collect=collect)["values"] for i in range(n)])
    tol = 1e-11 if batch.mode == 0 else 1e-6
    assert rel_err(vals, d["ref_values"][:n], batch.mode) <= tol
    # the symmetrising similarity is exact: g Ps g is symmetric to rounding for every mode (m = 0, 1, 2)
    assert collect["max_asym"] < 1e-13
