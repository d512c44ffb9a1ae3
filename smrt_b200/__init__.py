"""smrt_b200 — B200-native implementation of SMRT's DORT hot path (IBA / DMRT short-range emmodels + DORT solver).

    from smrt_b200 import make_model, make_snowpack, sensor_list
    m = make_model("iba", "dort")
    res = m.run(sensor_list.amsre("37V"), make_snowpack([100], "exponential", density=[320], temperature=[270],
                                                        corr_length=[5e-5]))
    res.TbV(), res.TbH()

See DESIGN.md (path, kernels, rooflines), INTEGRATION.md (how an SMRT maintainer binds it), include/smrt_dort_b200.h
(the C ABI).  The compute path is CUDA only: without the built library / a GPU every solve raises SMRTError.
"""
from .error import SMRTError, SMRTWarning  # noqa: F401
from .inputs import (SimpleIsotropicAtmosphere, Snowpack, make_atmosphere, make_interface, make_reflector,  # noqa: F401
                     make_snowpack, make_soil, sensor_list)
from .model import B200Runner, DORT, Model, make_model, run_ensemble, solve_batch  # noqa: F401
from .pack import ProblemBatch, pack_sea_ice_ensemble, pack_simulations, pack_snow_ensemble  # noqa: F401
from .result import ActiveResult, PassiveResult, Result, concat_results, make_result  # noqa: F401

__version__ = "0.1.0"
