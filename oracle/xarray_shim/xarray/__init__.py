"""TEST INFRASTRUCTURE ONLY — stand-in for the `xarray` package so that the unmodified reference
(/root/reference/smrt, which hard-imports xarray at core/result.py:45 and rtsolver/rtsolver_utils.py:11) can be
imported in the authoring container, where xarray is not installed and cannot be (no network).

It re-exports the product's own labelled array (smrt_b200.labelled), which implements the surface the reference
touches on the DORT path.  Only `oracle/` scripts put this directory on sys.path.
"""
from smrt_b200.labelled import DataArray, concat  # noqa: F401

__version__ = "0.0-shim"


def open_dataarray(*args, **kwargs):
    raise NotImplementedError("the xarray stand-in does not read netCDF")
