"""Result objects with the reference's access API (``smrt/core/result.py``): ``TbV() / TbH() / Tb(channel=...)``,
``sigmaVV() ... sigmaVH_dB()``, ``sigma(...)``, ``to_dataframe()``, ``other_data``, ``optical_depth()`` ...

Differences with the reference are internal only: the data block is a numpy-backed labelled array
(``smrt_b200.labelled.DataArray``; xarray is optional, ``.to_xarray()`` converts) and the N-d result of a whole
``Model.run`` is assembled in ONE shot from the batched solve instead of O(#simulations) ``xr.concat`` calls
(reference ``smrt/core/model.py:401-404``).
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from .error import SMRTError
from .labelled import DataArray, concat


def dB(x):
    """reference ``smrt/utils/__init__.py:13``"""
    return 10 * np.log10(x)


def _is_sequence(x):
    return isinstance(x, (list, tuple, np.ndarray, pd.Series)) and not isinstance(x, str)


def _strongsqueeze(x):
    # reference result.py:820-827
    x = x.squeeze()
    return float(x) if x.size == 1 else x


class Result:
    """Base class (reference ``result.py:93-304``)."""

    mode = None

    def __init__(self, radiance, coords=None, channel_map=None, other_data=None, mother_df=None):
        if self.mode is None:
            raise SMRTError("Result base class is abstract, uses a subclass instead. The subclass must define the "
                            "'mode' attribute")
        self.data = radiance if isinstance(radiance, DataArray) else DataArray(radiance, coords)
        self.other_data = dict(other_data or {})
        self.mother_df = mother_df
        self.data.attrs["mode"] = self.mode
        self.channel_map = channel_map or {}

    @property
    def coords(self):
        return self.data.coords

    def __getattr__(self, attr):
        data = self.__dict__.get("data")
        if attr != "data" and data is not None and attr in data.coords:
            return data.coords[attr]
        raise AttributeError(f"AttributeError: '{type(self)}' object has no attribute '{attr}'")

    def save(self, filename, netcdf_engine=None):
        self.data.to_netcdf(filename, engine=netcdf_engine)

    def sel_data(self, channel=None, **kwargs):
        raise NotImplementedError("must be implemented in a subclass")

    # ------------------------------------------------------------------------------------------------- dataframes
    def return_as_dataframe(self, name, channel_axis=None, **kwargs):
        def to_df(x, nm):
            return x.to_dataframe(name=nm) if x.dims else pd.DataFrame([float(x)], columns=[nm])

        if channel_axis in ("column", "index"):
            if not self.channel_map:
                raise SMRTError("No channel information is given in the result. Unable to index the result by channel.")
            df = pd.concat([to_df(self.sel_data(channel=ch, **kwargs), ch) for ch in self.channel_map], axis=1,
                           join="inner")
            if channel_axis == "index":
                droplevel = not df.index.name and len(df.index) == 1 and df.index[0] == 0
                df = df.stack()
                if isinstance(df, pd.Series):
                    df = pd.DataFrame(df, columns=[name])
                df.index.set_names("channel", level=-1)
                if droplevel:
                    df = df.droplevel(0)
        elif channel_axis is None:
            df = to_df(self.sel_data(**kwargs), name)
        else:
            raise SMRTError('channel_axis argument must be None, "column" or "index"')

        if self.mother_df is not None:
            if channel_axis == "column":
                df = df.reset_index(drop=True).join(self.mother_df.reset_index(drop=True))
                df.index = self.mother_df.index
            elif channel_axis is None:
                if not self.mother_df.index.is_unique:
                    raise SMRTError("The index of the snowpack DataFrame in input of Model.run must be unique for "
                                    "calling to_dataframe. The index is used to join the result and original DataFrame.")
                names = self.mother_df.index.names
                if names[0] is None:
                    nm = df.index.names[0]
                    if nm in df.columns:
                        raise SMRTError("The index of the snowpack DataFrame in input of Model.run shall be named to "
                                        "avoid naming conflict in to_dataframe.")
                    mother = self.mother_df.copy()
                    mother.index.name = nm
                    names = nm
                else:
                    mother = self.mother_df
                df = df.reset_index().join(mother, on=names).set_index(df.index.names)
        return df

    def to_series(self, **kwargs):
        return self.return_as_dataframe("out", channel_axis="column", **kwargs).iloc[0]

    # ---------------------------------------------------------------------------------------------- diagnostics
    def optical_depth(self):
        if "ka" not in self.other_data or "ks" not in self.other_data:
            raise SMRTError("optical_depth requires that the RT solver provides ka, ks and thickness.")
        ke = self.other_data["ka"] + self.other_data["ks"]
        return (ke * self.other_data["thickness"]).rename("optical_depth")

    def single_scattering_albedo(self):
        if "ke" not in self.other_data or "ks" not in self.other_data:
            raise SMRTError("single_scattering_albedo requires that the RT solver provides ke and ks.")
        return (self.other_data["ks"] / self.other_data["ke"]).rename("single_scattering_albedo")

    def single_scattering_albedo_using_absorption(self):
        if "ka" not in self.other_data or "ks" not in self.other_data:
            raise SMRTError("single_scattering_albedo requires that the RT solver provides ka and ks.")
        return self.other_data["ks"] / (self.other_data["ka"] + self.other_data["ks"])

    def ks(self):
        if "ks" not in self.other_data:
            raise SMRTError("This method requires that the selected RTsolver provides ks.")
        return self.other_data["ks"]

    def ka(self):
        if "ka" not in self.other_data:
            raise SMRTError("This method requires that the select RTsolver provides ka.")
        return self.other_data["ka"]


class PassiveResult(Result):
    """reference ``result.py:307-421``"""

    mode = "P"

    def sel_data(self, channel=None, **kwargs):
        if channel is not None:
            kwargs.update({k: v for k, v in self.channel_map[channel].items() if k in self.data.dims})
        return self.data.sel(drop=True, **kwargs)

    def Tb(self, channel=None, **kwargs):
        return _strongsqueeze(self.sel_data(channel=channel, **kwargs).rename("Tb"))

    def Tb_as_dataframe(self, channel_axis=None, **kwargs):
        return self.to_dataframe(channel_axis=None, **kwargs)

    def to_dataframe(self, channel_axis="auto", **kwargs):
        if channel_axis == "auto":
            channel_axis = "column" if self.channel_map else None
        return super().return_as_dataframe(name="Tb", channel_axis=channel_axis, **kwargs)

    def TbV(self, **kwargs):
        return _strongsqueeze(self.data.sel(polarization="V", **kwargs).rename("TbV"))

    def TbH(self, **kwargs):
        return _strongsqueeze(self.data.sel(polarization="H", **kwargs).rename("TbH"))

    def polarization_ratio(self, ratio="H_V", **kwargs):
        return _strongsqueeze(self.data.sel(polarization=ratio[0], **kwargs)
                              / self.data.sel(polarization=ratio[-1], **kwargs).rename("polarization_ratio"))

    def Tb_quasiV(self, **kwargs):
        theta = np.deg2rad(self.data.theta.values)
        return self.TbV(**kwargs) * np.cos(theta) ** 2 + self.TbH(**kwargs) * np.sin(theta) ** 2

    def Tb_quasiH(self, **kwargs):
        theta = np.deg2rad(self.data.theta.values)
        return self.TbH(**kwargs) * np.cos(theta) ** 2 + self.TbV(**kwargs) * np.sin(theta) ** 2

    def __repr__(self):
        return f"PassiveResult: TbV={self.TbV()}, TbH={self.TbH()}"


class ActiveResult(Result):
    """reference ``result.py:424-690``"""

    mode = "A"

    def sel_data(self, channel=None, return_backscatter=False, **kwargs):
        if channel is not None:
            kwargs.update({k: v for k, v in self.channel_map[channel].items() if k in self.data.dims})
        if return_backscatter:
            theta = kwargs.pop("theta", None)
            theta_inc = kwargs.pop("theta_inc", None)
            if theta is not None and theta_inc is not None and not np.all(theta_inc == theta):
                raise SMRTError("theta and theta_inc must be the same when returning backscatter")
            if theta is None:
                theta = theta_inc
            if theta is None:
                theta = self.data.theta_inc

            def select_theta(x, th, **kw):
                if "theta" in x.coords:
                    return x.sel(theta=th, theta_inc=th, **kw)
                return x.sel(theta_inc=th, **kw)

            if _is_sequence(theta):
                x = concat([select_theta(self.data, t, drop=True, **kwargs) for t in theta],
                           pd.Index(theta, name="theta_inc"))
            else:
                x = select_theta(self.data, theta, drop=True, **kwargs)
            th = theta.values if hasattr(theta, "values") else theta
            x = (4 * np.pi * np.cos(np.deg2rad(th))) * x  # sigma0 = 4 pi cos(theta) I  (result.py:485)
            return dB(x) if return_backscatter == "dB" else x
        return self.data.sel(drop=True, **kwargs)

    def sigma(self, channel=None, name="sigma", **kwargs):
        return _strongsqueeze(self.sel_data(channel=channel, return_backscatter="natural", **kwargs).rename(name))

    def sigma_dB(self, name="sigma_dB", channel=None, **kwargs):
        return _strongsqueeze(self.sel_data(channel=channel, return_backscatter="dB", **kwargs).rename(name))

    def sigma_as_dataframe(self, channel_axis=None, **kwargs):
        return super().return_as_dataframe(name="sigma", channel_axis=channel_axis, return_backscatter="natural",
                                           **kwargs)

    def sigma_dB_as_dataframe(self, channel_axis=None, **kwargs):
        return self.to_dataframe(channel_axis=channel_axis, **kwargs)

    def to_dataframe(self, channel_axis=None, **kwargs):
        if channel_axis == "auto":
            channel_axis = "column" if self.channel_map else None
        return super().return_as_dataframe(name="sigma", channel_axis=channel_axis, return_backscatter="dB", **kwargs)

    def to_series(self, **kwargs):
        return super().to_series(return_backscatter="dB", **kwargs)

    # NB: the reference labels axis 0 `polarization_inc` and axis 1 `polarization` although the solver lays the array out
    # as (outgoing, incident) — sigmaHV() reads element [1, 0].  Reproduced, not "fixed".
    def sigmaVV(self, name="sigmaVV", **kwargs):
        return self.sigma(polarization_inc="V", polarization="V", name=name, **kwargs)

    def sigmaVV_dB(self, name="sigmaVV_dB", **kwargs):
        return dB(self.sigmaVV(name=name, **kwargs))

    def sigmaHH(self, name="sigmaHH", **kwargs):
        return self.sigma(polarization_inc="H", polarization="H", name=name, **kwargs)

    def sigmaHH_dB(self, name="sigmaHH_dB", **kwargs):
        return dB(self.sigmaHH(name=name, **kwargs))

    def sigmaHV(self, name="sigmaHV", **kwargs):
        return self.sigma(polarization_inc="H", polarization="V", name=name, **kwargs)

    def sigmaHV_dB(self, name="sigmaHV_dB", **kwargs):
        return dB(self.sigmaHV(name=name, **kwargs))

    def sigmaVH(self, name="sigmaVH", **kwargs):
        return self.sigma(polarization_inc="V", polarization="H", name=name, **kwargs)

    def sigmaVH_dB(self, name="sigmaVH_dB", **kwargs):
        return dB(self.sigmaVH(name=name, **kwargs))

    def __repr__(self):
        return (f"ActiveResult:sigmaVV={self.sigmaVV_dB()} dB, sigmaHH={self.sigmaHH_dB()} dB, "
                f"sigmaHV={self.sigmaHV_dB()} dB")


def make_result(sensor, *args, **kwargs):
    """reference ``result.py:79-90``; `sensor` may also be the mode letter"""
    mode = sensor if isinstance(sensor, str) else sensor.mode
    channel_map = kwargs.pop("channel_map", None)
    if channel_map is None and not isinstance(sensor, str):
        channel_map = sensor.channel_map
    cls = ActiveResult if mode == "A" else PassiveResult
    return cls(*args, channel_map=channel_map, **kwargs)


def concat_results(result_list, coord):
    """Stack results along a new dimension (reference ``result.py:768-817``)."""
    if isinstance(coord, tuple):
        dim_name, dim_value = coord
        index = pd.Index(dim_value, name=dim_name)
    elif isinstance(coord, pd.Index):
        index = coord
        if index.name is None:
            index.name = "snowpack_index"
    else:
        raise SMRTError("unknown type for the coord argument")
    cls = type(result_list[0])
    if not all(type(r) is cls for r in result_list):
        raise SMRTError("The results are not all of the same type")
    if any(r.channel_map != result_list[0].channel_map for r in result_list):
        channel_map = {ch: dict(**r.channel_map[ch], dim_name=dv) for r, dv in zip(result_list, dim_value)
                       for ch in r.channel_map}
    else:
        channel_map = result_list[0].channel_map
    data = concat([r.data for r in result_list], index, join="outer")
    other = {v: concat([r.other_data[v] for r in result_list], index, join="outer")
             for v in result_list[0].other_data}
    return cls(data, channel_map=channel_map, other_data=other)
