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
    """Ratio in dB; anything below 1e-20 reads -200 dB (reference ``smrt/utils/__init__.py:13-22``)."""
    return 10 * np.log(np.maximum(x, 1e-20)) / np.log(10.0)


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
    # Behaviour pinned by the reference's smrt/core/test_result.py (run on the xarray stand-in) and by
    # tests/test_host_api.py; written against labelled.DataArray.
    @staticmethod
    def _table(selection, label) -> pd.DataFrame:
        """A selection as a pandas frame whose value column is `label` (coordinates that lost their dimension stay as
        extra columns, like xarray's to_dataframe; a fully selected scalar becomes a one-row frame)."""
        if selection.dims:
            return selection.to_dataframe(name=label)
        return pd.DataFrame({label: [float(selection)]})

    def _channel_table(self, **kwargs) -> pd.DataFrame:
        """One value column per channel of the channel map, rows = the coordinates common to every channel."""
        if not self.channel_map:
            raise SMRTError("No channel information is given in the result. Unable to index the result by channel.")
        tables = [self._table(self.sel_data(channel=ch, **kwargs), ch) for ch in self.channel_map]
        return pd.concat(tables, axis=1, join="inner")

    def _with_mother_columns(self, table: pd.DataFrame, wide: bool) -> pd.DataFrame:
        """Append the columns of the snowpack DataFrame given to Model.run (mother_df)."""
        mother = self.mother_df
        if wide:  # one row per snowpack, in the order of the mother frame: positional
            out = pd.concat([table.reset_index(drop=True), mother.reset_index(drop=True)], axis=1)
            out.index = mother.index
            return out
        if not mother.index.is_unique:
            raise SMRTError("The index of the snowpack DataFrame in input of Model.run must be unique for "
                            "calling to_dataframe. The index is used to join the result and original DataFrame.")
        key = mother.index.names
        if key[0] is None:  # unnamed index: it is the first level of the result's index
            key = table.index.names[0]
            if key in table.columns:
                raise SMRTError("The index of the snowpack DataFrame in input of Model.run shall be named to "
                                "avoid naming conflict in to_dataframe.")
            mother = mother.rename_axis(key)
        levels = table.index.names
        return table.reset_index().join(mother, on=key).set_index(levels)

    def return_as_dataframe(self, name, channel_axis=None, **kwargs):
        if channel_axis is None:
            table = self._table(self.sel_data(**kwargs), name)
            return table if self.mother_df is None else self._with_mother_columns(table, wide=False)
        if channel_axis == "column":
            table = self._channel_table(**kwargs)
            return table if self.mother_df is None else self._with_mother_columns(table, wide=True)
        if channel_axis == "index":
            wide = self._channel_table(**kwargs)
            trivial = wide.index.name is None and wide.index.nlevels == 1 and list(wide.index) == [0]
            long = wide.stack()  # the channel becomes the innermost (unnamed) index level
            long = long.to_frame(name) if isinstance(long, pd.Series) else long
            return long.droplevel(0) if trivial else long
        raise SMRTError('channel_axis argument must be None, "column" or "index"')

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
        """Select by channel and / or coordinates; with return_backscatter ("natural" or "dB") the intensity at
        theta == theta_inc is converted to sigma0 = 4 pi cos(theta) I."""
        selectors = dict(kwargs)
        if channel is not None:
            selectors.update((k, v) for k, v in self.channel_map[channel].items() if k in self.data.dims)
        if not return_backscatter:
            return self.data.sel(drop=True, **selectors)

        # the backscatter direction: one angle for both theta and theta_inc
        requested = [selectors.pop(k) for k in ("theta", "theta_inc") if k in selectors]
        requested = [a for a in requested if a is not None]
        if len(requested) == 2 and not np.all(requested[0] == requested[1]):
            raise SMRTError("theta and theta_inc must be the same when returning backscatter")
        angle = requested[0] if requested else self.data.theta_inc
        bistatic = "theta" in self.data.coords  # the solver kept a separate viewing-angle axis

        def at(a):
            both = dict(theta=a, theta_inc=a) if bistatic else dict(theta_inc=a)
            return self.data.sel(drop=True, **both, **selectors)

        if _is_sequence(angle):
            intensity = concat([at(a) for a in angle], pd.Index(list(angle), name="theta_inc"))
        else:  # a number, or the theta_inc coordinate itself (point-wise selection of the diagonal theta == theta_inc)
            intensity = at(angle)
        cos_angle = np.cos(np.deg2rad(np.asarray(angle, dtype=float)))
        sigma0 = (4 * np.pi * cos_angle) * intensity
        return dB(sigma0) if return_backscatter == "dB" else sigma0

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
    """Stack results of one type along a new leading dimension; ragged inner dimensions (layers, streams) are padded
    with NaN (the reference's ``concat_results``, ``smrt/core/result.py:768-817``, does an outer-join ``xr.concat``).

    `coord`: ``(name, values)`` or a ``pandas.Index`` (an unnamed one is called "snowpack_index").
    """
    results = list(result_list)
    if isinstance(coord, pd.Index):
        index = coord if coord.name is not None else coord.rename("snowpack_index")
    elif isinstance(coord, tuple):
        index = pd.Index(coord[1], name=coord[0])
    else:
        raise SMRTError("unknown type for the coord argument")
    kind = type(results[0])
    if any(type(r) is not kind for r in results):
        raise SMRTError("The results are not all of the same type")

    def stack(arrays):
        return concat(list(arrays), index, join="outer")

    channel_map = results[0].channel_map
    if any(r.channel_map != channel_map for r in results):
        # results of different sensors: every channel remembers where it came from.  The reference stores that value
        # under the literal key "dim_name" (result.py:795-800), which no selection ever uses; kept for parity.
        channel_map = {ch: {**r.channel_map[ch], "dim_name": value}
                       for r, value in zip(results, index) for ch in r.channel_map}
    other = {key: stack(r.other_data[key] for r in results) for key in results[0].other_data}
    return kind(stack(r.data for r in results), channel_map=channel_map, other_data=other)
