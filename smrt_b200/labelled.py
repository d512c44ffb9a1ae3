"""A small labelled N-d array: the subset of ``xarray.DataArray`` that SMRT's ``Result`` touches.

The reference wraps every result in an ``xarray.DataArray`` (reference ``smrt/core/result.py:102-105``) and uses
``sel``, ``rename``, ``squeeze``, coordinate attribute access, element-wise arithmetic, ``to_dataframe`` and
``xr.concat(..., join="outer")`` (``result.py:170, 318, 368, 474-477, 811-827``).  xarray is an optional dependency
here: the result arrays produced by the batched B200 solve are plain numpy blocks with named axes, so this class
implements exactly that surface on top of numpy.  If the real xarray is importable, ``to_xarray()`` converts.
"""

from __future__ import annotations

import numbers
from typing import Any, Iterable, Mapping, Sequence

import numpy as np


def _as_index_array(values) -> np.ndarray:
    if isinstance(values, range):
        return np.arange(values.start, values.stop, values.step)
    arr = np.asarray(values)
    if arr.ndim == 0:
        arr = arr.reshape(1)
    return arr


class Coordinate:
    """One named axis: behaves like a 1-d array with ``.values`` (as xarray coordinates do)."""

    __slots__ = ("name", "values")

    def __init__(self, name: str, values):
        self.name = name
        self.values = _as_index_array(values)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    def __getitem__(self, key):
        return self.values[key]

    def __eq__(self, other):  # element-wise, like xarray
        return self.values == (other.values if isinstance(other, Coordinate) else other)

    def __ne__(self, other):
        return self.values != (other.values if isinstance(other, Coordinate) else other)

    __hash__ = None

    def tolist(self):
        return self.values.tolist()

    def astype(self, dtype):
        return Coordinate(self.name, self.values.astype(dtype))

    @property
    def size(self):
        return self.values.size

    @property
    def shape(self):
        return self.values.shape

    def __float__(self):
        return float(self.values.reshape(-1)[0]) if self.values.size == 1 else float(self.values)

    def __repr__(self):
        return f"Coordinate({self.name!r}, {self.values!r})"

    # arithmetic on coordinates returns plain arrays
    def __mul__(self, o):
        return self.values * np.asarray(o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self.values / np.asarray(o)

    def __add__(self, o):
        return self.values + np.asarray(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.values - np.asarray(o)


class _Coords(dict):
    """Mapping name -> Coordinate, in axis order."""

    def __iter__(self):
        return iter(self.keys())


def _match(coord_values: np.ndarray, label) -> int:
    """Exact-match lookup (xarray ``sel`` without method=)."""
    if coord_values.dtype.kind in "fc" and isinstance(label, numbers.Number):
        hits = np.nonzero(coord_values == label)[0]
    else:
        hits = np.nonzero(coord_values == label)[0]
    if hits.size == 0:
        raise KeyError(f"{label!r} not found in coordinate {coord_values!r}")
    return int(hits[0])


class DataArray:
    """numpy values + ordered named dims + one coordinate vector per dim."""

    __array_priority__ = 100

    def __init__(self, data, coords=None, dims=None, name=None, attrs=None, aux=None):
        # aux: non-index coordinates {name: (dim, values)} (created by pointwise selection, as xarray does)
        self.aux = dict(aux) if aux else {}
        if isinstance(data, DataArray):
            coords = coords if coords is not None else [(d, data.coords[d].values) for d in data.dims]
            name = name if name is not None else data.name
            attrs = attrs if attrs is not None else dict(data.attrs)
            data = data.values
        self.values = np.asarray(data)
        self.name = name
        self.attrs = dict(attrs) if attrs else {}

        ndim = self.values.ndim
        pairs: list[tuple[str, np.ndarray]] = []
        if coords is None:
            names = list(dims) if dims is not None else [f"dim_{i}" for i in range(ndim)]
            pairs = [(n, np.arange(s)) for n, s in zip(names, self.values.shape)]
        elif isinstance(coords, Mapping):
            names = list(dims) if dims is not None else list(coords.keys())
            pairs = [(n, _as_index_array(coords[n]) if n in coords else np.arange(self.values.shape[i]))
                     for i, n in enumerate(names)]
        else:
            coords = list(coords)
            for i, c in enumerate(coords):
                if isinstance(c, tuple) and len(c) == 2 and isinstance(c[0], str):
                    pairs.append((c[0], _as_index_array(c[1])))
                elif isinstance(c, Coordinate):
                    pairs.append((c.name, c.values))
                else:
                    idx_name = getattr(c, "name", None)
                    nm = dims[i] if dims is not None else (idx_name if isinstance(idx_name, str) else f"dim_{i}")
                    pairs.append((nm, _as_index_array(c)))
        if len(pairs) != ndim:
            raise ValueError(f"got {len(pairs)} coordinates for a {ndim}-d array")
        for (n, v), s in zip(pairs, self.values.shape):
            if len(v) != s:
                raise ValueError(f"coordinate {n!r} has length {len(v)} but the axis has length {s}")
        self.dims = tuple(n for n, _ in pairs)
        self.coords = _Coords((n, Coordinate(n, v)) for n, v in pairs)

    # ------------------------------------------------------------------ basics
    @property
    def shape(self):
        return self.values.shape

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def size(self):
        return self.values.size

    @property
    def dtype(self):
        return self.values.dtype

    def __len__(self):
        return len(self.values)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __float__(self):
        if self.values.size != 1:
            raise TypeError("only size-1 arrays can be converted to float")
        return float(self.values.reshape(-1)[0])

    def __complex__(self):
        return complex(self.values.reshape(-1)[0])

    def __bool__(self):
        return bool(self.values)

    def item(self):
        return self.values.item()

    def __getattr__(self, attr):
        # coordinates are reachable as attributes (result.py:132-136 relies on it)
        if attr.startswith("__") or attr in ("values", "coords", "dims", "attrs", "name"):
            raise AttributeError(attr)
        coords = self.__dict__.get("coords")
        if coords is not None and attr in coords:
            return coords[attr]
        raise AttributeError(f"'DataArray' object has no attribute {attr!r}")

    def __repr__(self):
        dims = ", ".join(f"{d}: {s}" for d, s in zip(self.dims, self.shape))
        return f"<smrt_b200.DataArray {self.name or ''} ({dims})>\n{self.values!r}"

    def copy(self):
        return self._new(self.values.copy())

    def _new(self, values, dims=None, name="__keep__"):
        dims = self.dims if dims is None else dims
        return DataArray(values, [(d, self.coords[d].values) for d in dims],
                         name=self.name if name == "__keep__" else name, attrs=self.attrs,
                         aux={k: v for k, v in self.aux.items() if v[0] in dims})

    def rename(self, name):
        out = self._new(self.values)
        out.name = name
        return out

    def astype(self, dtype):
        return self._new(self.values.astype(dtype))

    # --------------------------------------------------------------- selection
    def isel(self, drop=False, **indexers):
        index: list[Any] = [slice(None)] * self.ndim
        keep = []
        for ax, d in enumerate(self.dims):
            if d in indexers:
                i = indexers[d]
                if isinstance(i, (int, np.integer)):
                    index[ax] = int(i)
                    continue
                index[ax] = np.asarray(i) if not isinstance(i, slice) else i
            keep.append(d)
        # apply one axis at a time to avoid numpy fancy-index broadcasting between axes
        vals = self.values
        new_coords = []
        aux = {}
        out_ax = 0
        for ax, d in enumerate(self.dims):
            sel = index[ax]
            if isinstance(sel, int):
                vals = np.take(vals, sel, axis=out_ax)
                continue
            if isinstance(sel, slice) and sel == slice(None):
                new_coords.append((d, self.coords[d].values))
                aux.update({k: v for k, v in self.aux.items() if v[0] == d})
            else:
                vals = vals[(slice(None),) * out_ax + (sel,)]
                new_coords.append((d, self.coords[d].values[sel]))
                aux.update({k: (d, np.asarray(v[1])[sel]) for k, v in self.aux.items() if v[0] == d})
            out_ax += 1
        return DataArray(vals, new_coords, name=self.name, attrs=self.attrs, aux=aux)

    def sel(self, drop=False, method=None, **labels):
        # xarray semantics: indexers that are themselves labelled 1-d arrays (a coordinate or a DataArray) select
        # POINTWISE along the indexer's own dimension; plain scalars / lists select orthogonally.
        pointwise = {}
        for d, lab in list(labels.items()):
            ldim = None
            if isinstance(lab, Coordinate) and lab.values.ndim == 1 and lab.values.size > 0:
                ldim = lab.name
            elif isinstance(lab, DataArray) and lab.ndim == 1:
                ldim = lab.dims[0]
            if ldim is not None:
                pointwise.setdefault(ldim, []).append((d, np.asarray(lab.values)))
                del labels[d]
        out = self
        for ldim, group in pointwise.items():
            out = out._sel_pointwise(ldim, group)
        if pointwise:
            return out._sel_orthogonal(**labels) if labels else out
        return self._sel_orthogonal(**labels)

    def _sel_pointwise(self, new_dim, group):
        n = len(group[0][1])
        axes, pos = [], []
        for d, labs in group:
            if d not in self.dims:
                raise KeyError(f"{d!r} is not a dimension of this array (dims={self.dims})")
            if len(labs) != n:
                raise ValueError("pointwise indexers must have the same length")
            cv = self.coords[d].values
            axes.append(self.dims.index(d))
            pos.append(np.array([_match(cv, x) for x in labs], dtype=int))
        k = len(axes)
        vals = np.moveaxis(self.values, axes, range(k))[tuple(pos)]
        rest = [d for d in self.dims if d not in [g[0] for g in group]]
        insert_at = sum(1 for d in self.dims[:min(axes)] if d in rest)
        vals = np.moveaxis(vals, 0, insert_at)
        dims = rest[:insert_at] + [new_dim] + rest[insert_at:]
        own = [labs for d, labs in group if d == new_dim]
        new_values = own[0] if own else group[0][1]
        coords = [(d, new_values if d == new_dim else self.coords[d].values) for d in dims]
        aux = {k_: v for k_, v in self.aux.items() if v[0] in rest}
        for d, labs in group:
            if d != new_dim:
                aux[d] = (new_dim, labs)
        return DataArray(vals, coords, name=self.name, attrs=self.attrs, aux=aux)

    def _sel_orthogonal(self, **labels):
        indexers = {}
        for d, lab in labels.items():
            if d not in self.dims:
                raise KeyError(f"{d!r} is not a dimension of this array (dims={self.dims})")
            cv = self.coords[d].values
            if isinstance(lab, slice):
                raise NotImplementedError("slice selection is not supported")
            if isinstance(lab, (list, tuple, np.ndarray, Coordinate, range)) and not isinstance(lab, str):
                arr = lab.values if isinstance(lab, Coordinate) else np.asarray(lab)
                if arr.ndim == 0:
                    indexers[d] = _match(cv, arr.item())
                else:
                    indexers[d] = np.array([_match(cv, x) for x in arr], dtype=int)
            elif isinstance(lab, DataArray):
                arr = lab.values
                indexers[d] = (_match(cv, arr.item()) if arr.ndim == 0
                               else np.array([_match(cv, x) for x in arr], dtype=int))
            else:
                indexers[d] = _match(cv, lab)
        return self.isel(**indexers)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self.ndim - len(key))
        return self.isel(**{d: k for d, k in zip(self.dims, key)})

    def squeeze(self, dim=None, drop=False):
        dims = [d for d, s in zip(self.dims, self.shape) if s == 1 and (dim is None or d == dim or d in np.atleast_1d(dim))]
        return self.isel(**{d: 0 for d in dims})

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        order = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.values, order), [(d, self.coords[d].values) for d in dims],
                         name=self.name, attrs=self.attrs)

    # -------------------------------------------------------------- reductions
    def _reduce(self, fn, dim=None, **kw):
        if dim is None:
            return DataArray(fn(self.values, **kw), name=self.name)
        dimsl = [dim] if isinstance(dim, str) else list(dim)
        axes = tuple(self.dims.index(d) for d in dimsl)
        keep = [d for d in self.dims if d not in dimsl]
        return DataArray(fn(self.values, axis=axes, **kw), [(d, self.coords[d].values) for d in keep],
                         name=self.name, attrs=self.attrs)

    def sum(self, dim=None):
        return self._reduce(np.nansum, dim)

    def mean(self, dim=None):
        return self._reduce(np.nanmean, dim)

    def max(self, dim=None):
        return self._reduce(np.nanmax, dim)

    def min(self, dim=None):
        return self._reduce(np.nanmin, dim)

    def cumsum(self, dim):
        ax = self.dims.index(dim)
        return self._new(np.cumsum(self.values, axis=ax))

    # -------------------------------------------------------------- arithmetic
    @staticmethod
    def _align(a: "DataArray", b: "DataArray"):
        """Broadcast two labelled arrays by dimension name (coordinates must agree on shared dims)."""
        dims = list(a.dims) + [d for d in b.dims if d not in a.dims]

        def expand(x):
            order = [x.dims.index(d) for d in dims if d in x.dims]
            v = np.transpose(x.values, order)
            shape = [x.shape[x.dims.index(d)] if d in x.dims else 1 for d in dims]
            return v.reshape(shape)

        coords = []
        for d in dims:
            src = a if d in a.dims else b
            coords.append((d, src.coords[d].values))
            if d in a.dims and d in b.dims and len(a.coords[d]) != len(b.coords[d]):
                raise ValueError(f"cannot align dimension {d!r}: lengths differ")
        return expand(a), expand(b), coords

    def _binary(self, other, op, reflexive=False):
        if isinstance(other, DataArray):
            av, bv, coords = DataArray._align(self, other)
            vals = op(bv, av) if reflexive else op(av, bv)
            return DataArray(vals, coords, name=self.name, attrs=self.attrs)
        if isinstance(other, Coordinate):
            other = other.values
        vals = op(other, self.values) if reflexive else op(self.values, other)
        return self._new(vals)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __radd__(self, o):
        return self._binary(o, np.add, True)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __rsub__(self, o):
        return self._binary(o, np.subtract, True)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def __rmul__(self, o):
        return self._binary(o, np.multiply, True)

    def __truediv__(self, o):
        return self._binary(o, np.divide)

    def __rtruediv__(self, o):
        return self._binary(o, np.divide, True)

    def __pow__(self, o):
        return self._binary(o, np.power)

    def __neg__(self):
        return self._new(-self.values)

    def __abs__(self):
        return self._new(np.abs(self.values))

    def __lt__(self, o):
        return self._binary(o, np.less)

    def __le__(self, o):
        return self._binary(o, np.less_equal)

    def __gt__(self, o):
        return self._binary(o, np.greater)

    def __ge__(self, o):
        return self._binary(o, np.greater_equal)

    def __eq__(self, o):
        return self._binary(o, np.equal)

    def __ne__(self, o):
        return self._binary(o, np.not_equal)

    __hash__ = None

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        arrays = [x.values if isinstance(x, (DataArray, Coordinate)) else x for x in inputs]
        das = [x for x in inputs if isinstance(x, DataArray)]
        if len(das) == 2 and len(inputs) == 2:
            av, bv, coords = DataArray._align(das[0], das[1])
            return DataArray(ufunc(av, bv, **kwargs), coords, name=das[0].name, attrs=das[0].attrs)
        res = ufunc(*arrays, **kwargs)
        ref = das[0]
        if isinstance(res, np.ndarray) and res.shape == ref.shape:
            return ref._new(res)
        return res

    def all(self, *args, **kwargs):
        return bool(np.all(self.values))

    def any(self, *args, **kwargs):
        return bool(np.any(self.values))

    # ------------------------------------------------------------------ export
    def to_series(self, name=None):
        import pandas as pd

        if self.ndim == 0:
            raise ValueError("cannot convert a 0-d array to a series")
        if self.ndim == 1:
            idx = pd.Index(self.coords[self.dims[0]].values, name=self.dims[0])
        else:
            idx = pd.MultiIndex.from_product([self.coords[d].values for d in self.dims], names=list(self.dims))
        return pd.Series(self.values.reshape(-1), index=idx, name=name or self.name)

    def to_dataframe(self, name=None):
        nm = name or self.name
        if nm is None:
            raise ValueError("a name is required to convert to a DataFrame")
        df = self.to_series(nm).to_frame(nm)
        for k, (d, v) in reversed(list(self.aux.items())):  # non-index coordinates come first, as in xarray
            shape = [len(v) if dd == d else 1 for dd in self.dims]
            df.insert(0, k, np.broadcast_to(np.asarray(v).reshape(shape), self.shape).reshape(-1))
        return df

    def to_xarray(self):
        import xarray as xr  # optional

        return xr.DataArray(self.values, coords=[(d, self.coords[d].values) for d in self.dims],
                            name=self.name, attrs=self.attrs)

    def to_netcdf(self, filename, engine=None):
        self.to_xarray().to_netcdf(filename, engine=engine)


def concat(arrays: Sequence[DataArray], dim, join: str = "outer") -> DataArray:
    """Stack labelled arrays along a NEW leading dimension (``xr.concat(list, pd.Index, join="outer")``).

    Dimensions whose coordinates differ between members are outer-joined: the union of labels in first-seen order,
    missing cells filled with NaN — the reference relies on this to pad the ragged ``stream_angles`` diagnostics
    (reference ``smrt/core/result.py:811-815``).
    """
    arrays = [a if isinstance(a, DataArray) else DataArray(a) for a in arrays]
    if hasattr(dim, "name") and not isinstance(dim, str):
        new_name, new_values = dim.name, np.asarray(dim)
    elif isinstance(dim, tuple):
        new_name, new_values = dim[0], _as_index_array(dim[1])
    else:
        new_name, new_values = str(dim), np.arange(len(arrays))
    if len(new_values) != len(arrays):
        raise ValueError("the new coordinate must have one label per array")

    dims = list(arrays[0].dims)
    for a in arrays[1:]:
        if list(a.dims) != dims:
            raise ValueError(f"cannot concatenate arrays with different dims: {a.dims} vs {tuple(dims)}")

    union = {}
    same = True
    for d in dims:
        first = arrays[0].coords[d].values
        if all(len(a.coords[d]) == len(first) and np.array_equal(a.coords[d].values, first) for a in arrays[1:]):
            union[d] = first
            continue
        same = False
        labels = list(first.tolist())
        seen = set(labels)
        for a in arrays[1:]:
            for x in a.coords[d].values.tolist():
                if x not in seen:
                    seen.add(x)
                    labels.append(x)
        try:
            labels = sorted(labels)
        except TypeError:
            pass
        union[d] = np.asarray(labels)

    if same:
        vals = np.stack([a.values for a in arrays], axis=0)
    else:
        if join not in ("outer",):
            raise NotImplementedError("only join='outer' is supported for ragged coordinates")
        shape = (len(arrays),) + tuple(len(union[d]) for d in dims)
        dtype = np.result_type(*[a.values.dtype for a in arrays], np.float64)
        vals = np.full(shape, np.nan, dtype=dtype)
        for i, a in enumerate(arrays):
            ix = []
            for d in dims:
                pos = {x: j for j, x in enumerate(union[d].tolist())}
                ix.append(np.array([pos[x] for x in a.coords[d].values.tolist()], dtype=int))
            vals[(i,) + tuple(np.ix_(*ix))] = a.values if dims else a.values
    coords = [(new_name, new_values)] + [(d, union[d]) for d in dims]
    return DataArray(vals, coords, name=arrays[0].name, attrs=arrays[0].attrs)
