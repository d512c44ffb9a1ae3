"""``make_model(emmodel, "dort").run(sensor, snowpacks)`` on the B200 — the drop-in boundary of SURVEY.md §8(b).

Same call surface as the reference (``smrt/core/model.py:120-127, 310-320``): ``make_model(emmodel, rtsolver,
emmodel_options=, rtsolver_options=)`` returns an object whose ``run(sensor, snowpack, ...)`` accepts a snowpack, a
list / dict / pandas Series / DataFrame of snowpacks and returns a ``Result`` with the reference's dims, coords and
``other_data``.  What differs is *how*: instead of one Python emmodel object per layer and one ``DORT`` instance per
(snowpack, frequency) simulation fanned out over joblib processes (``model.py:395-398, 584-619``), every simulation of
the call is packed into struct-of-arrays form and solved by ONE batched call into the CUDA library; the N-d result is
assembled from the output block in one shot.

Three seams are offered:

* ``smrt_b200.make_model(...)``                       — stand-alone replacement of ``smrt.make_model`` for the DORT path
* ``smrt_b200.B200Runner``                            — a *runner* for an unmodified SMRT: ``Model.run(..., runner=B200Runner())``
* ``smrt_b200.DORT``                                  — an *rtsolver plugin* class: ``smrt.make_model("iba", smrt_b200.DORT)``
"""

from __future__ import annotations

import itertools
import warnings
from collections.abc import Mapping
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from . import capi
from .error import SMRTError, SMRTWarning, smrt_warn
from .labelled import DataArray
from .pack import MODE_ACTIVE, MODE_PASSIVE, ProblemBatch, emmodel_code, pack_simulations
from .result import ActiveResult, PassiveResult, Result, concat_results

# DORT options of the reference (smrt/rtsolver/dort.py:148-161)
_DORT_DEFAULTS = dict(n_max_stream=32, m_max=2, stream_mode="most_refringent", phase_normalization="auto",
                      phase_symmetrization=False, error_handling="exception", process_coherent_layers=False,
                      prune_deep_snowpack=None, diagonalization_method="schur_forcedtriu",
                      diagonalization_cache=False, rayleigh_jeans_approximation=False)
_DIAG_METHODS = ("eig", "schur", "schur_forcedtriu", "half_rank_eig", "stamnes88")

_SHALLOW_MSG = ("DORT has detected that the snowpack is optically shallow (tau={tau:g}) and no substrate has been set, "
                "meaning that the space under the snowpack is 'empty' with snowpack shallow enough to affect the "
                "measured signal at the surface. This is usually not wanted and can produce wrong results. Either "
                "increase the thickness of the snowpack or set a substrate. If wanted, add a transparent substrate to "
                "supress this warning")


def _is_sequence(x):
    return isinstance(x, (Sequence, np.ndarray)) and not isinstance(x, str)


def check_dort_options(options: Optional[dict]) -> dict:
    """Validate rtsolver_options against the reference's DORT signature; reject what the device path does not do."""
    opts = dict(_DORT_DEFAULTS)
    for k, v in (options or {}).items():
        if k not in opts:
            raise TypeError(f"DORT.__init__() got an unexpected keyword argument '{k}'")
        opts[k] = v
    if opts["stream_mode"] not in (None, "most_refringent"):
        raise SMRTError(f"stream_mode={opts['stream_mode']!r} is not implemented on the B200 path "
                        "(only 'most_refringent')")
    if opts["process_coherent_layers"]:
        raise SMRTError("process_coherent_layers=True is not implemented on the B200 path")
    if opts["phase_symmetrization"]:
        raise SMRTError("phase_symmetrization=True is not implemented on the B200 path")
    if opts["diagonalization_method"] not in _DIAG_METHODS:
        raise SMRTError(f"Unknown method '{opts['diagonalization_method']}' to diagonalize the matrix")
    # diagonalization_method / diagonalization_cache: accepted and ignored — the answer does not depend on the
    # eigenbasis and the device path always uses the symmetrised half-rank solve
    if opts["error_handling"] not in ("exception", "nan"):
        raise SMRTError("error_handling must be 'exception' or 'nan'")
    if not (2 <= int(opts["n_max_stream"]) <= 256):
        raise SMRTError("n_max_stream must be between 2 and 256")
    if not (0 <= int(opts["m_max"]) <= 16):
        raise SMRTError("m_max larger than 16 is not implemented on the B200 path")
    return opts


class _PlanCache:
    """One plan per (device, option set); plans own GBs of workspace, so they are reused across run() calls and the
    cache is bounded: least recently used plans are closed once the total workspace exceeds `budget_bytes` (or more
    than `max_plans` are alive)."""

    def __init__(self, budget_bytes: float = 48e9, max_plans: int = 8):
        self._plans = {}  # key -> plan, in order of last use (dicts keep insertion order)
        self.budget_bytes = budget_bytes
        self.max_plans = max_plans

    def get(self, batch: ProblemBatch, opts: dict, device: int, max_batch: int = 0) -> capi.Plan:
        o = capi.make_options(batch, n_max_stream=opts["n_max_stream"], m_max=opts["m_max"],
                              phase_normalization=opts["phase_normalization"],
                              prune_deep_snowpack=opts["prune_deep_snowpack"],
                              rayleigh_jeans_approximation=opts["rayleigh_jeans_approximation"], device=device)
        key = (device, o.mode, o.n_max_stream, o.m_max, o.max_layers, o.n_theta, o.n_inc, o.normalization,
               o.rayleigh_jeans, o.prune_deep_snowpack)
        plan = self._plans.pop(key, None)
        if plan is not None and plan.options.max_batch < max(batch.B, max_batch):
            plan.close()
            plan = None
        if plan is None:
            o.max_batch = max(batch.B, max_batch, 1)  # (max_batch: the largest batch the caller will bring)
            plan = capi.Plan(o)
        self._plans[key] = plan  # most recently used: last
        self._evict(keep=key)
        return plan

    def _evict(self, keep):
        def total():
            return sum(p.workspace_bytes for p in self._plans.values())

        while len(self._plans) > 1 and (len(self._plans) > self.max_plans or total() > self.budget_bytes):
            oldest = next(k for k in self._plans if k != keep)
            self._plans.pop(oldest).close()

    def clear(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()


_PLANS = _PlanCache()


def solve_batch(batch: ProblemBatch, rtsolver_options: Optional[dict] = None, device: int = 0) -> capi.HostOutputs:
    """Solve a packed batch on one GPU through the C ABI (host buffers in, host buffers out)."""
    opts = check_dort_options(rtsolver_options)
    plan = _PLANS.get(batch, opts, device)
    return plan.solve_host(batch)


def _raise_or_nan(out: capi.HostOutputs, opts: dict):
    err = out.status & capi.ST_ERR_MASK
    if np.any(err == capi.ST_SUBSTRATE):
        # the reference raises a plain Warning (not an SMRTError) whatever error_handling says:
        # smrt/substrate/rough_choudhury79.py:29-31
        raise Warning(capi.STATUS_MESSAGES[capi.ST_SUBSTRATE])
    if np.any(err) and opts["error_handling"] == "exception":
        b = int(np.flatnonzero(err)[0])
        raise SMRTError(f"simulation #{b}: " + capi.STATUS_MESSAGES.get(int(err[b]), "DORT failed")
                        + "\nFor mass simulations, exceptions may be annoying: return NaN instead with "
                        "rtsolver_options=dict(error_handling='nan').")
    for b in np.flatnonzero(out.status & capi.ST_WARN_SHALLOW):
        smrt_warn(_SHALLOW_MSG.format(tau=float(out.optical_depth[b])))
        break  # once per run() call is enough


def _simulation_result(mode, sensor, batch: ProblemBatch, out: capi.HostOutputs, b: int) -> Result:
    """Result of ONE simulation, with the reference's coords and other_data (rtsolver_utils.py:322-344, 373-398)."""
    n = int(batch.nlayer[b])
    layer_index = ("layer", range(n))
    ns = int(out.n_streams[b])
    other = {
        "stream_angles": DataArray(out.stream_angles[b, :ns].copy(), coords=[range(ns)]),
        "effective_permittivity": DataArray(out.eps_eff[b, :n].copy(), coords=[layer_index]),
        "ks": DataArray(out.ks[b, :n].copy(), coords=[layer_index], name="ks"),
        "ke": DataArray(out.ks[b, :n] + out.ka[b, :n], coords=[layer_index], name="ke"),
        "ka": DataArray(out.ka[b, :n].copy(), coords=[layer_index], name="ka"),
        "thickness": DataArray(batch.thickness[b, :n].copy(), coords=[layer_index], name="thickness"),
    }
    if mode == "P":
        coords = [("polarization", ["V", "H"]), ("theta", sensor.theta_deg)]
        return PassiveResult(out.values[b].copy(), coords, channel_map=sensor.channel_map, other_data=other)
    pola = ["V", "H", "U"]
    coords = [("polarization_inc", pola), ("polarization", pola), ("theta_inc", sensor.theta_inc_deg)]
    return ActiveResult(out.values[b].copy(), coords, channel_map=sensor.channel_map, other_data=other)


def _merged_channel_map(maps, dimensions):
    """The channel map the reference ends up with after stacking per-simulation results dimension by dimension
    (``concat_results``: unchanged when every result of a group has the same map, else every channel is tagged with
    the coordinate it came from, under the literal key "dim_name")."""
    for name, values in reversed(dimensions):
        n = len(values)
        merged = []
        for g in range(0, len(maps), n):
            group = maps[g:g + n]
            first = group[0]
            if all(m is first or m == first for m in group):
                merged.append(first)
            else:
                merged.append({ch: {**m[ch], "dim_name": v} for m, v in zip(group, values) for ch in m})
        maps = merged
    return maps[0]


def _nd_result(mode, sims, rows, out, dimensions) -> Result:
    """The N-d Result of a whole run in ONE shot (the reference stacks one Result per simulation with an outer-join
    ``xr.concat`` per dimension, ``smrt/core/model.py:401-404``, ``result.py:768-817``): simulations are enumerated with
    the first dimension outermost, so the output block only needs reshaping; ragged layer / stream axes are padded
    with NaN like the outer join does."""
    sensor = sims[0][0]
    lead = [(name, list(values)) for name, values in dimensions]
    shape = tuple(len(v) for _, v in lead)
    B = len(sims)
    nl = np.asarray(rows.nlayer[:B])
    Lmax = max(int(nl.max()) if B else 0, 0)
    lay = np.arange(max(Lmax, 1))[None, :] < nl[:, None]

    def per_layer(a, name=None, dtype=float):
        block = np.where(lay[:, :Lmax], np.asarray(a)[:B, :Lmax], np.nan).astype(dtype)
        return DataArray(block.reshape(shape + (Lmax,)), coords=lead + [("layer", range(Lmax))], name=name)

    ns = np.asarray(out.n_streams[:B])
    smax = int(ns.max()) if B else 0
    angles = np.where(np.arange(smax)[None, :] < ns[:, None], out.stream_angles[:B, :smax], np.nan)
    other = {
        "stream_angles": DataArray(angles.reshape(shape + (smax,)), coords=lead + [("dim_0", range(smax))]),
        "effective_permittivity": per_layer(out.eps_eff, dtype=np.complex128),
        "ks": per_layer(out.ks, "ks"),
        "ke": per_layer(np.asarray(out.ks) + np.asarray(out.ka), "ke"),
        "ka": per_layer(out.ka, "ka"),
        "thickness": per_layer(rows.thickness, "thickness"),
    }
    channel_map = _merged_channel_map([s.channel_map for s, _ in sims], dimensions)
    if mode == "P":
        coords = lead + [("polarization", ["V", "H"]), ("theta", sensor.theta_deg)]
        return PassiveResult(np.asarray(out.values[:B]).reshape(shape + out.values.shape[1:]), coords,
                             channel_map=channel_map, other_data=other)
    pola = ["V", "H", "U"]
    coords = lead + [("polarization_inc", pola), ("polarization", pola), ("theta_inc", sensor.theta_inc_deg)]
    return ActiveResult(np.asarray(out.values[:B]).reshape(shape + out.values.shape[1:]), coords,
                        channel_map=channel_map, other_data=other)


class Model:
    """Batched B200 counterpart of ``smrt.core.model.Model`` for emmodel in {iba, dmrt_qca_shortrange,
    dmrt_qcacp_shortrange, nonscattering} and rtsolver "dort"."""

    _broadcast_capability = {"theta_inc", "polarization_inc", "theta", "phi", "polarization"}  # dort.py:140-146

    def __init__(self, emmodel, rtsolver="dort", emmodel_options=None, rtsolver_options=None, device: int = 0,
                 devices: Optional[Sequence[int]] = None):
        if rtsolver is not None and not (rtsolver == "dort" or getattr(rtsolver, "__name__", "") == "DORT"):
            raise SMRTError(f"rtsolver {rtsolver!r} is not implemented on the B200 path (only 'dort')")
        self.emmodel = emmodel
        for em in (list(emmodel.values()) if isinstance(emmodel, Mapping)  # one emmodel per medium (model.py:547-548)
                   else emmodel if _is_sequence(emmodel) else [emmodel]):
            if em is not None:
                emmodel_code(em)  # fail early on unsupported models
        self.rtsolver = rtsolver
        self.emmodel_options = list(emmodel_options) if _is_sequence(emmodel_options) else dict(emmodel_options or {})
        self.rtsolver_options = dict(rtsolver_options or {})
        self.device = device
        self.devices = list(devices) if devices else [device]  # GPUs that share the simulations of one run() call
        check_dort_options(self.rtsolver_options)

    def set_rtsolver_options(self, options=None, **kwargs):
        if options is not None:
            if not isinstance(options, Mapping):
                raise SMRTError("options must be a Mapping (eg. dict)")
            self.rtsolver_options = dict(options)
        self.rtsolver_options.update(kwargs)
        check_dort_options(self.rtsolver_options)

    def set_emmodel_options(self, options=None, **kwargs):
        if options is not None:
            if not isinstance(options, Mapping):
                raise SMRTError("options must be a Mapping (eg. dict)")
            self.emmodel_options = dict(options)
        self.emmodel_options.update(kwargs)

    # ------------------------------------------------------------------------------------------------------------
    def prepare_simulations(self, sensor, snowpack, snowpack_dimension, snowpack_column):
        """Same normalisation and ordering as reference ``model.py:415-527`` (sensor axes outermost, snowpacks
        innermost); returns (flat list of (sensor, snowpack), list of (dimension name, values))."""
        if isinstance(snowpack, Mapping):
            snowpack_dimension = "snowpack", list(snowpack.keys())
            snowpack = list(snowpack.values())
        if isinstance(snowpack, pd.DataFrame):
            try:
                snowpack = snowpack[snowpack_column]
            except KeyError:
                raise SMRTError(f"the snowpack DataFrame has no column named '{snowpack_column}'. "
                                "Check the snowpack_column argument.")
        if isinstance(snowpack, pd.Series):
            name = snowpack.index.name or "snowpack"
            snowpack_dimension = name, snowpack.index.tolist()
            snowpack = snowpack.tolist()
        if _is_sequence(snowpack):
            if snowpack_dimension is None:
                snowpack_dimension = "snowpack", None
            if snowpack_dimension[1] is None:
                snowpack_dimension = snowpack_dimension[0], range(len(snowpack))
            if len(snowpack) != len(snowpack_dimension[1]):
                raise SMRTError("The list of snowpacks must have the same length as the snowpack_dimension")
        if isinstance(snowpack_dimension, tuple) and not isinstance(snowpack_dimension[0], str):
            raise SMRTError("When the 'snowpack_dimension' argument is a tuple, the first argument must be a string")

        def configurations(s):
            return [(axis, values) for axis, values in s.configurations() if axis not in self._broadcast_capability]

        def recurse(s, confs, sps):
            if confs:
                for sub in s.iterate(confs[0][0]):
                    yield from recurse(sub, confs[1:], sps)
            elif _is_sequence(sps):
                for sp in sps:
                    yield (s, sp)
            else:
                yield (s, sps)

        if _is_sequence(sensor):
            if len(sensor) != len(snowpack):
                raise SMRTError("when sensor is a sequence, the length must be the same as snowpack sequence length")
            confs = configurations(sensor[0])
            sims = list(itertools.chain(*(recurse(se, confs, sp) for se, sp in zip(sensor, snowpack))))
        else:
            confs = configurations(sensor)
            sims = list(recurse(sensor, list(confs), snowpack))
        dimensions = list(confs)
        if snowpack_dimension is not None:
            dimensions.append(snowpack_dimension)
        return sims, dimensions

    def run(self, sensor, snowpack, atmosphere=None, snowpack_dimension=None, snowpack_column="snowpack",
            progressbar=False, parallel_computation="outer", runner=None):
        """Run the model — signature of reference ``Model.run`` (``model.py:310-320``).  `parallel_computation`,
        `progressbar` and `runner` are accepted for compatibility: the whole batch is one GPU call."""
        if atmosphere is not None:
            raise DeprecationWarning("The atmosphere argument of the run method is depreciated.")
        is_sensor = lambda s: hasattr(s, "mode") and hasattr(s, "frequency")  # noqa: E731
        if not (is_sensor(sensor) or (_is_sequence(sensor) and all(is_sensor(s) for s in sensor))):
            raise SMRTError("the first argument of 'run' must be a sensor or a sequence of sensor")
        opts = check_dort_options(self.rtsolver_options)
        sims, dimensions = self.prepare_simulations(sensor, snowpack, snowpack_dimension, snowpack_column)
        if not sims:
            raise SMRTError("nothing to simulate")
        modes = {s.mode for s, _ in sims}
        if dimensions and len(modes) == 1 and len(sims) == int(np.prod([len(d[1]) for d in dimensions])):
            # the usual case: ONE N-d block built directly from the output arrays of the batched solve
            mode = modes.pop()
            rows, out = self._solve_simulations(sims, list(range(len(sims))), opts)
            result = _nd_result(mode, sims, rows, out, dimensions)
        else:  # sensors of both modes in one call, or no dimension at all: per-simulation results, stacked
            results = self._run_simulations(sims, opts)
            for dimension in reversed(dimensions):
                n = len(dimension[1])
                results = [concat_results(results[i:i + n], dimension) for i in range(0, len(results), n)]
            assert len(results) == 1, f"Results size is {len(results)=}"
            result = results[0]
        if isinstance(snowpack, pd.DataFrame):
            result.mother_df = snowpack.drop(snowpack_column, axis=1)
        return result

    # simulations per pack / solve chunk of a large run: packing (Python) of chunk k+1 overlaps the GPU call of chunk k
    CHUNK_SIMULATIONS = 12288

    def _solve_simulations(self, sims, idx, opts, atmospheres=None):
        """Pack and solve the simulations sims[i], i in idx (one sensor mode).  Returns (rows, out): per-simulation
        arrays in the order of idx — rows.nlayer / rows.thickness from the packed batch, out = the solver's output block.

        Large runs are cut into chunks of whole snowpacks (a snowpack's simulations, one per frequency, stay together
        so that it is walked once): a packing thread prepares chunk k+1 while chunk k is on the GPU (the C call
        releases the GIL), and with several ``devices`` the chunks go round-robin to one worker thread per GPU."""
        group = [sims[i] for i in idx]
        atm = [atmospheres[i] for i in idx] if atmospheres is not None else None

        def pack(sel):
            batch = pack_simulations([group[k] for k in sel], self.emmodel, self.emmodel_options,
                                     atmospheres=[atm[k] for k in sel] if atm is not None else None)
            if batch.mode == MODE_ACTIVE and not np.array_equal(batch.theta, batch.theta_inc):
                raise SMRTError("only backscatter (theta == theta_inc) is implemented on the B200 path")
            return batch

        n = len(group)
        if n <= self.CHUNK_SIMULATIONS and len(self.devices) == 1:
            batch = pack(range(n))
            out = _PLANS.get(batch, opts, self.devices[0]).solve_host(batch)
            _raise_or_nan(out, opts)
            return SimpleNamespace(nlayer=batch.nlayer, thickness=batch.thickness), out

        # chunks of whole snowpacks, in order of first appearance
        first_seen, sp_of = {}, np.empty(n, dtype=np.int64)
        for k, (_, sp) in enumerate(group):
            sp_of[k] = first_seen.setdefault(id(sp), len(first_seen))
        per_sp = max(1, n // max(len(first_seen), 1))
        sp_per_chunk = max(1, self.CHUNK_SIMULATIONS // per_sp)
        if len(self.devices) > 1:  # at least two chunks per device
            sp_per_chunk = max(1, min(sp_per_chunk, -(-len(first_seen) // (2 * len(self.devices)))))
        # a short first chunk: nothing overlaps its packing
        head = max(1, sp_per_chunk // 6)
        chunk_of = np.where(sp_of < head, 0, 1 + (sp_of - head) // sp_per_chunk)
        selections = [np.flatnonzero(chunk_of == c) for c in range(int(chunk_of.max()) + 1)]
        largest = max(len(sel) for sel in selections)

        def solve(batch, device):
            return _PLANS.get(batch, opts, device, max_batch=largest).solve_host(batch)

        packer = ThreadPoolExecutor(1)
        workers = [ThreadPoolExecutor(1) for _ in self.devices]  # one thread per GPU: a plan is used by one thread
        try:
            pending = packer.submit(pack, selections[0])
            jobs = []
            for c, sel in enumerate(selections):
                batch = pending.result()
                if c + 1 < len(selections):
                    pending = packer.submit(pack, selections[c + 1])
                d = c % len(self.devices)
                jobs.append((sel, batch, workers[d].submit(solve, batch, self.devices[d])))
            L = max(b.L for _, b, _ in jobs)
            first = jobs[0][2].result()
            out = SimpleNamespace(
                values=np.empty((n,) + first.values.shape[1:]), ks=np.zeros((n, L)), ka=np.zeros((n, L)),
                eps_eff=np.zeros((n, L), dtype=np.complex128), n_streams=np.zeros(n, dtype=np.int32),
                stream_angles=np.full((n, first.stream_angles.shape[1]), np.nan), optical_depth=np.zeros(n),
                status=np.zeros(n, dtype=np.int32))
            rows = SimpleNamespace(nlayer=np.zeros(n, dtype=np.int32), thickness=np.zeros((n, L)))
            for sel, batch, job in jobs:
                o = job.result()
                out.values[sel] = o.values
                out.ks[sel, :batch.L], out.ka[sel, :batch.L], out.eps_eff[sel, :batch.L] = o.ks, o.ka, o.eps_eff
                out.n_streams[sel], out.stream_angles[sel] = o.n_streams, o.stream_angles
                out.optical_depth[sel], out.status[sel] = o.optical_depth, o.status
                rows.nlayer[sel], rows.thickness[sel, :batch.L] = batch.nlayer, batch.thickness
        finally:
            packer.shutdown(wait=False)
            for w in workers:
                w.shutdown(wait=False)
        _raise_or_nan(out, opts)
        return rows, out

    def _run_simulations(self, sims, opts, atmospheres=None):
        """Pack, solve in one batched GPU call per sensor mode, unpack into per-simulation Results."""
        # empty snowpacks give Tb = 0 without touching the solver (reference test/test_model.py:36-43 semantics are
        # reproduced by the kernel: nlayer = 0 -> zeros)
        groups = {}
        for i, (s, sp) in enumerate(sims):
            groups.setdefault(s.mode, []).append(i)
        results = [None] * len(sims)
        for mode, idx in groups.items():
            rows, out = self._solve_simulations(sims, idx, opts, atmospheres)
            for k, i in enumerate(idx):
                results[i] = _simulation_result(mode, sims[i][0], rows, out, k)
        return results


def make_model(emmodel=None, rtsolver="dort", emmodel_options=None, rtsolver_options=None, emmodel_kwargs=None,
               rtsolver_kwargs=None, device: int = 0, devices: Optional[Sequence[int]] = None) -> Model:
    """Same signature as reference ``smrt.core.model.make_model`` (``model.py:120-177``)."""
    if emmodel_kwargs is not None:
        raise DeprecationWarning("Use emmodel_options instead of emmodel_kwargs")
    if rtsolver_kwargs is not None:
        raise DeprecationWarning("Use rtsolver_options instead of rtsolver_kwargs")
    return Model(emmodel, rtsolver, emmodel_options=emmodel_options, rtsolver_options=rtsolver_options, device=device,
                 devices=devices)


# ---------------------------------------------------------------------------------------------------------------------
# seams inside an unmodified SMRT installation
# ---------------------------------------------------------------------------------------------------------------------
class B200Runner:
    """Runner for the reference's ``Model.run(..., runner=B200Runner())`` (protocol: ``runner(function, argument_list)
    -> list of Result``, ``smrt/core/model.py:395-398``; examples ``smrt/runner/sequential_runner.py:34-48``).

    It ignores `function` (the reference's per-simulation ``run_single_simulation``), packs every
    ``((sensor, snowpack), atmosphere, parallel_computation)`` tuple and makes ONE batched GPU call.  Results are
    returned as the *reference's* Result classes when SMRT is importable (so its ``concat_results`` accepts them),
    otherwise as ``smrt_b200`` Results.
    """

    def __init__(self, device: int = 0, progressbar: bool = False):
        self.device = device

    def __call__(self, function, argument_list):
        ref_model = getattr(function, "__self__", None)
        if ref_model is None:
            raise SMRTError("B200Runner must be given Model.run_single_simulation (a bound method)")
        args = list(argument_list)
        sims = [a[0] for a in args]
        atmospheres = [a[1] for a in args]
        rtsolver = ref_model.rtsolver
        if getattr(rtsolver, "__name__", "") != "DORT" and "DORT" not in [c.__name__ for c in
                                                                           getattr(rtsolver, "__mro__", [])]:
            raise SMRTError("B200Runner accelerates the DORT rtsolver only")
        rt_opts = dict(getattr(ref_model, "rtsolver_options", {}) or {})
        m = Model(ref_model.emmodel, "dort", emmodel_options=getattr(ref_model, "emmodel_options", None),
                  rtsolver_options=rt_opts, device=self.device)
        ours = m._run_simulations(sims, check_dort_options(rt_opts), atmospheres=atmospheres)
        try:  # hand back the reference's own Result types when available
            from smrt.core import result as ref_result
            import xarray as xr

            out = []
            for r in ours:
                cls = ref_result.ActiveResult if r.mode == "A" else ref_result.PassiveResult
                to_xr = lambda d: xr.DataArray(d.values, coords=[(k, d.coords[k].values) for k in d.dims])  # noqa
                out.append(cls(to_xr(r.data), channel_map=r.channel_map,
                               other_data={k: to_xr(v) for k, v in r.other_data.items()}))
            return out
        except ImportError:
            return ours


class DORT:
    """rtsolver plugin with the reference's contract (``C(**rtsolver_options)`` once per simulation, then
    ``solve(snowpack, emmodels, sensor, atmosphere, parallel_computation=)`` -> Result; ``smrt/core/model.py:598-617``,
    minimal form ``smrt/test/test_model.py:98-103``).  One problem per call: a compatibility shim — batching needs
    ``B200Runner`` or ``smrt_b200.make_model``."""

    _broadcast_capability = {"theta_inc", "polarization_inc", "theta", "phi", "polarization"}

    def __init__(self, device: int = 0, **rtsolver_options):
        self.options = check_dort_options(rtsolver_options)
        self.device = device

    @staticmethod
    def _instance_options(em, layer):
        """Options the reference has already applied to an emmodel INSTANCE (``smrt/core/model.py:529-582`` builds
        them with the model-level emmodel_options, which this seam never sees).  The one the device path implements is
        dense_snow_correction="auto": IBA and the DMRT models then hold the inverted medium of a layer with
        frac_volume > 0.5 (``iba.py:95-106``, ``dmrt_qca_shortrange.py:68-69``, ``core/layer.py:186-201``), which shows
        in the instance's own frac_volume."""
        f_em = getattr(em, "frac_volume", None)
        f_layer = float(layer.frac_volume)
        inverted = f_em is not None and f_layer > 0.5 and not np.isclose(float(f_em), f_layer)
        return dict(dense_snow_correction="auto" if inverted else None)

    def solve(self, snowpack, emmodels, sensor, atmosphere=None, parallel_computation=None):
        ems = [type(em) for em in emmodels]
        em_opts = [self._instance_options(em, layer) for em, layer in zip(emmodels, snowpack.layers)]
        batch = pack_simulations([(sensor, snowpack)], ems, em_opts, atmospheres=[atmosphere])
        plan = _PLANS.get(batch, self.options, self.device)
        out = plan.solve_host(batch)
        _raise_or_nan(out, self.options)
        return _simulation_result(sensor.mode, sensor, batch, out, 0)


def run_ensemble(batch: ProblemBatch, rtsolver_options: Optional[dict] = None, device: int = 0):
    """Array-in / array-out entry point for large ensembles: a packed ``ProblemBatch`` (e.g.
    from ``pack_snow_ensemble``) in, the raw output block out — no Python object per member."""
    opts = check_dort_options(rtsolver_options)
    out = _PLANS.get(batch, opts, device).solve_host(batch)
    if opts["error_handling"] == "exception":
        _raise_or_nan(out, opts)
    return out
