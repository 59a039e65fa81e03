"""`MonteCarlo` / `MvnSpacecraft` / `Results` — host-side mirror of ``mc/*.rs``.

`MonteCarlo::run_until_epoch` (mc/montecarlo.rs:188-273) fans the dispersed initial states
out over a rayon pool; here the whole ensemble is ONE call of `nyxb_propagate_batch`
(one GPU) or one call per shard (multi-GPU, see ``nyx_b200.dist``).  Dispersions are sampled
on the host (mc/montecarlo.rs:277-296 is serial in the reference too) so that CPU oracle and
GPU engine receive identical inputs.  The random stream is numpy's PCG64, not rand_pcg's
Pcg64Mcg + ziggurat: draw-for-draw parity with the reference RNG is out of scope (the
reference's own MC tests assert no numbers).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence, List, Optional, Tuple

import numpy as np

from . import abi
from .cosmic import Spacecraft, pack_spacecraft
from .frames import Almanac
from .propagator import PropagationError, Propagator, status_error


@dataclass
class DispersedState:
    """`DispersedState` (mc/generator.rs:56-67)."""

    state: Spacecraft
    actual_dispersions: List[Tuple[str, float]]


_PARAMS = ("X", "Y", "Z", "VX", "VY", "VZ", "Cr", "Cd", "PropMass")


@dataclass
class StateDispersion:
    """`StateDispersion` (mc/dispersion.rs): a Gaussian dispersion of one state parameter."""

    param: object                      # nyx_b200.param.StateParameter
    mean: Optional[float] = None
    std_dev: Optional[float] = None

    @classmethod
    def zero_mean(cls, param, std_dev: float) -> "StateDispersion":
        return cls(param, 0.0, float(std_dev))


_ANGLES = ("Inclination", "RAAN", "AoP", "TrueAnomaly", "AoL", "TrueLongitude")


def _param_jacobian(param, rv: np.ndarray, mu: float) -> np.ndarray:
    """d param / d (x, y, z, vx, vy, vz) at rv — the row `OrbitGrad::partial_for` supplies in the reference
    (mc/multivariate.rs:116-145, hyperdual partials); here Richardson-extrapolated central differences of
    `nyx_b200.param.evaluate` (relative accuracy ~1e-9, far below the Monte Carlo noise it feeds)."""
    from .param import evaluate

    def f(v):
        return float(evaluate(param, v.reshape(6, 1), mu)[0])

    def diff(a, b):
        d = a - b
        if param.name in _ANGLES:
            d = (d + 180.0) % 360.0 - 180.0
        return d

    row = np.zeros(6)
    for c in range(6):
        h = 1e-4 * max(1.0, abs(rv[c]))
        e = np.zeros(6)
        e[c] = h
        d1 = diff(f(rv + e), f(rv - e)) / (2 * h)
        d2 = diff(f(rv + 0.5 * e), f(rv - 0.5 * e)) / h
        row[c] = (4.0 * d2 - d1) / 3.0
    return row


class MvnSpacecraft:
    """Multivariate-normal spacecraft state generator (mc/multivariate.rs:61-331), Cartesian form.

    ``from_spacecraft_cov`` (multivariate.rs:213-296): x = sqrt_s_v · z + mean with
    sqrt_s_v = V·sqrt(S) from the SVD of the 9x9 covariance; x is added to the template's
    [r, v, Cr, Cd, prop mass] (multivariate.rs:298-331).
    """

    def __init__(self, template: Spacecraft, cov: np.ndarray, mean: Optional[np.ndarray] = None, dispersions=None):
        self.dispersions = list(dispersions) if dispersions else None
        cov = np.asarray(cov, dtype=np.float64)
        if cov.shape != (9, 9):
            raise ValueError("covariance must be 9x9")
        evals = np.linalg.eigvalsh(0.5 * (cov + cov.T))
        if (evals < -1e-14 * max(1.0, abs(evals).max())).any():
            raise ValueError("CovarianceMatrixNotPsd")
        _, s, vt = np.linalg.svd(cov)
        self.template = template
        self.mean = np.zeros(9) if mean is None else np.asarray(mean, dtype=np.float64)
        self.sqrt_s_v = vt.T * np.sqrt(s)[None, :]

    @classmethod
    def from_spacecraft_cov(cls, template: Spacecraft, cov, mean=None) -> "MvnSpacecraft":
        return cls(template, cov, mean)

    @classmethod
    def new(cls, template: Spacecraft, dispersions) -> "MvnSpacecraft":
        """`MvnSpacecraft::new` (mc/multivariate.rs:80-211): dispersions of orbital parameters (Cartesian components or
        elements) are rotated into the Cartesian state space through the pseudo-inverse of their Jacobian,
        cov = J^+ diag(sigma^2) J^+T, mean = J^+ means; Cr, Cd and the prop mass are dispersed independently.
        Deviations from the reference as coded: the non-orbital variances go to entries (6,6), (7,7), (8,8)
        of the 9x9 (the reference writes (7,7), (8,8), (9,9) — one is out of bounds) and use `std_dev` (the reference reads
        `mean`)."""
        from .param import StateError, StateParameter

        dispersions = list(dispersions)
        cov = np.zeros((9, 9))
        mean = np.zeros(9)
        non_orbital = {StateParameter.Cr: 6, StateParameter.Cd: 7, StateParameter.PropMass: 8, StateParameter.DryMass: 8}
        orbital = [d for d in dispersions if d.param not in non_orbital]
        for d in orbital:
            if d.param in (StateParameter.TotalMass, StateParameter.Isp, StateParameter.Thrust, StateParameter.GuidanceMode):
                raise StateError(f"ReadOnly: {d.param}")
        if orbital:
            rv = template.orbit.to_cartesian_pos_vel()
            mu = template.orbit.frame.mu_km3_s2()
            jac = np.array([_param_jacobian(d.param, rv, mu) for d in orbital])
            jac_inv = np.linalg.pinv(jac)
            covar = np.diag([(d.std_dev or 0.0) ** 2 for d in orbital])
            cov[:6, :6] = jac_inv @ covar @ jac_inv.T
            mean[:6] = jac_inv @ np.array([d.mean or 0.0 for d in orbital])
        for d in dispersions:
            if d.param in non_orbital:
                q = non_orbital[d.param]
                cov[q, q] = (d.std_dev or 0.0) ** 2
                mean[q] = d.mean or 0.0
        return cls(template, cov, mean, dispersions)

    @classmethod
    def zero_mean(cls, template: Spacecraft, dispersions) -> "MvnSpacecraft":
        """multivariate.rs:198-209"""
        return cls.new(template, [StateDispersion(d.param, 0.0, d.std_dev) for d in dispersions])

    @classmethod
    def from_cartesian_std(cls, template: Spacecraft, pos_std_km, vel_std_km_s, cr_std=0.0, cd_std=0.0,
                           prop_std_kg=0.0) -> "MvnSpacecraft":
        pos = np.broadcast_to(np.asarray(pos_std_km, dtype=float), (3,))
        vel = np.broadcast_to(np.asarray(vel_std_km_s, dtype=float), (3,))
        return cls(template, np.diag(np.concatenate([pos, vel, [cr_std, cd_std, prop_std_kg]]) ** 2))

    def sample_vectors(self, rng: np.random.Generator, num: int) -> np.ndarray:
        """[num, 9] perturbation vectors (one row per run, draw order = run index)."""
        z = rng.standard_normal((num, 9))
        return z @ self.sqrt_s_v.T + self.mean[None, :]

    def sample_vectors_reference(self, seed: int, num: int, skip: int = 0) -> np.ndarray:
        """[num, 9] perturbation vectors from the REFERENCE's stream: one serial `Pcg64Mcg::new(seed)` through rand_distr's ziggurat
        `StandardNormal`, nine normals per run in component order (mc/montecarlo.rs:277-296, multivariate.rs:298-302;
        `nyxb_reference_normals`, host code of libnyxb.so).  The same z_i as nyx draws for that seed; x_i = sqrt_s_v z_i + mean then
        equals nyx's whenever the two SVDs agree on the sign and order of the singular vectors."""
        lib = abi.load_library()
        z = np.empty((num, 9))
        seed = int(seed) & ((1 << 128) - 1)
        rc = lib.nyxb_reference_normals(seed & 0xFFFFFFFFFFFFFFFF, seed >> 64, int(skip), num, z.ctypes.data)
        if rc != 0:
            raise PropagationError(f"nyxb_reference_normals rc={rc}")
        return z @ self.sqrt_s_v.T + self.mean[None, :]

    def sample_on_device(self, seed: int, num: int, first_index: int = 0, device: int = 0):
        """`nyxb_mvn_sample`: the dispersed states of runs [first_index, first_index + num) drawn on the
        GPU from the counter-based stream keyed by (seed, run index).  Returns (state[9][n], dispersion[9][n])."""
        lib = abi.load_library()
        t = np.ascontiguousarray(self.template.to_vector(), dtype=np.float64)
        mean = np.ascontiguousarray(self.mean, dtype=np.float64)
        L = np.ascontiguousarray(self.sqrt_s_v, dtype=np.float64).reshape(81)
        out = np.empty((9, num))
        disp = np.empty((9, num))
        rc = lib.nyxb_mvn_sample(int(device), int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_index), num, t.ctypes.data, mean.ctypes.data,
                                 L.ctypes.data, out.ctypes.data, disp.ctypes.data)
        if rc != 0:
            raise PropagationError(f"nyxb_mvn_sample rc={rc}: {abi.last_error()}")
        return out, disp

    def apply(self, x: np.ndarray) -> DispersedState:
        vec = self.template.to_vector() + x
        state = self.template.with_vector(self.template.epoch(), vec)
        if self.dispersions is None:
            return DispersedState(state, [(p, float(-x[i])) for i, p in enumerate(_PARAMS)])
        # multivariate.rs:318-324: template.value(param) - state.value(param) for every requested dispersion
        from .param import evaluate

        mu = self.template.orbit.frame.mu_km3_s2()
        actual = []
        for d in self.dispersions:
            a = float(evaluate(d.param, self.template.orbit.to_cartesian_pos_vel().reshape(6, 1), mu, self.template)[0])
            b = float(evaluate(d.param, state.orbit.to_cartesian_pos_vel().reshape(6, 1), mu, state)[0])
            delta = a - b
            if d.param.name in _ANGLES:
                delta = (delta + 180.0) % 360.0 - 180.0
            actual.append((d.param.name, delta))
        return DispersedState(state, actual)


@dataclass
class Run:
    """`Run` (mc/results.rs:48-59): per-run result or error, never aborting the ensemble."""

    index: int
    dispersed_state: DispersedState
    result: object  # Spacecraft | PropagationError


class MonteCarloError(RuntimeError):
    """`MonteCarloError` (mc/mod.rs): `StateError` / `NoSuccessfulRuns`."""


_RESAMPLE_CHUNK_BYTES = 256 << 20  # bound on the [6][m][n] block one resampling launch returns


@dataclass
class Results:
    """`Results<Spacecraft, PropResult<Spacecraft>>` (mc/results.rs:62-81) in structure-of-arrays form.

    `runs[i].result` is the final `Spacecraft` (or the run's error).  When the ensemble was run with `traj_capacity > 0`
    the per-step states the reference keeps in each run's `Traj` (results.rs:74-81) live in `recording` — the step-major SoA
    sink (epochs[cap][n], states[6][cap][n], count[n]) the kernels appended to — and the report accessors
    (results.rs:88-240) resample it with ONE `nyxb_traj_resample` launch per report instead of one `Traj::at` per value."""

    runs: List[Run]
    scenario: str
    final_state_soa: np.ndarray = field(repr=False, default=None)  # [9][n]
    details: np.ndarray = field(repr=False, default=None)
    status: np.ndarray = field(repr=False, default=None)
    recording: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(repr=False, default=None)
    engine: object = field(repr=False, default=None)

    def ok_runs(self) -> List[Run]:
        return [r for r in self.runs if not isinstance(r.result, Exception)]

    def total_steps(self) -> int:
        return int(self.details["n_steps"].sum())

    # ---- per-run views --------------------------------------------------------------------------------------------
    def _mu(self) -> float:
        return self.runs[0].dispersed_state.state.orbit.frame.mu_km3_s2()

    def _ok_mask(self) -> np.ndarray:
        return np.array([not isinstance(r.result, Exception) for r in self.runs], dtype=bool)

    def _run_consts(self, idx):
        """(Cr, Cd, prop mass) of the runs `idx`: constants of motion on this path (spacecraft.rs:229-236), dispersed per run."""
        st = [self.runs[i].dispersed_state.state for i in idx]
        return (np.array([x.srp.coeff_reflectivity for x in st]), np.array([x.drag.coeff_drag for x in st]),
                np.array([x.mass.prop_mass_kg for x in st]))

    def _require_recording(self):
        if self.recording is None or self.engine is None:
            raise MonteCarloError("no recorded trajectories: run the Monte Carlo with traj_capacity > 0")
        return self.recording

    def _spans(self):
        """first / last recorded epoch of every run (Traj::first / last after finalize's sort, traj.rs:75-80, 128-137)."""
        t_ep, _, t_cnt = self._require_recording()
        n = t_ep.shape[1]
        k = np.clip(np.minimum(t_cnt, t_ep.shape[0]) - 1, 0, None)
        a, b = t_ep[0, :], t_ep[k, np.arange(n)]
        return np.minimum(a, b), np.maximum(a, b)

    def _resampled(self, step_ns: int, start_ns: Optional[int], end_ns: Optional[int]):
        """`Traj::every_between` (traj.rs:153-162) for all successful runs: {run index: (epochs[k], rv[6][k])}.
        TimeSeries::inclusive(max(start, first), min(end, last), step); the device evaluates a common grid per distinct
        series start (a Monte Carlo's runs all start at the template epoch: one group, one launch per chunk)."""
        step_ns = int(step_ns)
        if step_ns <= 0:
            raise ValueError("step must be positive")
        t_ep, t_st, t_cnt = self._require_recording()
        first, last = self._spans()
        s = first if start_ns is None else np.maximum(first, int(start_ns))
        e = last if end_ns is None else np.minimum(last, int(end_ns))
        ok = self._ok_mask() & (t_cnt > 0)
        out = {}
        for s0 in np.unique(s[ok]):
            cols = np.nonzero(ok & (s == s0))[0]
            k = np.where(e[cols] >= s0, (e[cols] - s0) // step_ns + 1, 0)
            m = int(k.max()) if len(cols) else 0
            if m == 0:
                for c in cols:
                    out[int(c)] = (np.empty(0, dtype=np.int64), np.empty((6, 0)))
                continue
            grid = int(s0) + step_ns * np.arange(m, dtype=np.int64)
            whole = len(cols) == t_ep.shape[1]
            rec = self.recording if whole else (np.ascontiguousarray(t_ep[:, cols]), np.ascontiguousarray(t_st[:, :, cols]),
                                                np.ascontiguousarray(t_cnt[cols]))
            nc = len(cols)
            chunk = max(1, _RESAMPLE_CHUNK_BYTES // (48 * nc))
            rv = np.empty((6, m, nc))
            st = np.empty((m, nc), dtype=np.int32)
            for j0 in range(0, m, chunk):   # the recording is uploaded once, later chunks reuse the resident copy
                o, q = self.engine.resample(grid[j0:j0 + chunk], rec if j0 == 0 else None, n=nc)
                rv[:, j0:j0 + chunk], st[j0:j0 + chunk] = o, q
            for q, c in enumerate(cols):
                kq = int(k[q])
                bad = np.nonzero(st[:kq, q])[0]   # the iterator ends at the first epoch without data (traj_it.rs:41-59)
                if len(bad):
                    kq = int(bad[0])
                out[int(c)] = (grid[:kq], rv[:, :kq, q])
        return out

    def _report(self, param, per_run, value_if_run_failed):
        """Shared shape of the report accessors (results.rs:98-123): run-major flat list; a failed run contributes
        `value_if_run_failed` once (or nothing); a parameter a state cannot provide contributes it once per state."""
        from .param import StateError, evaluate
        report: List[float] = []
        mu = self._mu()
        for run in self.runs:
            if isinstance(run.result, Exception) or run.index not in per_run:
                if value_if_run_failed is not None:
                    report.append(float(value_if_run_failed))
                continue
            rv = per_run[run.index]
            cr, cd, pm = self._run_consts([run.index])
            try:
                vals = evaluate(param, rv, mu, run.dispersed_state.state, cr=cr[0], cd=cd[0], prop_mass_kg=pm[0])
                report.extend(np.atleast_1d(vals).tolist())
            except StateError:
                if value_if_run_failed is not None:
                    report.extend([float(value_if_run_failed)] * (rv.shape[1] if rv.ndim == 2 else 1))
        return report

    # ---- mc/results.rs:88-240 -------------------------------------------------------------------------------------
    def every_value_of_between(self, param, step_ns: int, start_ns: int, end_ns: int, value_if_run_failed: Optional[float] = None):
        """results.rs:90-124: `param` of every run every `step` between `start` and `end` (clamped to each run's span)."""
        per_run = {i: rv for i, (_, rv) in self._resampled(step_ns, start_ns, end_ns).items()}
        return self._report(param, per_run, value_if_run_failed)

    def every_value_of(self, param, step_ns: int, value_if_run_failed: Optional[float] = None):
        """results.rs:128-160: from the start to the end of each trajectory."""
        per_run = {i: rv for i, (_, rv) in self._resampled(step_ns, None, None).items()}
        return self._report(param, per_run, value_if_run_failed)

    def first_values_of(self, param, value_if_run_failed: Optional[float] = None):
        """results.rs:164-191: `traj.first()` of every run — the dispersed initial state."""
        per_run = {r.index: r.dispersed_state.state.orbit.to_cartesian_pos_vel().reshape(6, 1) for r in self.ok_runs()}
        return self._report(param, per_run, value_if_run_failed)

    def last_values_of(self, param, value_if_run_failed: Optional[float] = None):
        """results.rs:195-222: `traj.last()` of every run — the final state."""
        per_run = {r.index: self.final_state_soa[:6, r.index].reshape(6, 1) for r in self.ok_runs()}
        return self._report(param, per_run, value_if_run_failed)

    def dispersion_values_of(self, param) -> List[float]:
        """results.rs:225-240: the dispersion actually applied to `param` in every run."""
        name = param.name if hasattr(param, "name") else str(param)
        report = []
        for run in self.runs:
            for dparam, val in run.dispersed_state.actual_dispersions:
                if dparam == name:
                    report.append(val)
                    break
            else:
                raise MonteCarloError(f"StateError: {name} unavailable in the dispersions")
        return report

    def to_parquet(self, path, fields=None, start_ns: Optional[int] = None, end_ns: Optional[int] = None,
                   step_ns: Optional[int] = None, metadata: Optional[dict] = None):
        """results.rs:242-427: one row per state of every successful run — all recorded states, or, when any of
        start/end/step is given, the states interpolated every `step` (default 1 min, results.rs:285-296).
        Columns: "Epoch (UTC)", "Monte Carlo Run Index", then the requested fields (default `Spacecraft::export_params`);
        a field no state can provide is dropped (results.rs:326-343)."""
        import pyarrow as pa
        import pyarrow.parquet as pq

        from .cosmic import epochs_to_utc_iso
        from .param import EXPORT_PARAMS, StateError, evaluate

        t_ep, t_st, t_cnt = self._require_recording()
        ok = self._ok_mask()
        if not ok.any():
            raise MonteCarloError(f"NoSuccessfulRuns: export of {len(self.runs)} runs")
        if start_ns is not None or end_ns is not None or step_ns is not None:
            first_ok = int(np.nonzero(ok)[0][0])
            first, last = self._spans()
            per_run = self._resampled(60 * 10**9 if step_ns is None else step_ns,
                                      int(first[first_ok]) if start_ns is None else start_ns,
                                      int(last[first_ok]) if end_ns is None else end_ns)
        else:
            per_run = {}
            for i in np.nonzero(ok)[0]:
                k = int(min(t_cnt[i], t_ep.shape[0]))
                order = np.argsort(t_ep[:k, i], kind="stable")
                per_run[int(i)] = (t_ep[:k, i][order], t_st[:, :k, i][:, order])
        idx = sorted(per_run)
        epochs = np.concatenate([per_run[i][0] for i in idx]) if idx else np.empty(0, dtype=np.int64)
        rv = np.concatenate([per_run[i][1] for i in idx], axis=1) if idx else np.empty((6, 0))
        counts = [len(per_run[i][0]) for i in idx]
        run_index = np.repeat(np.array(idx, dtype=np.int32), counts)
        cr, cd, pm = (np.repeat(a, counts) for a in self._run_consts(idx))
        frame = self.runs[0].dispersed_state.state.orbit.frame
        cols = [pa.array(epochs_to_utc_iso(epochs), type=pa.string()), pa.array(run_index, type=pa.int32())]
        schema = [pa.field("Epoch (UTC)", pa.string(), nullable=False), pa.field("Monte Carlo Run Index", pa.int32(), nullable=False)]
        tmpl = self.runs[idx[0]].dispersed_state.state
        for f in (EXPORT_PARAMS if fields is None else fields):
            try:
                vals = evaluate(f, rv, frame.mu_km3_s2(), tmpl, cr=cr, cd=cd, prop_mass_kg=pm)
            except StateError:
                continue
            cols.append(pa.array(vals, type=pa.float64()))
            schema.append(pa.field(str(f), pa.float64(), nullable=False, metadata={"unit": f.unit, "Frame": frame.name}))
        meta = {"Purpose": "Monte Carlo Trajectory data"}
        meta.update(metadata or {})
        table = pa.Table.from_arrays(cols, schema=pa.schema(schema, metadata=meta))
        pq.write_table(table, str(path))
        return path


class MonteCarlo:
    """`MonteCarlo` (mc/montecarlo.rs:48-327)."""

    def __init__(self, nominal_state: Spacecraft, random_variable: MvnSpacecraft, scenario: str, seed: Optional[int] = None):
        self.nominal_state = nominal_state
        self.random_state = random_variable
        self.scenario = scenario
        self.seed = seed
        self.stream = "numpy"   # or "reference": see generate_states

    @classmethod
    def new(cls, nominal_state, random_variable, scenario, seed=None) -> "MonteCarlo":
        return cls(nominal_state, random_variable, scenario, seed)

    def generate_states(self, skip: int, num_runs: int, seed: Optional[int] = None, stream: Optional[str] = None) -> List[Tuple[int, DispersedState]]:
        """mc/montecarlo.rs:277-296: one serial stream; `skip` discards the first draws.  `stream` (default: `self.stream`):
        "numpy" = numpy's PCG64 + its normal sampler (vectorised; what the benchmarks and tests of this repo feed to CPU oracle and
        GPU alike), "reference" = the reference's own Pcg64Mcg + ziggurat stream (`sample_vectors_reference`)."""
        sd = self.seed if seed is None else seed
        if (stream or self.stream) == "reference":
            if sd is None:
                raise MonteCarloError("the reference stream needs a seed (the reference draws one from the OS otherwise)")
            x = self.random_state.sample_vectors_reference(sd, num_runs, skip)
        else:
            rng = np.random.Generator(np.random.PCG64(sd))
            x = self.random_state.sample_vectors(rng, skip + num_runs)[skip:]
        return [(i, self.random_state.apply(x[i])) for i in range(num_runs)]

    def generate_states_on_device(self, skip: int, num_runs: int, seed: Optional[int] = None, device: int = 0):
        """Device counterpart of `generate_states`: run i gets the draw keyed by (seed, skip + i)."""
        sd = self.seed if seed is None else seed
        st, disp = self.random_state.sample_on_device(0 if sd is None else sd, num_runs, first_index=skip, device=device)
        return st, disp

    def run_until_epoch(self, prop: Propagator, almanac: Optional[Almanac], end_epoch_ns: int, num_runs: int,
                        device_dispersions: bool = False, traj_capacity: int = 0, devices: Optional[Sequence[int]] = None) -> Results:
        """mc/montecarlo.rs:188-203.  `traj_capacity` > 0 also records every accepted step of every run (what the reference
        always keeps in `PropResult.traj`, results.rs:74-81) for the report accessors of `Results`.  `devices` = CUDA ordinals:
        the ensemble is sharded over them inside ONE C-ABI call (`nyxb_propagate_batch_multi`; final states only)."""
        if devices is not None and len(devices) > 1:
            if device_dispersions or traj_capacity:
                raise MonteCarloError("multi-device runs return final states only")
            from .dist import propagate_batch_multi

            init_states = self.generate_states(0, num_runs, self.seed)
            st, cs, ep = pack_spacecraft(ds.state for _, ds in init_states)
            engs = prop.engines(self.nominal_state.orbit.frame, almanac, devices)
            try:
                out, out_ep, det, status = propagate_batch_multi(engs, st, cs, ep, end_epoch_ns)
            finally:
                for e in engs:
                    e.close()
            runs = []
            for (idx, ds) in init_states:
                err = status_error(status[idx])
                runs.append(Run(idx, ds, err if err is not None else ds.state.with_vector(int(out_ep[idx]), out[:, idx])))
            return Results(runs, self.scenario, out, det, status, None, None)
        if device_dispersions:
            return self._run_device_dispersions(prop, almanac, 0, end_epoch_ns, num_runs, traj_capacity)
        return self.resume_run_until_epoch(prop, almanac, 0, end_epoch_ns, num_runs, traj_capacity)

    @staticmethod
    def _propagate(eng, st, cs, ep, end_epoch_ns, traj_capacity):
        """One launch; with recording, the sink grows until no run overflows it (a truncated `Traj` would end early)."""
        cap = int(traj_capacity)
        if cap <= 0:
            return eng.propagate_batch(st, cs, ep, end_epoch_ns) + (None,)
        while True:
            out, out_ep, det, status, rec = eng.propagate_batch(st, cs, ep, end_epoch_ns, traj_capacity=cap)
            if int(det["n_steps"].max()) + 1 <= cap:
                return out, out_ep, det, status, rec
            cap = int(det["n_steps"].max()) + 1

    def _run_device_dispersions(self, prop, almanac, skip, end_epoch_ns, num_runs, traj_capacity=0) -> Results:
        """mc/montecarlo.rs:208-273 with the dispersions drawn on the GPU."""
        tmpl = self.random_state.template
        st, disp = self.generate_states_on_device(skip, num_runs, self.seed, prop.device)
        cs = np.empty((4, num_runs))
        cs[0], cs[1], cs[2], cs[3] = tmpl.mass.dry_mass_kg, tmpl.mass.extra_mass_kg, tmpl.srp.area_m2, tmpl.drag.area_m2
        ep = np.full(num_runs, tmpl.epoch(), dtype=np.int64)
        eng = prop.engine(self.nominal_state.orbit.frame, almanac)
        out, out_ep, det, status, rec = self._propagate(eng, st, cs, ep, end_epoch_ns, traj_capacity)
        runs = []
        for idx in range(num_runs):
            ds = DispersedState(tmpl.with_vector(tmpl.epoch(), st[:, idx]), [(p, float(-disp[q, idx])) for q, p in enumerate(_PARAMS)])
            err = status_error(status[idx])
            runs.append(Run(idx, ds, err if err is not None else ds.state.with_vector(int(out_ep[idx]), out[:, idx])))
        return Results(runs, self.scenario, out, det, status, rec, eng)

    def resume_run_until_epoch(self, prop: Propagator, almanac: Optional[Almanac], skip: int, end_epoch_ns: int,
                               num_runs: int, traj_capacity: int = 0) -> Results:
        """mc/montecarlo.rs:208-273"""
        init_states = self.generate_states(skip, num_runs, self.seed)
        st, cs, ep = pack_spacecraft(ds.state for _, ds in init_states)
        eng = prop.engine(self.nominal_state.orbit.frame, almanac)
        out, out_ep, det, status, rec = self._propagate(eng, st, cs, ep, end_epoch_ns, traj_capacity)
        runs = []
        for (idx, ds) in init_states:
            err = status_error(status[idx])
            res = err if err is not None else ds.state.with_vector(int(out_ep[idx]), out[:, idx])
            runs.append(Run(idx, ds, res))
        return Results(runs, self.scenario, out, det, status, rec, eng)

    def run_until_nth_event(self, prop: Propagator, almanac: Optional[Almanac], max_duration_ns: int, event, trigger: int,
                            num_runs: int, traj_capacity: int = 2048) -> Results:
        return self.resume_run_until_nth_event(prop, almanac, 0, max_duration_ns, event, trigger, num_runs, traj_capacity)

    def resume_run_until_nth_event(self, prop: Propagator, almanac: Optional[Almanac], skip: int, max_duration_ns: int, event,
                                   trigger: int, num_runs: int, traj_capacity: int = 2048) -> Results:
        """mc/montecarlo.rs:115-183: every run stops at the end of the step in which `event` crossed zero for the
        `trigger`-th time (ONE device launch with the stop condition + recording), then the event epoch of every run is
        located on the recording still on the device (ONE launch of `nyxb_event_locate`, event.rs:186-211).
        A run's result is (state at the event, Traj) or the error."""
        from .trajectory import Traj

        init_states = self.generate_states(skip, num_runs, self.seed)
        st, cs, ep = pack_spacecraft(ds.state for _, ds in init_states)
        eng = prop.engine(self.nominal_state.orbit.frame, almanac)
        end_ns = self.nominal_state.epoch() + int(max_duration_ns)
        cap = int(traj_capacity)
        while True:
            out, out_ep, det, status, (t_ep, t_st, t_cnt), crossings = eng.propagate_batch(
                st, cs, ep, end_ns, traj_capacity=cap, event=(event.kind, event.value, trigger))
            if int(det["n_steps"].max()) + 1 <= cap:
                break
            cap *= 4  # a run overflowed its sink: its bracket would be lost
        ev_ep, ev_st, ev_status = eng.locate_events(event.kind, event.value, event.epoch_precision_ns, n=num_runs, run_status=status)
        runs = []
        for (idx, ds) in init_states:
            code = int(status[idx]) & 0xFF
            if code == abi.ERR_EVENT_NOT_FOUND:
                res = PropagationError(f"NthEventError: nth={trigger}, found={int(crossings[idx])}")
            elif code:
                res = status_error(status[idx])
            else:
                k = int(t_cnt[idx])
                tr = Traj(ds.state, t_ep[:k, idx].copy(), np.ascontiguousarray(t_st[:, :k, idx].T)).finalize()
                if ev_status[idx] == 0:
                    res = (tr._sc(int(ev_ep[idx]), ev_st[:, idx]), tr)
                else:
                    res = PropagationError(f"event search failed in the bracketing step (status {int(ev_status[idx])})")
            runs.append(Run(idx, ds, res))
        return Results(runs, self.scenario, out, det, status, (t_ep, t_st, t_cnt), eng)

    def __str__(self):
        return f"{self.scenario} - Nyx Monte Carlo - seed: {self.seed}"
