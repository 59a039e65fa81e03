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
from typing import List, Optional, Tuple

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


class MvnSpacecraft:
    """Multivariate-normal spacecraft state generator (mc/multivariate.rs:61-331), Cartesian form.

    ``from_spacecraft_cov`` (multivariate.rs:213-296): x = sqrt_s_v · z + mean with
    sqrt_s_v = V·sqrt(S) from the SVD of the 9x9 covariance; x is added to the template's
    [r, v, Cr, Cd, prop mass] (multivariate.rs:298-331).
    """

    def __init__(self, template: Spacecraft, cov: np.ndarray, mean: Optional[np.ndarray] = None):
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
    def from_cartesian_std(cls, template: Spacecraft, pos_std_km, vel_std_km_s, cr_std=0.0, cd_std=0.0,
                           prop_std_kg=0.0) -> "MvnSpacecraft":
        pos = np.broadcast_to(np.asarray(pos_std_km, dtype=float), (3,))
        vel = np.broadcast_to(np.asarray(vel_std_km_s, dtype=float), (3,))
        return cls(template, np.diag(np.concatenate([pos, vel, [cr_std, cd_std, prop_std_kg]]) ** 2))

    def sample_vectors(self, rng: np.random.Generator, num: int) -> np.ndarray:
        """[num, 9] perturbation vectors (one row per run, draw order = run index)."""
        z = rng.standard_normal((num, 9))
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
        return DispersedState(state, [(p, float(-x[i])) for i, p in enumerate(_PARAMS)])


@dataclass
class Run:
    """`Run` (mc/results.rs:48-59): per-run result or error, never aborting the ensemble."""

    index: int
    dispersed_state: DispersedState
    result: object  # Spacecraft | PropagationError


@dataclass
class Results:
    """`Results` (mc/results.rs:62-72) in final-state-only form + SoA views for bulk consumers."""

    runs: List[Run]
    scenario: str
    final_state_soa: np.ndarray = field(repr=False, default=None)  # [9][n]
    details: np.ndarray = field(repr=False, default=None)
    status: np.ndarray = field(repr=False, default=None)

    def ok_runs(self) -> List[Run]:
        return [r for r in self.runs if not isinstance(r.result, Exception)]

    def total_steps(self) -> int:
        return int(self.details["n_steps"].sum())


class MonteCarlo:
    """`MonteCarlo` (mc/montecarlo.rs:48-327)."""

    def __init__(self, nominal_state: Spacecraft, random_variable: MvnSpacecraft, scenario: str, seed: Optional[int] = None):
        self.nominal_state = nominal_state
        self.random_state = random_variable
        self.scenario = scenario
        self.seed = seed

    @classmethod
    def new(cls, nominal_state, random_variable, scenario, seed=None) -> "MonteCarlo":
        return cls(nominal_state, random_variable, scenario, seed)

    def generate_states(self, skip: int, num_runs: int, seed: Optional[int] = None) -> List[Tuple[int, DispersedState]]:
        """mc/montecarlo.rs:277-296: one serial stream; `skip` discards the first draws."""
        rng = np.random.Generator(np.random.PCG64(self.seed if seed is None else seed))
        x = self.random_state.sample_vectors(rng, skip + num_runs)[skip:]
        return [(i, self.random_state.apply(x[i])) for i in range(num_runs)]

    def generate_states_on_device(self, skip: int, num_runs: int, seed: Optional[int] = None, device: int = 0):
        """Device counterpart of `generate_states`: run i gets the draw keyed by (seed, skip + i)."""
        sd = self.seed if seed is None else seed
        st, disp = self.random_state.sample_on_device(0 if sd is None else sd, num_runs, first_index=skip, device=device)
        return st, disp

    def run_until_epoch(self, prop: Propagator, almanac: Optional[Almanac], end_epoch_ns: int, num_runs: int,
                        device_dispersions: bool = False) -> Results:
        if device_dispersions:
            return self._run_device_dispersions(prop, almanac, 0, end_epoch_ns, num_runs)
        return self.resume_run_until_epoch(prop, almanac, 0, end_epoch_ns, num_runs)

    def _run_device_dispersions(self, prop, almanac, skip, end_epoch_ns, num_runs) -> Results:
        """mc/montecarlo.rs:208-273 with the dispersions drawn on the GPU."""
        tmpl = self.random_state.template
        st, disp = self.generate_states_on_device(skip, num_runs, self.seed, prop.device)
        cs = np.empty((4, num_runs))
        cs[0], cs[1], cs[2], cs[3] = tmpl.mass.dry_mass_kg, tmpl.mass.extra_mass_kg, tmpl.srp.area_m2, tmpl.drag.area_m2
        ep = np.full(num_runs, tmpl.epoch(), dtype=np.int64)
        eng = prop.engine(self.nominal_state.orbit.frame, almanac)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end_epoch_ns)
        runs = []
        for idx in range(num_runs):
            ds = DispersedState(tmpl.with_vector(tmpl.epoch(), st[:, idx]), [(p, float(-disp[q, idx])) for q, p in enumerate(_PARAMS)])
            err = status_error(status[idx])
            runs.append(Run(idx, ds, err if err is not None else ds.state.with_vector(int(out_ep[idx]), out[:, idx])))
        return Results(runs, self.scenario, out, det, status)

    def resume_run_until_epoch(self, prop: Propagator, almanac: Optional[Almanac], skip: int, end_epoch_ns: int,
                               num_runs: int) -> Results:
        """mc/montecarlo.rs:208-273"""
        init_states = self.generate_states(skip, num_runs, self.seed)
        st, cs, ep = pack_spacecraft(ds.state for _, ds in init_states)
        eng = prop.engine(self.nominal_state.orbit.frame, almanac)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end_epoch_ns)
        runs = []
        for (idx, ds) in init_states:
            err = status_error(status[idx])
            res = err if err is not None else ds.state.with_vector(int(out_ep[idx]), out[:, idx])
            runs.append(Run(idx, ds, res))
        return Results(runs, self.scenario, out, det, status)

    def run_until_nth_event(self, prop: Propagator, almanac: Optional[Almanac], max_duration_ns: int, event, trigger: int,
                            num_runs: int, traj_capacity: int = 2048) -> Results:
        return self.resume_run_until_nth_event(prop, almanac, 0, max_duration_ns, event, trigger, num_runs, traj_capacity)

    def resume_run_until_nth_event(self, prop: Propagator, almanac: Optional[Almanac], skip: int, max_duration_ns: int, event,
                                   trigger: int, num_runs: int, traj_capacity: int = 2048) -> Results:
        """mc/montecarlo.rs:115-183: every run stops at the end of the step in which `event` crossed zero for the
        `trigger`-th time (ONE device launch with the stop condition + recording), then the event epoch is located per
        run on its recorded trajectory (event.rs:186-211).  A run's result is (state at the event, Traj) or the error."""
        from .event import locate_event
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
                res = (locate_event(tr, event), tr)
            runs.append(Run(idx, ds, res))
        return Results(runs, self.scenario, out, det, status)

    def __str__(self):
        return f"{self.scenario} - Nyx Monte Carlo - seed: {self.seed}"
