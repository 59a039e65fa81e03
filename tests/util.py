"""Shared helpers for the parity tests."""
import json
from pathlib import Path

import numpy as np

import nyx_b200 as nb
from nyx_b200 import abi

GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_vectors.json").read_text())
S = 10**9


def opts_from_json(o) -> nb.IntegratorOptions:
    if o["kind"] == "fixed":
        return nb.IntegratorOptions.with_fixed_step_s(o["step_s"])
    if o["kind"] == "default":
        return nb.IntegratorOptions.default()
    return nb.IntegratorOptions.with_adaptive_step_s(o["min_step_s"], o["max_step_s"], o["tolerance"],
                                                     nb.ErrorControl[o["error_ctrl"]])


def leo_state(frame, epoch_ns=0) -> nb.Spacecraft:
    x = GOLDEN["initial_state"]
    return nb.Spacecraft.from_orbit(nb.Orbit.cartesian(*x, epoch_ns, frame))


def leo_ensemble(n, seed=0, sma=6678.0, pos_std=1.0, vel_std=1e-3, frame=None, mass=None, srp=None, drag=None):
    """Example-01 orbit (examples/01_orbit_prop/main.rs:52-53) + N(0, diag(1 km, 1 m/s)) dispersions."""
    frame = frame or nb.EARTH_J2000
    orbit = nb.Orbit.keplerian(sma, 0.015, 68.5, 65.2, 75.0, 0.0, 0, frame)
    template = nb.Spacecraft(orbit=orbit, mass=mass or nb.Mass(100.0, 20.0, 0.0), srp=srp or nb.SRPData(),
                             drag=drag or nb.DragData())
    mvn = nb.MvnSpacecraft.from_cartesian_std(template, pos_std, vel_std)
    mc = nb.MonteCarlo(template, mvn, "test", seed=seed)
    return mc, nb.pack_spacecraft(ds.state for _, ds in mc.generate_states(0, n))


def oracle_run(oracle, prop: nb.Propagator, frame, almanac, st, cs, ep, end_ns, step_ns=None):
    packed = prop.dynamics.pack(frame, almanac)
    return oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, end_ns, step_ns)


def max_dr_dv(a, b):
    d = a - b
    return float(np.sqrt((d[:3] ** 2).sum(0)).max()), float(np.sqrt((d[3:6] ** 2).sum(0)).max())
