#!/usr/bin/env python
"""The oracle's `GravityField::eom` restatement (oracle/nyx_oracle.c, following gravity_field.rs:148-268) against an INDEPENDENT
arbiter: the textbook gradient of the normalised spherical-harmonic potential in spherical coordinates with closed-form associated
Legendre functions at 40 digits (tests/arbiters.py) — VERDICT r01 item 2.  Writes profiles/r02_arbiter_harmonics.json.

    python scripts/arbiter_harmonics.py --points21 200 --points70 16"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def oracle_field_accel(packed_c, rb):
    from nyx_b200 import abi
    from oracle import pyoracle

    L = pyoracle.lib()
    L.nyx_oracle_grav_new.restype = C.c_void_p
    L.nyx_oracle_grav_new.argtypes = [C.POINTER(abi.GravityFieldC)]
    L.nyx_oracle_grav_accel.restype = None
    L.nyx_oracle_grav_accel.argtypes = [C.c_void_p, C.c_int64, abi.c_double_p, abi.c_double_p, abi.c_double_p]
    L.nyx_oracle_grav_free.argtypes = [C.c_void_p]
    gf = packed_c.gravity[0]
    h = L.nyx_oracle_grav_new(C.byref(gf))
    scratch = np.zeros((gf.degree + 3) ** 2)
    out = np.zeros(3)
    r = np.ascontiguousarray(rb, dtype=np.float64)
    L.nyx_oracle_grav_accel(h, 0, abi.as_double_p(r), abi.as_double_p(scratch), abi.as_double_p(out))
    L.nyx_oracle_grav_free(h)
    return out


def _case(job):
    import nyx_b200 as nb
    from tests.arbiters import mp_harmonic_accel

    fixture, deg, seed = job
    moon = fixture.startswith("luna")
    gd = nb.GravityFieldData.from_fixture(fixture, deg, deg, nb.IAU_MOON_FRAME if moon else nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    packed = dyn.pack(nb.MOON_J2000 if moon else nb.EARTH_J2000, None)
    gf = packed.c.gravity[0]
    gf.rot.kind = 0
    rng = np.random.default_rng(seed)
    d = rng.normal(size=3)
    rb = d / np.linalg.norm(d) * gf.r_eq_km * rng.uniform(1.02, 1.5)
    got = oracle_field_accel(packed.c, rb)
    want = np.array(mp_harmonic_accel(gd.c_nm, gd.s_nm, deg, deg, gf.mu_km3_s2, gf.r_eq_km, rb))
    return float(np.abs(got - want).max() / np.abs(want).max()), float(np.linalg.norm(rb) / gf.r_eq_km)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--points21", type=int, default=200)
    p.add_argument("--points70", type=int, default=16)
    p.add_argument("--out", default="profiles/r02_arbiter_harmonics.json")
    a = p.parse_args()
    jobs = [("jgm3_70x70", 21, s) for s in range(a.points21)] + [("jgm3_70x70", 70, 1000 + s) for s in range(a.points70)] + \
           [("luna_jggrx_80x80", 70, 2000 + s) for s in range(a.points70)]
    with mp.get_context("fork").Pool(mp.cpu_count()) as pool:
        res = pool.map(_case, jobs, chunksize=1)
    out = {}
    for name, sl in (("jgm3_21x21", slice(0, a.points21)), ("jgm3_70x70", slice(a.points21, a.points21 + a.points70)),
                     ("grail_70x70", slice(a.points21 + a.points70, None))):
        e = np.array([r[0] for r in res[sl]])
        if e.size == 0:
            continue
        out[name] = {"points": int(e.size), "max_rel_err": float(e.max()), "median_rel_err": float(np.median(e)),
                     "radius_range_r_eq": [float(min(r[1] for r in res[sl])), float(max(r[1] for r in res[sl]))]}
    out["arbiter"] = "textbook spherical-coordinate gradient, closed-form Legendre functions (mpmath.legenp), 40 digits: tests/arbiters.py"
    out["note"] = "relative to the largest component of the non-central acceleration; the oracle evaluates gravity_field.rs:148-268 in f64"
    txt = json.dumps(out, indent=1)
    print(txt)
    Path(a.out).write_text(txt + "\n")


if __name__ == "__main__":
    main()
