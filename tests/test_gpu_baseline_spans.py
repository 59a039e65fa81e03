"""FAST-mode (the mode bench.py measures) parity against the CPU oracle AT THE BASELINE SPANS — BASELINE.json configs[1..4]:
C2 3 days, C3 30 days, C4 7 days, C5 the full 2 880-epoch arc.  Every call goes through the C ABI.

Bound: the north-star's "sub-mm position over the benchmark span", max |dr| < 1e-6 km, asserted on a fixed sample of the same
ensembles bench.py builds (same seeds).  How to read it: FAST evaluates the same algorithm with fused multiply-adds and a
regrouped harmonic sum; the adaptive controller turns any rounding difference in the error norm into a slightly different step
sequence, and the reference algorithm's OWN sensitivity to that is (profiles/r02_oracle_sensitivity_*.json, CPU oracle against
itself over the same trajectories):
    C2 3 d   one ulp in the error norm: 30 of 4096 trajectories move, by up to 7.3e-7 km;
             the oracle built with -ffp-contract=fast (FMA) vs itself: median 1.9e-7, p99 6.9e-7, max 9.4e-7 km
    C3 30 d  FMA build: max 3.3e-8 km        C4 7 d  one ulp: max 1.6e-7 km, FMA build: max 1.6e-7 km
so for C2 the 1e-6 km bound sits at the reference's own step-sequence sensitivity; the distribution is asserted too (median,
p99), and printed.  Callers that need the reference's bits use STRICT (bit-identical, tests/test_gpu_fullsize.py)."""
import argparse

import numpy as np
import pytest

import nyx_b200 as nb

from .util import S

pytestmark = pytest.mark.gpu
DAY = 86400 * S
BOUND_KM = 1e-6


def _workload(name, n, span_days):
    import bench

    args = argparse.Namespace(workload=name, degree=21, span_days=span_days)
    return bench.build_workload(args, n, nb)


def _fast_vs_oracle(oracle, name, n, span_days, kernel=None):
    frame, dyn, alm, st, cs, ep = _workload(name, n, span_days)
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    end = int(span_days * DAY)
    eng = prop.engine(frame, alm)
    if kernel is not None:
        eng.set_kernel(kernel)
    out, oep, det, status = eng.propagate_batch(st, cs, ep, end)
    if kernel is not None:
        assert eng.last_kernel() == kernel
    ref, rep, rdet, rstatus = oracle.propagate_batch(dyn.pack(frame, alm).c, prop.opts.to_c(prop.method), st, cs, ep, end)
    assert (status == 0).all() and (rstatus == 0).all()
    assert np.array_equal(oep, rep)
    dr = np.sqrt(((out[:3] - ref[:3]) ** 2).sum(0))
    dv = np.sqrt(((out[3:6] - ref[3:6]) ** 2).sum(0))
    q = np.percentile(dr, [50, 90, 99])
    print(f"\n[{name} {span_days:g} d, {n} trajectories, {int(det['n_steps'].sum())} steps] |dr| km: median {q[0]:.3g}  p90 {q[1]:.3g}  "
          f"p99 {q[2]:.3g}  max {dr.max():.3g};  max |dv| {dv.max():.3g} km/s;  step counts differ on {(det['n_steps'] != rdet['n_steps']).sum()}")
    assert np.abs(det["n_steps"] - rdet["n_steps"]).max() <= 2
    return dr, dv


def test_c2_fast_3_days_vs_oracle(oracle):
    """BASELINE configs[1]: the first 512 runs of the 10 000-trajectory LEO ensemble, JGM-3 21x21, RK89 defaults, 3 days."""
    dr, dv = _fast_vs_oracle(oracle, "c2", 512, 3.0)
    assert dr.max() < BOUND_KM, dr.max()
    assert np.median(dr) < 3e-7 and np.percentile(dr, 99) < 8e-7
    assert dv.max() < 1.2e-9   # |dv| ~ n |dr| (mean motion 1.16e-3 rad/s)


def test_c3_fast_30_days_vs_oracle(oracle):
    """BASELINE configs[2]: JWST-like, Sun + Moon point masses + SRP with Earth / Moon shadows, 30 days (per-thread kernel)."""
    dr, dv = _fast_vs_oracle(oracle, "c3", 1024, 30.0)
    assert dr.max() < 2e-7, dr.max()   # FMA-built oracle vs itself: 3.3e-8 km
    assert dv.max() < 1e-12


@pytest.mark.parametrize("kernel", [nb.KERNEL_COOP, nb.KERNEL_TRANSPOSED], ids=["cooperative-32-lanes", "transposed-16-positions"])
def test_c4_fast_7_days_vs_oracle(oracle, kernel):
    """BASELINE configs[3]: low lunar orbit, GRAIL 70x70 + Earth / Sun point masses, 7 days — with the lane-cooperative kernel (what
    ensembles below 1 024 trajectories get) and with the transposed kernel (16 walker positions, one set context per CTA: what the
    dispatch picks from 1 024 trajectories)."""
    dr, dv = _fast_vs_oracle(oracle, "c4", 48, 7.0, kernel)
    assert dr.max() < BOUND_KM, dr.max()   # one ulp in the oracle's own error norm: 1.6e-7 km
    assert dv.max() < 1e-9


def _c5_fast_vs_oracle(n, n_msr, fixed_step_s):
    import multiprocessing as mp

    import bench

    args = argparse.Namespace(workload="c5", degree=21, span_days=n_msr * 60 / 86400.0, mode="fast")
    sc = bench.c5_scenario(args, nb, n, n_msr, 0, truth_on_cpu=True, fixed_step_s=fixed_step_s)
    sol = sc["odp"].process_arcs(sc["ests"], sc["arc"], record_estimates=True)
    assert (sol.status == 0).all()
    bench._C5_SC = sc
    with mp.get_context("fork").Pool(n) as pool:
        refs = pool.map(bench._c5_ref_full, list(range(n)))
    worst_r = worst_v = worst_final = 0.0
    for i, ref in enumerate(refs):
        assert np.array_equal(sol.msr_flags[:, i], ref["msr_flags"]), i
        d = np.abs(sol.est_state[:, :6, i] - ref["est_state"][:, :6])
        worst_r = max(worst_r, float(np.nanmax(d[:, :3])))
        worst_v = max(worst_v, float(np.nanmax(d[:, 3:6])))
        fr = float(np.abs(sol.final_state_soa[:3, i] - ref["state"][:3]).max())
        worst_final = max(worst_final, fr)
        print(f"\n[c5 filter {i}] final |dr| {fr:.3g} km, worst over the arc {float(np.nanmax(d[:, :3])):.3g} km, steps {sol.details['n_steps'][i]} / {ref['n_steps']}")
    print(f"\n[c5 {'fixed ' + str(fixed_step_s) + ' s' if fixed_step_s else 'adaptive'}] worst estimate difference over {n} filters x {n_msr} epochs: "
          f"{worst_r:.3g} km, {worst_v:.3g} km/s; final {worst_final:.3g} km")
    return worst_r, worst_v, worst_final


def test_c5_fast_full_arc_fixed_step_vs_oracle_filter(oracle):
    """BASELINE configs[4] geometry with FIXED 20 s DP78 steps between the measurements (the reference's own OD validation runs
    fixed steps, tests/orbit_determination/two_body.rs:72-203): identical step sequences, so this pins the arithmetic of the FAST
    filter kernel — dual-number harmonic gradient, STM, time / measurement updates — over the full 2 880-epoch arc."""
    worst_r, worst_v, worst_final = _c5_fast_vs_oracle(8, 2880, 20.0)
    assert worst_r < BOUND_KM and worst_v < 1e-9, (worst_r, worst_v)


def test_c5_fast_full_arc_adaptive_vs_oracle_filter(oracle):
    """BASELINE configs[4] as benchmarked (adaptive DP78).  The reference integrates the STM to first order in the step
    (spacecraft.rs:203-214), so the gain depends on the step sequence; the CPU oracle filter against ITSELF with one ulp on the
    error norm differs by up to 1.5e-3 km over this arc and 5.9e-5 km at its end (profiles/r02_oracle_sensitivity_c5.json).  FAST is
    held to twice that; identical accept / reject decisions for every measurement."""
    worst_r, worst_v, worst_final = _c5_fast_vs_oracle(8, 2880, None)
    assert worst_r < 3e-3 and worst_final < 1.2e-4, (worst_r, worst_final)
