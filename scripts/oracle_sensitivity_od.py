#!/usr/bin/env python
"""C5 (filter) companion of scripts/oracle_sensitivity.py: the CPU oracle filter against ITSELF with one ulp on every adaptive
error norm, over the full BASELINE arc.  The reference integrates the STM to first order in the step (Phi_k+1 = Phi_k (I + sum h b_i
A_i), dynamics/spacecraft.rs:203-214), so the covariance propagation, hence the gain, depends on the step sequence; while the
filter converges (residuals of kilometres) that moves the estimate by far more than the state propagation itself differs.

    python scripts/oracle_sensitivity_od.py --filters 4 --out profiles/r02_oracle_sensitivity_c5.json
"""
import argparse
import json
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _run(job):
    import bench
    from oracle import pyoracle

    i, scale = job
    pyoracle.set_error_scale(scale)
    return bench._c5_ref_full(i)


def main():
    import bench
    import nyx_b200 as nb

    p = argparse.ArgumentParser()
    p.add_argument("--filters", type=int, default=4)
    p.add_argument("--epochs", type=int, default=2880)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    args = argparse.Namespace(workload="c5", degree=21, span_days=a.epochs * 60 / 86400.0, mode="fast")
    bench._C5_SC = bench.c5_scenario(args, nb, a.filters, a.epochs, 0, truth_on_cpu=True)
    jobs = [(i, 1.0) for i in range(a.filters)] + [(i, 1.0 + 2.0 ** -52) for i in range(a.filters)]
    with mp.get_context("fork").Pool(min(len(jobs), mp.cpu_count())) as pool:
        res = pool.map(_run, jobs)
    rows = []
    for i in range(a.filters):
        A, B = res[i], res[a.filters + i]
        d = np.abs(A["est_state"][:, :3] - B["est_state"][:, :3])
        k = np.nanargmax(np.nanmax(d, axis=1))
        rows.append({"filter": i, "steps": [A["n_steps"], B["n_steps"]], "same_flags": bool(np.array_equal(A["msr_flags"], B["msr_flags"])),
                     "worst_over_arc_km": float(np.nanmax(d)), "at_epoch": int(k), "final_km": float(np.abs(A["state"][:3] - B["state"][:3]).max()),
                     "worst_after_epoch_500_km": float(np.nanmax(d[500:])) if a.epochs > 500 else None})
    out = {"workload": "c5", "epochs": a.epochs, "perturbation": "error norm * (1 + 2^-52)", "filters": rows,
           "worst_over_arc_km": max(r["worst_over_arc_km"] for r in rows), "worst_final_km": max(r["final_km"] for r in rows)}
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        Path(a.out).write_text(txt + "\n")


if __name__ == "__main__":
    main()
