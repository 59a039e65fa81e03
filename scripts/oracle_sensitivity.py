#!/usr/bin/env python
"""How sensitive is the REFERENCE algorithm itself to one ulp in its adaptive error norm?  

Runs the CPU oracle (parity build) on the first N trajectories of a bench workload over the full BASELINE span three times:
  A  untouched                                   (the checker)
  B  error norm of every attempt * (1 + 2^-52)   (ONE ulp: what any differently-rounded but equally valid evaluation does)
  C  the -O3 / FMA-contracted build of the same sources (a CPU stand-in for "FAST": same algorithm, fused multiply-adds)
and writes the distribution of |r_B - r_A| and |r_C - r_A| at the end epoch.  The FAST kernels cannot agree with the oracle
better than the oracle agrees with its own 1-ulp perturbation: this is the bound the tolerance-parity tests are read against.

    python scripts/oracle_sensitivity.py --workload c2 --n 512 --out profiles/r02_oracle_sensitivity_c2.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def summarize(d):
    q = np.percentile(d, [50, 90, 99, 100])
    edges = [0, 1e-8, 3e-8, 1e-7, 3e-7, 1e-6, 3e-6, 1e-5, np.inf]
    hist, _ = np.histogram(d, bins=edges)
    return {"median_km": float(q[0]), "p90_km": float(q[1]), "p99_km": float(q[2]), "max_km": float(q[3]),
            "hist_edges_km": [str(e) for e in edges], "hist": [int(h) for h in hist], "frac_below_1e-6": float((d < 1e-6).mean())}


def main():
    import bench
    import nyx_b200 as nb
    from oracle import pyoracle

    p = argparse.ArgumentParser()
    p.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"])
    p.add_argument("--n", type=int, default=512)
    p.add_argument("--span-days", type=float, default=None)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    span = a.span_days or {"c2": 3.0, "c3": 30.0, "c4": 7.0}[a.workload]
    args = argparse.Namespace(workload=a.workload, degree=21, span_days=span)
    frame, dyn, alm, st, cs, ep = bench.build_workload(args, a.n, nb)
    prop = nb.Propagator.default(dyn)
    packed = dyn.pack(frame, alm)
    opts = prop.opts.to_c(prop.method)
    end = int(span * 86400 * 10**9)

    def run(**kw):
        t0 = time.perf_counter()
        out, _, det, status = pyoracle.propagate_batch(packed.c, opts, st, cs, ep, end, **kw)
        assert (status == 0).all()
        return out, det, time.perf_counter() - t0

    A, detA, tA = run()
    pyoracle.set_error_scale(1.0 + 2.0 ** -52)
    B, detB, _ = run()
    pyoracle.set_error_scale(1.0)
    C, detC, tC = run(speed_build=True)
    dr = lambda X: np.sqrt(((X[:3] - A[:3]) ** 2).sum(0))
    res = {"workload": a.workload, "n": a.n, "span_days": span, "accepted_steps": int(detA["n_steps"].sum()),
           "one_ulp_error_norm": summarize(dr(B)), "fma_build": summarize(dr(C)),
           "step_count_differs_one_ulp": int((detA["n_steps"] != detB["n_steps"]).sum()),
           "step_count_differs_fma": int((detA["n_steps"] != detC["n_steps"]).sum()),
           "oracle_steps_per_s": {"parity_build": detA["n_steps"].sum() / tA, "speed_build": detC["n_steps"].sum() / tC,
                                  "threads": pyoracle.num_threads()}}
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        Path(a.out).write_text(txt + "\n")


if __name__ == "__main__":
    main()
