"""N>1 path on CPU: world_size-2 gloo processes shard an ensemble by contiguous run index, integrate their
shards (the CPU oracle stands in for the GPU engine here), and exchange final states with ONE all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nyx_b200.dist import all_gather_final_states, shard_bounds, shard_soa


def test_shard_bounds_cover_and_are_contiguous():
    for n in (0, 1, 7, 10_000, 10_001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nyx_b200 as nb
    from oracle import pyoracle
    from tests.util import S, leo_ensemble

    mc, (st, cs, ep) = leo_ensemble(n, seed=21)  # same host-side draw stream on every rank
    prop = nb.Propagator.default(nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()))
    packed = prop.dynamics.pack(nb.EARTH_J2000, None)
    lo, hi = shard_bounds(n, world, rank)
    out, _, det, status = pyoracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), shard_soa(st, world, rank),
                                                   shard_soa(cs, world, rank), shard_soa(ep, world, rank), 1800 * S, n_threads=1)
    assert out.shape == (9, hi - lo) and (status == 0).all()
    gathered = all_gather_final_states(torch.from_numpy(out), n)
    steps = torch.tensor([int(det["n_steps"].sum())])
    dist.all_reduce(steps)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), gathered.numpy())
    np.save(os.path.join(out_dir, f"steps{rank}.npy"), steps.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [33, 64])
def test_two_rank_gloo_sharded_run_matches_single_process(tmp_path, oracle, n):
    import nyx_b200 as nb
    from tests.util import S, leo_ensemble

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    mc, (st, cs, ep) = leo_ensemble(n, seed=21)
    prop = nb.Propagator.default(nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()))
    packed = prop.dynamics.pack(nb.EARTH_J2000, None)
    ref, _, det, _ = oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, 1800 * S)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert np.array_equal(got, ref)  # every rank holds the full, index-ordered result (no sort needed)
        assert int(np.load(tmp_path / f"steps{r}.npy")[0]) == int(det["n_steps"].sum())


def _od_worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nyx_b200 as nb
    from nyx_b200 import dist as nbdist
    from oracle import pyoracle, pyoracle_od
    from tests.od_util import leo_od_scenario, run_oracle_filter

    sc = leo_od_scenario(pyoracle, n=n, n_msr=8, seed=4)   # same scenario on every rank

    class OracleOD:   # the numpy oracle filter stands in for the GPU engine behind the same process_arcs contract
        def process_arcs(self, ests, arc):
            sub = dict(sc, ests=ests, arc=arc)
            res = [run_oracle_filter(pyoracle_od, sub, i) for i in range(len(ests))]
            st = np.stack([r["state"] for r in res], axis=1)
            cov = np.stack([r["covar"] for r in res], axis=0)
            return nb.ODSolution(st, np.array([r["epoch"] for r in res]), cov, st * 0, None, None, None, None, None, None, None, None)

    sol, states, covars = nbdist.sharded_process_arcs(OracleOD(), sc["ests"], sc["arc"])
    lo, hi = shard_bounds(n, world, rank)
    assert sol.final_state_soa.shape == (9, hi - lo) and states.shape == (9, n) and covars.shape == (n, 9, 9)
    np.save(os.path.join(out_dir, f"od_states{rank}.npy"), states)
    np.save(os.path.join(out_dir, f"od_covars{rank}.npy"), covars)
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_filters_match_single_process(tmp_path, oracle):
    """An ensemble of Kalman filters shards like an ensemble of propagations: contiguous filter indices per rank, one
    all-gather of (state, covariance); every rank ends with the full index-ordered result."""
    from oracle import pyoracle_od
    from tests.od_util import leo_od_scenario, run_oracle_filter

    n, world = 5, 2
    mp.spawn(_od_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    sc = leo_od_scenario(oracle, n=n, n_msr=8, seed=4)
    ref = [run_oracle_filter(pyoracle_od, sc, i) for i in range(n)]
    for r in range(world):
        st = np.load(tmp_path / f"od_states{r}.npy")
        cv = np.load(tmp_path / f"od_covars{r}.npy")
        for i in range(n):
            assert np.array_equal(st[:, i], ref[i]["state"]) and np.array_equal(cv[i], ref[i]["covar"])
