"""The N>1 host logic on CPU: two gloo ranks shard an ensemble by member block, broadcast the forcing once,
and the concatenated per-rank results equal the single-process result.  The kernel is replaced by the oracle
here (this is a test of the sharding / broadcast plumbing, which is all the multi-GPU path adds)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import oracle
    from rrmpg_b200 import distributed as rdist, synthetic
    from rrmpg_b200.models import HBVEdu
    r, _, w = rdist.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    T, N = 200, 37
    P = synthetic.random_params(HBVEdu(), N)
    f = synthetic.forcing(T)
    mat, layout = rdist.pack_forcing({"temp": f["temp"], "prec": f["prec"]})
    t = torch.as_tensor(mat if rank == 0 else np.zeros_like(mat))
    rdist.broadcast_forcing(t, src=0)                    # the path's single collective
    got = rdist.unpack_forcing(t.numpy(), layout)
    assert np.array_equal(got["temp"], f["temp"]) and np.array_equal(got["prec"], f["prec"])
    lo, hi = rdist.member_block(N, rank, world)
    q = oracle.hbvedu(got["temp"], got["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P[lo:hi])
    np.save(os.path.join(out_dir, f"q{rank}.npy"), q)
    slow = rdist.max_over_ranks(float(rank + 1))
    assert slow == float(world)
    rdist.barrier()
    rdist.host_barrier()     # the host-side (gloo) barrier bench.py's e2e leg waits on
    assert rdist.host_group() is not None
    dist.destroy_process_group()


def test_two_rank_member_sharding_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import oracle
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import HBVEdu
    f = synthetic.forcing(200)
    P = synthetic.random_params(HBVEdu(), 37)
    full = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    parts = np.concatenate([np.load(tmp_path / f"q{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(parts, full)


def test_pack_unpack_forcing_with_layer_arrays():
    from rrmpg_b200 import distributed as rdist
    rng = np.random.default_rng(0)
    d = {"prec": rng.random((50, 5)), "etp": rng.random(50), "mean_temp": rng.random((50, 5))}
    mat, layout = rdist.pack_forcing(d)
    assert mat.shape == (11, 50)
    back = rdist.unpack_forcing(mat, layout)
    for k in d:
        assert np.array_equal(back[k], d[k])


class _OracleGR4J:
    """Stand-in with the reference's GR4J.simulate signature whose member loop is the oracle (no GPU here)."""

    def simulate(self, prec, etp, s_init=0., r_init=0., return_storage=False, params=None):
        import oracle
        return oracle.gr4j(np.asarray(prec), np.asarray(etp), s_init, r_init, params)


def _worker_sharded(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rrmpg_b200 import distributed as rdist, synthetic
    from rrmpg_b200.models import GR4J
    rdist.init_process_group(backend="gloo")
    T, N = 150, 41
    P = synthetic.random_params(GR4J(), N)                       # same seeded global draw on every rank
    f = synthetic.forcing(T)
    series = {"prec": f["prec"], "etp": f["etp"]} if rank == 0 else None   # only rank 0 holds the forcing
    lo, hi, q = rdist.simulate_sharded(_OracleGR4J(), P, series, s_init=0.6, r_init=0.7)
    assert (lo, hi) == rdist.member_block(N, rank, world) and q.shape == (T, hi - lo)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), q)
    score = rdist.gather_members(q.mean(axis=0), N)              # a per-member vector, global order on every rank
    np.save(os.path.join(out_dir, f"g{rank}.npy"), score)
    rdist.barrier()
    dist.destroy_process_group()


def test_simulate_sharded_and_gather_members_gloo(tmp_path):
    world = 3
    mp.spawn(_worker_sharded, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import oracle
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import GR4J
    f = synthetic.forcing(150)
    P = synthetic.random_params(GR4J(), 41)
    full = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P)
    parts = np.concatenate([np.load(tmp_path / f"s{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(parts, full)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"g{r}.npy"), full.mean(axis=0))
