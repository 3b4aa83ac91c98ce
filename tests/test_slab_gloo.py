"""CPU, world_size 2 and 3 under gloo: the host logic of the multi-GPU slab path (interval exchange, halo selection,
pair ownership, merge of the per-rank lists) with the oracle standing in for the device kernels. The merged result must be
exactly the single-process reference result."""
import os
import pickle
import socket
import tempfile

import numpy as np
import pytest

from tests import slab_helpers as sh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, outdir, transport="nccl"):
    import torch.distributed as dist
    from scisim_b200.slab import Ball2DSlabs, partition_slab_major
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = sh.slab_major_scene(n, seed)
    firsts, counts = partition_slab_major(n, world)
    backend = sh.OracleSlabBackend(sh.slab_of(scene, firsts[rank], counts[rank]), firsts[rank], ghost_cap=n)
    drv = Ball2DSlabs(backend, rank, world, dist, check_non_neighbours=True, transport=transport)
    assert drv.transport == "nccl"   # "p2p" must fall back on every rank when a backend cannot export a mailbox
    pc, pa = drv.step(0, scene["dt"])
    q1, v1, res = backend.fetch()
    res["q1"], res["v1"], res["halo"] = q1, v1, drv.last_halo
    pickle.dump(res, open(os.path.join(outdir, "rank%d.pkl" % rank), "wb"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed", [(2, 600, 1), (3, 900, 2)])
def test_slab_merge_equals_single_process(world, n, seed):
    import torch.multiprocessing as mp
    from scisim_b200.slab import merge_active_sets
    from tests import oracle_binding as ob
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), n, seed, d), nprocs=world, join=True)
        parts = [pickle.load(open(os.path.join(d, "rank%d.pkl" % r), "rb")) for r in range(world)]
    scene = sh.slab_major_scene(n, seed)
    o = ob.Ball2DOracle(scene)
    q1, v1 = o.flow(0, scene["q"], scene["v"], scene["dt"])
    ref = o.active_set(scene["q"], q1, "allpairs")
    assert np.array_equal(np.concatenate([p["q1"] for p in parts]), q1)
    merged = merge_active_sets(parts, (scene["drum_x"].shape[0], scene["plane_x"].shape[0]))
    assert sum(p["halo"][0] + p["halo"][1] for p in parts) > 0, "the test scene must actually exchange ghosts"
    assert np.array_equal(merged["candidates"], ref["candidates"])
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(merged[k], ref[k]), k
    assert np.array_equal(merged["depth"], ref["depth"], equal_nan=True)


def test_partition_slab_major():
    from scisim_b200.slab import partition_slab_major
    firsts, counts = partition_slab_major(10, 4)
    assert counts == [3, 3, 2, 2] and firsts == [0, 3, 6, 8]


def test_p2p_request_falls_back_to_collectives_when_no_mailbox_can_be_mapped():
    """The driver is asked for the peer-memory transport with a backend that has no device memory: every rank must agree
    to fall back to the collective transport, and the result must still be the reference's."""
    import torch.multiprocessing as mp
    from scisim_b200.slab import merge_active_sets
    from tests import oracle_binding as ob
    world, n, seed = 2, 500, 3
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), n, seed, d, "p2p"), nprocs=world, join=True)
        parts = [pickle.load(open(os.path.join(d, "rank%d.pkl" % r), "rb")) for r in range(world)]
    scene = sh.slab_major_scene(n, seed)
    o = ob.Ball2DOracle(scene)
    q1, v1 = o.flow(0, scene["q"], scene["v"], scene["dt"])
    ref = o.active_set(scene["q"], q1, "allpairs")
    merged = merge_active_sets(parts, (scene["drum_x"].shape[0], scene["plane_x"].shape[0]))
    assert np.array_equal(merged["candidates"], ref["candidates"])
    for k in ("type", "i", "j"):
        assert np.array_equal(merged[k], ref[k]), k


# ---- arbitrary numbering: x-quantile partition, re-partition on demand, general merge ---------------------------------
def _sim_worker(rank, world, port, n, seed, outdir, mode):
    import torch.distributed as dist
    from scisim_b200.slab import Ball2DSlabSim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = sh.random_numbered_scene(n, seed, box=max(4.0, np.sqrt(n) * 1.2))
    factory = lambda s, gids, lim, cap: sh.OracleSlabBackend(s, 0, cap, gids=gids, x_limits=lim)
    sim = Ball2DSlabSim(scene, rank, world, dist, factory, transport="nccl", ghost_cap=n)
    q, v = scene["q"].copy(), scene["v"].copy()
    if mode == "stale":
        # partition a state, then upload (without re-partitioning) one in which the bodies have swapped sides: every
        # rank's bodies are now far outside its slab -> the step must ask for, and get, a fresh partition
        sim.upload(q, v)
        x = q[0::2]
        q = q.copy()
        q[0::2] = x.min() + x.max() - x
        sim.upload(q, v)
    else:
        sim.upload(q, v)
    pc, pa = sim.step(0, scene["dt"])
    merged = sim.gather_merged(0)
    if rank == 0:
        merged["n_partitions"] = sim.n_partitions
        merged["q0"] = q
        pickle.dump(merged, open(os.path.join(outdir, "merged.pkl"), "wb"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed,mode", [(2, 700, 11, "fresh"), (3, 1200, 12, "fresh"), (3, 900, 13, "stale")])
def test_quantile_slabs_of_a_randomly_numbered_scene(world, n, seed, mode):
    import torch.multiprocessing as mp
    from tests import oracle_binding as ob
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_sim_worker, args=(world, _free_port(), n, seed, d, mode), nprocs=world, join=True)
        m = pickle.load(open(os.path.join(d, "merged.pkl"), "rb"))
    scene = sh.random_numbered_scene(n, seed, box=max(4.0, np.sqrt(n) * 1.2))
    o = ob.Ball2DOracle(scene)
    q1, v1 = o.flow(0, m["q0"], scene["v"], scene["dt"])
    ref = o.active_set(m["q0"], q1, "allpairs")
    assert m["n_partitions"] == (2 if mode == "stale" else 1)
    assert np.array_equal(m["q1"], q1) and np.array_equal(m["v1"], v1)
    assert ref["candidates"].shape[0] > 0
    assert np.array_equal(m["candidates"], ref["candidates"])
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(m[k], ref[k]), k
    assert np.array_equal(m["depth"], ref["depth"], equal_nan=True)


def _agree_worker(rank, world, port, outdir, force_fallback):
    import torch.distributed as dist
    from scisim_b200 import slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ag = slab.HostAgreement(rank, world, dist)
    if force_fallback:
        ag.arr = None            # what a multi-node job (or a box without /dev/shm) ends up with: the 4-byte all_reduce
    used_shm = ag.arr is not None
    got = []
    # 200 steps with rank-dependent flags and rank-dependent delays: a fast rank runs ahead of the slow ones' reads
    rng = np.random.default_rng(100 + rank)
    for step in range(200):
        flag = (step % 7 == rank % 7) or (step % 13 == 0 and rank == world - 1)
        if rng.random() < 0.2:
            import time
            time.sleep(0.0005 * rng.random())
        got.append(ag.any(flag))
    pickle.dump({"got": got, "shm": used_shm}, open(os.path.join(outdir, "agree%d.pkl" % rank), "wb"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,force_fallback", [(2, False), (3, False), (3, True)])
def test_host_agreement_is_a_logical_or_over_the_ranks(world, force_fallback):
    """The per-step "does anyone need a re-partition?" exchange of the one-process-per-GPU driver: every rank must see the OR of all ranks'
    flags of THAT step, through the shared-memory path and through the all_reduce fallback, also when ranks run ahead of each other."""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_agree_worker, args=(world, _free_port(), d, force_fallback), nprocs=world, join=True)
        parts = [pickle.load(open(os.path.join(d, "agree%d.pkl" % r), "rb")) for r in range(world)]
    want = [any((step % 7 == r % 7) or (step % 13 == 0 and r == world - 1) for r in range(world)) for step in range(200)]
    for p in parts:
        assert p["got"] == want
    if not force_fallback:
        assert all(p["shm"] for p in parts) or not os.path.isdir("/dev/shm")
    assert any(want) and not all(want)
