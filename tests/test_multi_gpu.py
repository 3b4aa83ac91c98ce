"""GPU (one device is enough: the same device may be listed several times): sg_multi -- ONE process driving N slabs, the
way a SCISim process would -- against the single-scene oracle, bit for bit, on randomly numbered scenes: the x-quantile
partition, the peer-memory halo exchange (mailboxes connected by address), re-partitioning on demand, and the merge of
the per-slab lists into the reference's order all sit behind the calls of a single sim."""
import numpy as np
import pytest

from tests import slab_helpers as sh

pytestmark = pytest.mark.gpu


def _devices(world):
    import torch
    ng = torch.cuda.device_count()
    return [k % ng for k in range(world)]


def _check(a, ref):
    assert np.array_equal(a.candidates, ref["candidates"])
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(getattr(a, k), ref[k]), k
    assert np.array_equal(a.depth, ref["depth"], equal_nan=True)


@pytest.mark.parametrize("world,n,seed,kind", [(1, 3000, 21, 0), (2, 5000, 22, 0), (3, 9000, 23, 1), (4, 20000, 24, 1)])
def test_multi_flow_and_active_set_equal_oracle(oracle, world, n, seed, kind):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = sh.random_numbered_scene(n, seed, box=max(4.0, np.sqrt(n) * 1.2))
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    sim = sb.MultiGpuBall2DSim(st, _devices(world))
    umap = sb.SymplecticEulerMap() if kind == 0 else sb.VerletMap()
    o = ob.Ball2DOracle(s)
    q1, v1 = umap.flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = o.flow(kind, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    a = sim.computeActiveSet(s["q"], q1, resident=True)
    ref = o.active_set(s["q"], rq1, "grid")
    assert ref["candidates"].shape[0] > 0 and (ref["type"] != 0).any()
    _check(a, ref)
    info = sim.partition_info()
    assert info["n_partitions"] == 1 and int(info["n_owned"].sum()) == n
    if world > 1:
        assert info["ghosts"].sum() > 0, "the scene must actually exchange ghosts"
    # the same through the non-resident call: q0, q1 uploaded, nothing integrated
    a2 = sim.computeActiveSet(s["q"], q1)
    _check(a2, ref)
    sim.close()


@pytest.mark.parametrize("world,n,seed", [(2, 6000, 31), (3, 12000, 32)])
def test_multi_resident_steps_and_repartition(oracle, world, n, seed):
    """upload / step / fetch over several steps, the state advanced on the host; step 2 is uploaded mirrored in x, so every
    slab finds its bodies outside its limits: the step must come back correct, from a second partition."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = sh.random_numbered_scene(n, seed, box=max(4.0, np.sqrt(n) * 1.2))
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    sim = sb.MultiGpuBall2DSim(st, _devices(world))
    o = ob.Ball2DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    for step in range(4):
        if step == 2:
            x = q[0::2]
            q = q.copy()
            q[0::2] = x.min() + x.max() - x
        sim.upload(q, v)
        pc, pa = sim.step(sb.SymplecticEulerMap(), s["dt"])
        q1, v1, a = sim.fetch()
        rq1, rv1 = o.flow(0, q, v, s["dt"])
        ref = o.active_set(q, rq1, "grid")
        assert (pc, pa) == (ref["candidates"].shape[0], ref["type"].shape[0])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        _check(a, ref)
        assert sim.partition_info()["n_partitions"] == (1 if step < 2 else 2)
        q, v = q1, v1
    # the halo pack scans all bodies only in the first step after a (re-)partition; afterwards the candidate bands hold
    assert all(st[2] <= 1 for st in sim.slab_stats()), sim.slab_stats()
    sim.close()


def test_multi_too_thin_slabs_are_refused(oracle):
    """Slabs thinner than a body cannot be made safe by any re-partition: the call fails loudly instead of dropping contacts."""
    import scisim_b200 as sb
    s = sh.random_numbered_scene(400, 41, box=1.0)   # 4 slabs of width 0.5 for balls of radius up to 0.4
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    sim = sb.MultiGpuBall2DSim(st, _devices(4))
    sim.upload(s["q"], s["v"])
    with pytest.raises(sb.SciSimB200Error):
        sim.step(sb.SymplecticEulerMap(), s["dt"])
    sim.close()
