"""CPU: the host helpers of the slab decomposition exported by the C-ABI library (sg_slab_partition, sg_slab_limits,
sg_slab_merge_dest, sg_slab_merge_static_dest) -- no device work, so they run without a GPU.  The merge is checked
against a plain lexicographic sort of the concatenated lists, which is what the reference's std::set order is
(ball2d/Ball2DSim.cpp:580)."""
import numpy as np
import pytest

from scisim_b200 import slab


@pytest.mark.parametrize("n,world", [(1000, 1), (1000, 2), (1001, 3), (5000, 8), (7, 8), (0, 3)])
def test_partition_is_equal_count_and_ordered_in_x(n, world):
    rng = np.random.default_rng(n + world)
    q = rng.uniform(-50.0, 50.0, size=2 * n)
    rank_of, cuts, gids = slab.partition_quantiles(q, world)
    counts = [g.shape[0] for g in gids]
    assert sum(counts) == n and max(counts) - min(counts) <= 1
    x = q[0::2]
    for k in range(world):
        assert np.all(np.diff(gids[k].astype(np.int64)) > 0)          # ascending global index inside a slab
        for l in range(k + 1, world):
            if counts[k] and counts[l]:
                assert x[gids[k]].max() <= x[gids[l]].min()          # slabs are ordered in x
    if n >= world:
        assert np.all(np.diff(cuts) >= 0.0)
        for k in range(1, world):
            assert x[gids[k - 1]].max() <= cuts[k] <= x[gids[k]].min()


def test_partition_breaks_ties_by_index_and_is_deterministic():
    q = np.zeros(2 * 12)
    q[0::2] = np.array([1.0] * 12)            # every body at the same x
    rank_of, _, gids = slab.partition_quantiles(q, 3)
    assert [list(g) for g in gids] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]]
    r2, _, _ = slab.partition_quantiles(q, 3)
    assert np.array_equal(rank_of, r2)


def test_limits_keep_slabs_two_apart_separated():
    cuts = np.array([0.0, 10.0, 14.0, 30.0, 31.0])
    lim = [slab.slab_limits(cuts, k) for k in range(4)]
    assert lim[0][0] == -np.inf and lim[3][1] == np.inf
    for k in range(2):
        assert lim[k][1] <= lim[k + 2][0]        # the boxes of slabs k and k + 2 cannot overlap
    assert lim[1][0] == 5.0 and lim[1][1] == 22.0


@pytest.mark.parametrize("n,world,seed", [(200, 2, 1), (500, 3, 2), (3000, 8, 3)])
def test_merge_equals_lexicographic_sort(n, world, seed):
    rng = np.random.default_rng(seed)
    owner = rng.integers(0, world, size=n)
    # a random "pair list": a few partners j > i per body
    pairs = []
    for i in range(n - 1):
        k = int(rng.integers(0, 4))
        js = np.unique(rng.integers(i + 1, n, size=k)) if k else np.zeros(0, dtype=np.int64)
        pairs += [(i, int(j)) for j in js]
    pairs = np.array(pairs, dtype=np.uint32).reshape(-1, 2)
    ref = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    parts = []
    for k in range(world):
        mine = ref[owner[ref[:, 0]] == k]
        nst = int(rng.integers(0, 6))
        st_j = np.sort(rng.integers(0, 3, size=nst)).astype(np.uint32)
        own_ids = np.nonzero(owner == k)[0]
        st_i = np.array([rng.choice(own_ids) for _ in range(nst)], dtype=np.uint32) if own_ids.size else np.zeros(0, np.uint32)
        st_j = st_j[:st_i.shape[0]]
        o = np.lexsort((st_i, st_j))
        st_i, st_j = st_i[o], st_j[o]
        na = mine.shape[0] + st_i.shape[0]
        parts.append({"candidates": mine.copy(), "type": np.concatenate([np.zeros(mine.shape[0], np.uint32), np.full(st_i.shape[0], 2, np.uint32)]),
                      "i": np.concatenate([mine[:, 0], st_i]), "j": np.concatenate([mine[:, 1], st_j]),
                      "n": rng.normal(size=(na, 2)), "p": rng.normal(size=(na, 2)), "depth": rng.normal(size=na)})
    m = slab.merge_active_sets(parts, n_bodies=n)
    assert np.array_equal(m["candidates"], ref)
    nbb = ref.shape[0]
    assert np.array_equal(np.stack([m["i"][:nbb], m["j"][:nbb]], axis=1), ref)
    st = np.stack([m["j"][nbb:], m["i"][nbb:]], axis=1).astype(np.int64)
    assert np.all(m["type"][nbb:] == 2)
    assert np.all((np.diff(st[:, 0]) > 0) | ((np.diff(st[:, 0]) == 0) & (np.diff(st[:, 1]) >= 0)))
    # payload columns travel with their rows
    allrows = {(int(p["type"][e]), int(p["i"][e]), int(p["j"][e])): (p["n"][e], p["depth"][e]) for p in parts for e in range(p["type"].shape[0])}
    for e in range(0, m["type"].shape[0], 7):
        nn, dd = allrows[(int(m["type"][e]), int(m["i"][e]), int(m["j"][e]))]
        assert np.array_equal(m["n"][e], nn) and m["depth"][e] == dd


def test_merge_rejects_unsorted_part():
    bad = {"candidates": np.array([[5, 6], [2, 3]], dtype=np.uint32), "type": np.zeros(0, np.uint32), "i": np.zeros(0, np.uint32), "j": np.zeros(0, np.uint32),
           "n": np.zeros((0, 2)), "p": np.zeros((0, 2)), "depth": np.zeros(0)}
    with pytest.raises(RuntimeError):
        slab.merge_active_sets([bad], n_bodies=10)
