"""Shared helpers for the slab (multi-GPU) tests: scenes numbered slab-major and an oracle-backed backend that lets the
host logic of scisim_b200/slab.py run on CPU under gloo."""
import numpy as np

from scisim_b200 import scenes
from scisim_b200.slab import REC_BYTES


def slab_major_scene(n, seed, vmax=6.0, nplanes=2, ndrums=1):
    """A messy ball2d scene whose bodies are numbered in ascending x, so contiguous index blocks are spatial slabs."""
    s = scenes.ball2d_random(n, seed, nplanes=nplanes, ndrums=ndrums, vmax=vmax, box=max(2.0, np.sqrt(n) * 0.6))
    order = np.argsort(s["q"].reshape(-1, 2)[:, 0], kind="stable")
    s["q"] = s["q"].reshape(-1, 2)[order].ravel().copy()
    s["v"] = s["v"].reshape(-1, 2)[order].ravel().copy()
    s["r"] = s["r"][order].copy()
    s["m"] = s["m"][order].copy()
    return s


def slab_of(scene, first, count):
    s = dict(scene)
    s["q"] = scene["q"][2 * first:2 * (first + count)].copy()
    s["v"] = scene["v"][2 * first:2 * (first + count)].copy()
    s["r"] = scene["r"][first:first + count].copy()
    s["m"] = scene["m"][first:first + count].copy()
    return s


REC_DT = np.dtype([("q0", "<f8", 2), ("q1", "<f8", 2), ("r", "<f8"), ("gid", "<u4"), ("pad", "<u4")])
assert REC_DT.itemsize == REC_BYTES


def random_numbered_scene(n, seed, vmax=6.0, nplanes=2, ndrums=1, box=None):
    """A messy ball2d scene whose numbering has nothing to do with position (what a partitioner has to cope with)."""
    return scenes.ball2d_random(n, seed, nplanes=nplanes, ndrums=ndrums, vmax=vmax, box=box or max(2.0, np.sqrt(n) * 0.6))


class OracleSlabBackend:
    """CPU stand-in for GpuSlabBackend (same interface and buffer layout), built on the oracle. TEST INFRASTRUCTURE ONLY."""

    def __init__(self, scene_slab, gid_first, ghost_cap, gids=None, x_limits=None):
        import torch
        from tests import oracle_binding as ob
        self.torch, self.ob = torch, ob
        self.cap = ghost_cap
        self.recv = [torch.zeros((ghost_cap + 1) * REC_BYTES, dtype=torch.uint8) for _ in range(2)]
        self.ghosts = (0, 0)
        self.reinit(scene_slab, gid_first, gids, x_limits)

    def reinit(self, scene_slab, gid_first=0, gids=None, x_limits=None):
        self.s = scene_slab
        self.n_owned = scene_slab["r"].shape[0]
        self.gid = (np.asarray(gids, dtype=np.uint32) if gids is not None else gid_first + np.arange(self.n_owned, dtype=np.uint32))
        assert np.all(np.diff(self.gid.astype(np.int64)) > 0), "owned bodies are stored in ascending global index"
        self.xlim = np.asarray(x_limits, dtype=np.float64) if x_limits is not None else np.array([-np.inf, np.inf])
        self.o = self.ob.Ball2DOracle(scene_slab)
        self.q0 = scene_slab["q"].copy()
        self.v0 = scene_slab["v"].copy()
        self.ghost_recs = {0: np.zeros(0, REC_DT), 1: np.zeros(0, REC_DT)}
        self.violation = False

    def upload(self, q, v):
        self.q0, self.v0 = np.array(q, dtype=np.float64), np.array(v, dtype=np.float64)

    def flow(self, kind, dt):
        self.q1, self.v1 = self.o.flow(kind, self.q0, self.v0, dt)
        a, b = self.q0.reshape(-1, 2), self.q1.reshape(-1, 2)
        self.lo = np.minimum(b[:, 0], a[:, 0]) - self.s["r"]
        self.hi = np.maximum(b[:, 0], a[:, 0]) + self.s["r"]
        self.ghost_recs = {0: np.zeros(0, REC_DT), 1: np.zeros(0, REC_DT)}
        self.violation = bool(np.any(self.lo < self.xlim[0]) or np.any(self.hi > self.xlim[1]))
        if self.n_owned == 0:
            return self.torch.tensor([np.inf, -np.inf], dtype=self.torch.float64)
        return self.torch.tensor([self.lo.min(), self.hi.max()], dtype=self.torch.float64)

    def _select(self, interval):
        ilo, ihi = float(interval[0]), float(interval[1])
        return np.nonzero(~(self.hi < ilo) & ~(ihi < self.lo))[0]

    def pack(self, interval, side):
        idx = self._select(interval)
        assert idx.shape[0] <= self.cap
        rec = np.zeros(self.cap + 1, REC_DT)
        rec["gid"][0] = idx.shape[0]
        k = slice(1, 1 + idx.shape[0])
        rec["q0"][k] = self.q0.reshape(-1, 2)[idx]
        rec["q1"][k] = self.q1.reshape(-1, 2)[idx]
        rec["r"][k] = self.s["r"][idx]
        rec["gid"][k] = self.gid[idx]
        return self.torch.from_numpy(rec.view(np.uint8).copy())

    def count_overlapping(self, interval):
        return int(self._select(interval).shape[0])

    def recv_buffer(self, side):
        return self.recv[side]

    def unpack(self, side, buf):
        rec = buf.numpy().view(REC_DT)
        self.ghost_recs[side] = rec[1:1 + int(rec["gid"][0])].copy()

    def detect(self):
        from scisim_b200.slab import RebalanceNeeded
        if self.violation:
            raise RebalanceNeeded("oracle stand-in: a body left [%g, %g]" % (self.xlim[0], self.xlim[1]))
        L, R = self.ghost_recs[0], self.ghost_recs[1]
        self.ghosts = (L.shape[0], R.shape[0])
        nL, M = L.shape[0], self.n_owned
        # all local bodies in ascending GLOBAL index (the oracle lists pairs by local index)
        gid = np.concatenate([L["gid"], self.gid, R["gid"]]).astype(np.uint32)
        own = np.concatenate([np.zeros(nL, bool), np.ones(M, bool), np.zeros(R.shape[0], bool)])
        order = np.argsort(gid, kind="stable")
        assert np.all(np.diff(gid[order].astype(np.int64)) > 0), "a body arrived twice"
        gid, own = gid[order], own[order]
        loc = dict(self.s)
        loc["q"] = np.concatenate([L["q0"].ravel(), self.q0, R["q0"].ravel()]).reshape(-1, 2)[order].ravel().copy()
        q1 = np.concatenate([L["q1"].ravel(), self.q1, R["q1"].ravel()]).reshape(-1, 2)[order].ravel().copy()
        loc["r"] = np.concatenate([L["r"], self.s["r"], R["r"]])[order].copy()
        loc["m"] = np.ones(loc["r"].shape[0])
        a = self.ob.Ball2DOracle(loc).active_set(loc["q"], q1, "allpairs")
        ck = own[a["candidates"][:, 0]] if a["candidates"].shape[0] else np.zeros(0, bool)
        keep = own[a["i"]] if a["i"].shape[0] else np.zeros(0, bool)
        res = {"candidates": gid[a["candidates"][ck]].astype(np.uint32).reshape(-1, 2)}
        for k in ("type", "n", "p", "depth"):
            res[k] = a[k][keep]
        res["i"] = gid[a["i"][keep]]
        res["j"] = np.where(a["type"][keep] == 0, gid[np.minimum(a["j"][keep], gid.shape[0] - 1)], a["j"][keep]).astype(np.uint32)
        self.result = res
        return res["candidates"].shape[0], res["type"].shape[0]

    def fetch(self):
        return self.q1, self.v1, self.result
