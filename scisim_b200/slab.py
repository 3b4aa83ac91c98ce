"""Multi-GPU slab decomposition for the ball2d path, one process per GPU: the scene -- bodies numbered
arbitrarily -- is cut into equal-count x-quantile slabs, one per rank; ghost bodies are exchanged with the neighbouring
slabs every step.

The reference is single-process; this module is the host logic that has no counterpart there:
  * partition: sg_slab_partition (C, the same code the single-process sg_multi driver uses) ranks bodies by (x, index);
    a rank stores its bodies in ascending GLOBAL index, so its lists come out in the reference's order restricted to
    the pairs it owns
  * halo: a rank sends rank+-1 exactly its owned bodies whose swept AABB overlaps that rank's interval [min lo.x,
    max hi.x] -- by the kernels themselves through peer-mapped mailboxes ("p2p", NVLink, no collective in the step), or
    by all_gather + send/recv ("nccl", also gloo for the CPU tests)
  * non-neighbours: every owned body must stay inside its slab widened by half of each neighbouring slab
    (sg_slab_limits); a body outside could touch a body two slabs away, which no halo carries.  The device flags it
    (every step, either transport), sg_ball2d_slab_detect returns SG_ERR_REBALANCE, the ranks agree on it and the scene
    is re-partitioned from the current state (Ball2DSlabSim.step)
  * ownership: pair (i<j) is kept by the rank that owns body i, so the per-rank lists are disjoint and each ascending;
    merge_active_sets interleaves them body by body into the reference's std::set order (ball2d/Ball2DSim.cpp:580),
    then drums drum-major, then planes plane-major
All device work goes through a backend (GpuSlabBackend below; the CPU tests plug the oracle in as the backend to
exercise this host logic under gloo).
"""
import ctypes as C

import numpy as np

REC_BYTES = 48
# contact types whose j is a static geometry (drum, plane, cylinder), not a body: include/scisim_b200.h
STATIC_TYPES = np.array([1, 2, 14, 15, 16, 17, 18, 23, 24], dtype=np.uint32)


class RebalanceNeeded(RuntimeError):
    """A body left its slab's neighbourhood (SG_ERR_REBALANCE): the step's lists may miss contacts with a
    non-neighbouring slab; re-partition and repeat the step."""


def _lib():
    from . import _lib as L
    return L.load()


def partition_slab_major(n_total, world):
    """Owned global index range [first, first + count) of every rank: equal-count contiguous blocks (a scene that is
    already numbered slab-major)."""
    base, rem = divmod(n_total, world)
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    firsts = [sum(counts[:r]) for r in range(world)]
    return firsts, counts


def partition_quantiles(q, world):
    """Equal-count x-quantile slabs of an arbitrarily numbered 2-D scene (q = [x0,y0,x1,y1,...]).  Returns
    (rank_of[n], cuts[world+1], [ascending global indices of rank k for k in range(world)])."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    n = q.shape[0] // 2
    rank_of = np.zeros(n, dtype=np.uint32)
    cuts = np.zeros(world + 1, dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib().sg_slab_partition(n, vp(q), 2, world, vp(rank_of), vp(cuts))
    if rc != 0:
        raise RuntimeError("sg_slab_partition failed (%d)" % rc)
    gids = [np.nonzero(rank_of == k)[0].astype(np.uint32) for k in range(world)]
    return rank_of, cuts, gids


def slab_limits(cuts, rank):
    cuts = np.ascontiguousarray(cuts, dtype=np.float64)
    lim = np.zeros(2, dtype=np.float64)
    rc = _lib().sg_slab_limits(cuts.shape[0] - 1, cuts.ctypes.data_as(C.c_void_p), rank, lim.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("sg_slab_limits failed (%d)" % rc)
    return lim


def _merge_dest(n_bodies, firsts, stride):
    """dest arrays of sg_slab_merge_dest for per-part first-index columns (uint32 arrays, element stride `stride`)."""
    k = len(firsts)
    firsts = [np.ascontiguousarray(f, dtype=np.uint32) for f in firsts]
    lens = np.array([f.shape[0] // stride for f in firsts], dtype=np.uint64)
    dests = [np.zeros(int(l), dtype=np.uint64) for l in lens]
    fp = (C.c_void_p * k)(*[f.ctypes.data for f in firsts])
    dp = (C.c_void_p * k)(*[d.ctypes.data for d in dests])
    rc = _lib().sg_slab_merge_dest(int(n_bodies), k, fp, stride, lens.ctypes.data_as(C.c_void_p), dp)
    if rc != 0:
        raise RuntimeError("sg_slab_merge_dest failed (%d): a per-rank list is not ascending in its first index" % rc)
    return dests


def merge_active_sets(parts, n_static_geoms=None, n_bodies=None):
    """parts: per-rank dicts (any order) with type,i,j,n,p,depth,candidates in global indices, each list in the order
    sg_ball2d_slab_detect leaves it.  Returns the global active set in the reference's order: ball-ball ascending (i,j),
    then drums drum-major, then planes plane-major, each ball-ascending."""
    out = {}
    if not parts:
        return {"candidates": np.zeros((0, 2), np.uint32)}
    if n_bodies is None:
        n_bodies = 1
        for p in parts:
            for a in (p["candidates"], p["i"], p["j"][~np.isin(p["type"], STATIC_TYPES)]):
                if a.size:
                    n_bodies = max(n_bodies, int(a.max()) + 1)
    cands = [np.ascontiguousarray(p["candidates"], dtype=np.uint32).reshape(-1, 2) for p in parts]
    dests = _merge_dest(n_bodies, [c.ravel() for c in cands], 2)
    total = sum(c.shape[0] for c in cands)
    out["candidates"] = np.zeros((total, 2), dtype=np.uint32)
    for c, d in zip(cands, dests):
        out["candidates"][d] = c
    keys = tuple(k for k in ("type", "i", "j", "aux", "n", "p", "depth") if k in parts[0] and parts[0][k] is not None)
    static = lambda t: np.isin(t, STATIC_TYPES)
    nbb = [int(np.count_nonzero(~static(p["type"]))) for p in parts]
    for p, k in zip(parts, nbb):
        assert not np.any(static(p["type"][:k])), "body-body contacts come first in a rank's list"
    bb_dest = _merge_dest(n_bodies, [p["i"][:k] for p, k in zip(parts, nbb)], 1)
    n_bb = sum(nbb)
    # static contacts: (type, geometry, body)
    kk = len(parts)
    st = [[np.ascontiguousarray(p[c][k:], dtype=np.uint32) for p, k in zip(parts, nbb)] for c in ("type", "i", "j")]
    slens = np.array([a.shape[0] for a in st[0]], dtype=np.uint64)
    sdest = [np.zeros(int(l), dtype=np.uint64) for l in slens]
    arr = lambda lst: (C.c_void_p * kk)(*[a.ctypes.data for a in lst])
    rc = _lib().sg_slab_merge_static_dest(kk, arr(st[0]), arr(st[1]), arr(st[2]), slens.ctypes.data_as(C.c_void_p), arr(sdest))
    if rc != 0:
        raise RuntimeError("sg_slab_merge_static_dest failed (%d)" % rc)
    n_all = n_bb + int(slens.sum())
    for key in keys:
        proto = parts[0][key]
        shape = (n_all,) + tuple(proto.shape[1:])
        o = np.zeros(shape, dtype=proto.dtype)
        for p, k, db, ds in zip(parts, nbb, bb_dest, sdest):
            o[db] = p[key][:k]
            o[n_bb + ds] = p[key][k:]
        out[key] = o
    return out


class GpuSlabBackend:
    """One slab on one GPU through the sg_ball2d_slab_* calls; exchange buffers are torch CUDA tensors and every
    torch op is issued on the library's stream.  Buffers always travel at full size (ghost_cap + 1 records, the first
    one a header with the count), so nothing on the host waits for a count before detect()."""

    def __init__(self, ctx, scene_slab, gid_first, ghost_cap, gids=None, x_limits=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = ctx.lib
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.device)
        self.cap = int(ghost_cap)
        nbytes = (self.cap + 1) * REC_BYTES
        with torch.cuda.stream(self.stream):
            self.iv = torch.zeros(2, dtype=torch.float64, device=self.device)
            self.count = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.send = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self.recv = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.ghosts = (0, 0)
        self.reinit(scene_slab, gid_first, gids, x_limits)

    def reinit(self, scene_slab, gid_first=0, gids=None, x_limits=None):
        """(Re)loads the slab: bodies (r, m, q, v of the owned bodies in ascending global index), their global indices,
        the x-range they must stay inside, the static geometry.  Mailboxes and neighbour mappings stay as they are."""
        ctx = self.ctx
        s = scene_slab
        r = np.ascontiguousarray(s["r"], dtype=np.float64)
        m = np.ascontiguousarray(s["m"], dtype=np.float64)
        self.n_owned = r.shape[0]
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        ctx.check(self.lib.sg_ball2d_slab_init(ctx.h, self.n_owned, int(gid_first), self.cap, vp(r), vp(m)))
        if gids is not None or x_limits is not None:
            g = np.ascontiguousarray(gids, dtype=np.uint32) if gids is not None else None
            lim = np.ascontiguousarray(x_limits, dtype=np.float64) if x_limits is not None else None
            ctx.check(self.lib.sg_ball2d_slab_set_gids(ctx.h, vp(g) if g is not None else None, vp(lim) if lim is not None else None))
        g = np.ascontiguousarray(s["g"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_gravity(ctx.h, vp(g)))
        px, pn = np.ascontiguousarray(s["plane_x"], dtype=np.float64), np.ascontiguousarray(s["plane_n"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_planes(ctx.h, px.shape[0], vp(px), vp(pn)))
        dx, dr = np.ascontiguousarray(s["drum_x"], dtype=np.float64), np.ascontiguousarray(s["drum_r"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_drums(ctx.h, dx.shape[0], vp(dx), vp(dr)))
        self.upload(s["q"], s["v"])

    def upload(self, q, v):
        q, v = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.lib.sg_ball2d_upload(self.ctx.h, vp(q), vp(v)))

    def flow(self, kind, dt):
        self.ctx.check(self.lib.sg_ball2d_slab_flow(self.ctx.h, int(kind), float(dt), C.c_void_p(self.iv.data_ptr())))
        return self.iv

    # ---- peer-memory transport (NVLink, no collective in the step) ----
    def mailbox(self):
        """(device address, 64-byte CUDA IPC handle) of this rank's mailbox; creates it on first use."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self.ctx.check(self.lib.sg_ball2d_slab_mailbox(self.ctx.h, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def connect(self, side, ipc_handle=None, same_process_ptr=None, peer_device=-1):
        h = (C.c_ubyte * 64).from_buffer_copy(ipc_handle) if ipc_handle is not None else None
        self.ctx.check(self.lib.sg_ball2d_slab_connect(self.ctx.h, int(side), h, C.c_void_p(same_process_ptr) if same_process_ptr is not None else None, int(peer_device)))

    def exchange(self, phase=0):
        self.ctx.check(self.lib.sg_ball2d_slab_exchange(self.ctx.h, int(phase)))

    def disconnect(self):
        """Back to the collective transport: drops the mailbox and the neighbour mappings."""
        self.ctx.check(self.lib.sg_ball2d_slab_disconnect(self.ctx.h))

    def pack(self, interval, side):
        """Selects, in body order, the owned bodies overlapping `interval` (2-element tensor on this device) into the
        side's send buffer (header + records). Asynchronous."""
        buf = self.send[side]
        self.ctx.check(self.lib.sg_ball2d_slab_pack(self.ctx.h, C.c_void_p(interval.data_ptr()), C.c_void_p(buf.data_ptr()), self.cap, C.c_void_p(self.count.data_ptr())))
        return buf

    def count_overlapping(self, interval):
        """How many owned bodies reach `interval` (host-synchronous; diagnostics only -- the non-neighbour guarantee is
        the x-limit test the flow kernel makes every step)."""
        self.ctx.check(self.lib.sg_ball2d_slab_pack(self.ctx.h, C.c_void_p(interval.data_ptr()), None, 0, C.c_void_p(self.count.data_ptr())))
        with self.torch.cuda.stream(self.stream):
            return int(self.count.item())

    def recv_buffer(self, side):
        return self.recv[side]

    def unpack(self, side, buf):
        self.ctx.check(self.lib.sg_ball2d_slab_unpack(self.ctx.h, int(side), C.c_void_p(buf.data_ptr())))

    def detect(self):
        from ._lib import SG_ERR_REBALANCE, SgContacts
        c = SgContacts()
        g = (C.c_uint32 * 2)()
        rc = self.lib.sg_ball2d_slab_detect(self.ctx.h, C.byref(c), g)
        if rc == SG_ERR_REBALANCE:
            raise RebalanceNeeded(self.lib.sg_last_error(self.ctx.h).decode())
        self.ctx.check(rc)
        self.ghosts = (int(g[0]), int(g[1]))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self):
        from ._lib import SG_OUT_ALL, SgContacts
        from .host_api import ActiveSet
        q1, v1 = np.empty(2 * self.n_owned), np.empty(2 * self.n_owned)
        c = SgContacts()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.lib.sg_ball2d_fetch(self.ctx.h, SG_OUT_ALL, vp(q1), vp(v1), C.byref(c)))
        a = ActiveSet(c)
        return q1, v1, {"type": a.type, "i": a.i, "j": a.j, "n": a.n, "p": a.p, "depth": a.depth, "candidates": a.candidates}

    def run_on_stream(self):
        return self.torch.cuda.stream(self.stream)


class Ball2DSlabs:
    """Per-rank driver of one step: flow -> interval exchange -> halo exchange -> detection.  With the GPU backend
    nothing blocks the host until detect() reads the list sizes."""

    def __init__(self, backend, rank, world, dist, check_non_neighbours=False, transport="nccl"):
        """transport "nccl": intervals by all_gather, halos by send/recv (any backend, also gloo + the oracle backend);
        "p2p": the neighbours' mailboxes are mapped once with CUDA IPC (handles travel through `dist`), after that a
        step issues no collective at all -- intervals and halos are written over NVLink by the kernels themselves.
        check_non_neighbours: additionally count (host-synchronously, collective transport only) the owned bodies that
        reach a non-neighbour's interval -- a diagnostic on top of the x-limit test every step makes on the device."""
        self.b, self.rank, self.world, self.dist = backend, rank, world, dist
        self.check = check_non_neighbours
        self.last_halo = (0, 0)
        self.transport = transport if world > 1 else "nccl"
        if self.transport == "p2p":
            import torch
            dev = getattr(backend, "device", "cpu")   # a backend without device memory (the oracle stand-in) cannot map anything
            ok = 1
            try:
                _, handle = backend.mailbox()
                mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            except Exception:
                ok, mine = 0, torch.zeros(64, dtype=torch.uint8, device=dev)
            allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine)
            allh = allh.cpu().numpy().reshape(world, 64)
            if ok:
                try:
                    for side, peer in ((0, rank - 1), (1, rank + 1)):
                        if 0 <= peer < world:
                            backend.connect(side, ipc_handle=bytes(allh[peer]))
                except Exception:
                    ok = 0
            # every rank must take the same transport: fall back to NCCL everywhere if any mapping failed (no peer
            # access between two of the GPUs, CUDA IPC not permitted in this container, ...)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.transport = "nccl"
                if hasattr(backend, "disconnect"):
                    backend.disconnect()
            dist.barrier()

    def step(self, kind, dt):
        import contextlib
        import torch
        b, dist, rank, W = self.b, self.dist, self.rank, self.world
        cm = b.run_on_stream() if hasattr(b, "run_on_stream") else contextlib.nullcontext()
        with cm:
            iv = b.flow(kind, dt)
            if W > 1 and self.transport == "p2p":
                b.exchange(0)
            elif W > 1:
                all_iv = torch.empty(W * 2, dtype=torch.float64, device=iv.device)
                dist.all_gather_into_tensor(all_iv, iv)
                all_iv = all_iv.view(W, 2)
                peers = {0: rank - 1, 1: rank + 1}
                ops = []
                for side, peer in peers.items():
                    if 0 <= peer < W:
                        ops.append(dist.P2POp(dist.isend, b.pack(all_iv[peer], side), peer))
                        ops.append(dist.P2POp(dist.irecv, b.recv_buffer(side), peer))
                if self.check:
                    for peer in range(W):
                        if abs(peer - rank) > 1:
                            c = b.count_overlapping(all_iv[peer])
                            if c != 0:
                                raise RebalanceNeeded("rank %d: %d bodies reach the slab of non-neighbour rank %d" % (rank, c, peer))
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                for side, peer in peers.items():
                    if 0 <= peer < W:
                        b.unpack(side, b.recv_buffer(side))
            res = b.detect()
            self.last_halo = getattr(b, "ghosts", (0, 0))
            return res


class HostAgreement:
    """"Does any rank want to re-partition?" after every step, without touching the GPUs.  The ranks of one node share a small
    memory-mapped file under /dev/shm: rank r writes ( step number, flag ) into its slot of the step's parity buffer and spins until
    every slot carries that step (two buffers, because a fast rank may already be writing step s + 1 while a slow one still reads
    step s; it cannot reach step s + 2 before everybody has passed s + 1).  A device collective for these 4 bytes costs a kernel
    launch, a device-to-host copy and a stream synchronisation per step -- a tenth of a 0.8 ms step at 8 GPUs.  Falls back to
    dist.all_reduce when the file cannot be shared (several nodes, no /dev/shm)."""

    def __init__(self, rank, world, dist, device=None):
        import os
        self.rank, self.world, self.dist, self.device = rank, world, dist, device
        self.step_no = 0
        self.arr = None
        if world == 1:
            return
        name = [None]
        if rank == 0:
            name[0] = "/dev/shm/scisim_b200_agree_%d_%d" % (os.getpid(), id(self) & 0xffffff)
            try:
                with open(name[0], "wb") as f:
                    f.write(b"\0" * (2 * world * 8))
            except OSError:
                name[0] = None
        dist.broadcast_object_list(name, src=0)
        ok = 1
        try:
            if name[0] is None:
                raise OSError("no shared file")
            self.arr = np.memmap(name[0], dtype=np.int64, mode="r+", shape=(2, world))
        except Exception:
            ok = 0
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if not all(oks):
            self.arr = None      # every rank must take the same road
        dist.barrier()
        if rank == 0 and name[0] is not None:
            try:
                os.unlink(name[0])   # the mappings stay valid; nothing is left behind
            except OSError:
                pass

    def any(self, flag):
        if self.world == 1:
            return bool(flag)
        if self.arr is None:
            import torch
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.device if self.device is not None else "cpu")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return bool(int(t.item()))
        self.step_no += 1
        buf = self.arr[self.step_no & 1]
        want = 2 * self.step_no
        buf[self.rank] = want + (1 if flag else 0)   # one aligned 8-byte store
        spins = 0
        while True:
            vals = np.array(buf)                     # a snapshot
            if bool((vals >= want).all()):
                return bool(((vals - want) == 1).any())
            spins += 1
            if spins > 50_000_000:
                raise RuntimeError("rank %d: the other ranks did not reach step %d" % (self.rank, self.step_no))


class Ball2DSlabSim:
    """One rank's handle on a WHOLE ball2d scene (global arrays, arbitrary numbering) spread over `world` ranks.
    upload() partitions (the same partition on every rank: it is a deterministic function of q), step() runs one
    resident step and re-partitions when any rank asks for it, gather_merged() assembles the reference-order lists.

    backend_factory( scene_slab, gids, x_limits, ghost_cap ) -> backend (GpuSlabBackend for the product; the CPU tests
    pass the oracle stand-in)."""

    def __init__(self, scene, rank, world, dist, backend_factory, transport="p2p", ghost_cap=None, agree=True):
        self.scene = scene           # r, m, g, plane_x, plane_n, drum_x, drum_r (global); q, v come through upload()
        self.rank, self.world, self.dist = rank, world, dist
        self.factory = backend_factory
        self.want_transport = transport
        self.ghost_cap = ghost_cap
        self.agree = agree           # all ranks agree on "re-partition" after every step (HostAgreement: shared memory on one node, else a 4-byte all_reduce)
        self._agreement = None
        self.backend = None
        self.driver = None
        self.n = int(np.asarray(scene["r"]).shape[0])
        self.n_partitions = 0
        self.q = self.v = None
        self.gids = None

    def _slab_scene(self, gids, q, v):
        s = dict(self.scene)
        s["r"] = np.ascontiguousarray(np.asarray(self.scene["r"])[gids])
        s["m"] = np.ascontiguousarray(np.asarray(self.scene["m"])[gids])
        s["q"] = np.ascontiguousarray(q.reshape(-1, 2)[gids].ravel())
        s["v"] = np.ascontiguousarray(v.reshape(-1, 2)[gids].ravel())
        return s

    def _partition(self, q, v):
        rank_of, cuts, gids = partition_quantiles(q, self.world)
        self.cuts, self.all_gids, self.gids = cuts, gids, gids[self.rank]
        lim = slab_limits(cuts, self.rank)
        cap = self.ghost_cap or max(4096, max(g.shape[0] for g in gids) // 32)
        s = self._slab_scene(self.gids, q, v)
        if self.backend is None:
            self.ghost_cap = cap
            self.backend = self.factory(s, self.gids, lim, cap)
            self.driver = Ball2DSlabs(self.backend, self.rank, self.world, self.dist, transport=self.want_transport)
        else:
            self.backend.reinit(s, 0, self.gids, lim)
        self.n_partitions += 1

    def upload(self, q, v, repartition=False):
        """q, v: the GLOBAL state (identical arrays on every rank)."""
        q, v = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        if self.backend is None or repartition:
            self._partition(q, v)
        else:
            self.backend.upload(q.reshape(-1, 2)[self.gids].ravel(), v.reshape(-1, 2)[self.gids].ravel())
        self.q, self.v = q, v

    @property
    def transport(self):
        return self.driver.transport if self.driver is not None else None

    def _agree(self, flag):
        if self.world == 1 or not self.agree:
            return flag
        if self._agreement is None:
            self._agreement = HostAgreement(self.rank, self.world, self.dist, getattr(self.backend, "device", "cpu"))
        return self._agreement.any(flag)

    def step(self, kind, dt):
        """One resident step of this rank's slab: (candidates, contacts) it owns.  If any rank reports a body outside its
        slab's neighbourhood, all ranks re-partition from the uploaded state and repeat the step (once)."""
        for attempt in range(2):
            need = False
            try:
                res = self.driver.step(kind, dt)
            except RebalanceNeeded:
                need, res = True, (0, 0)
            if not self._agree(need):
                if need:
                    raise RebalanceNeeded("rank %d needs a re-partition but the ranks do not agree on steps (agree=False)" % self.rank)
                return res
            if attempt == 1:
                raise RuntimeError("the slabs of this scene are too thin for %d ranks: bodies reach beyond the neighbouring slab even after a fresh partition" % self.world)
            self._partition(self.q, self.v)
        return res

    def fetch(self):
        """(global indices of the owned bodies, q1, v1 of those bodies, this rank's lists)"""
        q1, v1, res = self.backend.fetch()
        return self.gids, q1, v1, res

    def gather_merged(self, dst=0):
        """Collects every rank's lists and state on rank `dst` and merges them into the reference's order (parity checks;
        the lists travel as pickled numpy arrays)."""
        gids, q1, v1, res = self.fetch()
        res = dict(res)
        res["gids"], res["q1"], res["v1"] = gids, q1, v1
        if self.world == 1:
            parts = [res]
        else:
            parts = [None] * self.world if self.rank == dst else None
            self.dist.gather_object(res, parts, dst=dst)
            if self.rank != dst:
                return None
        merged = merge_active_sets(parts, n_bodies=self.n)
        q1g, v1g = np.zeros(2 * self.n), np.zeros(2 * self.n)
        for p in parts:
            q1g.reshape(-1, 2)[p["gids"]] = p["q1"].reshape(-1, 2)
            v1g.reshape(-1, 2)[p["gids"]] = p["v1"].reshape(-1, 2)
        merged["q1"], merged["v1"] = q1g, v1g
        return merged


# ---- rigidbody3d (all-sphere scenes: BASELINE configs[3]) ---------------------------------------------------------------------------
class RB3DSlabBackend:
    """One slab of an all-sphere rigidbody3d scene on one GPU through the sg_rb3d_slab_* calls (peer-memory exchange only)."""

    def __init__(self, ctx, scene, gids, x_limits, ghost_cap):
        """scene: the WHOLE scene's static description (geometry list, geo_of_body, m, I0, g, planes); gids: this rank's bodies, ascending."""
        self.ctx, self.lib = ctx, ctx.lib
        self.cap = int(ghost_cap)
        s = scene
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        gt = np.ascontiguousarray(s["geo_type"], dtype=np.uint32)
        gr, gh = np.ascontiguousarray(s["geo_r"], dtype=np.float64), np.ascontiguousarray(s["geo_half"], dtype=np.float64)
        gm = np.ascontiguousarray(s["geo_mesh"], dtype=np.uint32)
        ctx.check(self.lib.sg_rb3d_set_geometry(ctx.h, gt.shape[0], vp(gt), vp(gr), vp(gh), vp(gm)))
        g = np.ascontiguousarray(s["g"], dtype=np.float64)
        ctx.check(self.lib.sg_rb3d_set_gravity(ctx.h, vp(g)))
        px, pn = np.ascontiguousarray(s["plane_x"], dtype=np.float64), np.ascontiguousarray(s["plane_n"], dtype=np.float64)
        ctx.check(self.lib.sg_rb3d_set_planes(ctx.h, px.shape[0], vp(px), vp(pn)))
        self.reinit(scene, gids, x_limits)

    def reinit(self, scene, gids, x_limits):
        ctx = self.ctx
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.gids = np.ascontiguousarray(gids, dtype=np.uint32)
        self.n_owned = self.gids.shape[0]
        geo = np.ascontiguousarray(np.asarray(scene["geo_of_body"])[self.gids], dtype=np.uint32)
        m = np.ascontiguousarray(np.asarray(scene["m"])[self.gids], dtype=np.float64)
        I0 = np.ascontiguousarray(np.asarray(scene["I0"]).reshape(-1, 3)[self.gids], dtype=np.float64)
        lim = np.ascontiguousarray(x_limits, dtype=np.float64)
        if np.any(np.asarray(scene["fixed"])[self.gids]):
            raise RuntimeError("kinematically scripted bodies are not supported in slab mode")
        ctx.check(self.lib.sg_rb3d_slab_init(ctx.h, self.n_owned, self.cap, vp(geo), vp(m), vp(I0), vp(self.gids), vp(lim)))

    def upload(self, q_owned, v_owned):
        q, v = np.ascontiguousarray(q_owned, dtype=np.float64), np.ascontiguousarray(v_owned, dtype=np.float64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.lib.sg_rb3d_upload(self.ctx.h, vp(q), vp(v)))

    def mailbox(self):
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self.ctx.check(self.lib.sg_rb3d_slab_mailbox(self.ctx.h, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def connect(self, side, ipc_handle=None, same_process_ptr=None, peer_device=-1):
        h = (C.c_ubyte * 64).from_buffer_copy(ipc_handle) if ipc_handle is not None else None
        self.ctx.check(self.lib.sg_rb3d_slab_connect(self.ctx.h, int(side), h, C.c_void_p(same_process_ptr) if same_process_ptr is not None else None, int(peer_device)))

    def disconnect(self):
        self.ctx.check(self.lib.sg_rb3d_slab_disconnect(self.ctx.h))

    def flow(self, kind, dt):
        self.ctx.check(self.lib.sg_rb3d_slab_flow(self.ctx.h, int(kind), float(dt)))

    def exchange(self, phase=0):
        self.ctx.check(self.lib.sg_rb3d_slab_exchange(self.ctx.h, int(phase)))

    def detect(self):
        from ._lib import SG_ERR_REBALANCE, SgContacts
        c = SgContacts()
        g = (C.c_uint32 * 2)()
        rc = self.lib.sg_rb3d_slab_detect(self.ctx.h, C.byref(c), g)
        if rc == SG_ERR_REBALANCE:
            raise RebalanceNeeded(self.lib.sg_last_error(self.ctx.h).decode())
        self.ctx.check(rc)
        self.ghosts = (int(g[0]), int(g[1]))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self):
        from ._lib import SG_OUT_ALL, SgContacts
        from .host_api import ActiveSet
        q1, v1 = np.empty(12 * self.n_owned), np.empty(6 * self.n_owned)
        c = SgContacts()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.ctx.check(self.lib.sg_rb3d_fetch(self.ctx.h, SG_OUT_ALL, vp(q1), vp(v1), C.byref(c)))
        a = ActiveSet(c)
        return q1, v1, {"type": a.type, "i": a.i, "j": a.j, "aux": a.aux, "n": a.n, "p": a.p, "depth": a.depth, "candidates": a.candidates}


def rb3d_owned_state(q, v, gids):
    """The owned bodies' ( q, v ) in the layout of a smaller scene: q = [3 n | 9 n], v = [3 n | 3 n]."""
    n = q.shape[0] // 12
    x, R = q[:3 * n].reshape(n, 3), q[3 * n:].reshape(n, 9)
    vl, w = v[:3 * n].reshape(n, 3), v[3 * n:].reshape(n, 3)
    return np.concatenate([x[gids].ravel(), R[gids].ravel()]), np.concatenate([vl[gids].ravel(), w[gids].ravel()])


def rb3d_scatter_state(n, parts):
    """Inverse of rb3d_owned_state over all ranks: parts = [(gids, q_owned, v_owned)]."""
    q, v = np.zeros(12 * n), np.zeros(6 * n)
    for gids, qo, vo in parts:
        k = gids.shape[0]
        q[:3 * n].reshape(n, 3)[gids] = qo[:3 * k].reshape(k, 3)
        q[3 * n:].reshape(n, 9)[gids] = qo[3 * k:].reshape(k, 9)
        v[:3 * n].reshape(n, 3)[gids] = vo[:3 * k].reshape(k, 3)
        v[3 * n:].reshape(n, 3)[gids] = vo[3 * k:].reshape(k, 3)
    return q, v


def partition_quantiles_3d(q, world):
    """x-quantile slabs of a rigidbody3d scene ( q = [3N positions | 9N rotations] )."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    n = q.shape[0] // 12
    rank_of = np.zeros(n, dtype=np.uint32)
    cuts = np.zeros(world + 1, dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib().sg_slab_partition(n, vp(q), 3, world, vp(rank_of), vp(cuts))
    if rc != 0:
        raise RuntimeError("sg_slab_partition failed (%d)" % rc)
    return rank_of, cuts, [np.nonzero(rank_of == k)[0].astype(np.uint32) for k in range(world)]


class RB3DSlabSim:
    """One rank's handle on a WHOLE all-sphere rigidbody3d scene spread over `world` ranks, one process per GPU (CUDA IPC mailboxes):
    the counterpart of Ball2DSlabSim.  upload() takes the global state and partitions it (deterministically, the same on every rank)."""

    def __init__(self, ctx, scene, rank, world, dist, ghost_cap=None, agree=True):
        self.ctx, self.scene, self.rank, self.world, self.dist = ctx, scene, rank, world, dist
        self.ghost_cap, self.agree = ghost_cap, agree
        self._agreement = None
        self.backend = None
        self.n = int(np.asarray(scene["m"]).shape[0])
        self.n_partitions = 0
        self.last_halo = (0, 0)

    def _partition(self, q, v):
        rank_of, cuts, gids = partition_quantiles_3d(q, self.world)
        self.gids = gids[self.rank]
        lim = slab_limits(cuts, self.rank)
        if self.backend is None:
            self.ghost_cap = self.ghost_cap or max(8192, max(g.shape[0] for g in gids) // 8)
            self.backend = RB3DSlabBackend(self.ctx, self.scene, self.gids, lim, self.ghost_cap)
            self._connect()
        else:
            self.backend.reinit(self.scene, self.gids, lim)
        self.backend.upload(*rb3d_owned_state(q, v, self.gids))
        self.n_partitions += 1

    def _connect(self):
        import torch
        _, handle = self.backend.mailbox()
        if self.world == 1:
            return
        dev = torch.device("cuda", self.ctx.device)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().reshape(self.world, 64)
        for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= peer < self.world:
                self.backend.connect(side, ipc_handle=bytes(allh[peer]))
        self.dist.barrier()

    def upload(self, q, v, repartition=False):
        q, v = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        self.q, self.v = q, v
        if self.backend is None or repartition:
            self._partition(q, v)
        else:
            self.backend.upload(*rb3d_owned_state(q, v, self.gids))

    def _agree(self, flag):
        if self.world == 1 or not self.agree:
            return flag
        if self._agreement is None:
            import torch
            self._agreement = HostAgreement(self.rank, self.world, self.dist, torch.device("cuda", self.ctx.device))
        return self._agreement.any(flag)

    def step(self, kind, dt):
        for attempt in range(2):
            need = False
            self.backend.flow(kind, dt)
            self.backend.exchange(0)
            try:
                res = self.backend.detect()
                self.last_halo = self.backend.ghosts
            except RebalanceNeeded:
                need, res = True, (0, 0)
            if not self._agree(need):
                if need:
                    raise RebalanceNeeded("rank %d needs a re-partition but the ranks do not agree on steps (agree=False)" % self.rank)
                return res
            if attempt == 1:
                raise RuntimeError("the slabs of this scene are too thin for %d ranks" % self.world)
            self._partition(self.q, self.v)
        return res

    def gather_merged(self, dst=0):
        q1, v1, res = self.backend.fetch()
        res = dict(res)
        res["gids"], res["q1"], res["v1"] = self.gids, q1, v1
        if self.world == 1:
            parts = [res]
        else:
            parts = [None] * self.world if self.rank == dst else None
            self.dist.gather_object(res, parts, dst=dst)
            if self.rank != dst:
                return None
        merged = merge_active_sets(parts, n_bodies=self.n)
        merged["q1"], merged["v1"] = rb3d_scatter_state(self.n, [(p["gids"], p["q1"], p["v1"]) for p in parts])
        return merged
