"""Multi-GPU slab decomposition for the ball2d path: one process per GPU, one slab of a
slab-major-numbered scene per process, ghost bodies exchanged with the neighbouring slabs every step.

The reference is single-process; this module is the host logic that has no counterpart there:
  * interval exchange: every rank publishes [min lo.x, max hi.x] of its owned swept AABBs (all_gather, 16 B/rank)
  * halo: a rank sends rank+-1 exactly its owned bodies whose swept AABB overlaps that rank's interval (NCCL
    send/recv over NVLink when the tensors are CUDA tensors); a body overlapping a non-neighbour's interval means
    the slabs need re-balancing and raises
  * ownership: pair (i<j) is kept by the rank that owns body i, so the per-rank lists are disjoint, each ascending,
    and their concatenation in rank order is the reference's std::set order; planes / drums are tested for owned
    bodies only and merged geometry-major
All device work goes through a backend (GpuSlabBackend below; the CPU tests plug the oracle in as the backend to
exercise this host logic under gloo).
"""
import ctypes as C

import numpy as np

REC_BYTES = 48


def partition_slab_major(n_total, world):
    """Owned global index range [first, first + count) of every rank: equal-count contiguous blocks."""
    base, rem = divmod(n_total, world)
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    firsts = [sum(counts[:r]) for r in range(world)]
    return firsts, counts


def merge_active_sets(parts, n_static_geoms):
    """parts: per-rank dicts (rank order) with type,i,j,n,p,depth,candidates in global indices. Returns the global
    active set in the reference's order: ball-ball (rank order == ascending (i,j)), then drums drum-major, then planes
    plane-major, each ball-ascending (ranks own ascending index ranges)."""
    out = {}
    out["candidates"] = np.concatenate([p["candidates"] for p in parts]) if parts else np.zeros((0, 2), np.uint32)
    keys = ("type", "i", "j", "n", "p", "depth")
    chunks = {k: [] for k in keys}
    for p in parts:
        sel = p["type"] == 0
        for k in keys:
            chunks[k].append(p[k][sel])
    for t in (1, 2):  # drums then planes
        for g in range(n_static_geoms[t - 1]):
            for p in parts:
                sel = (p["type"] == t) & (p["j"] == g)
                for k in keys:
                    chunks[k].append(p[k][sel])
    for k in keys:
        out[k] = np.concatenate(chunks[k]) if chunks[k] else None
    return out


class GpuSlabBackend:
    """One slab on one GPU through the sg_ball2d_slab_* calls; exchange buffers are torch CUDA tensors and every
    torch op is issued on the library's stream.  Buffers always travel at full size (ghost_cap + 1 records, the first
    one a header with the count), so nothing on the host waits for a count before detect()."""

    def __init__(self, ctx, scene_slab, gid_first, ghost_cap):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = ctx.lib
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.device)
        self.cap = int(ghost_cap)
        s = scene_slab
        r = np.ascontiguousarray(s["r"], dtype=np.float64)
        m = np.ascontiguousarray(s["m"], dtype=np.float64)
        self.n_owned = r.shape[0]
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        ctx.check(self.lib.sg_ball2d_slab_init(ctx.h, self.n_owned, int(gid_first), self.cap, vp(r), vp(m)))
        g = np.ascontiguousarray(s["g"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_gravity(ctx.h, vp(g)))
        px, pn = np.ascontiguousarray(s["plane_x"], dtype=np.float64), np.ascontiguousarray(s["plane_n"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_planes(ctx.h, px.shape[0], vp(px), vp(pn)))
        dx, dr = np.ascontiguousarray(s["drum_x"], dtype=np.float64), np.ascontiguousarray(s["drum_r"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_set_drums(ctx.h, dx.shape[0], vp(dx), vp(dr)))
        q, v = np.ascontiguousarray(s["q"], dtype=np.float64), np.ascontiguousarray(s["v"], dtype=np.float64)
        ctx.check(self.lib.sg_ball2d_upload(ctx.h, vp(q), vp(v)))
        nbytes = (self.cap + 1) * REC_BYTES
        with torch.cuda.stream(self.stream):
            self.iv = torch.zeros(2, dtype=torch.float64, device=self.device)
            self.count = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.send = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self.recv = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.ghosts = (0, 0)

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
        """How many owned bodies reach `interval` (host-synchronous; used for the non-neighbour check only)."""
        self.ctx.check(self.lib.sg_ball2d_slab_pack(self.ctx.h, C.c_void_p(interval.data_ptr()), None, 0, C.c_void_p(self.count.data_ptr())))
        with self.torch.cuda.stream(self.stream):
            return int(self.count.item())

    def recv_buffer(self, side):
        return self.recv[side]

    def unpack(self, side, buf):
        self.ctx.check(self.lib.sg_ball2d_slab_unpack(self.ctx.h, int(side), C.c_void_p(buf.data_ptr())))

    def detect(self):
        from ._lib import SgContacts
        c = SgContacts()
        g = (C.c_uint32 * 2)()
        self.ctx.check(self.lib.sg_ball2d_slab_detect(self.ctx.h, C.byref(c), g))
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
    """Per-rank driver of one step: flow -> interval all_gather -> halo exchange -> detection.  With the GPU backend
    nothing blocks the host until detect() reads the list sizes."""

    def __init__(self, backend, rank, world, dist, check_non_neighbours=False, transport="nccl"):
        """transport "nccl": intervals by all_gather, halos by send/recv (any backend, also gloo + the oracle backend);
        "p2p": the neighbours' mailboxes are mapped once with CUDA IPC (handles travel through `dist`), after that a
        step issues no collective at all -- intervals and halos are written over NVLink by the kernels themselves."""
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
                                raise RuntimeError("rank %d: %d bodies reach the slab of non-neighbour rank %d; re-balance the slabs" % (rank, c, peer))
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                for side, peer in peers.items():
                    if 0 <= peer < W:
                        b.unpack(side, b.recv_buffer(side))
            res = b.detect()
            self.last_halo = getattr(b, "ghosts", (0, 0))
            return res
