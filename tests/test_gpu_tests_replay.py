"""CPU-only: the bodies of the GPU tests that were written after the round's last GPU run (tests/test_z*_gpu.py), replayed with
an oracle-backed stand-in for the sim.

This says NOTHING about the kernels -- the stand-in answers with the oracle's own results, so every product-vs-oracle comparison
is trivially true.  What it does establish, before those files first execute on a B200: the test code itself runs (scene
generator arguments, dictionary keys, attribute names, shapes), and the assertions that depend on the SCENE rather than on the
product hold -- enough teleported contacts, Lees-Edwards kick contacts present, kinematic-object contacts present, bodies really
cross the portals during the trajectories, the unsupported cases really are unsupported.
"""
from types import SimpleNamespace

import numpy as np
import pytest

import scisim_b200 as sb


class _Tele(SimpleNamespace):
    pass


def _active(ref, dim):
    na = ref["type"].shape[0]
    nbb = ref["n_regular"] + ref["portal0"].shape[0]
    return SimpleNamespace(n_candidates=ref["candidates"].shape[0], candidates=ref["candidates"], n_active=na, n_body_body=nbb,
                           type=ref["type"], i=ref["i"], j=ref["j"], aux=ref.get("aux", np.zeros(na, np.uint32)), n=ref["n"], p=ref["p"], depth=ref["depth"],
                           n_plane=int((ref["type"] == 2).sum()), n_drum=int((ref["type"] == 1).sum()))  # ball2d's counters (SG_BALL_PLANE, SG_BALL_DRUM)


def _tele(ref, keys):
    t = _Tele(n_regular=ref["n_regular"], n_teleported=ref["portal0"].shape[0], box_body=ref["box_body"], box_portal=ref["box_portal"],
              portal0=ref["portal0"], portal1=ref["portal1"], x0=ref["x0"], x1=ref["x1"], kick=None, delta0=None, delta1=None)
    for k in keys:
        setattr(t, k, ref[k])
    return t


class OracleBackedSim:
    """The calls the replayed tests make on a sim, answered by the oracle (portals set by the test module's make_oracle)."""

    def __init__(self, oracle_obj, dim, tele_keys, flow_codes):
        self.o, self.dim, self.tele_keys, self.flow_codes = oracle_obj, dim, tele_keys, flow_codes
        self.last = None

    def updatePeriodicBoundaryConditionsStartOfStep(self, it, dt):
        return self.o.update_portals(it * dt)

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        return self.o.flow(self.flow_codes[kind], q0, v0, dt)

    def computeActiveSet(self, q0, q1, v=None, resident=False, **kw):
        ref = self.o.active_set_portals(q0, q1)
        if ref is None or not ref.get("supported", True):
            raise sb.SciSimB200Error("unsupported (replay)")
        self.last = ref
        return _active(ref, self.dim)

    def teleported(self):
        return _tele(self.last, self.tele_keys)

    def enforcePeriodicBoundaryConditions(self, *a):
        return self.o.enforce_portals(*a)

    # resident stepping (rigidbody3d test)
    def upload(self, q, v):
        self.q, self.v = q.copy(), v.copy()

    def step(self, umap, dt):
        self.q1, self.v1 = self._flow(umap.kind, self.q, self.v, dt)
        self.res = self.computeActiveSet(self.q, self.q1)
        return self.res.n_candidates, self.res.n_active

    def fetch(self):
        return self.q1, self.v1, self.res


def _patch(monkeypatch, mod, dim, tele_keys, flow_codes):
    monkeypatch.setattr(mod, "make_sim", lambda s, ctx: OracleBackedSim(mod.make_oracle(s), dim, tele_keys, flow_codes))


def test_replay_rb2d_portal_gpu_tests(monkeypatch, oracle):
    import tests.test_zz_rb2d_portals_gpu as m
    from scisim_b200 import SG_MAP_SYMPLECTIC_EULER, SG_MAP_VERLET
    _patch(monkeypatch, m, 2, ("kick", "delta0", "delta1"), {SG_MAP_SYMPLECTIC_EULER: 0, SG_MAP_VERLET: 1})
    for case in m.CASES:
        m.test_rb2d_portal_active_set_matches_oracle(None, oracle, case)
    m.test_rb2d_portal_flow_resident_and_enforce(None, oracle)
    m.test_rb2d_portal_unsupported_cases(None, oracle)
    m.test_rb2d_portal_trajectory(None, oracle)


def test_replay_rb3d_portal_gpu_tests(monkeypatch, oracle):
    import tests.test_zz_rb3d_portals_gpu as m
    from scisim_b200._lib import SG_MAP_DMV, SG_MAP_SPLIT_HAM
    _patch(monkeypatch, m, 3, (), {SG_MAP_SPLIT_HAM: 2, SG_MAP_DMV: 3})
    for case in m.CASES:
        m.test_rb3d_portal_active_set_matches_oracle(None, oracle, case)
    m.test_rb3d_portal_resident_step_and_enforce(None, oracle)
    m.test_rb3d_portal_trajectory(None, oracle)


def test_replay_ball2d_portal_trajectory(monkeypatch, oracle):
    import tests.test_zy_portal_trajectory_gpu as m
    from scisim_b200 import SG_MAP_SYMPLECTIC_EULER, SG_MAP_VERLET
    import tests.test_portals_gpu as base
    monkeypatch.setattr(m, "make_sim", lambda s, ctx: OracleBackedSim(base.make_oracle(s), 2, ("kick",), {SG_MAP_SYMPLECTIC_EULER: 0, SG_MAP_VERLET: 1}))
    m.test_portal_trajectory_many_steps(None, oracle)


class OracleBackedRB3D:
    """RigidBody3DSim's flow / updateMandMinv / resident stepping answered by the oracle (no portals)."""

    def __init__(self, s):
        from tests import oracle_binding as ob
        self.o, self.m_updated, self.q1 = ob.RB3DOracle(s), False, None

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        self.q1, v1 = self.o.flow(kind, q0, v0, dt, m_updated=self.m_updated)
        return self.q1, v1

    def updateMandMinv(self, q=None):
        if q is None and self.q1 is None:
            raise sb.SciSimB200Error("no configuration (replay)")
        self.m_updated = True
        return self.o.update_m_and_minv(self.q1 if q is None else q)

    def upload(self, q, v):
        self.q, self.v = q.copy(), v.copy()

    def step(self, umap, dt):
        self.res = self._flow(umap.kind, self.q, self.v, dt)

    def fetch(self):
        return self.res[0], self.res[1], None


def test_replay_rb3d_minertia_gpu_tests(monkeypatch, oracle):
    import tests.test_zw_rb3d_minertia_gpu as m
    monkeypatch.setattr(m, "make_sim", lambda s, ctx: OracleBackedRB3D(s))
    for n, seed in ((1, 1), (777, 2), (20000, 3)):
        m.test_update_m_and_minv_matches_oracle(None, oracle, n, seed)
    m.test_flows_after_the_first_read_the_updated_matrix(None, oracle)
    m.test_split_ham_with_the_updated_matrix(None, oracle)
    m.test_update_without_a_configuration_is_an_error(None, oracle)


class OracleBackedRB2D:
    """RigidBody2DSim without portals: flow, active set and the resident step answered by the oracle."""

    def __init__(self, s):
        from tests import oracle_binding as ob
        self.o, self.res = ob.RB2DOracle(s), None

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        return self.o.flow(kind, q0, v0, dt)

    def computeActiveSet(self, q0, q1, v=None, resident=False, **kw):
        ref = self.o.active_set(q0, q1, "grid")
        if not ref["supported"]:
            raise sb.SciSimB200Error("unsupported (replay)")
        return SimpleNamespace(n_candidates=ref["candidates"].shape[0], n_active=ref["type"].shape[0], **{k: ref[k] for k in ("candidates", "type", "i", "j", "aux", "n", "p", "depth")})

    def upload(self, q, v):
        self.q, self.v = q.copy(), v.copy()

    def step(self, umap, dt):
        q1, v1 = self._flow(umap.kind, self.q, self.v, dt)
        self.res = (q1, v1, self.computeActiveSet(self.q, q1))
        return self.res[2].n_candidates, self.res[2].n_active

    def fetch(self):
        if self.res is None:
            raise sb.SciSimB200Error("no step (replay)")
        return self.res


def test_replay_rb2d_resident_gpu_tests(monkeypatch, oracle):
    import tests.test_zx_rb2d_resident_gpu as m
    monkeypatch.setattr(m, "make_sim", lambda s, ctx: OracleBackedRB2D(s))
    m.test_rb2d_resident_step_matches_oracle(None, oracle)


class OracleBackedRB2DWithSnapshots:
    """RigidBody2DSim's resident stepping answered by the oracle; its snapshots written by the product's own writer (scisim_b200/csrc/sg_rb2d_snapshot.h compiled
    for the host, as in tests/test_rb2d_snapshot_cpu.py) from the stand-in's arrays; a restore re-reads ( q, v ) from the bytes and keeps the scene."""

    def __init__(self, s, portals, harness, t=None):
        from tests import oracle_binding as ob
        self.s, self.portals, self.harness, self.t, self.q1 = s, portals, harness, t, None
        self.o = ob.RB2DOracle(s)
        if portals is not None:
            self.o.set_portals(portals)
            if t is not None:
                self.o.update_portals(t)

    def nqdofs(self):
        return 3 * self.s["geo_of_body"].shape[0]

    def upload(self, q, v):
        self.q, self.v, self.q1 = np.array(q, dtype=np.float64).ravel(), np.array(v, dtype=np.float64).ravel(), None

    def updatePeriodicBoundaryConditionsStartOfStep(self, it, dt):
        self.t = it * dt
        return self.o.update_portals(self.t)

    def step(self, umap, dt):
        from scisim_b200 import SG_MAP_SYMPLECTIC_EULER
        self.q1, self.v1 = self.o.flow(0 if umap.kind == SG_MAP_SYMPLECTIC_EULER else 1, self.q, self.v, dt)
        ref = self.o.active_set_portals(self.q, self.q1) if self.portals is not None else self.o.active_set(self.q, self.q1, "grid")
        assert ref["supported"]
        na = ref["type"].shape[0]
        self.res = SimpleNamespace(n_candidates=ref["candidates"].shape[0], n_active=na, type=ref["type"], i=ref["i"], j=ref["j"], aux=ref.get("aux", np.zeros(na, np.uint32)),
                                   n=ref["n"], p=ref["p"], depth=ref["depth"])
        return self.res.n_candidates, na

    def fetch(self):
        return self.q1, self.v1, self.res

    def serializeState(self, which=1):
        from tests.test_rb2d_snapshot_cpu import product_bytes
        if which == 1 and self.q1 is None:
            raise sb.SciSimB200Error("which = 1 without a flow (replay)")
        q, v = (self.q, self.v) if which == 0 else (self.q1, self.v1)
        return product_bytes(self.harness, self.s, q, v, self.portals, t=self.t if self.portals is not None else None)


def test_replay_rb2d_state_io_gpu_tests(monkeypatch, oracle, tmp_path_factory):
    import os
    import tests.test_zzz_rb2d_state_io_gpu as m
    from tests.test_rb2d_snapshot_cpu import harness as harness_fixture
    if not os.path.exists(os.path.join(m.ROOT, "oracle", "_ref", "libref_rb2d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    lib = harness_fixture.__wrapped__(tmp_path_factory)
    made = []

    def make(s, ctx, portals=None):
        made.append(OracleBackedRB2DWithSnapshots(s, portals, lib))
        return made[-1]

    def restore(blob, ctx):
        nq = int(np.frombuffer(blob[:8], dtype=np.int64)[0])
        if len(blob) != len(made[-1].serializeState(which=0)):
            raise sb.SciSimB200Error("truncated (replay)")
        src = made[-1]
        sim = OracleBackedRB2DWithSnapshots(src.s, src.portals, lib, t=src.t)
        sim.upload(np.frombuffer(blob[8:8 + 8 * nq], dtype=np.float64), np.frombuffer(blob[16 + 8 * nq:16 + 16 * nq], dtype=np.float64))
        return sim

    monkeypatch.setattr(m, "_sim", make)
    monkeypatch.setattr(m, "_restore", restore)
    monkeypatch.setattr(m, "_context", lambda: SimpleNamespace(close=lambda: None))
    for scene in ("circles_boxes", "kinematic_circles", "lees_edwards"):
        m.test_rb2d_snapshot_is_the_references_own_and_resumes(oracle, None, scene)
    m.test_rb2d_snapshot_refusals(None)


class OracleBackedRB3DWithSnapshots:
    """RigidBody3DSim's resident stepping answered by the oracle; its snapshots written by the product's own writer (scisim_b200/csrc/sg_rb3d_snapshot.h compiled
    for the host, as in tests/test_rb3d_snapshot_cpu.py) with the mesh records the test attaches; a restore re-reads ( q, v ) from the bytes and keeps the scene."""

    def __init__(self, s, harness, records=None, m_updated=False):
        from tests import oracle_binding as ob
        self.s, self.harness, self.o, self.m_updated, self.q1 = s, harness, ob.RB3DOracle(s), m_updated, None
        self.records = dict(records or {})

    def setMeshSnapshot(self, k, record):
        m = self.s["meshes"][k]
        want = 1 + 8 + 4 + 8 + 24 * m["verts"].shape[0] + 8 + 8 + 24 + 24 + 72 + 8 + 24 * m["samples"].shape[0] + 8 + 24 * m["hull"].shape[0] + 24 + 12 + 24 + 8 + 8 * int(np.prod(m["dims"])) + 24
        if len(record) != want:   # the reference shim names its meshes "shim" ( 4 characters ) and gives them no faces
            raise sb.SciSimB200Error("not a whole record (replay)")
        self.records[k] = record

    def upload(self, q, v):
        self.q, self.v, self.q1 = np.array(q, dtype=np.float64).ravel(), np.array(v, dtype=np.float64).ravel(), None

    def updateMandMinv(self, q=None):
        self.m_updated = True

    def step(self, umap, dt):
        self.q1, self.v1 = self.o.flow(umap.kind, self.q, self.v, dt, m_updated=self.m_updated)
        ref = self.o.active_set(self.q, self.q1, "grid")
        assert ref["supported"]
        na = ref["type"].shape[0]
        self.res = SimpleNamespace(n_candidates=ref["candidates"].shape[0], n_active=na, **{k: ref[k] for k in ("type", "i", "j", "aux", "n", "p", "depth")})
        return self.res.n_candidates, na

    def fetch(self):
        return self.q1, self.v1, self.res

    def serializeState(self, which=1, m_updated=None):
        from tests.test_rb3d_snapshot_cpu import _transposed, product_bytes
        if m_updated is None:
            m_updated = True if which == 1 else self.m_updated
        q, v = (self.q, self.v) if which == 0 else (self.q1, self.v1)
        n = self.s["geo_of_body"].shape[0]
        I, Ii = self.o.update_m_and_minv(q)
        blocks = (I, Ii) if m_updated else (_transposed(I, n), _transposed(Ii, n))
        recs = [self.records.get(int(self.s["geo_mesh"][k]), b"") if int(self.s["geo_type"][k]) == 3 else b"" for k in range(len(self.s["geo_type"]))]
        out = product_bytes(self.harness, self.s, q, v, blocks, mesh_records=recs)
        if out is None:
            raise sb.SciSimB200Error("a mesh without its record (replay)")
        return out


def test_replay_rb3d_mesh_state_io_gpu_tests(monkeypatch, oracle, tmp_path_factory):
    import os
    import tests.test_zzz_rb3d_mesh_state_io_gpu as m
    from tests.test_rb3d_snapshot_cpu import harness as harness_fixture
    if not os.path.exists(os.path.join(m.ROOT, "oracle", "_ref", "libref_rb3d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    lib = harness_fixture.__wrapped__(tmp_path_factory)
    made = []

    def make(s, ctx):
        made.append(OracleBackedRB3DWithSnapshots(s, lib))
        return made[-1]

    def restore(blob, ctx):
        src = made[-1]
        n = int(np.frombuffer(blob[:4], dtype=np.uint32)[0])
        sim = OracleBackedRB3DWithSnapshots(src.s, lib, records=src.records, m_updated=True)
        sim.upload(np.frombuffer(blob[12:12 + 96 * n], dtype=np.float64), np.frombuffer(blob[20 + 96 * n:20 + 144 * n], dtype=np.float64))
        return sim

    monkeypatch.setattr(m, "_sim", make)
    monkeypatch.setattr(m, "_restore", restore)
    monkeypatch.setattr(m, "_context", lambda: SimpleNamespace(close=lambda: None))
    for scene in ("meshes", "mixed"):
        m.test_rb3d_mesh_snapshot_is_the_references_own_and_resumes(oracle, None, scene)
