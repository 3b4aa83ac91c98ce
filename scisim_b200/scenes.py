"""Synthetic scene generators for the configurations of BASELINE.json / SURVEY.md 8d (inputs only -- no physics).

All generators are seeded, FP64, and return plain numpy arrays in SCISim's own layouts.
"""
import numpy as np


def ball2d_lattice(nx=1000, ny=1000, r=0.5, spacing=0.99, jitter=0.002, seed=42, with_planes=True):
    """Config 2: nx*ny equal balls on a square lattice, axis neighbours overlapping by 1 %, under gravity,
    floor + two walls.  P_c ~ 4N (the 8-neighbourhood AABBs overlap), P_a ~ 2N."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nx * ny
    ix, iy = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), indexing="xy")
    q = np.empty((n, 2))
    q[:, 0] = ix.ravel() * spacing
    q[:, 1] = iy.ravel() * spacing
    q += rng.uniform(-jitter, jitter, size=(n, 2))
    scene = {
        "q": q.ravel().copy(), "v": np.zeros(2 * n), "r": np.full(n, r), "m": np.ones(n),
        "g": np.array([0.0, -9.81]), "dt": 1.0e-3, "map": "symplectic_euler",
        "plane_x": np.zeros((0, 2)), "plane_n": np.zeros((0, 2)), "drum_x": np.zeros((0, 2)), "drum_r": np.zeros(0),
    }
    if with_planes:
        scene["plane_x"] = np.array([[0.0, -r], [-r, 0.0], [(nx - 1) * spacing + r, 0.0]])
        scene["plane_n"] = np.array([[0.0, 1.0], [1.0, 0.0], [-1.0, 0.0]])
    return scene


def ball2d_gas(n=1 << 20, rmin=0.25, rmax=1.0, phi=0.55, seeds=(7, 8, 9), dt=1.0e-3, vmax=1.0, with_walls=True):
    """Config 3: polydisperse gas, radii log-uniform in [rmin,rmax], m = pi r^2, uniform centres in a square at
    area fraction phi (overlaps allowed), random velocities, no gravity, 4 walls.  Verlet."""
    r = np.exp(np.random.Generator(np.random.PCG64(seeds[0])).uniform(np.log(rmin), np.log(rmax), size=n))
    er2 = (rmax ** 2 - rmin ** 2) / (2.0 * np.log(rmax / rmin))
    side = np.sqrt(n * np.pi * er2 / phi)
    q = np.random.Generator(np.random.PCG64(seeds[1])).uniform(0.0, side, size=(n, 2))
    v = np.random.Generator(np.random.PCG64(seeds[2])).uniform(-vmax, vmax, size=(n, 2))
    scene = {
        "q": q.ravel().copy(), "v": v.ravel().copy(), "r": r, "m": np.pi * r * r,
        "g": np.array([0.0, 0.0]), "dt": dt, "map": "verlet", "side": side,
        "plane_x": np.zeros((0, 2)), "plane_n": np.zeros((0, 2)), "drum_x": np.zeros((0, 2)), "drum_r": np.zeros(0),
    }
    if with_walls:
        scene["plane_x"] = np.array([[0.0, 0.0], [0.0, 0.0], [side, 0.0], [0.0, side]])
        scene["plane_n"] = np.array([[0.0, 1.0], [1.0, 0.0], [-1.0, 0.0], [0.0, -1.0]])
    return scene


def ball2d_asset(name="different_friction", portal=False, map="symplectic_euler"):
    """BASELINE configs[0]: a scene bundled with the reference (assets/ball2d/...), from the committed fixture
    tests/golden/ball2d_assets.npz (written by tests/golden/make_ball2d_assets.py from the reference's XML; the GPU box has no
    reference tree).  name: "pool_break_ten_deep" (56 balls, no gravity, dt 0.1) or "different_friction" (6 079 balls, gravity
    at 23 degrees, 3 static planes, dt 1/10080).  The integrator is forced to symplectic_euler (every bundled scene says
    verlet; the parser accepts both, ball2dutils/Ball2DSceneParser.cpp:624-636).  portal=True keeps the scene's
    <planar_portal planeA planeB>: those two planes leave the static-plane list and form a planar portal (parser :369-585)."""
    import os
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ball2d_assets.npz"))
    g = lambda k: f["%s/%s" % (name, k)]
    px, pn = g("plane_x").reshape(-1, 2).copy(), g("plane_n").reshape(-1, 2).copy()
    scene = {"q": g("q").copy(), "v": g("v").copy(), "r": g("r").copy(), "m": g("m").copy(), "g": g("g").copy(), "dt": float(g("dt")), "map": map,
             "drum_x": np.zeros((0, 2)), "drum_r": np.zeros(0), "t": 0.0}
    pp = g("portal_planes").reshape(-1, 2)
    if portal and pp.shape[0]:
        used = sorted(set(int(k) for k in pp.ravel()))
        scene["portals"] = {"plane_a_x": px[pp[:, 0]].copy(), "plane_a_n": pn[pp[:, 0]].copy(), "plane_b_x": px[pp[:, 1]].copy(), "plane_b_n": pn[pp[:, 1]].copy(),
                            "v": np.zeros(pp.shape[0]), "bounds": np.zeros(pp.shape[0])}
        keep = [k for k in range(px.shape[0]) if k not in used]
        px, pn = px[keep], pn[keep]
    scene["plane_x"], scene["plane_n"] = np.ascontiguousarray(px), np.ascontiguousarray(pn)
    return scene


def ball2d_random(n, seed, box=None, rmin=0.05, rmax=0.4, nplanes=2, ndrums=1, vmax=40.0, dt=0.01):
    """Small messy scenes for parity tests: overlapping balls, fast movers (tunnelling CCD hits), oblique
    un-normalised planes and a drum."""
    rng = np.random.Generator(np.random.PCG64(seed))
    box = box if box is not None else max(1.0, np.sqrt(n) * 0.5)
    r = rng.uniform(rmin, rmax, size=n)
    q = rng.uniform(-box, box, size=(n, 2))
    v = rng.uniform(-vmax, vmax, size=(n, 2))
    v[rng.uniform(size=n) < 0.3] = 0.0
    scene = {
        "q": q.ravel().copy(), "v": v.ravel().copy(), "r": r, "m": rng.uniform(0.5, 3.0, size=n),
        "g": np.array([0.3, -9.81]), "dt": dt, "map": "symplectic_euler",
        "plane_x": rng.uniform(-box, box, size=(nplanes, 2)), "plane_n": rng.normal(size=(nplanes, 2)) * 3.0,
        "drum_x": rng.uniform(-0.1, 0.1, size=(ndrums, 2)), "drum_r": np.full(ndrums, box * 1.2),
    }
    return scene


def ball2d_periodic(n, seed, side=None, rmin=0.1, rmax=0.35, axes="xy", lees_edwards=0.0, t=0.0, vmax=2.0, dt=0.01, oblique=False):
    """Balls in a periodic box [0, side)^2 with planar portals (ball2d/Portals/PlanarPortal.h).

    axes: which directions are periodic ("x", "y" or "xy"); the others get walls (static planes).  lees_edwards != 0 turns
    the y pair (or the x pair when only x is periodic) into a Lees-Edwards portal with that tangential velocity and bounds
    side / 2 (so the tangential coordinate wraps with the box).  Plane points sit at the middle of each side, normals point
    into the box and are deliberately not unit length.  oblique rotates the whole set-up by 0.3 rad about the box centre so
    that no expression is axis aligned."""
    rng = np.random.Generator(np.random.PCG64(seed))
    side = float(side) if side is not None else max(2.0, np.sqrt(n) * 0.7)
    r = rng.uniform(rmin, rmax, size=n)
    q = rng.uniform(0.0, side, size=(n, 2))
    v = rng.uniform(-vmax, vmax, size=(n, 2))
    h = 0.5 * side
    pairs = {"x": (([0.0, h], [2.0, 0.0]), ([side, h], [-3.0, 0.0])), "y": (([h, 0.0], [0.0, 1.5]), ([h, side], [0.0, -0.5]))}
    pax, pan, pbx, pbn, pv, pb = [], [], [], [], [], []
    plane_x, plane_n = [], []
    le_axis = "y" if "y" in axes else "x"
    for ax in "xy":
        (xa, na), (xb, nb) = pairs[ax]
        if ax in axes:
            pax.append(xa); pan.append(na); pbx.append(xb); pbn.append(nb)
            is_le = lees_edwards != 0.0 and ax == le_axis
            pv.append(lees_edwards if is_le else 0.0)
            pb.append(h if is_le else 0.0)
        else:
            plane_x += [xa, xb]; plane_n += [na, nb]
    arr = lambda a, w: np.array(a, dtype=np.float64).reshape(-1, w)
    portals = {"plane_a_x": arr(pax, 2), "plane_a_n": arr(pan, 2), "plane_b_x": arr(pbx, 2), "plane_b_n": arr(pbn, 2),
               "v": np.array(pv, dtype=np.float64), "bounds": np.array(pb, dtype=np.float64)}
    planes_x, planes_n = arr(plane_x, 2), arr(plane_n, 2)
    if oblique:
        c, s_ = np.cos(0.3), np.sin(0.3)
        R = np.array([[c, -s_], [s_, c]])
        ctr = np.array([h, h])
        rot_p = lambda a: (a - ctr) @ R.T + ctr
        rot_v = lambda a: a @ R.T
        q, v = rot_p(q), rot_v(v)
        for k in ("plane_a_x", "plane_b_x"):
            portals[k] = rot_p(portals[k])
        for k in ("plane_a_n", "plane_b_n"):
            portals[k] = rot_v(portals[k])
        planes_x, planes_n = (rot_p(planes_x), rot_v(planes_n)) if planes_x.shape[0] else (planes_x, planes_n)
    return {
        "q": np.ascontiguousarray(q).ravel().copy(), "v": np.ascontiguousarray(v).ravel().copy(), "r": r, "m": rng.uniform(0.5, 3.0, size=n),
        "g": np.array([0.0, 0.0]), "dt": dt, "map": "symplectic_euler", "side": side, "t": float(t),
        "plane_x": np.ascontiguousarray(planes_x), "plane_n": np.ascontiguousarray(planes_n), "drum_x": np.zeros((0, 2)), "drum_r": np.zeros(0),
        "portals": {k: np.ascontiguousarray(a) for k, a in portals.items()},
    }

# ---- rigidbody3d -------------------------------------------------------------------------------------
def _random_rotations(rng, n):
    """n proper rotation matrices (row-major 9-vectors) from random unit quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 9))
    R[:, 0] = 1 - 2 * (y * y + z * z); R[:, 1] = 2 * (x * y - w * z); R[:, 2] = 2 * (x * z + w * y)
    R[:, 3] = 2 * (x * y + w * z); R[:, 4] = 1 - 2 * (x * x + z * z); R[:, 5] = 2 * (y * z - w * x)
    R[:, 6] = 2 * (x * z - w * y); R[:, 7] = 2 * (y * z + w * x); R[:, 8] = 1 - 2 * (x * x + y * y)
    return R


def _rb3d_pack(x, R, v, w):
    n = x.shape[0]
    return np.concatenate([x.ravel(), R.reshape(n, 9).ravel()]), np.concatenate([v.ravel(), w.ravel()])


def _rb3d_scene(geo_type, geo_r, geo_half, geo_mesh, meshes, geo_of_body, fixed, m, I0, q, v, g, plane_x, plane_n, dt, umap):
    return {"geo_type": np.asarray(geo_type, np.uint32), "geo_r": np.asarray(geo_r, np.float64), "geo_half": np.asarray(geo_half, np.float64).reshape(-1, 3),
            "geo_mesh": np.asarray(geo_mesh, np.uint32), "meshes": meshes, "geo_of_body": np.asarray(geo_of_body, np.uint32),
            "fixed": np.asarray(fixed, np.uint8), "m": np.asarray(m, np.float64), "I0": np.asarray(I0, np.float64).reshape(-1, 3),
            "q": q, "v": v, "g": np.asarray(g, np.float64), "plane_x": np.asarray(plane_x, np.float64).reshape(-1, 3),
            "plane_n": np.asarray(plane_n, np.float64).reshape(-1, 3), "dt": dt, "map": umap}


def rb3d_sphere_lattice(nx=160, ny=160, nz=160, r=0.5, spacing=0.99, jitter=0.002, seed=11, rho=1.0):
    """Config 4: nx*ny*nz equal spheres on a cubic lattice (neighbours along the axes overlap by 1 %), R = I, v = w = 0,
    gravity along -y, floor + 4 walls, split_ham (rigidbody3d has no Verlet map)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nx * ny * nz
    iz, iy, ix = np.meshgrid(np.arange(nz, dtype=np.float64), np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing="ij")
    x = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1) * spacing + rng.uniform(-jitter, jitter, size=(n, 3))
    R = np.tile(np.eye(3).ravel(), (n, 1))
    q, v = _rb3d_pack(x, R, np.zeros((n, 3)), np.zeros((n, 3)))
    mass = rho * 4.0 / 3.0 * np.pi * r ** 3
    I = 0.4 * mass * r * r
    ex, ez = (nx - 1) * spacing + r, (nz - 1) * spacing + r
    plane_x = [[0, -r, 0], [-r, 0, 0], [ex, 0, 0], [0, 0, -r], [0, 0, ez]]
    plane_n = [[0, 1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]]
    return _rb3d_scene([1], [r], [[0, 0, 0]], [0], [], np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.full(n, mass), np.full((n, 3), I),
                       q, v, [0.0, -9.81, 0.0], plane_x, plane_n, 1.0e-3, "split_ham")


def rb3d_random_spheres(n, seed, spin=False, nfixed_frac=0.1, nplanes=2, box=None):
    """Messy all-sphere scenes: several radii, overlapping, some kinematically scripted, optional spin."""
    rng = np.random.Generator(np.random.PCG64(seed))
    box = box if box is not None else max(1.0, n ** (1.0 / 3.0) * 0.45)
    radii = np.array([0.25, 0.4, 0.6])
    gi = rng.integers(0, 3, size=n)
    fixed = (rng.uniform(size=n) < nfixed_frac).astype(np.uint8)
    x = rng.uniform(-box, box, size=(n, 3))
    R = _random_rotations(rng, n)
    v = rng.uniform(-5, 5, size=(n, 3))
    w = rng.uniform(-3, 3, size=(n, 3)) if spin else np.zeros((n, 3))
    v[fixed == 1] = 0.0
    w[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    I0 = 0.4 * m[:, None] * (radii[gi] ** 2)[:, None] * np.ones((1, 3))
    q, vv = _rb3d_pack(x, R, v, w)
    return _rb3d_scene([1, 1, 1], radii, np.zeros((3, 3)), [0, 0, 0], [], gi, fixed, m, I0, q, vv, [0.2, -9.81, 0.1],
                       rng.uniform(-box, box, size=(nplanes, 3)), rng.normal(size=(nplanes, 3)) * 2.0, 0.01, "split_ham")


def rb3d_random_boxes(n, seed, spin=True, nfixed_frac=0.1, nplanes=2, box=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    box = box if box is not None else max(1.0, n ** (1.0 / 3.0) * 0.6)
    ngeo = 4
    half = rng.uniform(0.3, 0.6, size=(ngeo, 3))
    gi = rng.integers(0, ngeo, size=n)
    fixed = (rng.uniform(size=n) < nfixed_frac).astype(np.uint8)
    x = rng.uniform(-box, box, size=(n, 3))
    R = _random_rotations(rng, n)
    v = rng.uniform(-2, 2, size=(n, 3))
    w = rng.uniform(-2, 2, size=(n, 3)) if spin else np.zeros((n, 3))
    v[fixed == 1] = 0.0
    w[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    h = half[gi]
    I0 = m[:, None] / 3.0 * np.stack([h[:, 1] ** 2 + h[:, 2] ** 2, h[:, 0] ** 2 + h[:, 2] ** 2, h[:, 0] ** 2 + h[:, 1] ** 2], axis=1)
    q, vv = _rb3d_pack(x, R, v, w)
    return _rb3d_scene([0] * ngeo, np.zeros(ngeo), half, [0] * ngeo, [], gi, fixed, m, I0, q, vv, [0.0, -9.81, 0.0],
                       rng.uniform(-box, box, size=(nplanes, 3)), rng.normal(size=(nplanes, 3)), 0.005, "dmv")


def torus_mesh(R=1.0, r=0.4, grid=40, nsamples=600, seed=3, pad=0.25):
    """Synthetic stand-in for the reference's dragon.h5 (absent from the mount): an analytic torus
    (axis z) sampled into a signed distance grid, random surface samples, a coarse vertex set and 'hull' vertices
    (the outer ring).  Plain numpy dict in RigidBodyTriangleMesh's field layout."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ext = np.array([R + r + pad, R + r + pad, r + pad])
    origin = -ext
    dims = np.array([grid, grid, max(8, int(grid * ext[2] / ext[0]))], dtype=np.uint32)
    delta = 2.0 * ext / (dims.astype(np.float64) - 1.0)
    gx = origin[0] + delta[0] * np.arange(dims[0])
    gy = origin[1] + delta[1] * np.arange(dims[1])
    gz = origin[2] + delta[2] * np.arange(dims[2])
    Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
    sdf = np.sqrt((np.sqrt(X * X + Y * Y) - R) ** 2 + Z * Z) - r
    u = rng.uniform(0, 2 * np.pi, size=nsamples)
    t = rng.uniform(0, 2 * np.pi, size=nsamples)
    samples = np.stack([(R + r * np.cos(t)) * np.cos(u), (R + r * np.cos(t)) * np.sin(u), r * np.sin(t)], axis=1)
    uu, tt = np.meshgrid(np.linspace(0, 2 * np.pi, 24, endpoint=False), np.linspace(0, 2 * np.pi, 12, endpoint=False))
    verts = np.stack([(R + r * np.cos(tt)) * np.cos(uu), (R + r * np.cos(tt)) * np.sin(uu), r * np.sin(tt)], axis=2).reshape(-1, 3)
    keep = np.cos(tt).ravel() > 0.3
    hull = verts[keep]
    return {"verts": verts, "samples": samples, "hull": hull, "cell_delta": delta, "dims": dims, "origin": origin, "sdf": sdf.ravel()}


def rb3d_random_meshes(n, seed, spin=True, nfixed_frac=0.1, nplanes=1, box=None, grid=(36, 30), nsamples=(500, 400)):
    rng = np.random.Generator(np.random.PCG64(seed))
    box = box if box is not None else max(1.5, n ** (1.0 / 3.0) * 1.1)
    meshes = [torus_mesh(1.0, 0.4, grid[0], nsamples[0], seed=3), torus_mesh(0.8, 0.3, grid[1], nsamples[1], seed=4)]
    gi = rng.integers(0, 2, size=n)
    fixed = (rng.uniform(size=n) < nfixed_frac).astype(np.uint8)
    x = rng.uniform(-box, box, size=(n, 3))
    R = _random_rotations(rng, n)
    v = rng.uniform(-1, 1, size=(n, 3))
    w = rng.uniform(-1, 1, size=(n, 3)) if spin else np.zeros((n, 3))
    v[fixed == 1] = 0.0
    w[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    I0 = m[:, None] * rng.uniform(0.2, 0.6, size=(n, 3))
    q, vv = _rb3d_pack(x, R, v, w)
    return _rb3d_scene([3, 3], [0, 0], np.zeros((2, 3)), [0, 1], meshes, gi, fixed, m, I0, q, vv, [0.0, -9.81, 0.0],
                       rng.uniform(-box, box, size=(nplanes, 3)), rng.normal(size=(nplanes, 3)), 1.0 / 10800.0, "dmv")


def rb3d_mixed_segregated(nper=200, seed=21):
    """Config 5 in miniature: spheres, boxes and meshes in three spatially separated blocks (mixed-type pairs make
    the reference exit)."""
    a = rb3d_random_spheres(nper, seed, spin=False, nplanes=0)
    b = rb3d_random_boxes(nper, seed + 1, nplanes=0)
    c = rb3d_random_meshes(max(8, nper // 8), seed + 2, nplanes=0)
    scenes_ = [a, b, c]
    shift = [np.array([0.0, 0.0, 0.0]), np.array([60.0, 0.0, 0.0]), np.array([0.0, 0.0, 80.0])]
    xs, Rs, vs, ws, gi, fixed, m, I0 = [], [], [], [], [], [], [], []
    geo_type, geo_r, geo_half, geo_mesh, meshes = [], [], [], [], []
    for s, sh in zip(scenes_, shift):
        n = s["geo_of_body"].shape[0]
        xs.append(s["q"][:3 * n].reshape(n, 3) + sh); Rs.append(s["q"][3 * n:].reshape(n, 9))
        vs.append(s["v"][:3 * n].reshape(n, 3)); ws.append(s["v"][3 * n:].reshape(n, 3))
        gi.append(s["geo_of_body"] + len(geo_type)); fixed.append(s["fixed"]); m.append(s["m"]); I0.append(s["I0"])
        geo_type += list(s["geo_type"]); geo_r += list(s["geo_r"]); geo_half += list(s["geo_half"])
        geo_mesh += [int(k) + len(meshes) for k in s["geo_mesh"]]
        meshes += s["meshes"]
    q, v = _rb3d_pack(np.vstack(xs), np.vstack(Rs), np.vstack(vs), np.vstack(ws))
    return _rb3d_scene(geo_type, geo_r, geo_half, geo_mesh, meshes, np.concatenate(gi), np.concatenate(fixed), np.concatenate(m), np.vstack(I0),
                       q, v, [0.0, -9.81, 0.0], [[0.0, -6.0, 0.0]], [[0.0, 1.0, 0.0]], 1.0 / 10800.0, "dmv")


# ---- rigidbody2d -------------------------------------------------------------------------------------
def rb2d_random(n, seed, kinds=("circle", "box"), nfixed_frac=0.08, nplanes=2, box=None, vmax=8.0, spin=True):
    """Messy rigidbody2d scenes: circles (several radii) and / or rotated boxes, [x,y,theta] DoFs, oblique planes.
    Kinematic bodies are circles only, and kinematic circles are kept away from boxes' reach by giving boxes no fixed
    flag (the reference exits on kinematic box-box and on a kinematic circle meeting a box)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    box_ = box if box is not None else max(1.5, np.sqrt(n) * 0.55)
    geo_type, geo_r, geo_half = [], [], []
    if "circle" in kinds:
        for r in (0.2, 0.35, 0.5):
            geo_type.append(0); geo_r.append(r); geo_half.append([0.0, 0.0])
    if "box" in kinds:
        for h in ((0.3, 0.2), (0.45, 0.45), (0.6, 0.25)):
            geo_type.append(1); geo_r.append(0.0); geo_half.append(list(h))
    geo_type = np.array(geo_type, np.uint32)
    gi = rng.integers(0, geo_type.shape[0], size=n)
    is_circle = geo_type[gi] == 0
    fixed = ((rng.uniform(size=n) < nfixed_frac) & is_circle & (len(kinds) == 1)).astype(np.uint8)
    q = np.empty((n, 3))
    q[:, :2] = rng.uniform(-box_, box_, size=(n, 2))
    q[:, 2] = rng.uniform(-np.pi, np.pi, size=n)
    v = np.empty((n, 3))
    v[:, :2] = rng.uniform(-vmax, vmax, size=(n, 2))
    v[:, 2] = rng.uniform(-3, 3, size=n) if spin else 0.0
    v[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    inertia = m * rng.uniform(0.05, 0.3, size=n)
    M = np.stack([m, m, inertia], axis=1).ravel()
    pn = rng.normal(size=(nplanes, 2))
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    return {"geo_type": geo_type, "geo_r": np.array(geo_r), "geo_half": np.array(geo_half).reshape(-1, 2), "geo_of_body": gi.astype(np.uint32),
            "fixed": fixed, "M": M, "q": q.ravel().copy(), "v": v.ravel().copy(), "g": np.array([0.0, -9.81]),
            "plane_x": rng.uniform(-box_, box_, size=(nplanes, 2)), "plane_n": pn, "dt": 0.01, "map": "symplectic_euler"}


def rb2d_periodic(n, seed, side=None, axes="xy", lees_edwards=0.0, t=0.0, oblique=False, boxes=False, nfixed_frac=0.0, vmax=2.0, dt=0.01):
    """rigidbody2d in a periodic box [0, side)^2 with planar portals (rigidbody2d/PlanarPortal.h): circles everywhere,
    optionally rotated boxes in the middle of the domain (a box reaching a portal makes the reference exit).  Portal
    layout as ball2d_periodic, but with unit normals: RigidBody2DStaticPlane does not normalise.  nfixed_frac marks some
    interior circles as kinematically scripted (a kinematic body in a teleported collision makes the reference exit)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    side = float(side) if side is not None else max(3.0, np.sqrt(n) * 0.9)
    h = 0.5 * side
    geo_type, geo_r, geo_half = [0, 0, 0], [0.2, 0.35, 0.5], [[0.0, 0.0]] * 3
    if boxes:
        geo_type += [1, 1]; geo_r += [0.0, 0.0]; geo_half += [[0.3, 0.2], [0.45, 0.3]]
    geo_type = np.array(geo_type, np.uint32)
    gi = rng.integers(0, geo_type.shape[0], size=n)
    is_box = geo_type[gi] == 1
    q = np.empty((n, 3))
    q[:, :2] = rng.uniform(0.0, side, size=(n, 2))
    inner = rng.uniform(0.3 * side, 0.7 * side, size=(n, 2))
    q[is_box, :2] = inner[is_box]
    q[:, 2] = rng.uniform(-np.pi, np.pi, size=n)
    v = np.empty((n, 3))
    v[:, :2] = rng.uniform(-vmax, vmax, size=(n, 2))
    v[:, 2] = rng.uniform(-3, 3, size=n)
    interior = np.all(np.abs(q[:, :2] - h) < 0.3 * side, axis=1)
    fixed = ((rng.uniform(size=n) < nfixed_frac) & ~is_box & interior & (not boxes)).astype(np.uint8)
    v[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    M = np.stack([m, m, m * rng.uniform(0.05, 0.3, size=n)], axis=1).ravel()
    pairs = {"x": (([0.0, h], [1.0, 0.0]), ([side, h], [-1.0, 0.0])), "y": (([h, 0.0], [0.0, 1.0]), ([h, side], [0.0, -1.0]))}
    pax, pan, pbx, pbn, pv, pb, plane_x, plane_n = [], [], [], [], [], [], [], []
    le_axis = "y" if "y" in axes else "x"
    for ax in "xy":
        (xa, na), (xb, nb) = pairs[ax]
        if ax in axes:
            pax.append(xa); pan.append(na); pbx.append(xb); pbn.append(nb)
            is_le = lees_edwards != 0.0 and ax == le_axis
            pv.append(lees_edwards if is_le else 0.0)
            pb.append(h if is_le else 0.0)
        else:
            plane_x += [xa, xb]; plane_n += [na, nb]
    arr = lambda a, w: np.array(a, dtype=np.float64).reshape(-1, w)
    portals = {"plane_a_x": arr(pax, 2), "plane_a_n": arr(pan, 2), "plane_b_x": arr(pbx, 2), "plane_b_n": arr(pbn, 2),
               "v": np.array(pv, dtype=np.float64), "bounds": np.array(pb, dtype=np.float64)}
    planes_x, planes_n = arr(plane_x, 2), arr(plane_n, 2)
    if oblique:
        c, s_ = np.cos(0.3), np.sin(0.3)
        R = np.array([[c, -s_], [s_, c]])
        ctr = np.array([h, h])
        rot_p = lambda a: (a - ctr) @ R.T + ctr
        q[:, :2] = rot_p(q[:, :2]); q[:, 2] += 0.3
        v[:, :2] = v[:, :2] @ R.T
        for k in ("plane_a_x", "plane_b_x"):
            portals[k] = rot_p(portals[k])
        for k in ("plane_a_n", "plane_b_n"):
            portals[k] = portals[k] @ R.T
        if planes_x.shape[0]:
            planes_x, planes_n = rot_p(planes_x), planes_n @ R.T
    return {"geo_type": geo_type, "geo_r": np.array(geo_r), "geo_half": np.array(geo_half, dtype=np.float64).reshape(-1, 2), "geo_of_body": gi.astype(np.uint32),
            "fixed": fixed, "M": M, "q": q.ravel().copy(), "v": v.ravel().copy(), "g": np.array([0.0, 0.0]),
            "plane_x": np.ascontiguousarray(planes_x), "plane_n": np.ascontiguousarray(planes_n), "dt": dt, "map": "symplectic_euler", "side": side, "t": float(t),
            "portals": {k: np.ascontiguousarray(a) for k, a in portals.items()}}


def rb3d_periodic_spheres(n, seed, side=None, axes="xz", nfixed_frac=0.0, tilt=False, mult=None):
    """Spheres in a box [0, side)^3 with planar portals (rigidbody3d/Portals/PlanarPortal.h) along the axes named in `axes`
    (walls = static planes along the others).  Plane points sit in the middle of each face, normals point into the box and
    are not unit length; tilt rotates everything about z so that no plane is axis aligned (and none faces -y, where Eigen's
    FromTwoVectors takes its SVD branch).  nfixed_frac marks spheres as kinematically scripted.  mult: the integer portal
    multipliers; None picks, per axis, the signs that make the pair of plane frames a pure translation (the tangents of two
    opposite planes, both built by FromTwoVectors( UnitY, n ), mirror one tangential coordinate otherwise)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    side = float(side) if side is not None else max(3.0, n ** (1.0 / 3.0) * 0.9)
    translation = {"x": (1, 1, -1), "y": (1, 1, 1), "z": (1, -1, 1)}
    h = 0.5 * side
    radii = np.array([0.25, 0.4, 0.6])
    gi = rng.integers(0, 3, size=n)
    fixed = (rng.uniform(size=n) < nfixed_frac).astype(np.uint8)
    x = rng.uniform(0.0, side, size=(n, 3))
    R = _random_rotations(rng, n)
    v = rng.uniform(-2, 2, size=(n, 3))
    w = np.zeros((n, 3))
    v[fixed == 1] = 0.0
    m = rng.uniform(0.5, 2.0, size=n)
    I0 = 0.4 * m[:, None] * (radii[gi] ** 2)[:, None] * np.ones((1, 3))
    faces = {"x": (([0.0, h, h], [2.0, 0.0, 0.0]), ([side, h, h], [-3.0, 0.0, 0.0])),
             "y": (([h, 0.0, h], [0.0, 1.5, 0.0]), ([h, side, h], [0.0, -0.5, 0.0])),
             "z": (([h, h, 0.0], [0.0, 0.0, 1.0]), ([h, h, side], [0.0, 0.0, -2.0]))}
    pax, pan, pbx, pbn, pm, plane_x, plane_n = [], [], [], [], [], [], []
    for ax in "xyz":
        (xa, na), (xb, nb) = faces[ax]
        if ax in axes:
            pax.append(xa); pan.append(na); pbx.append(xb); pbn.append(nb); pm.append(mult if mult is not None else translation[ax])
        else:
            plane_x += [xa, xb]; plane_n += [na, nb]
    arr = lambda a: np.array(a, dtype=np.float64).reshape(-1, 3)
    portals = {"plane_a_x": arr(pax), "plane_a_n": arr(pan), "plane_b_x": arr(pbx), "plane_b_n": arr(pbn),
               "mult": np.array(pm, dtype=np.int32).reshape(-1, 3)}
    planes_x, planes_n = arr(plane_x), arr(plane_n)
    if tilt:
        c, s_ = np.cos(0.35), np.sin(0.35)
        Rz = np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]])
        ctr = np.array([h, h, h])
        rot_p = lambda a: (a - ctr) @ Rz.T + ctr
        x = rot_p(x); v = v @ Rz.T
        for k in ("plane_a_x", "plane_b_x"):
            portals[k] = rot_p(portals[k])
        for k in ("plane_a_n", "plane_b_n"):
            portals[k] = portals[k] @ Rz.T
        if planes_x.shape[0]:
            planes_x, planes_n = rot_p(planes_x), planes_n @ Rz.T
    q, vv = _rb3d_pack(x, R, v, w)
    s = _rb3d_scene([1, 1, 1], radii, np.zeros((3, 3)), [0, 0, 0], [], gi, fixed, m, I0, q, vv, [0.0, 0.0, 0.0], planes_x, planes_n, 0.01, "split_ham")
    s["portals"] = {k: np.ascontiguousarray(a) for k, a in portals.items()}
    s["side"] = side
    return s
