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
