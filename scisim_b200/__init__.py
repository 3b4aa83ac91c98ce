"""scisim_b200 -- B200-native collision-detection + unconstrained-flow back end for SCISim.

The product is libscisim_b200.so (hand-written sm_100a CUDA behind the C ABI in include/scisim_b200.h).
This package is the thin host-side mirror of the reference's interfaces for that path, used by the
tests and bench.py; the C++14 shim for SCISim itself is in scisim_b200/host/ (see INTEGRATION.md).
"""
from ._lib import (SG_BALL_BALL, SG_BALL_BALL_KICK_TELEPORTED, SG_BALL_BALL_TELEPORTED, SG_BALL_DRUM, SG_BALL_PLANE, SG_MAP_SYMPLECTIC_EULER, SG_MAP_VERLET, SG_OUT_ALL,
                   SG_OUT_CANDIDATES, SG_OUT_DEPTHS, SG_OUT_NORMALS, SG_OUT_POINTS, SciSimB200Error, load)
from .host_api import (ActiveSet, Ball2DSim, MultiGpuBall2DSim, Ball2DState, Context, DMVMap, ExponentialEulerMap, PlanarPortal, PlanarPortal3D, RigidBody2DSim, RigidBody2DState, RigidBody3DSim, RigidBody3DState, SpatialGridDetector, SplitHamMap,
                       SymplecticEulerMap, TriangleMesh, VerletMap)

__all__ = ["ActiveSet", "Ball2DSim", "MultiGpuBall2DSim", "Ball2DState", "Context", "DMVMap", "ExponentialEulerMap", "RigidBody2DSim", "RigidBody2DState", "RigidBody3DSim", "RigidBody3DState", "SplitHamMap", "TriangleMesh", "SpatialGridDetector", "SymplecticEulerMap", "VerletMap",
           "SciSimB200Error", "load", "SG_OUT_ALL", "SG_OUT_CANDIDATES", "SG_OUT_DEPTHS", "SG_OUT_NORMALS", "SG_OUT_POINTS",
           "SG_BALL_BALL", "SG_BALL_BALL_TELEPORTED", "SG_BALL_BALL_KICK_TELEPORTED", "SG_BALL_DRUM", "SG_BALL_PLANE", "PlanarPortal", "PlanarPortal3D", "SG_MAP_SYMPLECTIC_EULER", "SG_MAP_VERLET"]
