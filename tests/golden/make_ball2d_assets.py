"""Generates tests/golden/ball2d_assets.npz from the reference's own bundled ball2d scenes (BASELINE.json configs[0]):
    python tests/golden/make_ball2d_assets.py            (needs /root/reference; the .npz is committed, the GPU box has no reference tree)
The <ball>, <static_plane>, <gravity>, <integrator dt> and <planar_portal> nodes are read with regular expressions, the way
ball2dutils/Ball2DSceneParser.cpp reads them (:240-290 balls, :340-367 planes, :369-585 portals: the two planes of a portal leave the
static-plane list, :590-640 integrator with a rational dt, gravity node).  Numbers go through Python's float(), which rounds decimal
strings exactly like the strtod behind the reference's StringUtilities::extractFromString."""
import os
import re
import sys

import numpy as np

REF = "/root/reference/assets/ball2d"
SCENES = {"pool_break_ten_deep": "examples_gr/pool_break_ten_deep.xml", "different_friction": "tests_python_serialization/different_friction.xml"}


def attr(node, name):
    m = re.search(r'\b%s="([^"]*)"' % name, node)
    return m.group(1) if m else None


def rational(s):
    if "/" in s:
        a, b = s.split("/")
        return float(int(a)) / float(int(b))      # Rational::operator double: numerator / denominator (scisim/Math/Rational.h:83-86)
    return float(s)


def parse(path):
    txt = open(path).read()
    txt = re.sub(r"<!--.*?-->", "", txt, flags=re.S)
    balls = re.findall(r"<ball\s[^>]*>", txt)
    q = np.array([[float(attr(b, "x")), float(attr(b, "y"))] for b in balls])
    v = np.array([[float(attr(b, "vx")), float(attr(b, "vy"))] for b in balls])
    m = np.array([float(attr(b, "m")) for b in balls])
    r = np.array([float(attr(b, "r")) for b in balls])
    fixed = np.array([int(attr(b, "fixed")) for b in balls], dtype=np.uint8)
    planes = re.findall(r"<static_plane\s[^>]*>", txt)
    px = np.array([[float(t) for t in attr(p, "x").split()] for p in planes]).reshape(-1, 2)
    pn = np.array([[float(t) for t in attr(p, "n").split()] for p in planes]).reshape(-1, 2)
    g = re.search(r"<gravity\s[^>]*>", txt)
    grav = np.array([float(attr(g.group(0), "fx")), float(attr(g.group(0), "fy"))]) if g else np.zeros(2)
    integ = re.search(r"<integrator\s[^>]*>", txt).group(0)
    portals = [(int(attr(p, "planeA")), int(attr(p, "planeB"))) for p in re.findall(r"<planar_portal\s[^>]*>", txt)]
    return {"q": q.ravel(), "v": v.ravel(), "m": m, "r": r, "fixed": fixed, "plane_x": px, "plane_n": pn, "g": grav, "dt": np.array(rational(attr(integ, "dt"))),
            "integrator": np.array(attr(integ, "type")), "portal_planes": np.array(portals, dtype=np.int64).reshape(-1, 2)}


def main():
    out = {}
    for name, rel in SCENES.items():
        s = parse(os.path.join(REF, rel))
        for k, a in s.items():
            out["%s/%s" % (name, k)] = a
        print(name, "balls", s["r"].shape[0], "planes", s["plane_x"].shape[0], "portals", s["portal_planes"].shape[0], "dt", float(s["dt"]), "g", s["g"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ball2d_assets.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
