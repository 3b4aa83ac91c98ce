#!/usr/bin/env python
"""Regenerates the golden fixtures under tests/golden/ from the read-only reference checkout.

Run in the build container (needs /root/reference); the GPU box only ever sees the committed outputs.

  ccd_cases.json      the 11 known-answer ball-ball CCD cases of scisimtests/narrowphase_tests.cpp:7-455
                      (inputs, expected quadratic coefficients where the test states them, hit/miss, and the
                      expected time of impact with the tolerance the reference test uses)
  aabb_fixtures.npz   the 3 AABB sets (5 000 / 10 000 / 20 000 boxes, minx,miny,maxx,maxy per box) embedded as
                      C array literals at ball2dtests/collision_detection_tests.cpp:12,14,16; the expected pair
                      set is, as in the reference test, whatever brute force over all pairs gives.
"""
import json
import math
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_ccd():
    src = open(os.path.join(REF, "scisimtests/narrowphase_tests.cpp")).read()
    bodies = re.split(r"static int executeCCDTest(\d\d)\(\)", src)[1:]
    cases = []
    # expected times of impact that the reference writes as C++ expressions
    toi = {
        "03": (0.0, 0.0),
        "04": ((2.0 / 505.0) * (310.0 - math.sqrt(4190.0)), 1.0e-9),
        "05": ((2.0 * (550.0 - math.sqrt(22766.0))) / 1537.0, 1.0e-9),
        "06": (0.5, 1.0e-9),
        "08": (0.0, 1.0e-9),
        "09": (0.677219, 1.0e-7),
    }
    for num, body in zip(bodies[0::2], bodies[1::2]):
        body = body.split("\n}\n")[0]

        def vec(name):
            m = re.search(r"Vector2s %s\{\s*([-0-9.e]+),\s*([-0-9.e]+)\s*\}" % name, body)
            return [float(m.group(1)), float(m.group(2))]

        def sc(name):
            return float(re.search(r"scalar %s\{\s*([-0-9.e]+)\s*\}" % name, body).group(1))

        case = {"name": "ball_ball_ccd_" + num, "q0a": vec("q0a"), "q1a": vec("q1a"), "ra": sc("ra"),
                "q0b": vec("q0b"), "q1b": vec("q1b"), "rb": sc("rb")}
        m = re.search(r"cexpected\{\s*([-0-9.e]+),\s*([-0-9.e]+),\s*([-0-9.e]+)\s*\}", body)
        case["cexpected"] = [float(x) for x in m.groups()] if m else None
        case["coeff_tol"] = 1.0e-9
        # "if( !collision_xxx.first )" => a hit is expected; "if( collision_xxx.first )" => a miss is expected
        case["hit"] = re.search(r"if\( !collision_\w+\.first \)", body) is not None
        if num in toi:
            assert case["hit"]
            case["toi"], case["toi_tol"] = toi[num]
        cases.append(case)
    assert len(cases) == 11
    json.dump(cases, open(os.path.join(HERE, "ccd_cases.json"), "w"), indent=1)
    print("ccd_cases.json:", len(cases), "cases,", sum(c["hit"] for c in cases), "hits")


def parse_aabbs():
    out = {}
    with open(os.path.join(REF, "ball2dtests/collision_detection_tests.cpp")) as f:
        for line in f:
            m = re.match(r"const scalar (test_\d\d)_data\[\] = \{(.*)\};", line)
            if m:
                vals = np.array([float(x) for x in m.group(2).split(",")], dtype=np.float64)
                assert vals.size % 4 == 0
                out["spatial_grid_" + m.group(1)[-2:]] = vals.reshape(-1, 4)
    assert sorted(v.shape[0] for v in out.values()) == [5000, 10000, 20000]
    np.savez_compressed(os.path.join(HERE, "aabb_fixtures.npz"), **out)
    print("aabb_fixtures.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    parse_ccd()
    parse_aabbs()
