"""CPU: the multi-core ball2d step (oracle/ball2d_parallel.h -- NOT reference behaviour, the optional second CPU baseline of
SURVEY.md 8d) returns exactly what the single-threaded restatement of the reference returns: q1, v1, the candidate list in
std::set order, the ball-ball contact list, the number of drum / plane contacts -- for any thread count."""
import os

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob


@pytest.mark.parametrize("threads", ["1", "3", "8"])
@pytest.mark.parametrize("maker,kind", [(lambda: scenes.ball2d_random(4000, 5, nplanes=3, ndrums=2), 0), (lambda: scenes.ball2d_lattice(70, 50), 0),
                                         (lambda: scenes.ball2d_gas(n=6000), 1), (lambda: scenes.ball2d_random(1, 1), 0), (lambda: scenes.ball2d_random(2, 2), 1)],
                         ids=["random", "lattice", "gas", "one", "two"])
def test_parallel_step_equals_restatement(oracle, maker, kind, threads):
    s = maker()
    o = ob.Ball2DOracle(s)
    old = os.environ.get("ORC_THREADS")
    os.environ["ORC_THREADS"] = threads
    try:
        par = o.parallel_step(kind, s["q"], s["v"], s["dt"])
        cnt = o.parallel_step(kind, s["q"], s["v"], s["dt"], keep_lists=False)
    finally:
        if old is None:
            del os.environ["ORC_THREADS"]
        else:
            os.environ["ORC_THREADS"] = old
    assert par["threads"] == int(threads)
    q1, v1 = o.flow(kind, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert np.array_equal(par["q1"], q1) and np.array_equal(par["v1"], v1)
    assert np.array_equal(par["candidates"], ref["candidates"])
    bb = ref["type"] == 0
    assert np.array_equal(par["active"], np.stack([ref["i"][bb], ref["j"][bb]], axis=1))
    assert par["n_static"] == int((~bb).sum())
    assert (cnt["n_candidates"], cnt["n_active"], cnt["n_static"]) == (par["n_candidates"], par["n_active"], par["n_static"])
    assert par["n_candidates"] == ref["candidates"].shape[0] and par["n_active"] == int(bb.sum())
