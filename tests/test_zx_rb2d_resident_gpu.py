"""GPU parity (pytest -m gpu): SG_IN_RESIDENT for rigidbody2d -- computeActiveSet on the device copies the last flow() left.

Added after the round's last full GPU run (the same flag is verified for ball2d and rigidbody3d in their own files); kept in a
file that sorts behind the executed ones so that `pytest -x` reaches every verified test first.
"""
import numpy as np
import pytest

from scisim_b200 import scenes
from tests.test_rb2d_gpu import make_sim

pytestmark = pytest.mark.gpu


def test_rb2d_active_set_on_resident_flow_result(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb2d_random(3000, 9, kinds=("circle", "box"))
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"], resident=True)
    q1, v1 = sim._flow(0, s["q"], s["v"], s["dt"])
    a = sim.computeActiveSet(s["q"], q1, resident=True)
    b = sim.computeActiveSet(s["q"], q1)
    assert a.n_active == b.n_active > 0 and a.n_candidates == b.n_candidates
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
