"""The drop-in at the reference's own seam, end to end (tests/scisim_plugin_example.cpp): the host shim compiled against the reference's OWN headers
(SCISIM_B200_WITH_SCISIM: GpuSymplecticEulerMap / GpuVerletMap are UnconstrainedMaps of the reference) and linked with the reference's own Ball2DSim
(oracle/_ref/libref_ball2d.so).  CPU: it builds, links, and refuses to run without a GPU.  GPU: the reference's own Ball2DSim::flow( call_back, iteration, dt, umap )
steps with the GPU map plugged in and ends in the same bits as with the reference's own map; the reference's own computeActiveSet equals the shim's, constraint
by constraint (type, indices, normal, point)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "scisim_plugin_example")


def _build_if_possible():
    if os.path.isdir("/root/reference/ball2d"):
        from scisim_b200 import build
        build.build_library()
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "../tests/_build/scisim_plugin_example"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if not os.path.exists(EXE):
        pytest.skip("tests/_build/scisim_plugin_example not built (needs the reference tree; oracle/Makefile.ref)")


def test_plugin_example_builds_against_the_reference_headers_and_fails_loudly_without_a_gpu():
    _build_if_possible()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the run itself is the gpu test below")
    except ImportError:
        pass
    out = subprocess.run([EXE, "100", "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode != 0 and "no CPU fallback" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["3000", "4"], ["2000", "3", "verlet"]])
def test_reference_sim_steps_with_the_gpu_map_plugged_in(args):
    if not os.path.exists(EXE):
        pytest.skip("tests/_build/scisim_plugin_example not built (needs the reference tree; oracle/Makefile.ref)")
    out = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and out.stdout.startswith("plugin ok"), out.stdout
