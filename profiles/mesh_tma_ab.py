"""A/B timing of the mesh narrow phase with and without TMA-staged SDF bricks (run under gpurun):
   python profiles/mesh_tma_ab.py [n_meshes]   -- runs itself twice, second time with SG_RB3D_NO_TMA=1."""
import json, os, subprocess, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(n, grid):
    import scisim_b200 as sb
    from scisim_b200 import scenes
    from tests.test_rb3d_gpu import make_sim
    s = scenes.rb3d_random_meshes(n, 5, grid=(grid, grid - 6), nsamples=(2000, 1500))
    ctx = sb.Context(0)
    sim = make_sim(s, ctx)
    sim.upload(s["q"], s["v"])
    for _ in range(3):
        sim.step(sb.DMVMap(), s["dt"])
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(10):
        sim.step(sb.DMVMap(), s["dt"])
    prof = ctx.profile()
    ctx.profile_enable(False)
    st = sim.meshStats()
    pc, pa = sim.step(sb.DMVMap(), s["dt"])
    return {"n": n, "candidates": int(pc), "active": int(pa), "staged_sweeps": st[0], "direct_sweeps": st[1],
            "us_per_launch": {k: round(1e3 * v[1] / max(v[0], 1), 2) for k, v in prof.items() if k.startswith("rb3d_mesh")}}

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else 36
    if os.environ.get("_MESH_AB_CHILD"):
        print(json.dumps(run(n, grid)))
    else:
        for no_tma, tiles in (("1", "1"), ("1", "12"), ("0", "6"), ("0", "12"), ("0", "24"), ("0", "48")):
            env = dict(os.environ, _MESH_AB_CHILD="1", SG_RB3D_NO_TMA=no_tma, SG_RB3D_TILES=tiles)
            out = subprocess.run([sys.executable, __file__, str(n), str(grid)], env=env, capture_output=True, text=True)
            print("SG_RB3D_NO_TMA=%s SG_RB3D_TILES=%s" % (no_tma, tiles), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:])
