import os, sys, time
os.environ["NCCL_DEBUG"]="WARN"
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
import bench
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); lr=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import scisim_b200 as sb
from scisim_b200.slab import Ball2DSlabs, GpuSlabBackend
scene=bench.scene_for_rank(rank, world); n=scene["r"].shape[0]
ctx=sb.Context(lr)
b=GpuSlabBackend(ctx, scene, rank*n, ghost_cap=8192)
s=Ball2DSlabs(b, rank, world, dist)
for _ in range(3): s.step(0, scene["dt"])
for flush in (False, True):
    ms=[]; wall=[]
    for _ in range(8):
        if flush: ctx.flush_l2()
        t0=time.perf_counter(); ctx.timer_begin(); s.step(0, scene["dt"]); ms.append(ctx.timer_end()); wall.append((time.perf_counter()-t0)*1e3)
    print("rank",rank,"flush",flush,"event ms",[round(x,2) for x in ms],"wall ms",[round(x,2) for x in wall], "halo", s.last_halo, flush=True)
# phase timing (host-synchronised)
def t(f):
    ctx.synchronize(); torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); ctx.synchronize(); torch.cuda.synchronize(); return (time.perf_counter()-t0)*1e3, r
with b.run_on_stream():
    a,iv=t(lambda: b.flow(0, scene["dt"]))
    all_iv=torch.empty(world*2, dtype=torch.float64, device=iv.device)
    c,_=t(lambda: dist.all_gather_into_tensor(all_iv, iv))
    all_iv=all_iv.view(world,2)
    peer=1-rank; side=1 if rank==0 else 0
    d,buf=t(lambda: b.pack(all_iv[peer], side))
    def ex():
        ops=[dist.P2POp(dist.isend, buf, peer), dist.P2POp(dist.irecv, b.recv_buffer(side), peer)]
        for r in dist.batch_isend_irecv(ops): r.wait()
    e,_=t(ex)
    f,_=t(lambda: b.unpack(side, b.recv_buffer(side)))
    g,_=t(lambda: b.detect())
print("rank",rank,"phases ms: flow %.3f allgather %.3f pack %.3f exchange %.3f unpack %.3f detect %.3f"%(a,c,d,e,f,g), flush=True)
dist.destroy_process_group()
