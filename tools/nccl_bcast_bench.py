"""NCCL broadcast bandwidth at the panel sizes of the large-window factorisation (run under torchrun)."""
import os, json, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
out = {}
for mb in (1, 4, 8, 16, 33, 67):
    t = torch.zeros(mb * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for _ in range(3):
        dist.broadcast(t, src=0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        dist.broadcast(t, src=i % world)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    out[f"{mb}MB"] = {"us": round(us, 1), "GBps": round(mb * 1.048576e-3 / (us * 1e-6), 1)}
if rank == 0:
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
