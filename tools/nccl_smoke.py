import os, sys, time
import torch, torch.distributed as td
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
t0 = time.time()
td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
print(rank, "init done", time.time() - t0, flush=True)
x = torch.ones(62001, dtype=torch.float64, device="cuda") * (rank + 1)
td.all_reduce(x)
torch.cuda.synchronize()
print(rank, "allreduce ok", x[0].item(), time.time() - t0, flush=True)
for n in (62001,):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5): td.all_reduce(x)
    a.record()
    for _ in range(50): td.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    print(rank, "allreduce %d doubles: %.1f us" % (n, a.elapsed_time(b) * 1000 / 50), flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    td.all_reduce(x)
g.replay(); torch.cuda.synchronize()
print(rank, "graph-captured allreduce ok", flush=True)
os._exit(0)
