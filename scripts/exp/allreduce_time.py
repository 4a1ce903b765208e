"""Experiment: device time of the per-step gradient all-reduce (31 MB fp32) at N ranks."""
import os, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.randn(7_760_000, device="cuda")
for op, name in ((dist.ReduceOp.SUM, "sum"), (dist.ReduceOp.AVG, "avg")):
    for _ in range(5): dist.all_reduce(x, op=op)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): dist.all_reduce(x, op=op)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"all_reduce {name} 31 MB x{world}: {e0.elapsed_time(e1)/50*1e3:.1f} us", flush=True)
# interleaved with compute on the same stream (as in the step)
y = torch.randn(4096, 4096, device="cuda")
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(50):
    for _ in range(4): y @ y
e1.record(); torch.cuda.synchronize(); t_c = e0.elapsed_time(e1) / 50
e0.record()
for _ in range(50):
    for _ in range(4): y @ y
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
e1.record(); torch.cuda.synchronize(); t_ca = e0.elapsed_time(e1) / 50
if rank == 0: print(f"compute {t_c*1e3:.0f} us, compute + all_reduce {t_ca*1e3:.0f} us -> exposed {1e3*(t_ca-t_c):.0f} us")
dist.destroy_process_group()
