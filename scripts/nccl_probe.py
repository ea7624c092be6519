"""Diagnostic: time the all-gather of the trajectory slab alone and print the transport NCCL picked."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
for rows in (1024, 8192):
    slab = torch.randn(rows, 198, dtype=torch.float64, device=dev)
    out = torch.empty(world * rows, 198, dtype=torch.float64, device=dev)
    for _ in range(5): dist.all_gather_into_tensor(out, slab)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_gather_into_tensor(out, slab)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"all_gather {rows}x198 f64 ({slab.numel()*8/1e6:.2f} MB/rank): {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5): dist.all_gather_into_tensor(out, slab)
        s.synchronize(); e0.record(s)
        for _ in range(20): dist.all_gather_into_tensor(out, slab)
        e1.record(s)
    s.synchronize()
    if rank == 0: print(f"  on a side stream: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
dist.destroy_process_group()
