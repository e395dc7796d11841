"""all_gather_into_tensor latency at the bench's per-rank sizes (development aid; NCCL env variants)."""
import os, sys, torch, torch.distributed as dist
W = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
tag = os.environ.get("TAG", "default")
for nbytes in (2 * 4096 * 4 + 32, 4096 * 512 * 2, 2 * 4096 * 512 * 2):
    x = torch.empty(nbytes // 2, dtype=torch.float16, device=dev)
    y = torch.empty(W * (nbytes // 2), dtype=torch.float16, device=dev)
    for _ in range(10):
        dist.all_gather_into_tensor(y, x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_gather_into_tensor(y, x); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    if rank == 0:
        print(f"[{tag}] W={W} per-rank {nbytes} B: median {1e3 * ts[len(ts) // 2]:.1f} us  min {1e3 * ts[0]:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
