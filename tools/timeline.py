"""Kernel/collective timeline of one loss step on rank 0 (torch.profiler / CUPTI; development aid).
torchrun --nproc-per-node W tools/timeline.py  ->  gpurun_out/timeline_W{W}.txt"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
sys.argv = [sys.argv[0]]
import bench
from nans_clip_b200.loss import clip_contrastive_loss

W = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
group = None
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
n_loc = bench.N_GLOBAL // W
img, txt = bench.synth_features(n_loc, rank * n_loc, bench.D)
I = img.to(dev).requires_grad_(True); T = txt.to(dev).requires_grad_(True)
s = torch.tensor(bench.LOGIT_SCALE, device=dev, requires_grad=True)


def step():
    I.grad = None; T.grad = None; s.grad = None
    loss, _ = clip_contrastive_loss(I, T, s, group=group)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
if W > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity
import time
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    flush.zero_()
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print(f"unsynced loop: CPU issue {1e6 * (t1 - t0) / 20:.0f} us/step, total {1e6 * (t2 - t0) / 20:.0f} us/step")
if W > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6):      # like bench.py: no sync between steps, L2 flush before each
        flush.zero_()
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last step only: events after the last big gap
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/timeline_W{W}{os.environ.get('TL_TAG', '')}.txt", "w") as f:
        t0 = None; prev_end = None
        for e in evs:
            st, en = e.time_range.start, e.time_range.end
            if prev_end is not None and st - prev_end > 100000:
                f.write("---- gap %.1f us ----\n" % (st - prev_end)); t0 = st
            if t0 is None:
                t0 = st
            f.write("%9.1f %8.1f  %s\n" % (st - t0, en - st, e.name[:100]))
            prev_end = max(prev_end or en, en)
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith(("cudaLaunch", "cuLaunch", "nccl", "c10d"))]
    print("wrote timeline; cuda events", len(evs))
if W > 1:
    dist.barrier(); dist.destroy_process_group()
