"""Where the host time of one loss step goes (cProfile of rank 0 at a size where the GPU is never
the bottleneck).  Development aid."""
import cProfile, os, pstats, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from nans_clip_b200.loss import clip_contrastive_loss

W = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
group = None
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
n_loc, D = 512, 512
I = torch.nn.functional.normalize(torch.randn(n_loc, D, device=dev), dim=-1).requires_grad_(True)
T = torch.nn.functional.normalize(torch.randn(n_loc, D, device=dev), dim=-1).requires_grad_(True)
s = torch.tensor(14.28, device=dev, requires_grad=True)


def step():
    I.grad = None; T.grad = None; s.grad = None
    loss, _ = clip_contrastive_loss(I, T, s, group=group)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
if rank == 0:
    print(f"W={W} host issue {1e6 * (t1 - t0) / 200:.0f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
if rank == 0:
    st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
if W > 1:
    dist.barrier(); dist.destroy_process_group()
