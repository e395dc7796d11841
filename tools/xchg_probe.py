"""Where does the multi-GPU forward spend its time?  (development aid; torchrun --nproc-per-node W)
Per rank: the forward kernel with everything already arrived (no waiting), then one real step with the
per-CTA flag-wait time recorded by the producer warps (NANS_XCHG_PROBE)."""
import copy, ctypes, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
sys.argv = [sys.argv[0]]
import bench
from nans_clip_b200 import exchange, kernels as K
from nans_clip_b200.loss import clip_contrastive_loss

W = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD
n_loc = bench.N_GLOBAL // W
img, txt = bench.synth_features(n_loc, rank * n_loc, bench.D)
I = img.to(dev).requires_grad_(True); T = txt.to(dev).requires_grad_(True)
s = torch.tensor(bench.LOGIT_SCALE, device=dev, requires_grad=True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def step():
    I.grad = None; T.grad = None; s.grad = None
    loss, _ = clip_contrastive_loss(I, T, s, group=group)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier()
ex = exchange.for_group(group)
desc = ex.desc
# (1) forward alone on the data of the last step: a descriptor whose step counter is one behind
prev = (ex.epoch - 1).clone()
d2 = exchange.make_desc(W, rank, ex.bases, n_loc, bench.D, prev.data_ptr())
loc16 = torch.stack([I.detach().half(), T.detach().half()])
s_dev = s.detach().reshape(1)
ns = K.fwd_xchg_slots(n_loc, W, bench.D)
ws = K.fwd_workspace(n_loc, ns, dev)
ts = []
for _ in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); K.fwd_xchg(d2, loc16[0], loc16[1], s_dev, False, ws); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
alone = sorted(ts)[len(ts) // 2]
# (2) real steps with the wait probe
grid = 4096
probe = torch.zeros(grid, dtype=torch.int64, device=dev)
os.environ["NANS_XCHG_PROBE"] = str(probe.data_ptr())
dist.barrier()
for _ in range(4):
    flush.zero_(); probe.zero_(); torch.cuda.synchronize(); dist.barrier()
    step()
    torch.cuda.synchronize()
w = probe[probe > 0].float() / 1e3
out = torch.tensor([alone, float(w.max()) if w.numel() else 0.0, float(w.mean()) if w.numel() else 0.0, float(w.numel())], device=dev)
outs = [torch.zeros_like(out) for _ in range(W)]
dist.all_gather(outs, out)
if rank == 0:
    for r, o in enumerate(outs):
        print(f"rank {r}: forward alone {o[0]*1e3:.0f} us | flag waits per CTA: max {o[1]:.0f} us, mean {o[2]:.0f} us over {int(o[3])} CTAs")
dist.barrier(); dist.destroy_process_group()
