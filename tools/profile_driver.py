"""Runs each kernel a few times at a profiling-friendly size (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
g = torch.Generator(device=dev).manual_seed(0)
base = torch.randn(n, d, device=dev, generator=g)
I = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(n, d, device=dev, generator=g), dim=-1)
T = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(n, d, device=dev, generator=g), dim=-1)
s_dev = torch.tensor([14.2857], device=dev)
for it in range(3):
    I16, _, _ = K.l2norm_cast(I, torch.float16, normalize=True)
    T16, _, _ = K.l2norm_cast(T, torch.float16, normalize=True)
    slots = K.fwd_phase_slots(n, n, d)
    ws = K.fwd_workspace(n, slots, dev)
    K.fwd_phase(I16, T16, T16, I16, col_global_begin=0, label_begin=0, s_dev=s_dev, with_acc=False, ws=ws, slot_begin=0)
    lse, sc, _ = K.fwd_finalize(n, slots, 0, s_dev, False, ws)
    if which in ("all", "bwd"):
        K.bwd(I16, T16, T16, I16, label_begin=0, s_dev=s_dev, lse_all=lse, grad_out=torch.ones(1, device=dev),
              grad_mult=1.0, row_begin=0, row_count=n, out_dtype=torch.float32)
    if which in ("all", "topk"):
        if it == 0:
            G32 = torch.nn.functional.normalize(torch.randn(8 * n, d, device=dev, generator=g), dim=-1)
            G16 = G32.half()
        K.topk_ip(I16, G16, I, G32, 10, 16, 0)
torch.cuda.synchronize()
print("ok", float(sc[0]))
