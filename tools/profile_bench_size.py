"""One forward + one backward launch at the bench size (N=32768, D=512) for `ncu --set full`."""
import sys
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K
sys.argv = [sys.argv[0]]
import bench

dev = torch.device("cuda:0")
n, d = bench.N_GLOBAL, bench.D
img, txt = bench.synth_features(n, 0, d)
I16, _, _ = K.l2norm_cast(img.to(dev), torch.float16, normalize=False)
T16, _, _ = K.l2norm_cast(txt.to(dev), torch.float16, normalize=False)
s_dev = torch.tensor([bench.LOGIT_SCALE], device=dev)
for it in range(2):
    slots = K.fwd_phase_slots(n, n, d)
    ws = K.fwd_workspace(n, slots, dev)
    K.fwd_phase(I16, T16, T16, I16, col_global_begin=0, label_begin=0, s_dev=s_dev, with_acc=False, ws=ws, slot_begin=0)
    lse, sc, _ = K.fwd_finalize(n, slots, 0, s_dev, False, ws)
    K.bwd(I16, T16, T16, I16, label_begin=0, s_dev=s_dev, lse_all=lse, grad_out=torch.ones(1, device=dev),
          grad_mult=1.0, row_begin=0, row_count=n, out_dtype=torch.float32)
torch.cuda.synchronize()
print("ok", float((sc[0] + sc[1]) / (2 * n)))
