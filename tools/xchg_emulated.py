"""One step of the 8-rank exchange with the ranks EMULATED on one GPU at the bench shape (n_loc=4096, D=512):
every phase runs for all ranks before the next (no kernel waits for a later launch).  Used under ncu to look at
the exchange-mode forward (flag polling + fence.proxy.async + TMA in the same kernel as the tcgen05 tiles) and
at the push kernels; prints CUDA-event times of rank 0's kernels when run plain."""
import sys, math
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import exchange as X, kernels as K

W, n_loc, D, s = 8, 4096, 512, 14.2857
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
N = W * n_loc
g = torch.Generator(device=dev).manual_seed(0)
base = torch.randn(N, D, device=dev, generator=g)
I = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(N, D, device=dev, generator=g), dim=-1)
T = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(N, D, device=dev, generator=g), dim=-1)
bufs = [torch.zeros(X.layout_bytes(W, n_loc, D), dtype=torch.uint8, device=dev) for _ in range(W)]
eps = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(W)]
descs = [X.make_desc(W, r, [b.data_ptr() for b in bufs], n_loc, D, eps[r].data_ptr()) for r in range(W)]
s_dev = torch.tensor([s], device=dev)
ns = K.fwd_xchg_slots(n_loc, W, D)
ev = lambda: torch.cuda.Event(enable_timing=True)
for step in range(3):
    t = [ev() for _ in range(8)]
    loc = []
    for r in range(W):
        if r == 0: t[0].record()
        loc.append(K.xchg_cast_push(descs[r], I[r * n_loc:(r + 1) * n_loc], T[r * n_loc:(r + 1) * n_loc], torch.float16))
        if r == 0: t[1].record()
    wss = [K.fwd_workspace(n_loc, ns, dev) for _ in range(W)]
    for r in range(W):
        if r == 0: t[2].record()
        K.fwd_xchg(descs[r], loc[r][0], loc[r][1], s_dev, False, wss[r])
        if r == 0: t[3].record()
    for r in range(W):
        K.fwd_finalize_push(descs[r], ns, s_dev, False, wss[r])
    fin = [K.exchange_finish_xchg(descs[r], dev) for r in range(W)]
    lse_all, out, mm, stp = fin[0]
    t[4].record()
    K.bwd_xchg(descs[0], stp, loc[0][0], loc[0][1], s_dev=s_dev, lse_all=lse_all, lse_minmax=mm,
               grad_out=torch.ones(1, device=dev), grad_mult=1.0, row_begin=0, row_count=n_loc, out_dtype=torch.float32)
    t[5].record()
    torch.cuda.synchronize()
print(f"rank 0 of {W} emulated ranks, n_loc={n_loc}, D={D}: cast+push (local memory) {t[0].elapsed_time(t[1])*1e3:.0f} us, "
      f"forward (all flags up) {t[2].elapsed_time(t[3])*1e3:.0f} us, backward {t[4].elapsed_time(t[5])*1e3:.0f} us, "
      f"loss {float(out[0]):.5f} ({ns} column splits)")
