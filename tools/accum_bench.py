"""One optimizer step of the gradient-accumulation path (A get_loss calls + backward each), full
recomputation vs the incremental forward (accum.py).  Single GPU; config-3-like widths.
usage: python tools/accum_bench.py [A] [B] [D]"""
import os, sys, time
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import accum
from nans_clip_b200.loss import clip_contrastive_loss

A = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
D = int(sys.argv[3]) if len(sys.argv) > 3 else 768
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
base = torch.randn(A * B, D, device=dev, generator=g)
I = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(A * B, D, device=dev, generator=g), dim=-1)
T = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(A * B, D, device=dev, generator=g), dim=-1)
s = torch.tensor(14.2857, device=dev, requires_grad=True)


def one_step(incremental: bool):
    cache_i = [I[a * B:(a + 1) * B].clone() for a in range(A)]   # fresh lists = a new optimizer step
    cache_t = [T[a * B:(a + 1) * B].clone() for a in range(A)]
    for j in range(A):
        ci = cache_i[j].clone().requires_grad_(True)
        ct = cache_t[j].clone().requires_grad_(True)
        if incremental:
            loss, _ = accum.incremental_accum_loss(ci, ct, s, cache_i, cache_t, j)
        else:
            loss, _ = clip_contrastive_loss(ci, ct, s, full_image_features=torch.cat(cache_i[:j] + [ci.detach()] + cache_i[j + 1:]),
                                            full_text_features=torch.cat(cache_t[:j] + [ct.detach()] + cache_t[j + 1:]),
                                            row_begin=j * B)
        loss.backward()
    return float(loss)


for inc in (False, True):
    one_step(inc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        last = one_step(inc)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(f"A={A} B={B} D={D} N={A * B}: {'incremental' if inc else 'full recompute'} {ms:.2f} ms per optimizer step "
          f"({A} calls), loss {last:.6f}", flush=True)
