"""GPU bring-up: each kernel against a plain fp32 torch computation on the same bf16-exact inputs.
Prints one line per check and keeps going after a failed comparison (a CUDA fault still aborts).
Usage: python tools/bringup.py [l2norm] [fwd] [bwd] [topk]
"""
import math
import sys
import time
import traceback

import torch

sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K  # noqa: E402
from nans_clip_b200.loss import clip_contrastive_loss  # noqa: E402

dev = torch.device("cuda:0")


def feats(n, d, seed, corr=0.5, dt=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, d, generator=g)
    a = corr * base + (1 - corr) * torch.randn(n, d, generator=g)
    b = corr * base + (1 - corr) * torch.randn(n, d, generator=g)
    a = torch.nn.functional.normalize(a, dim=-1).to(dt).float()
    b = torch.nn.functional.normalize(b, dim=-1).to(dt).float()
    return a, b


def ref_loss(I, T, s):
    I = I.double().requires_grad_(True)
    T = T.double().requires_grad_(True)
    s = torch.tensor(float(s), dtype=torch.float64, requires_grad=True)
    logits = s * I @ T.t()
    gt = torch.arange(len(I))
    loss = (torch.nn.functional.cross_entropy(logits, gt) + torch.nn.functional.cross_entropy(logits.t(), gt)) / 2
    loss.backward()
    acc1 = (logits.argmax(-1) == gt).float().mean()
    acc2 = (logits.t().argmax(-1) == gt).float().mean()
    return loss.item(), I.grad.float(), T.grad.float(), s.grad.item(), acc1.item(), acc2.item(), logits.detach()


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def report(name, ok, msg):
    print(f"[{'PASS' if ok else 'FAIL'}] {name}: {msg}", flush=True)


def t_l2norm():
    for dt_in in (torch.float32, torch.float16, torch.bfloat16):
        for D in (64, 512, 768, 1024, 2048 + 64):
            x = torch.randn(1000, D, device=dev).to(dt_in)
            y16, y32, inv = K.l2norm_cast(x, torch.bfloat16, want_fp32=True, want_inv_norm=True)
            r = x.float() / x.float().norm(dim=-1, keepdim=True)
            e32 = (y32 - r).abs().max().item()
            e16 = (y16.float() - r.bfloat16().float()).abs().max().item()
            einv = ((inv - 1 / x.float().norm(dim=-1)).abs().max() * 1).item()
            report(f"l2norm {dt_in} D={D}", e32 < 1e-6 and e16 < 1e-2 and einv < 1e-3,
                   f"max|y32-ref|={e32:.2e} max|y16-ref16|={e16:.2e} inv={einv:.2e}")
    x = torch.randn(777, 512, device=dev)
    _, _, inv = K.l2norm_cast(x, None, want_inv_norm=True)
    dy = torch.randn(777, 512, device=dev)
    dx = K.l2norm_bwd(x, inv, dy)
    xr = x.clone().requires_grad_(True)
    (xr / xr.norm(dim=-1, keepdim=True)).backward(dy)
    report("l2norm_bwd", rel(dx, xr.grad) < 1e-5, f"rel={rel(dx, xr.grad):.2e}")


def t_fwd_bwd(do_bwd, dts=(torch.float16, torch.bfloat16)):
    cases = [(128, 64, 14.285), (256, 512, 14.285), (1000, 512, 100.0), (4096, 512, 14.285),
             (1111, 768, 50.0), (640, 1024, 14.285), (2048, 256, 1.0), (300, 72, 20.0)]
    for (n, d, s) in cases:
        for dt in dts:
            try:
                I, T = feats(n, d, seed=n + d, dt=dt)
                L, gI, gT, gs, a1, a2, _ = ref_loss(I, T, s)
                Ic = I.to(dev).requires_grad_(True)
                Tc = T.to(dev).requires_grad_(True)
                sc = torch.tensor(s, device=dev, requires_grad=True)
                loss, acc = clip_contrastive_loss(Ic, Tc, sc, report_acc=True, feat_dtype=dt)
                torch.cuda.synchronize()
                eL = abs(loss.item() - L) / max(abs(L), 1e-3)
                ok = eL < 1e-3 and abs(acc["i2t"].item() - a1) < 2e-3 and abs(acc["t2i"].item() - a2) < 2e-3
                report(f"fwd n={n} D={d} s={s} {dt}", ok,
                       f"loss={loss.item():.6f} ref={L:.6f} rel={eL:.2e} acc=({acc['i2t'].item():.4f},{acc['t2i'].item():.4f}) ref=({a1:.4f},{a2:.4f})")
                if do_bwd:
                    loss.backward()
                    torch.cuda.synchronize()
                    eI, eT = rel(Ic.grad.cpu(), gI), rel(Tc.grad.cpu(), gT)
                    es = abs(sc.grad.item() - gs) / max(abs(gs), 1e-6)
                    report(f"bwd n={n} D={d} s={s} {dt}", eI < 1e-3 and eT < 1e-3 and es < 1e-3,
                           f"dI rel={eI:.2e} dT rel={eT:.2e} ds={sc.grad.item():.6e} ref={gs:.6e} rel={es:.2e}")
            except Exception:
                traceback.print_exc()
                report(f"fwd/bwd n={n} D={d} s={s} {dt}", False, "exception")
                raise


def t_topk():
    for (Q, G, D, k, kc) in [(100, 1000, 64, 10, 16), (300, 5000, 512, 10, 16), (1000, 70000, 512, 10, 32),
                             (130, 300, 256, 10, 16), (5, 7, 512, 10, 16)]:
        g = torch.Generator().manual_seed(Q + G)
        Gm = torch.nn.functional.normalize(torch.randn(G, D, generator=g), dim=-1).bfloat16().float()
        Qm = torch.nn.functional.normalize(torch.randn(Q, D, generator=g) + 0.5 * Gm[(torch.arange(Q) * 33) % G], dim=-1).bfloat16().float()
        sc = Qm.double() @ Gm.double().t()
        kk = min(k, G)
        rs, ri = torch.sort(sc, dim=1, descending=True, stable=True)
        rs, ri = rs[:, :kk], ri[:, :kk]
        s, i = K.topk_ip(Qm.to(dev).bfloat16(), Gm.to(dev).bfloat16(), Qm.to(dev), Gm.to(dev), k, kc, 1000000)
        torch.cuda.synchronize()
        i = i.cpu()[:, :kk] - 1000000
        s = s.cpu()[:, :kk]
        mism = (i != ri)
        # excuse positions whose neighbouring reference gaps are below 1e-4
        gap_ok = torch.ones_like(mism)
        d = (rs[:, :-1] - rs[:, 1:]).abs() > 1e-4
        gap_ok[:, :-1] &= d
        gap_ok[:, 1:] &= d
        bad = (mism & gap_ok).sum().item()
        report(f"topk Q={Q} G={G} D={D} k={k} kc={kc}", bad == 0 and (s - rs.float()).abs().max().item() < 1e-4,
               f"mismatch={mism.sum().item()} inexcusable={bad} max score err={(s - rs.float()).abs().max().item():.2e}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["l2norm", "fwd", "bwd", "topk"]
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    if "l2norm" in which:
        t_l2norm()
    if "fwd" in which or "bwd" in which:
        dts = tuple(d for d, nm in ((torch.float16, "f16"), (torch.bfloat16, "bf16")) if nm in which) or (torch.float16, torch.bfloat16)
        t_fwd_bwd("bwd" in which, dts)
    if "topk" in which:
        t_topk()
    print(f"done in {time.time() - t0:.1f}s", flush=True)
