"""Quick device timings of each kernel at the BASELINE sizes (CUDA events, L2 flushed between
iterations).  Development aid; bench.py is the contract."""
import sys
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K
from nans_clip_b200.loss import clip_contrastive_loss

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def feats(n, d, dt):
    x = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=-1)
    return x.to(dt)


which = sys.argv[1:] or ["l2norm", "fwd", "bwd", "topk"]
if "l2norm" in which:
    x = torch.randn(1000000, 512, device=dev)
    ms = timeit(lambda: K.l2norm_cast(x, torch.bfloat16))
    print(f"l2norm 1M x512 fp32->bf16: {ms:.3f} ms  {1e6*512*6/ms/1e6:.0f} GB/s")
for dt in (torch.float16, torch.bfloat16):
    for (n, d) in [(32768, 512), (4096, 512), (32768, 768), (32768, 1024)]:
        if "fwd" in which or "bwd" in which:
            I = feats(n, d, dt).requires_grad_(True); T = feats(n, d, dt).requires_grad_(True)
            s = torch.tensor(14.285, device=dev, requires_grad=True)
            def f():
                return clip_contrastive_loss(I, T, s, feat_dtype=dt)[0]
            ms = timeit(f)
            print(f"fwd {dt} N={n} D={d}: {ms:.3f} ms  alg {2*n*n*d/ms/1e9:.0f} TF/s  hw {4*n*n*d/ms/1e9:.0f} TF/s")
            if "bwd" in which:
                def fb():
                    l = f(); l.backward()
                ms2 = timeit(fb)
                print(f"fwd+bwd {dt} N={n} D={d}: {ms2:.3f} ms  alg {6*n*n*d/ms2/1e9:.0f} TF/s  pairs/s {n/ms2*1e3:.3e}")
if "topk" in which:
    Q, G, D = 30000, 125000, 512
    q32 = torch.nn.functional.normalize(torch.randn(Q, D, device=dev), dim=-1)
    g32 = torch.nn.functional.normalize(torch.randn(G, D, device=dev), dim=-1)
    q16, g16 = q32.bfloat16(), g32.bfloat16()
    ms = timeit(lambda: K.topk_ip(q16, g16, q32, g32, 10, 16, 0))
    print(f"topk Q={Q} G={G}: {ms:.3f} ms  {2*Q*G*D/ms/1e9:.0f} TF/s  {Q/ms*1e3:.3e} q/s")
    G = 1000000
    g32 = torch.nn.functional.normalize(torch.randn(G, D, device=dev), dim=-1); g16 = g32.bfloat16()
    ms = timeit(lambda: K.topk_ip(q16, g16, q32, g32, 10, 16, 0), iters=3, warm=1)
    print(f"topk Q={Q} G={G}: {ms:.3f} ms  {2*Q*G*D/ms/1e9:.0f} TF/s  {Q/ms*1e3:.3e} q/s")
