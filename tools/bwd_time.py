"""Backward kernel alone (CUDA events, L2 flushed) at N=32768 for several D; development aid.
usage: python tools/bwd_time.py [D ...]      (env NLOC=4096: time one rank's strip of an 8-rank job)"""
import os
import sys
import torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
n = 32768
nloc = int(os.environ.get('NLOC', n))
for d in [int(a) for a in sys.argv[1:]] or [512, 768, 1024]:
    I = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=-1).half()
    T = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=-1).half()
    s_dev = torch.tensor([14.2857], device=dev)
    slots = K.fwd_phase_slots(n, n, d)
    ws = K.fwd_workspace(n, slots, dev)
    K.fwd_phase(I, T, T, I, col_global_begin=0, label_begin=0, s_dev=s_dev, with_acc=False, ws=ws, slot_begin=0)
    lse, sc, _ = K.fwd_finalize(n, slots, 0, s_dev, False, ws)
    g = torch.ones(1, device=dev)

    def bwd():
        K.bwd(I[:nloc], T[:nloc], T, I, label_begin=0, s_dev=s_dev, lse_all=lse, grad_out=g, grad_mult=1.0,
              row_begin=0, row_count=nloc, out_dtype=torch.float32)

    def fwd():
        K.fwd_phase(I, T, T, I, col_global_begin=0, label_begin=0, s_dev=s_dev, with_acc=False, ws=ws, slot_begin=0)

    for name, fn, alg in (("bwd", bwd, 4.0), ("fwd", fwd, 2.0)):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        rows = nloc if name == "bwd" else n
        print(f"{name} rows={rows} N={n} D={d}: {ms:.3f} ms  alg {alg * rows * n * d / ms / 1e9:.0f} TF/s", flush=True)
