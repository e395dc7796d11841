"""Per-shard retrieval time at a given gallery size (development aid): python tools/topk_only.py [G]
Honours NANS_TOPK_FLOOR_COLS / NANS_TOPK_FLOOR_SPLITS; prints the median of 8 CUDA-event timings."""
import sys, torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K
dev = torch.device("cuda:0")
Q, G, D = 30000, int(sys.argv[1]) if len(sys.argv) > 1 else 125000, 512
g = torch.Generator(device=dev).manual_seed(1)
g32 = torch.nn.functional.normalize(torch.randn(G, D, device=dev, generator=g), dim=-1)
q32 = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g) + 0.5 * (D ** 0.5) * g32[(torch.arange(Q, device=dev) * 33) % G], dim=-1)
q16, g16 = q32.half(), g32.half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(11):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); K.topk_ip(q16, g16, q32, g32, 10, 16, 0); b.record()
    torch.cuda.synchronize()
    if it >= 3:
        ts.append(a.elapsed_time(b))
ts.sort()
import os
print(f"G={G} floor_cols={os.environ.get('NANS_TOPK_FLOOR_COLS', 'default')} splits={os.environ.get('NANS_TOPK_FLOOR_SPLITS', '1')}: {ts[len(ts)//2]:.3f} ms")
