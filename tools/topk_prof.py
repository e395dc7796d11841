import sys, torch
sys.path.insert(0, ".")
from nans_clip_b200 import kernels as K
dev = torch.device("cuda:0")
Q, G, D = 18944, 262144, 512
q32 = torch.nn.functional.normalize(torch.randn(Q, D, device=dev), dim=-1)
g32 = torch.nn.functional.normalize(torch.randn(G, D, device=dev), dim=-1)
q16, g16 = q32.half(), g32.half()
for _ in range(2):
    K.topk_ip(q16, g16, q32, g32, 10, 16, 0)
torch.cuda.synchronize()
print("ok")
