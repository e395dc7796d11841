"""BASELINE.md §3 leg B1 / BASELINE.json configs[0], run in the BUILD container (the only place
/root/reference exists): the UNMODIFIED reference CLIP (RN50 + RBT3, D = 1024, random init) forward +
reference get_loss (aggregate=False) + backward on a synthetic batch of 64 @224px, on the host cores.
Also the share of that step the hot path of this repo covers (the loss tail, timed alone).

    python tools/config1_cpu_baseline.py  ->  profiles/r02_config1_cpu_reference.json
"""
import json, os, sys, time, types
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_golden as MG  # the shims of SURVEY.md 8c (flash-attn v1 stub, Tensor.cuda -> identity)

MG.install_shims()
from cn_clip.clip.model import CLIP                      # noqa: E402
from cn_clip.training.train import get_loss              # noqa: E402

cfg_dir = Path(MG.REF) / "cn_clip" / "clip" / "model_configs"
vis = json.load(open(cfg_dir / "RN50.json"))
txt = json.load(open(cfg_dir / "RBT3-chinese.json"))
info = dict(vis)
info.update(txt)
if isinstance(info.get("vision_layers"), str):
    info["vision_layers"] = eval(info["vision_layers"])
tok = types.SimpleNamespace(vocab={"[PAD]": 0})
torch.manual_seed(0)
model = CLIP(**info, tokenizer=tok)
model.train()
g = torch.Generator().manual_seed(0)
images = torch.randn(64, 3, 224, 224, generator=g)
texts = torch.randint(1, 21128, (64, 52), generator=g)
texts[:, 0] = 101
args = MG.make_args()
crit = nn.CrossEntropyLoss()
threads = torch.get_num_threads()


def step():
    model.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    total, acc = get_loss(model, images, texts, crit, crit, args)
    t1 = time.perf_counter()
    total.backward()
    return float(total), t1 - t0, time.perf_counter() - t1


step()
runs = [step() for _ in range(3)]
fwd = sum(r[1] for r in runs) / 3
bwd = sum(r[2] for r in runs) / 3

# the loss tail alone (what the fused kernels replace) on the towers' outputs
with torch.no_grad():
    I, T, s = model(images, texts, 0)


class Stub(nn.Module):
    def __init__(self):
        super().__init__()
        self.i, self.t = nn.Parameter(I.clone()), nn.Parameter(T.clone())
        self.logit_scale = nn.Parameter(model.logit_scale.detach().clone())

    def forward(self, images, texts, mask_ratio=0):
        return self.i, self.t, self.logit_scale.exp()


stub = Stub()
t0 = time.perf_counter()
for _ in range(200):
    stub.zero_grad(set_to_none=True)
    l, _ = get_loss(stub, None, None, crit, crit, args)
    l.backward()
tail = (time.perf_counter() - t0) / 200
out = {"config": "BASELINE.json configs[0]: reference CLIP RN50+RBT3 (D=1024, random init) forward + get_loss "
                 "(aggregate=False) + backward, synthetic batch 64 @224px, CPU",
       "where": "build container (no GPU), unmodified /root/reference imported with the two shims of SURVEY.md 8c",
       "cores": threads, "loss": runs[-1][0], "ln_64": 4.1589,
       "forward_s": fwd, "backward_s": bwd, "pairs_per_s": 64 / (fwd + bwd),
       "loss_tail_alone_s": tail, "loss_tail_share_of_step": tail / (fwd + bwd),
       "note": "encoder-dominated: the hot path of this repo (normalise + logits + 2 x CE + backward) is "
               f"{100 * tail / (fwd + bwd):.3f} % of this step; reported baseline only (SURVEY.md 8d row 1)"}
(ROOT / "profiles" / "r02_config1_cpu_reference.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
