"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

It imports the unmodified reference (n571e/NanS-CLIP at /root/reference) with the two
non-invasive shims of SURVEY.md §8c — a stub for the flash-attn v1 module the reference imports
at module load, and `Tensor.cuda` -> identity because `get_loss` hard-codes `.cuda()`
(cn_clip/training/train.py:110) — and records, for seeded synthetic inputs:

  loss_w1_*.npz      get_loss (aggregate=False), loss / acc / autograd gradients
  loss_dist_*.npz    get_loss under torch.distributed gloo, W ranks, both gather modes
  loss_kd_*.npz      get_loss with args.distillation under gloo (W = 2, both gather modes) and local
  loss_accum_*.npz   get_loss on the gradient-accumulation path (train.py:34-51)
  loss_accum4_*.npz  the same with 4 chunks of 256 rows and a re-forwarded chunk that differs from its cache
                     (features stored as their bf16 bit patterns: they are bf16-exact by construction)
  tail_*.npz         CLIP.forward normalise lines + get_similarity (model.py:412-431)
  topk_*.npz         make_topk_predictions.py / _tr.py run as scripts on small JSONL files
  lora_loss_*.npz    train_lora.py `contrastive_loss` (label-smoothed InfoNCE), value + gradients
                     (third shim: a stub `lmdb` module, which train_lora.py imports at load)

Nothing in the test-suite reads /root/reference: tests only read the .npz files written here.
"""
from __future__ import annotations

import json
import os
import runpy
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = Path(__file__).resolve().parent
REF = "/root/reference"


def install_shims():
    if "flash_attn.flash_attention" not in sys.modules:
        stub = types.ModuleType("flash_attn.flash_attention")
        stub.FlashMHA = type("FlashMHA", (nn.Module,), {})
        sys.modules["flash_attn.flash_attention"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)


def synth(n, d, seed, corr):
    """correlated unit-norm pairs, rounded to bf16 so that every implementation
    sees identical 16-bit-exact values."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, d, generator=g)
    img = corr * base + (1 - corr) * torch.randn(n, d, generator=g)
    txt = corr * base + (1 - corr) * torch.randn(n, d, generator=g)
    img = (img / img.norm(dim=-1, keepdim=True)).bfloat16().float()
    txt = (txt / txt.norm(dim=-1, keepdim=True)).bfloat16().float()
    return img, txt


class StubModel(nn.Module):
    """Stands in for DDP(CLIP): returns fixed (already normalised) features and exp(logit_scale),
    exactly the triple CLIP.forward returns (model.py:415)."""

    def __init__(self, img, txt, logit_scale_log):
        super().__init__()
        self.img = nn.Parameter(img.clone())
        self.txt = nn.Parameter(txt.clone())
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale_log)))

    def forward(self, images, texts, mask_ratio=0):
        return self.img, self.txt, self.logit_scale.exp()


def make_args(**kw):
    a = types.SimpleNamespace(accum_freq=1, mask_ratio=0, distillation=False, aggregate=False,
                              gather_with_grad=False, local_device_rank=0,
                              report_training_batch_acc=True, kd_loss_weight=0.5)
    a.__dict__.update(kw)
    return a


def np_(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
def gen_loss_w1():
    from cn_clip.training.train import get_loss
    cases = [("a", 48, 64, 0.0, 2.6593), ("b", 96, 128, 0.5, 2.6593), ("c", 33, 72, 0.5, 4.6052),
             ("d", 200, 512, 0.5, 0.0), ("e", 1, 64, 0.5, 2.6593),
             ("f", 64, 1024, 0.5, 2.6593)]   # f: the loss shape of BASELINE.json configs[0] (batch 64, D = 1024)
    for name, n, d, corr, ls in cases:
        img, txt = synth(n, d, 1234 + n, corr)
        model = StubModel(img, txt, ls)
        loss_img, loss_txt = nn.CrossEntropyLoss(), nn.CrossEntropyLoss()
        total, acc = get_loss(model, None, None, loss_img, loss_txt, make_args())
        total.backward()
        s = float(np.exp(ls))
        np.savez(HERE / f"loss_w1_{name}.npz", img=np_(img), txt=np_(txt), logit_scale_log=ls,
                 loss=np_(total), i2t=np_(acc["i2t"]), t2i=np_(acc["t2i"]), dI=np_(model.img.grad),
                 dT=np_(model.txt.grad), dlogit_scale_log=np_(model.logit_scale.grad), s=s)
        print("loss_w1", name, float(total))


def _dist_worker(rank, W, port, gather_with_grad, n_loc, d, seed, corr, ls, out):
    install_shims()
    from cn_clip.training.train import get_loss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=W)
    img, txt = synth(W * n_loc, d, seed, corr)
    model = StubModel(img[rank * n_loc:(rank + 1) * n_loc], txt[rank * n_loc:(rank + 1) * n_loc], ls)
    args = make_args(aggregate=True, gather_with_grad=gather_with_grad)
    total, acc = get_loss(model, None, None, nn.CrossEntropyLoss(), nn.CrossEntropyLoss(), args)
    total.backward()
    out.put((rank, np_(total), np_(acc["i2t"]), np_(acc["t2i"]), np_(model.img.grad),
             np_(model.txt.grad), np_(model.logit_scale.grad)))
    dist.barrier()
    dist.destroy_process_group()


class StubTeacher(nn.Module):
    """teacher_model.module.get_feature(images) (train.py:25-32): fixed features of another width."""

    def __init__(self, feat):
        super().__init__()
        self.module = self
        self.feat = feat

    def get_feature(self, images):
        return self.feat


def _kd_worker(rank, W, port, gather_with_grad, n_loc, d, dt, seed, ls, out):
    install_shims()
    from cn_clip.training.train import get_loss
    if W > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=W)
    img, txt = synth(W * n_loc, d, seed, 0.5)
    teacher = torch.randn(W * n_loc, dt, generator=torch.Generator().manual_seed(seed + 7))
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    model = StubModel(img[sl], txt[sl], ls)
    args = make_args(aggregate=W > 1, gather_with_grad=gather_with_grad, distillation=True, kd_loss_weight=0.5)
    total, acc = get_loss(model, None, None, nn.CrossEntropyLoss(), nn.CrossEntropyLoss(), args,
                          teacher_model=StubTeacher(teacher[sl]))
    total.backward()
    out.put((rank, np_(total), np_(model.img.grad), np_(model.txt.grad), np_(model.logit_scale.grad)))
    if W > 1:
        dist.barrier()
        dist.destroy_process_group()


def gen_loss_kd():
    """The knowledge-distillation branch (train.py:25-32, 62-63, 90-100, 106-107, 123-124), including the
    reference's pairing of a rank-ordered student gather with a local-first teacher gather under
    gather_with_grad."""
    port = 29641
    for name, W, gwg in [("w1", 1, False), ("w2", 2, False), ("w2g", 2, True)]:
        n_loc, d, dt, seed, ls = 24, 64, 96, 777, 2.6593
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_kd_worker, args=(r, W, port, gwg, n_loc, d, dt, seed, ls, q)) for r in range(W)]
        port += 1
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=300) for _ in range(W)], key=lambda x: x[0])
        for p in procs:
            p.join()
        img, txt = synth(W * n_loc, d, seed, 0.5)
        teacher = torch.randn(W * n_loc, dt, generator=torch.Generator().manual_seed(seed + 7))
        np.savez(HERE / f"loss_kd_{name}.npz", img=np_(img), txt=np_(txt), teacher=np_(teacher), logit_scale_log=ls,
                 W=W, n_loc=n_loc, gather_with_grad=gwg, kd_loss_weight=0.5,
                 loss=np.stack([r[1] for r in res]), dI=np.stack([r[2] for r in res]),
                 dT=np.stack([r[3] for r in res]), dlogit_scale_log=np.stack([r[4] for r in res]))
        print("loss_kd", name, [float(r[1]) for r in res])


def gen_loss_dist():
    port = 29611
    for name, W, gwg, n_loc, d, corr, ls in [("w2", 2, False, 40, 64, 0.5, 2.6593),
                                             ("w2g", 2, True, 40, 64, 0.5, 2.6593),
                                             ("w4", 4, False, 24, 128, 0.0, 4.6052),
                                             ("w4g", 4, True, 24, 128, 0.5, 3.0)]:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        seed = 4321 + W
        procs = [ctx.Process(target=_dist_worker, args=(r, W, port, gwg, n_loc, d, seed, corr, ls, q))
                 for r in range(W)]
        port += 1
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=300) for _ in range(W)], key=lambda x: x[0])
        for p in procs:
            p.join()
        img, txt = synth(W * n_loc, d, seed, corr)
        np.savez(HERE / f"loss_dist_{name}.npz", img=np_(img), txt=np_(txt), logit_scale_log=ls, W=W,
                 n_loc=n_loc, gather_with_grad=gwg, s=float(np.exp(ls)),
                 loss=np.stack([r[1] for r in res]), i2t=np.stack([r[2] for r in res]),
                 t2i=np.stack([r[3] for r in res]), dI=np.stack([r[4] for r in res]),
                 dT=np.stack([r[5] for r in res]), dlogit_scale_log=np.stack([r[6] for r in res]))
        print("loss_dist", name, [float(r[1]) for r in res])


def gen_loss_accum():
    """accum_freq = 3: three cached (no-grad) chunks, chunk j re-forwarded with grad."""
    from cn_clip.training.train import get_loss
    A, B, d, ls = 3, 16, 64, 2.6593
    img, txt = synth(A * B, d, 777, 0.5)
    for j in range(A):
        cache_i = [img[a * B:(a + 1) * B].clone() for a in range(A)]
        cache_t = [txt[a * B:(a + 1) * B].clone() for a in range(A)]
        model = StubModel(img[j * B:(j + 1) * B], txt[j * B:(j + 1) * B], ls)
        args = make_args(accum_freq=A)
        total, acc = get_loss(model, None, None, nn.CrossEntropyLoss(), nn.CrossEntropyLoss(), args,
                              cache_i, cache_t, j)
        total.backward()
        np.savez(HERE / f"loss_accum_j{j}.npz", img=np_(img), txt=np_(txt), logit_scale_log=ls, A=A,
                 B=B, j=j, s=float(np.exp(ls)), loss=np_(total), i2t=np_(acc["i2t"]),
                 t2i=np_(acc["t2i"]), dI=np_(model.img.grad), dT=np_(model.txt.grad),
                 dlogit_scale_log=np_(model.logit_scale.grad))
        print("loss_accum", j, float(total))


def gen_loss_accum_big():
    """accum_freq = 4 with 256-row chunks (what the incremental accumulate path needs) and a
    re-forwarded chunk that DIFFERS from its cached copy (the reference re-forwards with masking /
    dropout, train.py:35 vs :210): calls j = 2 then j = 0 of one optimizer step."""
    from cn_clip.training.train import get_loss
    A, B, d, ls = 4, 256, 64, 2.9
    img, txt = synth(A * B, d, 4242, 0.5)
    cache_i = [img[a * B:(a + 1) * B].clone() for a in range(A)]
    cache_t = [txt[a * B:(a + 1) * B].clone() for a in range(A)]
    g = torch.Generator().manual_seed(99)
    for j in (2, 0):
        ni = img[j * B:(j + 1) * B] + 0.05 * torch.randn(B, d, generator=g)
        nt = txt[j * B:(j + 1) * B] + 0.05 * torch.randn(B, d, generator=g)
        ni = (ni / ni.norm(dim=-1, keepdim=True)).bfloat16().float()
        nt = (nt / nt.norm(dim=-1, keepdim=True)).bfloat16().float()
        model = StubModel(ni, nt, ls)
        args = make_args(accum_freq=A)
        total, acc = get_loss(model, None, None, nn.CrossEntropyLoss(), nn.CrossEntropyLoss(), args,
                              cache_i, cache_t, j)
        total.backward()
        bits = lambda x: x.bfloat16().view(torch.int16).numpy()   # the features are bf16-exact: store the bit patterns
        assert all(bool((x.bfloat16().float() == x).all()) for x in (img, txt, ni, nt))
        np.savez_compressed(HERE / f"loss_accum4_j{j}.npz", img_bf16=bits(img), txt_bf16=bits(txt),
                            new_img_bf16=bits(ni), new_txt_bf16=bits(nt),
                            logit_scale_log=ls, A=A, B=B, j=j, s=float(np.exp(ls)), loss=np_(total),
                            i2t=np_(acc["i2t"]), t2i=np_(acc["t2i"]), dI=np_(model.img.grad), dT=np_(model.txt.grad),
                            dlogit_scale_log=np_(model.logit_scale.grad))
        print("loss_accum4", j, float(total))


def gen_tail():
    from cn_clip.clip.model import CLIP
    torch.manual_seed(0)
    clip = CLIP(embed_dim=64, image_resolution=32, vision_layers=1, vision_width=64,
                vision_patch_size=16, vocab_size=21128, text_attention_probs_dropout_prob=0.0,
                text_hidden_act="gelu", text_hidden_dropout_prob=0.0, text_hidden_size=64,
                text_initializer_range=0.02, text_intermediate_size=128,
                text_max_position_embeddings=64, text_num_attention_heads=1,
                text_num_hidden_layers=1, text_type_vocab_size=2)
    for name, n, d, scale in [("a", 37, 64, 1.0), ("b", 8, 512, 30.0), ("c", 5, 1024, 1e-3)]:
        g = torch.Generator().manual_seed(99 + n)
        raw_i = torch.randn(n, d, generator=g) * scale
        raw_t = torch.randn(n, d, generator=g) * scale
        clip.encode_image = lambda image, mask_ratio=0, _x=raw_i: _x
        clip.encode_text = lambda text, _x=raw_t: _x
        with torch.no_grad():
            I, T, s = clip.forward(torch.zeros(1), torch.zeros(1))
            lpi, lpt = clip.get_similarity(torch.zeros(1), torch.zeros(1))
        # gradient of a fixed linear functional through the normalisation
        ri = raw_i.clone().requires_grad_(True)
        clip.encode_image = lambda image, mask_ratio=0, _x=ri: _x
        I2, _, _ = clip.forward(torch.zeros(1), torch.zeros(1))
        gI = torch.randn(n, d, generator=g)
        (I2 * gI).sum().backward()
        np.savez(HERE / f"tail_{name}.npz", raw_i=np_(raw_i), raw_t=np_(raw_t), I=np_(I), T=np_(T),
                 s=np_(s), logit_scale_log=np_(clip.logit_scale), lpi=np_(lpi), lpt=np_(lpt),
                 gI=np_(gI), d_raw_i=np_(ri.grad))
        print("tail", name, float(s))


def _run_script(script, argv):
    old = sys.argv
    sys.argv = [script] + argv
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old


def gen_topk():
    for name, G, Q, d, k, bs, ties in [("a", 300, 20, 32, 10, 128, False),
                                       ("b", 64, 9, 16, 10, 32768, True),
                                       ("c", 7, 4, 8, 10, 4, False)]:
        g = torch.Generator().manual_seed(555 + G)
        gal = torch.randn(G, d, generator=g)
        gal = (gal / gal.norm(dim=-1, keepdim=True)).bfloat16().float()
        if ties:  # duplicated gallery rows -> exactly equal scores: pins the stable tie order
            gal[5] = gal[40]
            gal[41] = gal[40]
            gal[3] = gal[17]
        qry = torch.randn(Q, d, generator=g) + 2.0 * gal[(torch.arange(Q) * 33) % G]
        qry = (qry / qry.norm(dim=-1, keepdim=True)).bfloat16().float()
        image_ids = [1000000 + 7 * i for i in range(G)]
        text_ids = [5000 + i for i in range(Q)]
        with tempfile.TemporaryDirectory() as tmp:
            fi, ft = os.path.join(tmp, "img.jsonl"), os.path.join(tmp, "txt.jsonl")
            with open(fi, "w") as f:
                for i in range(G):
                    f.write(json.dumps({"image_id": image_ids[i], "feature": gal[i].tolist()}) + "\n")
            with open(ft, "w") as f:
                for i in range(Q):
                    f.write(json.dumps({"text_id": text_ids[i], "feature": qry[i].tolist()}) + "\n")
            out1, out2 = os.path.join(tmp, "t2i.jsonl"), os.path.join(tmp, "i2t.jsonl")
            _run_script(f"{REF}/cn_clip/eval/make_topk_predictions.py",
                        ["--image-feats", fi, "--text-feats", ft, "--top-k", str(k),
                         "--eval-batch-size", str(bs), "--output", out1])
            _run_script(f"{REF}/cn_clip/eval/make_topk_predictions_tr.py",
                        ["--image-feats", fi, "--text-feats", ft, "--top-k", str(k),
                         "--eval-batch-size", str(bs), "--output", out2])
            t2i = [json.loads(l) for l in open(out1)]
            i2t = [json.loads(l) for l in open(out2)]
        kk = min(k, G)
        kq = min(k, Q)
        np.savez(HERE / f"topk_{name}.npz", gallery=np_(gal), queries=np_(qry),
                 image_ids=np.array(image_ids), text_ids=np.array(text_ids), k=k, eval_batch_size=bs,
                 t2i_text_ids=np.array([o["text_id"] for o in t2i]),
                 t2i_image_ids=np.array([o["image_ids"] for o in t2i]).reshape(Q, kk),
                 i2t_image_ids=np.array([o["image_id"] for o in i2t]),
                 i2t_text_ids=np.array([o["text_ids"] for o in i2t]).reshape(G, kq))
        print("topk", name, t2i[0])


def gen_lora_loss():
    """train_lora.py:95-110 imported from the reference (its module-level `import lmdb` is stubbed:
    only the dataset class uses it)."""
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    import train_lora
    cases = {"a": (96, 64, 11, 0.5, 1 / 0.07, 0.05), "b": (300, 128, 12, 0.0, 20.0, 0.1),
             "c": (257, 72, 13, 0.5, 100.0, 0.05), "d": (64, 512, 14, 0.5, 1.0, 0.3)}
    for name, (n, d, seed, corr, scale, eps) in cases.items():
        img, txt = synth(n, d, seed, corr)
        g = torch.Generator().manual_seed(seed + 100)
        # un-normalised inputs: the function normalises them itself (train_lora.py:97-98)
        img = (img * (0.5 + torch.rand(n, 1, generator=g))).requires_grad_(True)
        txt = (txt * (0.5 + 2 * torch.rand(n, 1, generator=g))).requires_grad_(True)
        sc = torch.tensor(scale, requires_grad=True)
        loss = train_lora.contrastive_loss(img, txt, sc, label_smoothing=eps)
        loss.backward()
        np.savez(HERE / f"lora_loss_{name}.npz", img=np_(img), txt=np_(txt), scale=np.float32(scale),
                 eps=np.float32(eps), loss=np_(loss), dI=np_(img.grad), dT=np_(txt.grad), ds=np_(sc.grad))
        print("lora_loss", name, float(loss))


if __name__ == "__main__":
    install_shims()
    if len(sys.argv) > 1 and sys.argv[1] == "lora":
        gen_lora_loss()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "w1":
        gen_loss_w1()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "kd":
        gen_loss_kd()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "accum4":
        gen_loss_accum_big()
        sys.exit(0)
    gen_lora_loss()
    gen_loss_w1()
    gen_loss_accum()
    gen_loss_accum_big()
    gen_tail()
    gen_topk()
    gen_loss_dist()
    gen_loss_kd()
    print("golden fixtures written to", HERE)
