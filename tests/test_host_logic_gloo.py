"""Multi-rank host logic of loss.py / retrieval.py under torch.distributed gloo on CPU
(world sizes 2 and 4), with the kernel entry points replaced by the CPU stand-ins of
tests/_fake_kernels.py.  Compared against the golden fixtures the REFERENCE produced under gloo."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def _init(rank, W, port):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from nans_clip_b200 import kernels as K
    import _fake_kernels
    _fake_kernels.install(K)


def _loss_worker(rank, W, port, fixture, q):
    try:
        _init(rank, W, port)
        from nans_clip_b200.loss import clip_contrastive_loss
        g = np.load(fixture)
        n_loc, gwg = int(g["n_loc"]), bool(g["gather_with_grad"])
        img = torch.from_numpy(g["img"])[rank * n_loc:(rank + 1) * n_loc].clone().requires_grad_(True)
        txt = torch.from_numpy(g["txt"])[rank * n_loc:(rank + 1) * n_loc].clone().requires_grad_(True)
        ls = torch.tensor(float(g["logit_scale_log"]), requires_grad=True)
        loss, acc = clip_contrastive_loss(img, txt, ls.exp(), group=dist.group.WORLD, gather_with_grad=gwg,
                                          report_acc=True, feat_dtype=torch.float32)
        loss.backward()
        ok = []
        ok.append(abs(float(loss) - float(g["loss"][rank])) <= 5e-6 * max(1.0, abs(float(g["loss"][rank]))))
        ok.append(abs(float(acc["i2t"]) - float(g["i2t"][rank])) < 1e-6)
        ok.append(abs(float(acc["t2i"]) - float(g["t2i"][rank])) < 1e-6)
        for got, want in ((img.grad, g["dI"][rank]), (txt.grad, g["dT"][rank])):
            want = torch.from_numpy(want)
            ok.append(float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-9)
        want = float(g["dlogit_scale_log"][rank])
        ok.append(abs(float(ls.grad) - want) <= 1e-4 * abs(want) + 1e-8)
        q.put((rank, ok, float(loss)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure instead of a hang
        import traceback
        q.put((rank, [False], traceback.format_exc()))


def _spawn(worker, W, port, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, W, port, *args, q)) for r in range(W)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(W)]
    for p in procs:
        p.join(timeout=60)
    return sorted(res, key=lambda x: x[0])


@pytest.mark.parametrize("name,port", [("w2", 29701), ("w2g", 29702), ("w4", 29703), ("w4g", 29704)])
def test_distributed_loss_matches_reference(name, port):
    g = np.load(GOLDEN / f"loss_dist_{name}.npz")
    res = _spawn(_loss_worker, int(g["W"]), port, str(GOLDEN / f"loss_dist_{name}.npz"))
    for rank, ok, info in res:
        assert all(ok), f"rank {rank}: {ok} {info}"


def _kd_worker(rank, W, port, fixture, q):
    """get_loss with args.distillation against the reference's own output (loss_kd_*.npz): student /
    teacher gather orders and the gradient through the gather under gather_with_grad."""
    try:
        if W > 1:
            _init(rank, W, port)
        else:
            sys.path.insert(0, str(ROOT))
            sys.path.insert(0, str(ROOT / "tests"))
            from nans_clip_b200 import kernels as K
            import _fake_kernels
            _fake_kernels.install(K)
        import types
        import torch.nn as nn
        from nans_clip_b200.training import train as TR
        TR.FEAT_DTYPE = torch.float32   # CPU stand-in kernels
        g = np.load(fixture)
        n_loc = int(g["n_loc"])
        sl = slice(rank * n_loc, (rank + 1) * n_loc)

        class Stub(nn.Module):
            def __init__(self):
                super().__init__()
                self.img = nn.Parameter(torch.from_numpy(g["img"])[sl].clone())
                self.txt = nn.Parameter(torch.from_numpy(g["txt"])[sl].clone())
                self.logit_scale = nn.Parameter(torch.tensor(float(g["logit_scale_log"])))

            def forward(self, images, texts, mask_ratio=0):
                return self.img, self.txt, self.logit_scale.exp()

        teacher = types.SimpleNamespace()
        teacher.module = types.SimpleNamespace(get_feature=lambda images: torch.from_numpy(g["teacher"])[sl])
        args = types.SimpleNamespace(accum_freq=1, mask_ratio=0, distillation=True, aggregate=W > 1,
                                     gather_with_grad=bool(g["gather_with_grad"]), local_device_rank=0,
                                     report_training_batch_acc=False, kd_loss_weight=float(g["kd_loss_weight"]))
        model = Stub()
        total, _ = TR.get_loss(model, None, None, nn.CrossEntropyLoss(), nn.CrossEntropyLoss(), args,
                               teacher_model=teacher)
        total.backward()
        ok = [abs(float(total) - float(g["loss"][rank])) <= 5e-6 * abs(float(g["loss"][rank]))]
        for got, want in ((model.img.grad, g["dI"][rank]), (model.txt.grad, g["dT"][rank])):
            want = torch.from_numpy(want)
            ok.append(float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-9)
        want = float(g["dlogit_scale_log"][rank])
        ok.append(abs(float(model.logit_scale.grad) - want) <= 1e-4 * abs(want) + 1e-8)
        q.put((rank, ok, float(total)))
        if W > 1:
            dist.barrier()
            dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


@pytest.mark.parametrize("name,port", [("w1", 29731), ("w2", 29732), ("w2g", 29733)])
def test_distillation_branch_matches_reference(name, port):
    g = np.load(GOLDEN / f"loss_kd_{name}.npz")
    for rank, ok, info in _spawn(_kd_worker, int(g["W"]), port, str(GOLDEN / f"loss_kd_{name}.npz")):
        assert all(ok), f"rank {rank}: {ok} {info}"


def _smoothing_worker(rank, W, port, gwg, q):
    """Label smoothing across ranks: the column sums / diagonal sum are all-reduced, the fix-up uses
    the global sums and the rank's rows (oracle: smoothed CE on the concatenated batch)."""
    try:
        _init(rank, W, port)
        from nans_clip_b200.loss import clip_contrastive_loss
        from oracle import clip_loss as OL
        n_loc, D, s, eps = 40, 24, 15.0, 0.1
        gen = torch.Generator().manual_seed(5)
        img = torch.nn.functional.normalize(torch.randn(W * n_loc, D, generator=gen), dim=-1)
        txt = torch.nn.functional.normalize(img + 0.7 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
        want = OL.global_smoothed_loss_and_grads(img, txt, s, eps, torch.float64)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        I = img[sl].clone().requires_grad_(True)
        T = txt[sl].clone().requires_grad_(True)
        sc = torch.tensor(s, requires_grad=True)
        loss, _ = clip_contrastive_loss(I, T, sc, group=dist.group.WORLD, gather_with_grad=gwg,
                                        feat_dtype=torch.float32, label_smoothing=eps)
        loss.backward()
        mult = float(W) if gwg else 1.0
        ok = [abs(float(loss) - float(want["loss"])) <= 1e-5 * abs(float(want["loss"]))]
        for got, w in ((I.grad, want["dI"][sl]), (T.grad, want["dT"][sl])):
            ok.append(float((got.double() - mult * w.double()).abs().max()) <= 1e-4 * mult * float(w.abs().max()))
        ok.append(abs(float(sc.grad) - float(want["ds"])) <= 1e-4 * abs(float(want["ds"])) + 1e-9)
        q.put((rank, ok, float(loss)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


@pytest.mark.parametrize("gwg,port", [(False, 29721), (True, 29722)])
def test_label_smoothing_under_gloo(gwg, port):
    for rank, ok, info in _spawn(_smoothing_worker, 2, port, gwg):
        assert all(ok), f"rank {rank}: {ok} {info}"


def _accum_worker(rank, W, port, gwg, q):
    """Incremental accumulate path (accum.py) across ranks: 4 chunks of 128 rows per rank, calls
    j = 1, 3, 0 of one optimizer step with re-forwarded chunks that differ from the cache; oracle =
    the full loss on the spliced, concatenated batch."""
    try:
        _init(rank, W, port)
        from nans_clip_b200 import accum
        from oracle import clip_loss as OL
        A, B, D, s = 4, 128, 24, 12.0
        gen = torch.Generator().manual_seed(17)
        n_loc = A * B
        img = torch.nn.functional.normalize(torch.randn(W * n_loc, D, generator=gen), dim=-1)
        txt = torch.nn.functional.normalize(img + 0.8 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
        mine = slice(rank * n_loc, (rank + 1) * n_loc)
        cache_i = [img[mine][a * B:(a + 1) * B].clone() for a in range(A)]
        cache_t = [txt[mine][a * B:(a + 1) * B].clone() for a in range(A)]
        assert accum.eligible(cache_i, cache_t, B, W, 0.0)
        ok = []
        for j in (1, 3, 0):
            # every rank re-forwards its chunk j: the global batch differs from the cache in W blocks
            new_i = torch.nn.functional.normalize(img + 0.1 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
            new_t = torch.nn.functional.normalize(txt + 0.1 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
            cur_i, cur_t = img.clone(), txt.clone()
            for r in range(W):
                blk = slice(r * n_loc + j * B, r * n_loc + (j + 1) * B)
                cur_i[blk], cur_t[blk] = new_i[blk], new_t[blk]
            want = OL.global_loss_and_grads(cur_i, cur_t, s, torch.float64)
            blk = slice(rank * n_loc + j * B, rank * n_loc + (j + 1) * B)
            ci = new_i[blk].clone().requires_grad_(True)
            ct = new_t[blk].clone().requires_grad_(True)
            sc = torch.tensor(s, requires_grad=True)
            loss, acc = accum.incremental_accum_loss(ci, ct, sc, cache_i, cache_t, j, group=dist.group.WORLD,
                                                     gather_with_grad=gwg, report_acc=True, feat_dtype=torch.float32)
            loss.backward()
            mult = float(W) if gwg else 1.0
            ok.append(abs(float(loss) - float(want["loss"])) <= 1e-5 * abs(float(want["loss"])))
            ok.append(abs(float(acc["i2t"]) - float(want["i2t"])) < 1e-6 and abs(float(acc["t2i"]) - float(want["t2i"])) < 1e-6)
            for got, w in ((ci.grad, want["dI"][blk]), (ct.grad, want["dT"][blk])):
                ok.append(float((got.double() - mult * w.double()).abs().max()) <= 1e-4 * mult * float(w.abs().max()))
            ok.append(abs(float(sc.grad) - float(want["ds"])) <= 1e-4 * abs(float(want["ds"])) + 1e-9)
        q.put((rank, ok, ""))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


@pytest.mark.parametrize("gwg,port", [(False, 29731), (True, 29732)])
def test_incremental_accumulate_under_gloo(gwg, port):
    for rank, ok, info in _spawn(_accum_worker, 2, port, gwg):
        assert all(ok), f"rank {rank}: {ok} {info}"


def _strip_split_worker(rank, W, port, gwg, q):
    """n_loc = 256: the tile-aligned path (local block, then one single-strip launch per gathered
    tensor sharing their slots) against the oracle on the concatenated batch."""
    try:
        _init(rank, W, port)
        from nans_clip_b200.loss import clip_contrastive_loss
        from oracle import clip_loss as OL
        n_loc, D, s = 256, 32, 20.0
        gen = torch.Generator().manual_seed(99)
        base = torch.randn(W * n_loc, D, generator=gen)
        img = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
        txt = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(W * n_loc, D, generator=gen), dim=-1)
        want = OL.global_loss_and_grads(img, txt, s, torch.float64)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        I = img[sl].clone().requires_grad_(True)
        T = txt[sl].clone().requires_grad_(True)
        sc = torch.tensor(s, requires_grad=True)
        loss, acc = clip_contrastive_loss(I, T, sc, group=dist.group.WORLD, gather_with_grad=gwg,
                                          report_acc=True, feat_dtype=torch.float32)
        loss.backward()
        mult = float(W) if gwg else 1.0
        ok = [abs(float(loss) - float(want["loss"])) <= 1e-5 * abs(float(want["loss"]))]
        for got, w in ((I.grad, want["dI"][sl]), (T.grad, want["dT"][sl])):
            ok.append(float((got.double() - mult * w.double()).abs().max()) <= 1e-4 * mult * float(w.abs().max()))
        ok.append(abs(float(sc.grad) - float(want["ds"])) <= 1e-4 * abs(float(want["ds"])) + 1e-9)
        q.put((rank, ok, float(loss)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


@pytest.mark.parametrize("gwg,port", [(False, 29711), (True, 29712)])
def test_tile_aligned_strip_split_path(gwg, port):
    for rank, ok, info in _spawn(_strip_split_worker, 2, port, gwg):
        assert all(ok), f"rank {rank}: {ok} {info}"


def _topk_worker(rank, W, port, q):
    try:
        _init(rank, W, port)
        from nans_clip_b200 import retrieval
        g = torch.Generator().manual_seed(7)
        G, Q, D, k = 203, 17, 32, 10
        gal = torch.randn(G, D, generator=g)
        gal[50] = gal[150]  # a tie that straddles two shards
        qry = torch.randn(Q, D, generator=g)
        lo, hi = G * rank // W, G * (rank + 1) // W

        class FakeShard(retrieval.GalleryShard):
            def __init__(self, gallery, device, feat_dtype, index_offset, gallery16=None):
                self.g32 = gallery.float()
                self.g16 = gallery.float()
                self.feat_dtype = torch.float32
                self.index_offset = index_offset

        retrieval.GalleryShard = FakeShard
        s, i = retrieval.topk_retrieve(qry, gal[lo:hi], k, group=dist.group.WORLD, device="cpu")
        sc = qry @ gal.t()
        rs, ri = torch.sort(sc, dim=1, descending=True, stable=True)
        q.put((rank, [bool(torch.equal(i, ri[:, :k])), bool(torch.allclose(s, rs[:, :k]))], ""))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


def test_sharded_retrieval_merge_under_gloo():
    for rank, ok, info in _spawn(_topk_worker, 2, 29711):
        assert all(ok), f"rank {rank}: {ok} {info}"
