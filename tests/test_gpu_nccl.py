"""Two real ranks over NCCL (needs >= 2 GPUs; skipped otherwise): the drop-in loss against the
golden fixtures the reference produced under gloo, and sharded retrieval against the oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, W, port, fixture, q):
    try:
        sys.path.insert(0, str(ROOT))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=W, device_id=torch.device("cuda", rank))
        from nans_clip_b200.loss import clip_contrastive_loss
        from nans_clip_b200.retrieval import topk_retrieve
        g = np.load(fixture)
        n_loc, gwg = int(g["n_loc"]), bool(g["gather_with_grad"])
        dev = torch.device("cuda", rank)
        img = torch.from_numpy(g["img"])[rank * n_loc:(rank + 1) * n_loc].to(dev).requires_grad_(True)
        txt = torch.from_numpy(g["txt"])[rank * n_loc:(rank + 1) * n_loc].to(dev).requires_grad_(True)
        ls = torch.tensor(float(g["logit_scale_log"]), device=dev, requires_grad=True)
        ok = []
        for dt, tol in ((torch.float16, 1e-3), (torch.bfloat16, 3e-3)):   # product default at the 1e-3 bar; bf16 extra
            img.grad = txt.grad = ls.grad = None
            loss, acc = clip_contrastive_loss(img, txt, ls.exp(), group=dist.group.WORLD, gather_with_grad=gwg,
                                              report_acc=True, feat_dtype=dt)
            loss.backward()
            ok += [abs(float(loss) - float(g["loss"][rank])) <= 1e-3 * abs(float(g["loss"][rank])),
                   abs(float(acc["i2t"]) - float(g["i2t"][rank])) < 1e-6]
            for got, want in ((img.grad, g["dI"][rank]), (txt.grad, g["dT"][rank])):
                want = torch.from_numpy(want)
                ok.append(float((got.cpu() - want).norm() / want.norm()) < tol)
            want = float(g["dlogit_scale_log"][rank])
            ok.append(abs(float(ls.grad) - want) <= 1e-3 * abs(want))
        # sharded retrieval
        gen = torch.Generator().manual_seed(5)
        gal = torch.nn.functional.normalize(torch.randn(5003, 128, generator=gen), dim=-1).half().float()
        qry = torch.nn.functional.normalize(torch.randn(77, 128, generator=gen), dim=-1).half().float()
        lo, hi = 5003 * rank // W, 5003 * (rank + 1) // W
        s, i = topk_retrieve(qry, gal[lo:hi], 10, group=dist.group.WORLD)
        rs, ri = torch.sort(qry @ gal.t(), dim=1, descending=True, stable=True)
        ok.append(bool(torch.equal(i.cpu(), ri[:, :10])))
        q.put((rank, ok, ""))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


@pytest.mark.parametrize("name", ["w2", "w2g"])
def test_two_ranks_nccl(name):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29801 + (name == "w2g")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(ROOT / "tests" / "golden" / f"loss_dist_{name}.npz"), q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert all(ok), f"rank {rank}: {ok} {info}"


def _accum_worker(rank, W, port, q):
    """Incremental accumulate path over NCCL: 4 chunks of 128 rows per rank (W * B = 256), calls
    j = 1, 3, 0 of one optimizer step; oracle = full loss on the spliced, concatenated batch."""
    try:
        sys.path.insert(0, str(ROOT))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=W, device_id=dev)
        from nans_clip_b200 import accum
        from oracle import clip_loss as OL
        A, B, D, s = 4, 128, 64, 14.2857
        gen = torch.Generator().manual_seed(17)
        n_loc = A * B
        img = torch.nn.functional.normalize(torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
        txt = torch.nn.functional.normalize(img + 0.8 * torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
        mine = slice(rank * n_loc, (rank + 1) * n_loc)
        cache_i = [img[mine][a * B:(a + 1) * B].to(dev) for a in range(A)]
        cache_t = [txt[mine][a * B:(a + 1) * B].to(dev) for a in range(A)]
        ok = [accum.eligible(cache_i, cache_t, B, W, 0.0)]
        for j in (1, 3, 0):
            new_i = torch.nn.functional.normalize(img + 0.1 * torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
            new_t = torch.nn.functional.normalize(txt + 0.1 * torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
            cur_i, cur_t = img.clone(), txt.clone()
            for r in range(W):
                blk = slice(r * n_loc + j * B, r * n_loc + (j + 1) * B)
                cur_i[blk], cur_t[blk] = new_i[blk], new_t[blk]
            want = OL.global_loss_and_grads(cur_i, cur_t, s, torch.float64)
            blk = slice(rank * n_loc + j * B, rank * n_loc + (j + 1) * B)
            ci = new_i[blk].to(dev).requires_grad_(True)
            ct = new_t[blk].to(dev).requires_grad_(True)
            sc = torch.tensor(s, device=dev, requires_grad=True)
            loss, acc = accum.incremental_accum_loss(ci, ct, sc, cache_i, cache_t, j, group=dist.group.WORLD,
                                                     report_acc=True)
            loss.backward()
            ok.append(abs(float(loss) - float(want["loss"])) <= 1e-3 * abs(float(want["loss"])))
            ok.append(abs(float(acc["i2t"]) - float(want["i2t"])) < 1e-6)
            for got, w in ((ci.grad, want["dI"][blk]), (ct.grad, want["dT"][blk])):
                ok.append(float((got.cpu().double() - w.double()).norm() / w.double().norm()) < 1e-3)
            ok.append(abs(float(sc.grad) - float(want["ds"])) <= 1e-3 * abs(float(want["ds"])))
        q.put((rank, ok, ""))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


def test_incremental_accumulate_two_ranks_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_accum_worker, args=(r, 2, 29811, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert all(ok), f"rank {rank}: {ok} {info}"


def _push_worker(rank, W, port, q):
    """The NVLink push exchange between REAL ranks (CUDA IPC peer buffers, flag-driven forward): several
    steps with changing features in both gather modes against the fp64 oracle on the global batch, the
    NCCL path on the same inputs, and the too-late-backward guard."""
    try:
        sys.path.insert(0, str(ROOT))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=W, device_id=dev)
        from nans_clip_b200 import exchange
        from nans_clip_b200.loss import clip_contrastive_loss
        from oracle import clip_loss as OL
        ok, info = [], []
        for n_loc, D, s in ((512, 256, 14.2857), (256, 512, 20.0)):
            for step in range(3):
                gwg = step == 1
                gen = torch.Generator().manual_seed(100 * n_loc + step)
                base = torch.randn(W * n_loc, D, generator=gen)
                I = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
                T = torch.nn.functional.normalize(0.5 * base + 0.5 * torch.randn(W * n_loc, D, generator=gen), dim=-1).half().float()
                want = OL.global_loss_and_grads(I, T, s, torch.float64)
                sl = slice(rank * n_loc, (rank + 1) * n_loc)
                outs = {}
                for mode in ("push", "nccl"):
                    os.environ["NANS_EXCHANGE"] = mode
                    a = I[sl].to(dev).requires_grad_(True)
                    b = T[sl].to(dev).requires_grad_(True)
                    sc = torch.tensor(s, device=dev, requires_grad=True)
                    loss, acc = clip_contrastive_loss(a, b, sc, group=dist.group.WORLD, gather_with_grad=gwg, report_acc=True)
                    loss.backward()
                    outs[mode] = (float(loss), a.grad.cpu(), b.grad.cpu(), float(sc.grad), float(acc["i2t"]))
                os.environ["NANS_EXCHANGE"] = "push"
                ex = exchange.for_group(dist.group.WORLD)
                ok.append(ex is not None and not ex.broken and ex.shape == (n_loc, D))   # the push path really ran
                mult = float(W) if gwg else 1.0
                for mode, (l, dI, dT, ds, a_i2t) in outs.items():
                    # absolute floors as in test_gpu_parity.grad_ok: a saturated softmax (second shape) leaves
                    # a loss of ~1e-5 and gradients of fp32-rounding size
                    N = W * n_loc
                    floor = 3e-5 * s / (2 * N) * (N ** 0.5) * mult
                    ok.append(abs(l - float(want["loss"])) <= 1e-3 * abs(float(want["loss"])) + 1e-6 * s)
                    ok.append(float((dI.double() - mult * want["dI"][sl]).norm()) <= 1e-3 * float((mult * want["dI"][sl]).norm()) + floor)
                    ok.append(float((dT.double() - mult * want["dT"][sl]).norm()) <= 1e-3 * float((mult * want["dT"][sl]).norm()) + floor)
                    ok.append(abs(ds - float(want["ds"])) <= 1e-3 * abs(float(want["ds"])) + 1e-6)
                    ok.append(abs(a_i2t - float(want["i2t"])) <= 1.5 / (W * n_loc))
                info.append((n_loc, step, outs["push"][0], outs["nccl"][0], float(want["loss"])))
        # backward after two later forwards must refuse (the slot was overwritten), not compute garbage
        a = I[sl].to(dev).requires_grad_(True)
        first, _ = clip_contrastive_loss(a, T[sl].to(dev), torch.tensor(s, device=dev), group=dist.group.WORLD)
        for _ in range(2):
            clip_contrastive_loss(I[sl].to(dev), T[sl].to(dev), torch.tensor(s, device=dev), group=dist.group.WORLD)
        try:
            first.backward()
            ok.append(False)
        except RuntimeError as exc:
            ok.append("push exchange" in str(exc))
        torch.cuda.synchronize()
        q.put((rank, ok, repr(info)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, [False], traceback.format_exc()))


def test_push_exchange_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_push_worker, args=(r, 2, 29821, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert all(ok), f"rank {rank}: {ok} {info}"
