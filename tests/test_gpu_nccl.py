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
        loss, acc = clip_contrastive_loss(img, txt, ls.exp(), group=dist.group.WORLD, gather_with_grad=gwg,
                                          report_acc=True, feat_dtype=torch.bfloat16)
        loss.backward()
        ok = [abs(float(loss) - float(g["loss"][rank])) <= 1e-3 * abs(float(g["loss"][rank])),
              abs(float(acc["i2t"]) - float(g["i2t"][rank])) < 1e-6]
        for got, want in ((img.grad, g["dI"][rank]), (txt.grad, g["dT"][rank])):
            want = torch.from_numpy(want)
            ok.append(float((got.cpu() - want).norm() / want.norm()) < 3e-3)
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
