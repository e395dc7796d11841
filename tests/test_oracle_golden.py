"""The oracle (oracle/) against the golden fixtures recorded from the REFERENCE ITSELF
(tests/golden/make_golden.py ran n571e/NanS-CLIP's own get_loss / CLIP.forward /
make_topk_predictions.py).  This is what pins the oracle; CPU only."""
import glob

import numpy as np
import pytest
import torch

from oracle import clip_loss as OL
from oracle import topk as OT


def _load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e", "f"])
def test_local_loss_matches_reference(golden_dir, name):
    g = _load(golden_dir / f"loss_w1_{name}.npz")
    img, txt = _t(g["img"]).requires_grad_(True), _t(g["txt"]).requires_grad_(True)
    ls = torch.tensor(float(g["logit_scale_log"]), requires_grad=True)
    loss, acc = OL.local_loss(img, txt, ls.exp(), report_acc=True)
    loss.backward()
    # same torch ops as the reference on the same machine: bit-for-bit up to reduction order
    assert torch.allclose(loss.detach(), _t(g["loss"]), rtol=1e-6, atol=1e-7)
    assert float(acc["i2t"]) == float(g["i2t"]) and float(acc["t2i"]) == float(g["t2i"])
    assert torch.allclose(img.grad, _t(g["dI"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(txt.grad, _t(g["dT"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(ls.grad, _t(g["dlogit_scale_log"]), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", ["w2", "w2g", "w4", "w4g"])
def test_rank_loss_matches_reference_under_gloo(golden_dir, name):
    g = _load(golden_dir / f"loss_dist_{name}.npz")
    W, n_loc, gwg = int(g["W"]), int(g["n_loc"]), bool(g["gather_with_grad"])
    img, txt = _t(g["img"]), _t(g["txt"])
    ib = [img[r * n_loc:(r + 1) * n_loc] for r in range(W)]
    tb = [txt[r * n_loc:(r + 1) * n_loc] for r in range(W)]
    s = torch.tensor(float(np.exp(g["logit_scale_log"])), dtype=torch.float32)
    for r in range(W):
        loss, acc, dI, dT, ds = OL.rank_loss(ib, tb, s, r, gwg, report_acc=True)
        assert torch.allclose(loss, _t(g["loss"][r]), rtol=2e-6, atol=1e-7)
        assert abs(float(acc["i2t"]) - float(g["i2t"][r])) < 1e-7
        assert abs(float(acc["t2i"]) - float(g["t2i"][r])) < 1e-7
        scale = float(_t(g["dI"][r]).abs().max()) + 1e-12
        assert float((dI - _t(g["dI"][r])).abs().max()) <= 2e-5 * scale
        assert float((dT - _t(g["dT"][r])).abs().max()) <= 2e-5 * scale
        # the fixture holds d/d(log scale) = s * d/ds
        ref_ds = float(g["dlogit_scale_log"][r]) / float(s)
        assert abs(float(ds) - ref_ds) <= 2e-5 * abs(ref_ds) + 1e-9


def test_gather_with_grad_is_w_times_plain(golden_dir):
    """the reference's own outputs show dI(gather_with_grad) = W * dI(plain)."""
    a, b = _load(golden_dir / "loss_dist_w2.npz"), _load(golden_dir / "loss_dist_w2g.npz")
    assert np.allclose(b["dI"], 2 * a["dI"], rtol=1e-4, atol=1e-8)
    assert np.allclose(b["dT"], 2 * a["dT"], rtol=1e-4, atol=1e-8)
    assert np.allclose(b["dlogit_scale_log"], a["dlogit_scale_log"], rtol=1e-5)
    assert np.allclose(a["loss"], b["loss"], rtol=1e-6)


@pytest.mark.parametrize("j", [0, 1, 2])
def test_accumulate_path_matches_reference(golden_dir, j):
    g = _load(golden_dir / f"loss_accum_j{j}.npz")
    A, B = int(g["A"]), int(g["B"])
    img, txt = _t(g["img"]), _t(g["txt"])
    cache_i = [img[a * B:(a + 1) * B] for a in range(A)]
    cache_t = [txt[a * B:(a + 1) * B] for a in range(A)]
    ci = cache_i[j].clone().requires_grad_(True)
    ct = cache_t[j].clone().requires_grad_(True)
    ls = torch.tensor(float(g["logit_scale_log"]), requires_grad=True)
    loss, acc = OL.local_loss(OL.accum_splice(cache_i, ci, j), OL.accum_splice(cache_t, ct, j), ls.exp(), True)
    loss.backward()
    assert torch.allclose(loss.detach(), _t(g["loss"]), rtol=1e-6)
    assert torch.allclose(ci.grad, _t(g["dI"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(ct.grad, _t(g["dT"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(ls.grad, _t(g["dlogit_scale_log"]), rtol=1e-5)


def _bf16(bits):
    return torch.from_numpy(bits).view(torch.bfloat16).float()


@pytest.mark.parametrize("j", [2, 0])
def test_accumulate_path_with_a_changed_chunk_matches_reference(golden_dir, j):
    """4 chunks of 256 rows, the re-forwarded chunk differs from its cached copy (fixtures of the
    incremental accumulate path)."""
    g = _load(golden_dir / f"loss_accum4_j{j}.npz")
    A, B = int(g["A"]), int(g["B"])
    img, txt = _bf16(g["img_bf16"]), _bf16(g["txt_bf16"])
    cache_i = [img[a * B:(a + 1) * B] for a in range(A)]
    cache_t = [txt[a * B:(a + 1) * B] for a in range(A)]
    ci = _bf16(g["new_img_bf16"]).requires_grad_(True)
    ct = _bf16(g["new_txt_bf16"]).requires_grad_(True)
    ls = torch.tensor(float(g["logit_scale_log"]), requires_grad=True)
    loss, acc = OL.local_loss(OL.accum_splice(cache_i, ci, j), OL.accum_splice(cache_t, ct, j), ls.exp(), True)
    loss.backward()
    assert torch.allclose(loss.detach(), _t(g["loss"]), rtol=1e-6)
    assert torch.allclose(ci.grad, _t(g["dI"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(ct.grad, _t(g["dT"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(ls.grad, _t(g["dlogit_scale_log"]), rtol=1e-5)
    assert abs(float(acc["i2t"]) - float(g["i2t"])) < 1e-6


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_forward_tail_matches_reference(golden_dir, name):
    g = _load(golden_dir / f"tail_{name}.npz")
    raw_i = _t(g["raw_i"]).requires_grad_(True)
    I, T, s = OL.forward_tail(raw_i, _t(g["raw_t"]), _t(g["logit_scale_log"]))
    assert torch.equal(I.detach(), _t(g["I"])) and torch.equal(T, _t(g["T"]))
    assert torch.allclose(s, _t(g["s"]))
    lpi, lpt = OL.logits_pair(I.detach(), T, s)
    assert torch.allclose(lpi, _t(g["lpi"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(lpt, _t(g["lpt"]), rtol=1e-6, atol=1e-6)
    (I * _t(g["gI"])).sum().backward()
    assert torch.allclose(raw_i.grad, _t(g["d_raw_i"]), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_topk_matches_reference_script(golden_dir, name):
    g = _load(golden_dir / f"topk_{name}.npz")
    gal, qry = _t(g["gallery"]), _t(g["queries"])
    k = int(g["k"])
    image_ids = g["image_ids"].tolist()
    text_ids = g["text_ids"].tolist()
    # text -> image (make_topk_predictions.py)
    _, pos = OT.topk_vectorised(gal, qry, k)
    got = np.array(image_ids)[pos.numpy()]
    assert np.array_equal(got, g["t2i_image_ids"])
    assert np.array_equal(g["t2i_text_ids"], np.array(text_ids))
    # image -> text (make_topk_predictions_tr.py)
    _, pos = OT.topk_vectorised(qry, gal, k)
    assert np.array_equal(np.array(text_ids)[pos.numpy()], g["i2t_text_ids"])
    # the literal per-query loop agrees with the vectorised restatement
    for q in range(min(4, qry.shape[0])):
        ids, _ = OT.topk_literal(image_ids, gal.numpy(), qry[q].numpy(), k, int(g["eval_batch_size"]))
        assert ids == g["t2i_image_ids"][q].tolist()


def test_tie_order_is_stable_gallery_order(golden_dir):
    """topk_b has duplicated gallery rows: equal scores must come out in ascending position."""
    g = _load(golden_dir / "topk_b.npz")
    ids = g["t2i_image_ids"]
    base = 1000000
    pos = (ids - base) // 7
    sc = _t(g["queries"]) @ _t(g["gallery"]).t()
    for q in range(ids.shape[0]):
        s = sc[q, pos[q]]
        for a in range(ids.shape[1] - 1):
            assert s[a] >= s[a + 1]
            if s[a] == s[a + 1]:
                assert pos[q, a] < pos[q, a + 1]
    assert any((sc[q, pos[q]][:-1] == sc[q, pos[q]][1:]).any() for q in range(ids.shape[0]))


def test_analytic_known_answers():
    """Cases the reference has no tests for."""
    # N = 1 -> loss 0
    x = torch.nn.functional.normalize(torch.randn(1, 16), dim=-1)
    assert float(OL.local_loss(x, x, torch.tensor(14.0))[0]) == 0.0
    # all rows identical -> loss = ln N
    n = 37
    x = torch.nn.functional.normalize(torch.randn(1, 32), dim=-1).repeat(n, 1)
    assert abs(float(OL.local_loss(x, x, torch.tensor(10.0))[0]) - np.log(n)) < 1e-5
    # orthonormal rows, I = T: loss = ln(1 + (N-1) e^{-s})
    n, s = 16, 3.0
    q, _ = torch.linalg.qr(torch.randn(64, n))
    x = q.t().contiguous()
    want = np.log(1 + (n - 1) * np.exp(-s))
    assert abs(float(OL.local_loss(x, x, torch.tensor(s))[0]) - want) < 1e-5


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_lora_label_smoothed_loss_matches_reference(golden_dir, name):
    """train_lora.py:95-110 run from the reference itself -> oracle restatement, and the O(N D)
    decomposition the CUDA path relies on (csrc/smooth.cu): smoothed = plain + closed-form terms."""
    g = np.load(golden_dir / f"lora_loss_{name}.npz")
    img = torch.from_numpy(g["img"]).requires_grad_(True)
    txt = torch.from_numpy(g["txt"]).requires_grad_(True)
    sc = torch.tensor(float(g["scale"]), requires_grad=True)
    eps = float(g["eps"])
    loss = OL.lora_contrastive_loss(img, txt, sc, eps)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert torch.allclose(img.grad, torch.from_numpy(g["dI"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(txt.grad, torch.from_numpy(g["dT"]), rtol=1e-4, atol=1e-7)
    assert abs(float(sc.grad) - float(g["ds"])) <= 1e-4 * abs(float(g["ds"])) + 1e-8
    # decomposition on normalised features (fp64)
    I = OL.normalize(img.detach().double())
    T = OL.normalize(txt.detach().double())
    s, n = float(sc), I.shape[0]
    plain = OL.global_loss_and_grads(I, T, s, torch.float64)
    smooth = OL.global_smoothed_loss_and_grads(I, T, s, eps, torch.float64)
    isum, tsum, diag = I.sum(0), T.sum(0), (I * T).sum()
    corr = eps / n * diag - eps / n ** 2 * (isum @ tsum)
    assert abs(float(plain["loss"] + s * corr) - float(smooth["loss"])) < 1e-10
    assert abs(float(plain["ds"] + corr) - float(smooth["ds"])) < 1e-10
    assert torch.allclose(plain["dI"] + s * eps / n * T - s * eps / n ** 2 * tsum, smooth["dI"], atol=1e-12)
    assert torch.allclose(plain["dT"] + s * eps / n * I - s * eps / n ** 2 * isum, smooth["dT"], atol=1e-12)
    assert abs(float(smooth["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


def test_all_fixtures_are_covered(golden_dir):
    names = sorted(p.split("/")[-1] for p in glob.glob(str(golden_dir / "*.npz")))
    # 25 fixtures checked against the oracle in this file + the 3 loss_kd_* ones, which pin the drop-in's
    # distillation branch directly (tests/test_host_logic_gloo.py; the KD term is outside the fused path)
    assert len(names) == 28, names
