"""The C-ABI boundary without a GPU: the library builds/loads, exports every symbol the header
declares, and fails loudly (never silently) when no sm_100 device is present."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from nans_clip_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "nans_clip.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(nans_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ["nans_l2norm_cast", "nans_l2norm_bwd", "nans_clip_loss_fwd_phase",
                 "nans_clip_loss_fwd_finalize", "nans_clip_loss_fwd", "nans_clip_loss_bwd",
                 "nans_topk_ip", "nans_topk_merge", "nans_last_error", "nans_version",
                 "nans_device_check"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in include/nans_clip.h but not exported"
    assert sorted(_lib.EXPORTS) == declared_functions(), "ctypes signatures out of sync with the header"
    assert lib.nans_version() == 100


def test_workspace_queries_need_no_device():
    lib = _lib.load()
    assert lib.nans_clip_loss_fwd_phase_slots(4096, 32768, 512) >= 1
    assert lib.nans_clip_loss_fwd_workspace_bytes(4096, 4) > 4 * 5 * 2 * 4096 * 4
    assert lib.nans_clip_loss_bwd_workspace_bytes(4096, 32768, 512) >= 2 * 4096 * 512 * 4
    assert lib.nans_topk_ip_workspace_bytes(30000, 125000, 512, 16) >= 2 * 30000 * 16 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.nans_device_check() == -2
    assert "no CPU path" in _lib.last_error() or "sm_100" in _lib.last_error()
    rc = lib.nans_l2norm_cast(None, 0, 4, 64, 64, None, 1, None, None, 1, None)
    assert rc == -2
    with pytest.raises(_lib.NansError):
        _lib.check(rc)
    rc = lib.nans_topk_merge(None, None, 1, 4, 10, None, None, None)
    assert rc == -2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_refuses_cpu_tensors():
    from nans_clip_b200.loss import clip_contrastive_loss
    x = torch.nn.functional.normalize(torch.randn(8, 64), dim=-1)
    with pytest.raises(RuntimeError):
        clip_contrastive_loss(x, x, torch.tensor(10.0))
    from nans_clip_b200.retrieval import GalleryShard
    with pytest.raises(Exception):
        GalleryShard(x).search(x, 5)


def test_forward_slot_and_workspace_queries_are_consistent():
    """Host-only planning queries of the forward: 1 <= slots <= min(column tiles, 32) for two-strip and
    single-strip launches alike, the two single-strip flags agree, workspace grows with slots."""
    lib = _lib.load()
    for n_loc in (1, 100, 256, 4096, 32768):
        for ncols in (1, 255, 256, 4096, 28672, 32768, 1000000):
            both = lib.nans_clip_loss_fwd_phase_slots(n_loc, ncols, 512)
            one = lib.nans_clip_loss_fwd_phase_slots_flags(n_loc, ncols, 512, 2)
            assert one == lib.nans_clip_loss_fwd_phase_slots_flags(n_loc, ncols, 512, 4)
            assert both == lib.nans_clip_loss_fwd_phase_slots_flags(n_loc, ncols, 512, 0)
            tiles = -(-ncols // 256)
            assert 1 <= both <= max(1, min(tiles, 32)) and 1 <= one <= max(1, min(tiles, 32))
            w1 = lib.nans_clip_loss_fwd_workspace_bytes(n_loc, both)
            w2 = lib.nans_clip_loss_fwd_workspace_bytes(n_loc, both + one)
            assert w2 > w1 >= 2 * n_loc * 4 * 5 * both
    assert lib.nans_clip_loss_fwd_phase_slots(0, 10, 512) == 0
    assert lib.nans_clip_loss_bwd_workspace_bytes(4096, 32768, 512) >= 2 * 4096 * 512 * 4
    assert lib.nans_topk_ip_workspace_bytes(30000, 1000000, 512, 16) > 0


def test_exchange_layout_needs_no_device_and_is_well_formed():
    """nans_xchg_layout (host only): regions in order, aligned, non-overlapping; sizes follow the job."""
    from nans_clip_b200 import exchange as X
    for W, n_loc, D in ((2, 256, 64), (8, 4096, 512), (8, 8192, 768), (16, 256, 1024)):
        d = X.make_desc(W, 1, [4096 * (r + 1) for r in range(W)], n_loc, D, 64)
        N, pad = W * n_loc, (n_loc + 3) // 4 * 4
        assert (d.world, d.rank, d.n_loc, d.D) == (W, 1, n_loc, D)
        assert d.feat_off == 0 and d.lse_len == 2 * pad + 8
        assert d.lse_off >= 2 * 2 * N * D * 2 and d.lse_off % 1024 == 0
        assert d.fflag_off >= d.lse_off + 2 * W * d.lse_len * 4 and d.fflag_off % 1024 == 0
        assert d.lflag_off >= d.fflag_off + 2 * W * (n_loc // 64) * 4 and d.lflag_off % 1024 == 0
        assert d.bytes == d.lflag_off + 1024 == X.layout_bytes(W, n_loc, D)
        assert [d.base[r] for r in range(W)] == [4096 * (r + 1) for r in range(W)]
    with pytest.raises(_lib.NansError):
        X.layout_bytes(4, 300, 64)        # rows per rank must be whole 256-column tiles
    with pytest.raises(_lib.NansError):
        X.layout_bytes(17, 256, 64)       # more peers than NANS_MAX_PEERS
    assert X.eligible(4096, 512, 8) and not X.eligible(4096, 512, 1) and not X.eligible(4000, 512, 8)
    assert not X.eligible(4096, 2048, 8)  # D > 1024: the narrow backward's exchange mode does not cover it
