"""bench.py's reference arm runs without a GPU: check the JSON-line contract it prints (one line on
stdout, the keys the driver reads) and that ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--contract-test-n", "2048", "--no-secondary", *args], capture_output=True, text=True, env=env, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_contract_line():
    out = _run({"OMP_NUM_THREADS": "1"})          # torchrun's default: the arm must still use the host cores
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "contrastive_fwd_bwd_pairs_per_s" and d["unit"] == "pairs/s"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["steps"] == 2 and d["config"]["global_batch"] == 2048
    # measured, not extrapolated: ms_per_step x steps is time this process really spent
    assert d["ms_per_step"] * d["steps"] / 1e3 < 120


def test_both_arms_print_the_same_workload_string():
    sys.path.insert(0, str(ROOT))
    import importlib
    bench = importlib.import_module("bench")
    for W in (1, 2, 4, 8):
        assert f"{32768 // W} rows/rank" in bench.workload_string(W) and "configs[1]" in bench.workload_string(W)


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"}, "--gpus", "8").strip() == ""


def test_rank_shards_of_the_synthetic_batch_concatenate_to_the_full_batch():
    """bench.py's parity_check compares every rank's gradient slice with the single-GPU run on the full
    batch: that only means something if rank r's synthetic shard IS rows [r n_loc, (r+1) n_loc) of it."""
    sys.path.insert(0, str(ROOT))
    import importlib
    import torch
    bench = importlib.import_module("bench")
    full_i, full_t = bench.synth_features(16384, 0, 64)
    for W in (2, 4, 8):
        n_loc = 16384 // W
        parts = [bench.synth_features(n_loc, r * n_loc, 64) for r in range(W)]
        assert torch.equal(torch.cat([p[0] for p in parts]), full_i)
        assert torch.equal(torch.cat([p[1] for p in parts]), full_t)
    assert torch.equal(full_i.bfloat16().float(), full_i)                # 16-bit exact, unit norm up to rounding
    assert float((full_i.norm(dim=-1) - 1).abs().max()) < 1e-2
