"""bench.py — the reference's headline metric for the hot path, measured on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--only loss|retrieval|...]
    (N > 1: launched by torch.distributed.run, one rank per GPU over NCCL)

The ONE JSON line describes BASELINE.json configs[1] (`metric` / `value` / `e2e` / `roofline`):
ViT-B/16-width contrastive loss forward + backward, GLOBAL batch 32768 x D=512 sharded over the N
ranks (strong scaling: the global batch is fixed), synthetic unit-norm features as SURVEY.md §8d.
A step is one fwd+bwd of the loss through the public call (`clip_contrastive_loss`: 16-bit cast,
feature exchange, fused forward, lse/scalar exchange, fused backward).
  value : pairs/s = 32768 / step time, inputs resident in HBM, CUDA events, max over ranks
  e2e   : the same step with HOST (pinned) fp32 features: H2D of every step's features inside the timed
          region (pipelined: step k+1's copy runs under step k's compute; the strictly serial figure is
          reported beside it as e2e.serial_ms_per_step), loss scalar read back every step
  roofline : the dominant kernel (fused backward), algorithmic flops / its measured launch time,
             against the measured bf16 burst peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle (fp32 torch port of the reference path) on this box's host cores (N=1 only)
  parity_check : N > 1 only — every rank's loss, d(scale) and gradient slice against the single-GPU
                 fused path on the full batch (which tests/ hold to the oracle); exits non-zero on failure

The other BASELINE.json configs and kernels ride in the same line under `workloads` (time-boxed, each
with its own roofline): `retrieval` (configs[4]: top-10, Q=30000 x G=1e6, gallery sharded x N),
`loss_d1024` (configs[3]), `accum_n65536_d768_a8` (configs[2], the gradient-accumulation call site),
`l2norm` (kernel (1), HBM GB/s).  `--only X` runs one of them alone.

`--impl reference` times the reference's own algorithm on the host cores, rank 0 only, on the same
config string: the oracle port (the reference is PyTorch code: its arithmetic IS these torch ops;
/root/reference itself cannot travel to the GPU box), really at N=32768, every step measured.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_GLOBAL = 32768
D = 512
LOGIT_SCALE = 14.285714  # exp(ln(1/0.07)), model.py:356
RET_Q, RET_G, RET_K = 30000, 1000000, 10
WORKLOADS = ("retrieval", "loss_d1024", "accum_n65536_d768_a8", "l2norm")


def workload_string(W: int) -> str:
    """The same string in both arms (the driver compares them)."""
    return (f"ViT-B/16-width contrastive loss fwd+bwd, global batch {N_GLOBAL}, D={D}, "
            f"{N_GLOBAL // W} rows/rank (BASELINE.json configs[1])")


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def synth_features(n_rows, row0, d, seed=1235, corr=0.5):
    """correlated pairs, row-normalised, rounded to bf16 and back (16-bit exact).
    Rows are generated in blocks keyed by (seed, block) so every rank builds only its shard."""
    blk = 4096
    outs_i, outs_t = [], []
    b0, b1 = row0 // blk, (row0 + n_rows + blk - 1) // blk
    for b in range(b0, b1):
        g = torch.Generator().manual_seed(seed * 100003 + b)
        base = torch.randn(blk, d, generator=g)
        img = corr * base + (1 - corr) * torch.randn(blk, d, generator=g)
        txt = corr * base + (1 - corr) * torch.randn(blk, d, generator=g)
        outs_i.append(img)
        outs_t.append(txt)
    img = torch.cat(outs_i)[row0 - b0 * blk: row0 - b0 * blk + n_rows]
    txt = torch.cat(outs_t)[row0 - b0 * blk: row0 - b0 * blk + n_rows]
    img = (img / img.norm(dim=-1, keepdim=True)).bfloat16().float()
    txt = (txt / txt.norm(dim=-1, keepdim=True)).bfloat16().float()
    return img, txt


def synth_features_device(n_rows, d, dev, seed, corr=0.5):
    """The same recipe on the device (the secondary workloads; identical on every rank for one seed)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.randn(n_rows, d, device=dev, generator=g)
    img = corr * base + (1 - corr) * torch.randn(n_rows, d, device=dev, generator=g)
    txt = corr * base + (1 - corr) * torch.randn(n_rows, d, device=dev, generator=g)
    del base
    img = torch.nn.functional.normalize(img, dim=-1).bfloat16().float()
    txt = torch.nn.functional.normalize(txt, dim=-1).bfloat16().float()
    return img, txt


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, world, rank, local
    return None, 1, 0, local


def timed_steps(step_fn, steps, warmup, flush, dist, dev):
    """W warm-ups, then K steps each bracketed by CUDA events on the current stream; the L2 is
    flushed (256 MB write) before every step, outside the events.  Returns total ms (max over ranks)."""
    for _ in range(warmup):
        flush.zero_()
        step_fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ================================================================================================
# CPU legs (the only places that execute oracle/): reported baselines, never the product path
# ================================================================================================
def host_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; a CPU leg is rank 0 alone on the host cores,
    # so it takes all of them (what a plain `python bench.py --impl reference` gets by default)
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        n = max(1, os.cpu_count() or 1)
    torch.set_num_threads(n)
    return n


def cpu_loss_steps(steps, warmup, budget_s, n=None):
    """The oracle (port of model.py:412-415 + train.py:87-115 + autograd backward) REALLY at
    N = 32768, D = 512 on the host cores: `warmup` untimed steps (at most 1 is enough to fault the
    20 GB of temporaries in), then up to `steps` timed ones inside `budget_s` (at least 2).
    Returns (per-step seconds list, threads)."""
    from oracle import clip_loss as OL
    threads = host_threads()
    img, txt = synth_features(n or N_GLOBAL, 0, D)
    t0 = time.perf_counter()
    OL.global_loss_and_grads(img[:2048], txt[:2048], LOGIT_SCALE)
    est = None
    for _ in range(min(warmup, 1)):
        t1 = time.perf_counter()
        OL.global_loss_and_grads(img, txt, LOGIT_SCALE)
        est = time.perf_counter() - t1
    times = []
    while len(times) < steps:
        if len(times) >= 2 and est is not None and (time.perf_counter() - t0) + est > budget_s:
            break
        t1 = time.perf_counter()
        OL.global_loss_and_grads(img, txt, LOGIT_SCALE)
        times.append(time.perf_counter() - t1)
        est = times[-1]
    return times, threads


def cpu_loss_baseline(reps=2):
    times, threads = cpu_loss_steps(reps, 1, 1e9)
    per = sum(times) / len(times)
    return {"value": N_GLOBAL / per, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"fp32 torch fwd+bwd of the oracle port REALLY at N={N_GLOBAL}, D={D} (no extrapolation): "
                      f"1 warm-up + {len(times)} timed steps, {per:.2f} s/step"}


def cpu_topk_baseline(n_vec=512, n_lit=8):
    """BASELINE.md §3 legs B3 (the literal per-query loop of make_topk_predictions.py:71-85) and B4 (its
    vectorised restatement) at G = 1e6, on a bounded sample of the 30000 queries."""
    from oracle import topk as OT
    threads = host_threads()
    g = torch.Generator().manual_seed(77)
    gal = torch.nn.functional.normalize(torch.randn(RET_G, D, generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(n_vec, D, generator=g) + 0.5 * gal[(torch.arange(n_vec) * 33) % RET_G], dim=-1)
    OT.topk_vectorised(gal, q[:64], RET_K)
    t1 = time.perf_counter()
    _, vi = OT.topk_vectorised(gal, q, RET_K, block=128)
    per_vec = (time.perf_counter() - t1) / n_vec
    ids = list(range(1000000, 1000000 + RET_G))
    gal_np = gal.numpy()
    t1 = time.perf_counter()
    same = True
    for i in range(n_lit):
        got, _ = OT.topk_literal(ids, gal_np, q[i].numpy(), RET_K)
        same &= [g_ - 1000000 for g_ in got] == vi[i].tolist()
    per_lit = (time.perf_counter() - t1) / n_lit
    return {"value": 1.0 / per_vec, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"B4: vectorised fp32 Q@G^T + stable sort, {n_vec} of {RET_Q} queries x G={RET_G}",
            "literal_loop": {"value": 1.0 / per_lit, "unit": "queries/s", "cores": threads,
                             "sample": f"B3: literal loop of make_topk_predictions.py:71-85, {n_lit} queries x G={RET_G} "
                                       f"({per_lit:.2f} s/query; 30000 queries would take {per_lit * RET_Q / 3600:.1f} h)",
                             "same_ids_as_vectorised": bool(same)}}


def run_reference_arm(args):
    """The reference's own CPU path for the same metric/config (rank 0 only; other ranks exit 0)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    W = max(1, args.gpus)
    if args.only == "retrieval":
        base = cpu_topk_baseline()
        line = {"impl": "reference", "metric": "retrieval_queries_per_s", "value": base["value"], "unit": base["unit"],
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * RET_Q / base["value"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"top-{RET_K} retrieval Q={RET_Q} G={RET_G} D={D}"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    budget = float(os.environ.get("NANS_REF_BUDGET_S", "240"))
    n = args.contract_test_n or N_GLOBAL
    times, threads = cpu_loss_steps(args.steps, args.warmup, budget, n)
    per = sum(times) / len(times)
    value = n / per
    sample = (f"oracle port (fp32 torch, train.py:87-115 + autograd) REALLY at N={n}, D={D}: "
              f"{min(args.warmup, 1)} warm-up + {len(times)} timed steps of {args.steps} requested "
              f"(time budget {budget:.0f} s), {per:.2f} s/step, no extrapolation")
    line = {"impl": "reference", "metric": "contrastive_fwd_bwd_pairs_per_s", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(W) if n == N_GLOBAL else f"CONTRACT TEST ONLY: global batch {n}",
                       "global_batch": n, "D": D, "logit_scale": LOGIT_SCALE,
                       "where": f"host cores of the box ({threads} threads), rank 0 only"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_secondary:
        try:
            line["workloads"] = {"retrieval": cpu_topk_baseline(n_vec=256, n_lit=8)}
        except Exception as exc:  # the headline line must survive a secondary leg
            line["workloads"] = {"retrieval": {"error": repr(exc)}}
    emit(line)


# ================================================================================================
# the loss step (configs[1]) and its relatives
# ================================================================================================
def parity_check(dist, W, rank, dev, group, feat_dt):
    """N > 1: the multi-rank path against the single-GPU fused path on the full batch (rank 0 computes
    it and broadcasts), in both gather modes.  The single-GPU path is what tests/ hold to the oracle and
    to the reference's fixtures at 1e-3; here every rank compares its loss, d(scale) and its dI / dT slice
    (x W under gather_with_grad)."""
    from nans_clip_b200.loss import clip_contrastive_loss
    n_loc = N_GLOBAL // W
    ref = torch.empty((2, N_GLOBAL, D), dtype=torch.float32, device=dev)
    sc = torch.empty(2, dtype=torch.float32, device=dev)
    if rank == 0:
        img, txt = synth_features(N_GLOBAL, 0, D)
        a = img.to(dev).requires_grad_(True)
        b = txt.to(dev).requires_grad_(True)
        s = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
        loss, _ = clip_contrastive_loss(a, b, s, group=None, feat_dtype=feat_dt)
        loss.backward()
        ref[0], ref[1] = a.grad, b.grad
        sc[0], sc[1] = loss.detach(), s.grad
        del a, b
    dist.broadcast(ref, 0)
    dist.broadcast(sc, 0)
    img, txt = synth_features(n_loc, rank * n_loc, D)
    worst = torch.zeros(4, dtype=torch.float64, device=dev)
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    for gwg in (False, True):
        a = img.to(dev).requires_grad_(True)
        b = txt.to(dev).requires_grad_(True)
        s = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
        loss, _ = clip_contrastive_loss(a, b, s, group=group, gather_with_grad=gwg, feat_dtype=feat_dt)
        loss.backward()
        mult = float(W) if gwg else 1.0
        errs = torch.stack([
            (loss.detach() - sc[0]).abs() / sc[0].abs(),
            (s.grad - sc[1]).abs() / sc[1].abs(),
            (a.grad - mult * ref[0, sl]).norm() / (mult * ref[0, sl]).norm(),
            (b.grad - mult * ref[1, sl]).norm() / (mult * ref[1, sl]).norm()]).double()
        worst = torch.maximum(worst, torch.nan_to_num(errs, nan=1e9))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    w = [float(x) for x in worst.tolist()]
    tol = 1e-3
    return {"against": "single-GPU fused path on the full batch (rank 0), both gather modes, max over ranks",
            "loss_rel": w[0], "dscale_rel": w[1], "dI_rel": w[2], "dT_rel": w[3], "tol": tol,
            "ok": bool(max(w) <= tol)}


def loss_step_bench(args, dist, W, rank, dev, n_global, d, with_e2e, kernel_rooflines, sampler_index):
    """Times the loss step at (n_global, d) sharded over the W ranks.  Returns a dict (rank 0 uses it)."""
    from nans_clip_b200 import kernels as K
    from nans_clip_b200.loss import clip_contrastive_loss
    n_loc = n_global // W
    group = dist.group.WORLD if dist is not None else None
    feat_dt = torch.float16
    peaks = measured_peaks()
    if d == D and n_global == N_GLOBAL:
        img, txt = synth_features(n_loc, rank * n_loc, d)
    else:
        fi, ft = synth_features_device(n_global, d, dev, 4000 + d)
        img, txt = fi[rank * n_loc:(rank + 1) * n_loc].clone(), ft[rank * n_loc:(rank + 1) * n_loc].clone()
        del fi, ft
    img_d = img.to(dev).requires_grad_(True)
    txt_d = txt.to(dev).requires_grad_(True)
    s = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        img_d.grad = txt_d.grad = s.grad = None
        loss, _ = clip_contrastive_loss(img_d, txt_d, s, group=group, feat_dtype=feat_dt)
        loss.backward()
        return loss

    l0 = K.LAUNCHES
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = (K.LAUNCHES - l0) // max(args.warmup, 1)
    run_step, graphed = step, False
    exchange_kind = "none (1 GPU)"
    push_active = False
    if dist is not None:
        from nans_clip_b200 import exchange
        ex = exchange.for_group(group)
        push_active = ex is not None and not ex.broken and ex.shape == (n_loc, d)
        exchange_kind = ("NVLink push into peer-mapped buffers (CUDA IPC; rows and flags by copy-engine copies under "
                         "the forward), flag-driven forward, no collective per step"
                         if push_active else "NCCL all-gather (ProcessGroupNCCL), overlapped with the local block")
    # A whole step (cast + push, forward, lse exchange, backward) is plain kernels on one stream with the
    # step counter in device memory, so it can be captured once and replayed.  NCCL collectives inside a
    # captured step hung on this image (round 1), so the NCCL path is timed eagerly.
    graphs = []
    if args.graph and (dist is None or push_active):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            # The push exchange double-buffers by step parity and its copy-engine copies carry host
            # addresses, so ONE captured step is only valid for every other step: two graphs are captured
            # (even step, odd step) and replayed alternately, each still ONE step between flush and events.
            # thread_local: ProcessGroupNCCL's watchdog thread polls CUDA events while this thread captures.
            for _ in range(2 if push_active else 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    step()
                graphs.append(g)
            torch.cuda.synchronize()
            graphed = True
        except Exception as exc:  # fall back to eager launches, say so in the line
            print(f"[bench] CUDA graph capture failed, timing eager launches: {exc!r}", file=sys.stderr)
            graphs, graphed = [], False
        if dist is not None:   # every rank must time the same thing
            flag = torch.tensor([1 if graphed else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                graphs, graphed = [], False
    replayed = [0]
    if graphed:
        def run_step():
            graphs[replayed[0] % len(graphs)].replay()
            replayed[0] += 1
        for _ in range(2):
            run_step()
        torch.cuda.synchronize()
    sampler = ClockSampler(sampler_index).start()
    ms = timed_steps(run_step, args.steps, 0, flush, dist, dev) / args.steps
    clocks = sampler.stop()
    if graphed and len(graphs) == 2 and replayed[0] % 2:
        run_step()   # leave the exchange on the parity the host expects for the eager steps that follow
    out = {"ms_per_step": ms, "launches_per_step": launches_per_step, "graphed": graphed, "clocks": clocks,
           "n_loc": n_loc, "exchange": exchange_kind, "graphs": len(graphs)}
    if graphed:   # also the eager number, for the record
        out["eager_ms_per_step"] = timed_steps(step, args.steps, 2, flush, dist, dev) / args.steps
    if with_e2e:
        img_h, txt_h = img.pin_memory(), txt.pin_memory()
        loss_h = torch.zeros(1, dtype=torch.float32).pin_memory()
        seen = []

        def step_e2e_body():
            i = img_h.to(dev, non_blocking=True).requires_grad_(True)     # H2D of this step's features
            t = txt_h.to(dev, non_blocking=True).requires_grad_(True)
            s.grad = None
            loss, _ = clip_contrastive_loss(i, t, s, group=group, feat_dtype=feat_dt)
            loss.backward()
            loss_h.copy_(loss.detach().reshape(1), non_blocking=True)     # D2H of the step's result
            return i, t

        def step_e2e():
            step_e2e_body()
            torch.cuda.current_stream().synchronize()
            seen.append(float(loss_h[0]))                                 # the host reads the loss every step

        e2e_step, e2e_graphed = step_e2e, False
        if graphed:
            # the same end-to-end step (H2D copies from pinned memory, the public call, backward, D2H of the
            # loss) captured and replayed: at N > 1 the eager version is bound by ~0.5 ms of host launches
            try:
                egraphs = []
                for _ in range(len(graphs)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        step_e2e_body()
                    egraphs.append(g)
                torch.cuda.synchronize()
                ecount = [0]

                def step_e2e_graph():
                    egraphs[ecount[0] % len(egraphs)].replay()
                    ecount[0] += 1
                    torch.cuda.current_stream().synchronize()
                    seen.append(float(loss_h[0]))

                e2e_step, e2e_graphed = step_e2e_graph, True
            except Exception as exc:
                print(f"[bench] e2e graph capture failed, timing the eager e2e step: {exc!r}", file=sys.stderr)
            if dist is not None:
                flag = torch.tensor([1 if e2e_graphed else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) == 0:
                    e2e_step, e2e_graphed = step_e2e, False
        out["e2e_ms"] = timed_steps(e2e_step, args.steps, min(args.warmup, 3), flush, dist, dev) / args.steps
        if e2e_graphed and len(egraphs) == 2 and ecount[0] % 2:
            e2e_step()
        if e2e_graphed:
            out["e2e_eager_ms"] = timed_steps(step_e2e, args.steps, 2, flush, dist, dev) / args.steps
        out["e2e_graphed"] = e2e_graphed
        out["e2e_loss_seen"] = seen[-1] if seen else None
        out["e2e_serial_ms"] = out["e2e_ms"]
        # ---- the same end-to-end step with the input copies PIPELINED (what a data loader with pinned memory
        # and non_blocking copies does): the H2D copy of step k + 1's features runs on a copy stream while step
        # k computes.  Every step's copy and every step's loss read-back are inside the timed region (K steps,
        # K copies: the first step copies its own inputs and waits for them, the last one prefetches nothing).
        try:
            out["e2e_ms"] = pipelined_e2e(args, dist, dev, flush, graphs if graphed else [], img_h, txt_h, loss_h, s,
                                          group, feat_dt, seen)
            out["e2e_pipelined"] = True
        except Exception as exc:
            print(f"[bench] pipelined e2e failed, reporting the serial e2e step: {exc!r}", file=sys.stderr)
            out["e2e_pipelined"] = False
        out["e2e_loss_seen"] = seen[-1] if seen else None
    if kernel_rooflines:
        # ---- forward and backward kernels timed alone, on this rank ----
        with torch.no_grad():
            I16, _, _ = K.l2norm_cast(img_d.detach(), feat_dt, normalize=False)
            T16, _, _ = K.l2norm_cast(txt_d.detach(), feat_dt, normalize=False)
            if W > 1:
                I_all = torch.empty((n_global, d), dtype=feat_dt, device=dev)
                T_all = torch.empty((n_global, d), dtype=feat_dt, device=dev)
                dist.all_gather_into_tensor(I_all, I16)
                dist.all_gather_into_tensor(T_all, T16)
            else:
                I_all, T_all = I16, T16
            s_dev = s.detach().reshape(1)
            slots = K.fwd_phase_slots(n_loc, n_global, d)
            ws = K.fwd_workspace(n_loc, slots, dev)

            def fwd_only():
                K.fwd_phase(I16, T16, T_all, I_all, col_global_begin=0, label_begin=rank * n_loc, s_dev=s_dev,
                            with_acc=False, ws=ws, slot_begin=0)

            fwd_only()
            lse, _sc, _ = K.fwd_finalize(n_loc, slots, rank * n_loc, s_dev, False, ws)
            if W > 1:
                lse_g = torch.empty((W * 2, n_loc), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(lse_g, lse.contiguous())
                lse_all = lse_g.view(W, 2, n_loc).permute(1, 0, 2).reshape(2, n_global)
            else:
                lse_all = lse
            pad = (n_global + 3) // 4 * 4
            lse_pad = torch.empty((2, pad), dtype=torch.float32, device=dev)[:, :n_global]
            lse_pad.copy_(lse_all)
            gout = torch.ones(1, device=dev)
            # lse min / max in the kernel's order-preserving int encoding (the step gets them from the
            # exchange-finish launch): the timed launch is the backward kernel alone, as in the step
            bits = torch.stack([lse_pad.min(), lse_pad.max()]).view(torch.int32)
            mm = torch.where(bits >= 0, bits, bits ^ 0x7fffffff).contiguous()

            def bwd_only():
                K.bwd(I16, T16, T_all, I_all, label_begin=rank * n_loc, s_dev=s_dev, lse_all=lse_pad,
                      grad_out=gout, grad_mult=1.0, row_begin=0, row_count=n_loc, out_dtype=torch.float32,
                      lse_minmax=mm)

            kb = max(3, min(args.steps, 10))
            bwd_ms = timed_steps(bwd_only, kb, 2, flush, None, dev) / kb
            fwd_ms = timed_steps(fwd_only, kb, 2, flush, None, dev) / kb
        # algorithmic work per rank: backward 4 * n_loc * N * D (dI and dT; the logit
        # recompute is not counted), forward 2 * n_loc * N * D
        bwd_alg = 4.0 * n_loc * n_global * d
        fwd_alg = 2.0 * n_loc * n_global * d
        out["roofline"] = {"kernel": "clip_bwd_np_kernel", "bound": "tensor", "achieved": bwd_alg / (bwd_ms / 1e3) / 1e12,
                           "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                           "frac": bwd_alg / (bwd_ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": None,
                           "launch_ms": bwd_ms, "algorithmic_flops_per_launch": bwd_alg,
                           "hardware_tflops": 2 * bwd_alg / (bwd_ms / 1e3) / 1e12,
                           "peak_source": f"{peaks['source']} burst bf16 (kernel timed alone)"}
        out["roofline_fwd"] = {"kernel": "clip_fwd_kernel", "bound": "tensor", "achieved": fwd_alg / (fwd_ms / 1e3) / 1e12,
                               "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": fwd_alg / (fwd_ms / 1e3) / 1e12 / peaks["bf16_tflops"], "launch_ms": fwd_ms,
                               "hardware_tflops": 2 * fwd_alg / (fwd_ms / 1e3) / 1e12}
    del flush
    return out


def pipelined_e2e(args, dist, dev, flush, graphs, img_h, txt_h, loss_h, s, group, feat_dt, seen):
    """K end-to-end steps with double-buffered device inputs: returns ms per step (max over ranks)."""
    from nans_clip_b200.loss import clip_contrastive_loss
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    bufs = [(torch.empty(img_h.shape, device=dev).requires_grad_(True), torch.empty(txt_h.shape, device=dev).requires_grad_(True))
            for _ in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(p):   # H2D of one step's features into input buffer set p, on the copy stream
        copy_stream.wait_event(consumed[p])          # the step that last read this set has finished
        with torch.cuda.stream(copy_stream), torch.no_grad():
            bufs[p][0].copy_(img_h, non_blocking=True)
            bufs[p][1].copy_(txt_h, non_blocking=True)
        copied[p].record(copy_stream)

    def compute_body(p):
        i, t = bufs[p]
        i.grad = t.grad = s.grad = None
        loss, _ = clip_contrastive_loss(i, t, s, group=group, feat_dtype=feat_dt)
        loss.backward()
        loss_h.copy_(loss.detach().reshape(1), non_blocking=True)     # D2H of the step's result

    pgraphs = []
    if graphs:   # same launch mode as the device-timed value: one captured step per input buffer set (= per
        for p in range(2):   # parity of the double-buffered exchange), replayed alternately
            consumed[p].record(main)
            issue_copy(p)
        torch.cuda.synchronize()
        for p in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                compute_body(p)
            pgraphs.append(g)
        torch.cuda.synchronize()
    count = [0]

    def run(k, first, last):
        p = count[0] & 1
        if first:
            issue_copy(p)                            # this step's own inputs: nothing to hide them behind
        if not last:
            issue_copy(p ^ 1)                        # the next step's inputs, under this step's compute
        main.wait_event(copied[p])
        if pgraphs:
            pgraphs[p].replay()
        else:
            compute_body(p)
        consumed[p].record(main)
        count[0] += 1
        main.synchronize()
        seen.append(float(loss_h[0]))                # the host reads the loss every step

    for p in range(2):
        consumed[p].record(main)
    K_, Wm = args.steps, min(args.warmup, 3)
    for k in range(Wm):                              # warm-up: same loop, untimed
        flush.zero_()
        run(k, k == 0, k == Wm - 1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    evs = []
    for k in range(K_):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(k, k == 0, k == K_ - 1)
        e.record()
        evs.append((a, e))
    torch.cuda.synchronize()
    if pgraphs and count[0] % 2:   # leave the exchange on the parity the host expects
        run(0, True, True)
    if dist is not None:
        dist.barrier()
    total = torch.tensor([sum(a.elapsed_time(e) for a, e in evs)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    return float(total.item()) / K_


def bench_loss_d1024(args, dist, W, rank, dev, local):
    """BASELINE.json configs[3]: ViT-H/14 width, N = 32768, D = 1024 over the W ranks."""
    n, d = N_GLOBAL, 1024
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup, sub.graph = max(3, min(args.steps, 6)), 3, False
    r = loss_step_bench(sub, dist, W, rank, dev, n, d, False, True, local)
    peaks = measured_peaks()
    alg = 6.0 * n * n * d
    return {"config": f"BASELINE.json configs[3]: contrastive loss fwd+bwd, global batch {n}, D={d}, {n // W} rows/rank",
            "value": n / (r["ms_per_step"] / 1e3), "unit": "pairs/s", "ms_per_step": r["ms_per_step"], "steps": sub.steps,
            "algorithmic_tflops": alg / (r["ms_per_step"] / 1e3) / 1e12,
            "frac_of_bf16_peak_algorithmic": alg / (r["ms_per_step"] / 1e3) / 1e12 / (W * peaks["bf16_tflops"]),
            "roofline": r["roofline"], "roofline_fwd": r["roofline_fwd"], "clocks": r["clocks"]}


def bench_accum(args, dist, W, rank, dev, local):
    """BASELINE.json configs[2]: ViT-L/14 width, global batch 65536, D = 768, gradient-accumulation path
    with A = 8 (train.py:205-247): one optimizer step = A calls of the drop-in get_loss, each with the
    re-forwarded chunk j (B = N / (W A) rows per rank) + backward.  FLIP masking only changes the towers."""
    import types
    import torch.nn as nn
    from nans_clip_b200.training.train import get_loss
    n, d, A = 65536, 768, 8
    n_loc = n // W
    B = n_loc // A
    fi, ft = synth_features_device(n, d, dev, 6500)
    img, txt = fi[rank * n_loc:(rank + 1) * n_loc].clone(), ft[rank * n_loc:(rank + 1) * n_loc].clone()
    del fi, ft
    ls = torch.tensor(2.6593, device=dev)

    class Chunk(nn.Module):
        def __init__(self):
            super().__init__()
            self.j = 0
            self.logit_scale = nn.Parameter(ls.clone())
            self.img = nn.Parameter(img.clone())
            self.txt = nn.Parameter(txt.clone())

        def forward(self, images, texts, mask_ratio=0):
            sl = slice(self.j * B, (self.j + 1) * B)
            return self.img[sl], self.txt[sl], self.logit_scale.exp()

    model = Chunk()
    a = types.SimpleNamespace(accum_freq=A, mask_ratio=0.5, distillation=False, aggregate=W > 1, gather_with_grad=False,
                              local_device_rank=local, report_training_batch_acc=False)
    crit = nn.CrossEntropyLoss()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def optimizer_step():
        # the cache pass of train.py:205-232 is the towers' (out of scope): the cached features exist
        cache_i = [img[c * B:(c + 1) * B].clone() for c in range(A)]
        cache_t = [txt[c * B:(c + 1) * B].clone() for c in range(A)]
        model.zero_grad(set_to_none=True)
        for j in range(A):
            model.j = j
            total, _ = get_loss(model, None, None, crit, crit, a, cache_i, cache_t, j)
            total.backward()

    steps = max(2, min(args.steps, 3))
    sampler = ClockSampler(local).start()
    ms = timed_steps(optimizer_step, steps, 2, flush, dist, dev) / steps
    clocks = sampler.stop()
    peaks = measured_peaks()
    # algorithmic work of the REFERENCE's A calls: each recomputes the whole forward
    per_call = 2.0 * n * n * d + 4.0 * (n / A) * n * d
    return {"config": f"BASELINE.json configs[2]: accumulate path, global batch {n}, D={d}, A={A}, {n_loc} rows/rank "
                      f"({B}-row chunks), through the drop-in get_loss",
            "ms_per_optimizer_step": ms, "ms_per_call": ms / A, "steps": steps,
            "value": n / (ms / 1e3), "unit": "pairs/s (global batch / optimizer step of A get_loss calls)",
            "reference_algorithmic_tflop_per_call": per_call / 1e12,
            "frac_of_bf16_peak_reference_algorithmic": per_call * A / (ms / 1e3) / 1e12 / (W * peaks["bf16_tflops"]),
            "note": "the incremental forward (accum.py) does 4 N_l N D / A + one full forward per optimizer step "
                    "instead of the reference's A full forwards: the fraction is against the reference's flops",
            "clocks": clocks}


def bench_l2norm(args, dev, local):
    """Kernel (1) alone: [1e6, 512] fp32 -> fp16, normalise + cast (+ inv_norm), HBM-bound."""
    from nans_clip_b200 import kernels as K
    rows = 1000000
    x = torch.randn(rows, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peaks = measured_peaks()

    def run():
        K.l2norm_cast(x, torch.float16, normalize=True, want_inv_norm=True)

    steps = max(5, min(args.steps, 20))
    sampler = ClockSampler(local).start()
    ms = timed_steps(run, steps, 3, flush, None, dev) / steps
    clocks = sampler.stop()
    nbytes = rows * D * (4 + 2) + rows * 4
    return {"config": f"kernel (1): L2-normalise + fp16 cast of [{rows}, {D}] fp32 (+ inv_norm), kernel alone",
            "ms_per_launch": ms, "steps": steps,
            "roofline": {"kernel": "l2norm_cast_kernel", "bound": "hbm", "achieved": nbytes / (ms / 1e3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_launch": nbytes, "traffic": None,
                         "peak_source": f"{peaks['source']} copy bandwidth (a copy writes as much as it reads; this "
                                        "kernel writes a third of what it reads, so > 1.0 is possible)"},
            "clocks": clocks}


def bench_retrieval_core(args, dist, W, rank, dev, local, steps):
    """BASELINE.json configs[4]: top-10 of 30000 queries over a 1M-row gallery sharded x W.  SURVEY.md
    §8d data: every query carries its planted match (gallery row (q * 33) mod G mixed in at 0.5), ids
    offset by 1e6.  Also times an ASCENDING-score gallery (the worst case of threshold-gated insertion)."""
    from nans_clip_b200 import kernels as K
    from nans_clip_b200.retrieval import GalleryShard
    peaks = measured_peaks()
    lo, hi = RET_G * rank // W, RET_G * (rank + 1) // W
    g = torch.Generator(device=dev).manual_seed(4242)
    gal_full = torch.nn.functional.normalize(torch.randn(RET_G, D, device=dev, generator=g), dim=-1).bfloat16().float()
    planted = (torch.arange(RET_Q, device=dev) * 33) % RET_G
    qry = torch.nn.functional.normalize(torch.randn(RET_Q, D, device=dev, generator=g) + 0.5 * gal_full[planted] * (D ** 0.5),
                                        dim=-1).bfloat16().float()
    gal = gal_full[lo:hi].clone()
    del gal_full
    shard = GalleryShard(gal, dev, torch.float16, 1000000 + lo)
    q16, _, _ = K.l2norm_cast(qry, torch.float16, normalize=False)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    qry_h = qry.cpu().pin_memory()

    def merge(s, i):
        if W > 1:
            all_s = torch.empty((W * RET_Q, RET_K), dtype=s.dtype, device=dev)
            all_i = torch.empty((W * RET_Q, RET_K), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(all_s, s)
            dist.all_gather_into_tensor(all_i, i)
            s, i = K.topk_merge(all_s.view(W, RET_Q, RET_K), all_i.view(W, RET_Q, RET_K))
        return s, i

    def step():
        return merge(*K.topk_ip(q16, shard.g16, qry, shard.g32, RET_K, 16, shard.index_offset))[1]

    def step_e2e():
        q = qry_h.to(dev, non_blocking=True)
        return merge(*shard.search(q, RET_K))[1].cpu()

    l0 = K.LAUNCHES
    for _ in range(3):
        idx = step()
    torch.cuda.synchronize()
    lps = (K.LAUNCHES - l0) // 3
    # the planted match must be every query's top-1 (cosine ~0.45 against ~0.2 for the best of 1e6 strangers)
    top1_ok = bool((idx[:, 0] == planted + 1000000).all())
    sampler = ClockSampler(local).start()
    ms = timed_steps(step, steps, 0, flush, dist, dev) / steps
    clocks = sampler.stop()
    ne = max(2, steps // 4)
    e2e_ms = timed_steps(step_e2e, ne, 1, flush, dist, dev) / ne
    # ascending gallery: this shard's rows re-ordered by their score against the mean query direction
    u = torch.nn.functional.normalize(qry.mean(dim=0), dim=0)
    order = torch.argsort(shard.g32 @ u)
    asc = GalleryShard(shard.g32[order].contiguous(), dev, torch.float16, 1000000 + lo)

    def step_asc():
        return merge(*K.topk_ip(q16, asc.g16, qry, asc.g32, RET_K, 16, asc.index_offset))[1]

    na = max(2, steps // 4)
    asc_ms = timed_steps(step_asc, na, 1, flush, dist, dev) / na
    flops = 2.0 * RET_Q * (hi - lo) * D
    return {"config": f"BASELINE.json configs[4]: top-{RET_K} text->image retrieval, Q={RET_Q}, G={RET_G}, D={D}, gallery "
                      f"sharded x{W}, planted matches (SURVEY 8d), fp16 candidate pass (k_cand 16) + fp32 rescoring",
            "value": RET_Q / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms, "steps": steps,
            "planted_top1_found": top1_ok,
            "e2e": {"value": RET_Q / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": RET_Q * D * 4, "d2h_bytes_per_step": RET_Q * RET_K * 8},
            "ascending_gallery": {"ms_per_step": asc_ms, "value": RET_Q / (asc_ms / 1e3), "unit": "queries/s",
                                  "what": "gallery rows sorted by ascending score against the mean query direction"},
            "gpu_launches_per_step": lps,
            "roofline": {"kernel": "topk_floor_kernel+topk_sweep_kernel+topk_finalize_kernel", "bound": "tensor",
                         "achieved": flops / (ms / 1e3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": None,
                         "algorithmic_flops_per_launch": flops,
                         "peak_source": f"{peaks['source']} burst bf16"},
            "clocks": clocks}


def bench_main(args):
    from nans_clip_b200 import kernels as K

    dist, W, rank, local = dist_setup()
    dev = torch.device("cuda", local)
    assert N_GLOBAL % W == 0
    peaks = measured_peaks()
    line = None
    rc = 0
    if args.only in (None, "loss"):
        pc = parity_check(dist, W, rank, dev, dist.group.WORLD, torch.float16) if W > 1 else None
        r = loss_step_bench(args, dist, W, rank, dev, N_GLOBAL, D, True, True, local)
        ms, e2e_ms, n_loc = r["ms_per_step"], r["e2e_ms"], r["n_loc"]
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("clip_bwd_np_kernel", {}).get(f"W{W}")
            except Exception:
                traffic = None
        r["roofline"]["traffic"] = traffic
        step_alg = 6.0 * N_GLOBAL * N_GLOBAL * D
        line = {
            "metric": "contrastive_fwd_bwd_pairs_per_s", "value": N_GLOBAL / (ms / 1e3), "unit": "pairs/s",
            "n_gpus": W, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload_string(W),
                       "global_batch": N_GLOBAL, "D": D, "operand_dtype": "fp16 (fp32 accumulate)",
                       "logit_scale": LOGIT_SCALE, "parallelism": f"dp{W}",
                       "l2": "flushed (256 MB write) before every timed step",
                       "launch": ("cuda-graph replay, one captured step per launch" +
                                  (" (two graphs, even / odd step of the double-buffered exchange, alternating)"
                                   if r["graphs"] == 2 else "") if r["graphed"] else "eager"),
                       "exchange": r["exchange"],
                       "step_algorithmic_tflop": step_alg / 1e12},
            "algorithmic_tflops": step_alg / (ms / 1e3) / 1e12,
            # the step is timed alone between L2 flushes (milliseconds): the burst peak is its denominator
            "frac_of_bf16_peak_algorithmic": step_alg / (ms / 1e3) / 1e12 / (W * peaks["bf16_tflops"]),
            "e2e": {"value": N_GLOBAL / (e2e_ms / 1e3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * n_loc * D * 4, "d2h_bytes_per_step": 4,
                    "launch": ("cuda-graph replay of the compute of a step" if r["graphed"] else "eager"),
                    "pipeline": ("the H2D copy of step k+1's features (pinned host -> device, copy stream, double-buffered "
                                 "inputs) runs while step k computes; every step's copy and loss read-back are inside the "
                                 "timed region (K steps, K copies)") if r.get("e2e_pipelined") else "none (serial)",
                    "serial_ms_per_step": r.get("e2e_serial_ms"), "serial_eager_ms_per_step": r.get("e2e_eager_ms"),
                    "loss_read_back": r.get("e2e_loss_seen")},
            "gpu_launches": r["launches_per_step"] * args.steps,
            "gpu_launches_per_step": r["launches_per_step"],
            "roofline": r["roofline"], "roofline_fwd": r["roofline_fwd"],
            "clocks": r["clocks"],
        }
        if "eager_ms_per_step" in r:
            line["eager_ms_per_step"] = r["eager_ms_per_step"]
        if pc is not None:
            line["parity_check"] = pc
            if not pc["ok"]:
                rc = 3
    workloads = {}
    todo = [w for w in WORKLOADS if (args.only == w or (args.only is None and not args.no_secondary))]
    for name in todo:
        try:
            torch.cuda.empty_cache()
            if name == "retrieval":
                workloads[name] = bench_retrieval_core(args, dist, W, rank, dev, local, max(4, min(args.steps, 10)))
            elif name == "loss_d1024":
                workloads[name] = bench_loss_d1024(args, dist, W, rank, dev, local)
            elif name == "accum_n65536_d768_a8":
                workloads[name] = bench_accum(args, dist, W, rank, dev, local)
            elif name == "l2norm":
                workloads[name] = bench_l2norm(args, dev, local)
        except Exception as exc:  # a secondary workload must not take the headline line down
            import traceback
            traceback.print_exc()
            workloads[name] = {"error": repr(exc)}
            if dist is not None:
                raise  # ranks would desynchronise: fail loudly instead
    if rank == 0:
        if line is None:  # --only <secondary>: that workload's dict is the line
            name = todo[0]
            line = {"metric": name, "n_gpus": W, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic", **workloads[name]}
            line.setdefault("value", None)
            line["config"] = {"workload": workloads[name].get("config", name)}
        else:
            if workloads:
                line["workloads"] = workloads
            if W == 1 and not args.no_cpu_baseline:
                cpu = cpu_loss_baseline()
                line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
                if "retrieval" in workloads and "error" not in workloads["retrieval"]:
                    workloads["retrieval"]["cpu_baseline"] = cpu_topk_baseline(n_vec=256, n_lit=4)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _quiet_stdout() -> None:
    """Libraries write to fd 1 behind Python's back (NCCL prints "NCCL version ..." there on the
    first collective): keep the original stdout for the JSON line and point fd 1 at stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--only", choices=["loss", *WORKLOADS], default=None,
                    help="run one workload alone (default: the loss line with every other workload inside it)")
    ap.add_argument("--workload", choices=["loss", "retrieval"], default=None, help="alias of --only (round-1 flag)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `workloads` section")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contract-test-n", type=int, default=0,
                    help="tests/test_bench_contract.py only: reference arm at a reduced global batch (the line's "
                         "config.workload says so); never used by a measurement")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="time CUDA-graph replays of the captured step (the default: 1 GPU, or N > 1 with the push "
                         "exchange active; the eager time is reported beside it as eager_ms_per_step)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    args = ap.parse_args()
    if args.workload and not args.only:
        args.only = args.workload
    if args.graph is None:
        args.graph = True   # the step is replayed as a captured graph wherever capture works (eager otherwise)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    bench_main(args)


if __name__ == "__main__":
    main()
