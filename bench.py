"""bench.py — the reference's headline metric for the hot path, measured on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload loss|retrieval]
    (N > 1: launched by torch.distributed.run, one rank per GPU over NCCL)

Workload "loss" (default; BASELINE.json configs[1]): ViT-B/16-width contrastive loss forward +
backward, GLOBAL batch 32768 x D=512 sharded over the N ranks (strong scaling: the global batch
is fixed), synthetic unit-norm features as SURVEY.md §8d.  A step is one fwd+bwd of the loss
(`clip_contrastive_loss`: 16-bit cast, feature all-gather, fused forward, scalar/lse exchange,
fused backward).
  value : pairs/s = 32768 / step time, inputs resident in HBM, CUDA events, max over ranks
  e2e   : the same step through the public call with HOST (pinned) fp32 features: H2D of the
          step's features inside the timed region, loss scalar read back
  roofline : the dominant kernel (fused backward), algorithmic flops / its measured launch time,
             against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle (fp32 torch port of the reference path) on this box's host cores

`--impl reference` times the reference's own algorithm (the oracle port: the reference is
PyTorch code, its arithmetic IS these torch ops) on the host cores, rank 0 only.
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

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, world, rank, local
    return None, 1, 0, 0


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


def cpu_loss_baseline(budget_s=20.0):
    """The oracle (port of train.py:87-115 + autograd) on the host cores, bounded sample."""
    from oracle import clip_loss as OL
    n = 8192
    img, txt = synth_features(n, 0, D)
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    OL.global_loss_and_grads(img[:1024], txt[:1024], LOGIT_SCALE)  # warm-up
    reps, spent = 0, 0.0
    while reps < 1 or (spent < budget_s / 2 and reps < 5):
        t1 = time.perf_counter()
        OL.global_loss_and_grads(img, txt, LOGIT_SCALE)
        spent += time.perf_counter() - t1
        reps += 1
    per = spent / reps
    full = per * (N_GLOBAL / n) ** 2  # cost is quadratic in the global batch
    return {"value": N_GLOBAL / full, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"fp32 torch fwd+bwd at N={n}, D={D}: {per:.3f} s/step over {reps} reps; "
                      f"extrapolated x{(N_GLOBAL // n) ** 2} (cost ~ N^2) to N={N_GLOBAL}",
            "measured_pairs_per_s_at_sample": n / per, "wall_s": time.perf_counter() - t0}


def cpu_topk_baseline(budget_s=15.0):
    from oracle import topk as OT
    g = torch.Generator().manual_seed(77)
    gal = torch.nn.functional.normalize(torch.randn(RET_G, D, generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(512, D, generator=g), dim=-1)
    t0 = time.perf_counter()
    OT.topk_vectorised(gal, q[:64], RET_K)
    t1 = time.perf_counter()
    OT.topk_vectorised(gal, q, RET_K, block=128)
    per = (time.perf_counter() - t1) / q.shape[0]
    return {"value": 1.0 / per, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"vectorised fp32 Q@G^T + stable sort, {q.shape[0]} of {RET_Q} queries x G={RET_G}",
            "wall_s": time.perf_counter() - t0}


def run_reference_arm(args):
    """The reference's own CPU path for the same metric/config (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone on the host
    # cores, so it takes all of them (what a plain `python bench.py --impl reference` gets by default)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    if args.workload == "retrieval":
        base = cpu_topk_baseline()
        metric, cfg = "retrieval_queries_per_s", {"workload": f"top-{RET_K} retrieval Q={RET_Q} G={RET_G} D={D}"}
    else:
        base = cpu_loss_baseline()
        metric, cfg = "contrastive_fwd_bwd_pairs_per_s", {"workload": f"contrastive loss fwd+bwd, global batch {N_GLOBAL}, D={D}"}
    line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": base["unit"],
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (N_GLOBAL if args.workload == "loss" else RET_Q) / base["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def bench_loss(args):
    from nans_clip_b200 import kernels as K
    from nans_clip_b200.loss import clip_contrastive_loss

    dist, W, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    assert N_GLOBAL % W == 0
    n_loc = N_GLOBAL // W
    group = dist.group.WORLD if dist is not None else None
    feat_dt = torch.float16
    peaks = measured_peaks()

    img, txt = synth_features(n_loc, rank * n_loc, D)
    img_h, txt_h = img.pin_memory(), txt.pin_memory()
    img_d = img.to(dev).requires_grad_(True)
    txt_d = txt.to(dev).requires_grad_(True)
    s = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        img_d.grad = txt_d.grad = s.grad = None
        loss, _ = clip_contrastive_loss(img_d, txt_d, s, group=group, feat_dtype=feat_dt)
        loss.backward()
        return loss

    def step_e2e():
        i = img_h.to(dev, non_blocking=True).requires_grad_(True)
        t = txt_h.to(dev, non_blocking=True).requires_grad_(True)
        s.grad = None
        loss, _ = clip_contrastive_loss(i, t, s, group=group, feat_dtype=feat_dt)
        loss.backward()
        return float(loss.item())  # D2H read of the step's result

    sampler = ClockSampler(local)
    l0 = K.LAUNCHES
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = (K.LAUNCHES - l0) // max(args.warmup, 1)
    run_step = step
    graphed = False
    if args.graph and dist is None:  # NCCL collectives inside a captured step hung at 2 GPUs: 1 GPU only
        # capture one whole step (casts, gathers, fused forward, exchange, fused backward) in a CUDA
        # graph and replay it: same kernels and collectives, no per-launch host latency
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            with torch.cuda.graph(g):
                step()
            torch.cuda.synchronize()
            run_step = g.replay
            graphed = True
            for _ in range(2):
                run_step()
            torch.cuda.synchronize()
        except Exception as exc:  # fall back to eager launches, say so in the line
            print(f"[bench] CUDA graph capture failed, timing eager launches: {exc!r}", file=sys.stderr)
            run_step = step
    sampler.start()
    total_ms = timed_steps(run_step, args.steps, 0, flush, dist, dev)
    clocks = sampler.stop()
    ms = total_ms / args.steps
    e2e_ms = timed_steps(step_e2e, args.steps, min(args.warmup, 3), flush, dist, dev) / args.steps

    # ---- dominant kernel (fused backward) timed alone, on this rank ----
    with torch.no_grad():
        I16, _, _ = K.l2norm_cast(img_d.detach(), feat_dt, normalize=False)
        T16, _, _ = K.l2norm_cast(txt_d.detach(), feat_dt, normalize=False)
        if W > 1:
            I_all = torch.empty((N_GLOBAL, D), dtype=feat_dt, device=dev)
            T_all = torch.empty((N_GLOBAL, D), dtype=feat_dt, device=dev)
            dist.all_gather_into_tensor(I_all, I16)
            dist.all_gather_into_tensor(T_all, T16)
        else:
            I_all, T_all = I16, T16
        s_dev = s.detach().reshape(1)
        slots = K.fwd_phase_slots(n_loc, N_GLOBAL, D)
        ws = K.fwd_workspace(n_loc, slots, dev)

        def fwd_only():
            K.fwd_phase(I16, T16, T_all, I_all, col_global_begin=0, label_begin=rank * n_loc, s_dev=s_dev,
                        with_acc=False, ws=ws, slot_begin=0)

        fwd_only()
        lse, _sc, _ = K.fwd_finalize(n_loc, slots, rank * n_loc, s_dev, False, ws)
        if W > 1:
            lse_g = torch.empty((W * 2, n_loc), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(lse_g, lse.contiguous())
            lse_all = lse_g.view(W, 2, n_loc).permute(1, 0, 2).reshape(2, N_GLOBAL)
        else:
            lse_all = lse
        pad = (N_GLOBAL + 3) // 4 * 4
        lse_pad = torch.empty((2, pad), dtype=torch.float32, device=dev)[:, :N_GLOBAL]
        lse_pad.copy_(lse_all)
        gout = torch.ones(1, device=dev)

        def bwd_only():
            K.bwd(I16, T16, T_all, I_all, label_begin=rank * n_loc, s_dev=s_dev, lse_all=lse_pad,
                  grad_out=gout, grad_mult=1.0, row_begin=0, row_count=n_loc, out_dtype=torch.float32)

        kb = max(3, min(args.steps, 10))
        bwd_ms = timed_steps(bwd_only, kb, 2, flush, None, dev) / kb
        fwd_ms = timed_steps(fwd_only, kb, 2, flush, None, dev) / kb
    # algorithmic work per rank: backward 4 * n_loc * N * D (dI and dT; the logit
    # recompute is not counted), forward 2 * n_loc * N * D
    bwd_alg = 4.0 * n_loc * N_GLOBAL * D
    fwd_alg = 2.0 * n_loc * N_GLOBAL * D
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("clip_bwd_np_kernel", {}).get(f"W{W}")
        except Exception:
            traffic = None

    line = None
    if rank == 0:
        cpu = cpu_loss_baseline() if (W == 1 and not args.no_cpu_baseline) else None
        step_alg = 6.0 * N_GLOBAL * N_GLOBAL * D
        line = {
            "metric": "contrastive_fwd_bwd_pairs_per_s", "value": N_GLOBAL / (ms / 1e3), "unit": "pairs/s",
            "n_gpus": W, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"ViT-B/16-width contrastive loss fwd+bwd, global batch {N_GLOBAL}, D={D}, "
                                   f"{n_loc} rows/rank (BASELINE.json configs[1])",
                       "global_batch": N_GLOBAL, "D": D, "operand_dtype": "fp16 (fp32 accumulate)",
                       "logit_scale": LOGIT_SCALE, "parallelism": f"dp{W}",
                       "l2": "flushed (256 MB write) before every timed step",
                       "launch": "cuda-graph replay of one captured step" if graphed else "eager",
                       "step_algorithmic_tflop": step_alg / 1e12},
            "algorithmic_tflops": step_alg / (ms / 1e3) / 1e12 ,
            "frac_of_bf16_peak_algorithmic": step_alg / (ms / 1e3) / 1e12 / (W * peaks["bf16_tflops_sustained"]),
            "e2e": {"value": N_GLOBAL / (e2e_ms / 1e3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * n_loc * D * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": {"kernel": "clip_bwd_np_kernel", "bound": "tensor", "achieved": bwd_alg / (bwd_ms / 1e3) / 1e12,
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": bwd_alg / (bwd_ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": traffic,
                         "launch_ms": bwd_ms, "algorithmic_flops_per_launch": bwd_alg,
                         "hardware_tflops": 2 * bwd_alg / (bwd_ms / 1e3) / 1e12,
                         "peak_source": f"{peaks['source']} burst bf16 (kernel timed alone)"},
            "roofline_fwd": {"kernel": "clip_fwd_kernel", "bound": "tensor", "achieved": fwd_alg / (fwd_ms / 1e3) / 1e12,
                             "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": fwd_alg / (fwd_ms / 1e3) / 1e12 / peaks["bf16_tflops"], "launch_ms": fwd_ms,
                             "hardware_tflops": 2 * fwd_alg / (fwd_ms / 1e3) / 1e12},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_retrieval(args):
    from nans_clip_b200 import kernels as K
    from nans_clip_b200.retrieval import GalleryShard

    dist, W, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = measured_peaks()
    lo, hi = RET_G * rank // W, RET_G * (rank + 1) // W
    g = torch.Generator(device=dev).manual_seed(4242 + rank)
    gal = torch.nn.functional.normalize(torch.randn(hi - lo, D, device=dev, generator=g), dim=-1).bfloat16().float()
    gq = torch.Generator(device=dev).manual_seed(99)
    qry = torch.nn.functional.normalize(torch.randn(RET_Q, D, device=dev, generator=gq), dim=-1).bfloat16().float()
    shard = GalleryShard(gal, dev, torch.float16, lo)
    q16, _, _ = K.l2norm_cast(qry, torch.float16, normalize=False)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    qry_h = qry.cpu().pin_memory()

    def step():
        s, i = K.topk_ip(q16, shard.g16, qry, shard.g32, RET_K, 16, lo)
        if W > 1:
            all_s = torch.empty((W * RET_Q, RET_K), dtype=s.dtype, device=dev)
            all_i = torch.empty((W * RET_Q, RET_K), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(all_s, s)
            dist.all_gather_into_tensor(all_i, i)
            s, i = K.topk_merge(all_s.view(W, RET_Q, RET_K), all_i.view(W, RET_Q, RET_K))
        return i

    def step_e2e():
        q = qry_h.to(dev, non_blocking=True)
        if W > 1:
            s, i = shard.search(q, RET_K)
            all_s = torch.empty((W * RET_Q, RET_K), dtype=s.dtype, device=dev)
            all_i = torch.empty((W * RET_Q, RET_K), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(all_s, s)
            dist.all_gather_into_tensor(all_i, i)
            s, i = K.topk_merge(all_s.view(W, RET_Q, RET_K), all_i.view(W, RET_Q, RET_K))
        else:
            s, i = shard.search(q, RET_K)
        return i.cpu()

    sampler = ClockSampler(local)
    l0 = K.LAUNCHES
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    lps = (K.LAUNCHES - l0) // max(args.warmup, 1)
    sampler.start()
    ms = timed_steps(step, args.steps, 0, flush, dist, dev) / args.steps
    clocks = sampler.stop()
    e2e_ms = timed_steps(step_e2e, max(2, args.steps // 4), 1, flush, dist, dev) / max(2, args.steps // 4)
    flops = 2.0 * RET_Q * (hi - lo) * D
    if rank == 0:
        line = {"metric": "retrieval_queries_per_s", "value": RET_Q / (ms / 1e3), "unit": "queries/s", "n_gpus": W,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": f"top-{RET_K} text->image retrieval, Q={RET_Q}, G={RET_G}, D={D}, gallery "
                                       f"sharded x{W} (BASELINE.json configs[4])",
                           "l2": "flushed (256 MB write) before every timed step", "k_cand": 16,
                           "operand_dtype": "fp16 candidate pass + fp32 rescoring"},
                "e2e": {"value": RET_Q / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": RET_Q * D * 4, "d2h_bytes_per_step": RET_Q * RET_K * 8},
                "gpu_launches": lps * args.steps, "gpu_launches_per_step": lps,
                "roofline": {"kernel": "topk_sweep_kernel+finalize", "bound": "tensor",
                             "achieved": flops / (ms / 1e3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": None,
                             "peak_source": f"{peaks['source']} burst bf16"},
                "clocks": clocks}
        if W == 1 and not args.no_cpu_baseline:
            cpu = cpu_topk_baseline()
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--workload", choices=["loss", "retrieval"], default="loss")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="1 GPU only: time a CUDA-graph replay of the step (value only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    (bench_loss if args.workload == "loss" else bench_retrieval)(args)


if __name__ == "__main__":
    main()
