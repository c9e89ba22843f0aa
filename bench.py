#!/usr/bin/env python
"""Benchmark of the Boosted_DETR hot path (BASELINE.json metric: images/sec fwd+bwd+matcher).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            the restated reference CPU path (oracle/) on host cores

A step = one training pass (forward of all boosted blocks, Hungarian matching loss at every block,
backward of everything, gradient all-reduce when N > 1) over one synthetic batch of BASELINE config 2:
6 enc/dec pairs, d_model 256, 8 heads, 100 queries, 20x20 feature map (640x640 images / stride 32),
16 images per GPU, 20 padded targets, C=82, A=3.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = dict(N=6, B=16, rows=20, cols=20, Q=100, T=20, D=256, H=8)
METRIC = "images/sec fwd+bwd+matcher"
# dram__bytes_read.sum + dram__bytes_write.sum of one gemm_umma_kernel<128,...> launch at [6400,256]x[256,256]
# (profiles/r1b_ncu_full_kernels.csv, ncu --set full): 13.39 MB read (A 6.55 MB + cold weights / L2 flush residue), 0 written
# back before the kernel ended; algorithmic bytes are 13.4 MB (read A, write C) -- no wasted re-reads.
NCU_DRAM_BYTES_GEMM = 13.39e6


def synth_batch(rank, B, C, A, cfg=CFG):
    from util import synth_targets
    rng = np.random.default_rng(1234 + rank)
    cat, attr, box, n = synth_targets(rng, B, cfg["T"], C, A)
    x = np.tanh(rng.standard_normal((B, cfg["rows"], cfg["cols"], cfg["D"])))
    x = ((x - x.mean((0, 1, 2))) / x.std((0, 1, 2))).astype(np.float32)      # mimics BN o tanh o BN (backbone.py:90-95)
    return {"features": x, "category": cat, "attribute": attr, "bbox": box, "num_objects": n}


def make_model(cfg=CFG, seed=0):
    from boosted_detr_b200.boosted_model import BoostedDETR
    from boosted_detr_b200.parameters import baseline_params
    p = baseline_params(2)
    p.update(num_decoder_blocks=cfg["N"], num_encoder_blocks=cfg["N"], num_object_preds=cfg["Q"],
             image_size=(cfg["rows"] * 32, cfg["cols"] * 32))
    model = BoostedDETR(**p, attribute_weight=1.0, seed=seed).build()
    # throughput runs: queries N(0, 0.02^2) instead of the zero init
    w = model.get_weights_dict()
    rng = np.random.default_rng(seed)
    w["DecoderPrep/init_decoder_features"] = rng.normal(0, 0.02, w["DecoderPrep/init_decoder_features"].shape).astype(np.float32)
    model.set_weights_dict(w)
    model.dropout_seed = 2024                                               # training-mode dropout (rate .1) is on
    # the reference's optimizer (Boosted_DETR_COCO.ipynb cell 26): the timed step ends with its update
    from boosted_detr_b200.optimizers import SGD, CosineDecayRestarts
    model.compile(optimizer=SGD(learning_rate=CosineDecayRestarts(1e-3, 4000, m_mul=.95, alpha=.1), momentum=.9, nesterov=True, clipnorm=0.1))
    return model


def algorithmic_flops_per_step(cfg, B, C, A, training=True):
    """SURVEY.md §8d (2 flops/MAC, batch-invariant work hoisted)."""
    L, D, Q, N = cfg["rows"] * cfg["cols"], cfg["D"], cfg["Q"], cfg["N"]
    f_enc = 12 * L * D * D + 4 * L * L * D
    f_dec = 4 * L * D * D + 4 * Q * L * D + 8 * Q * D * D
    f_heads = 6 * Q * D * D + 2 * Q * D * (C + A + 4)
    f_self = 8 * Q * D * D + 4 * Q * Q * D
    return (3 if training else 1) * (B * N * (f_enc + f_dec + f_heads) + (N - 1) * f_self)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def time_kernel(fn, reps=50, iters=7, flush=None):
    """Average device time of one launch: `reps` back-to-back launches captured in a CUDA graph (no host launch
    gaps; a single launch bracketed by events is floored at ~10 us by the event/launch overhead), median of
    `iters` replays, CUDA events on the launching stream, L2 flushed before each replay."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return float(np.median(ts)) * 1e-3        # seconds


def roofline_block(cfg, B, flush, peaks):
    """Dominant kernel of the step (the Dense GEMM: ~330 of the ~820 launches and the largest share of device
    time in profiles/), timed alone, live; the attention core and the cost-matrix kernel are reported beside it."""
    import torch
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.device import ptr, stream_ptr
    L, D, H = cfg["rows"] * cfg["cols"], cfg["D"], cfg["H"]
    mode = _lib.load().bdetr_get_mode()
    M = B * L
    x = torch.randn(M, D, device="cuda"); wt = torch.randn(D, D, device="cuda"); bias = torch.randn(D, device="cuda")
    y = torch.empty(M, D, device="cuda")
    t_gemm = time_kernel(lambda: _lib.call("bdetr_gemm", M, D, D, ptr(x), 0, ptr(wt), 0, ptr(bias), 0, 0, ptr(y), stream_ptr()), flush=flush)
    flops_gemm = 2.0 * M * D * D
    q, k, v = (torch.randn(B, L, D, device="cuda") for _ in range(3))
    o, lse = torch.empty(B, H, L, D // H, device="cuda"), torch.empty(B, H, L, device="cuda")
    t_attn = time_kernel(lambda: _lib.call("bdetr_attention_core_fwd", B, H, L, L, D // H, ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), stream_ptr()), flush=flush)
    flops_attn = 4.0 * B * L * L * D
    # cost matrix at BASELINE config 4 (the size its HBM target is stated for)
    from util import synth_preds, synth_targets
    rng = np.random.default_rng(0)
    Bm, T, Q, C, A = 256, 100, 300, 82, 3
    tr = synth_targets(rng, Bm, T, C, A); pr = synth_preds(rng, Bm, Q, C, A)
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (*tr, *pr)]
    cost = torch.empty(Bm, T, Q, device="cuda")
    t_cost = time_kernel(lambda: _lib.call("bdetr_cost_matrix_fwd", Bm, T, Q, C, A, ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[4]), ptr(d[5]), ptr(d[6]),
                                           1000.0, 1.0, 1.0, ptr(cost), stream_ptr()), reps=20, flush=flush)
    bytes_cost = 4.0 * Bm * (Q * C + Q * A + 4 * Q + T * C + T * A + 4 * T + T * Q)
    tpeak = peaks.get("bf16_tflops", 1590.0)
    hpeak = peaks.get("hbm_gbs", 6650.0)
    src = "MEASURED_PEAKS.json (burst figures, kernel timed alone)" if peaks else "fallback 1590 TFLOP/s / 6650 GB/s"
    ach = flops_gemm / t_gemm / 1e12
    return {"bound": "tensor", "kernel": "gemm_umma_kernel (Dense forward [%d,%d]x[%d,%d], tf32 operands)" % (M, D, D, D) if mode else
            "gemm_simt_kernel (Dense forward [%d,%d]x[%d,%d], fp32 FFMA)" % (M, D, D, D),
            "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
            "traffic": NCU_DRAM_BYTES_GEMM if (mode and (M, D) == (6400, 256)) else None, "peak_source": src,
            "us_per_launch": t_gemm * 1e6, "mode": "tf32" if mode else "fp32",
            "note": "peak is the measured bf16 cuBLAS figure; TF32 tensor peak is half of it. K=256 makes this GEMM "
                    "latency / L2-ingest bound (see profiles/README.md)",
            "other_kernels": {
                "attention_core_fwd": {"bound": "tensor", "achieved": flops_attn / t_attn / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                                       "frac": flops_attn / t_attn / 1e12 / tpeak, "us_per_launch": t_attn * 1e6, "shape": [B, H, L, L]},
                "cost_matrix_fwd_config4": {"bound": "hbm", "achieved": bytes_cost / t_cost / 1e9, "peak": hpeak, "unit": "GB/s",
                                            "frac": bytes_cost / t_cost / 1e9 / hpeak, "us_per_launch": t_cost * 1e6,
                                            "algorithmic_bytes": bytes_cost}}}


def cpu_reference_steps(cfg, B, C, A, steps, warmup, threads):
    """The restated reference CPU path (oracle/reference_path.py, torch CPU fp32) on the host cores."""
    import torch
    from oracle import reference_path as R
    torch.set_num_threads(threads)
    from boosted_detr_b200.layers import Layer, glorot_normal, he_normal
    # weights: same shapes/initialisers as the product, built on the host (no GPU needed)
    rng = np.random.default_rng(0)
    L, D, Q, N = cfg["rows"] * cfg["cols"], cfg["D"], cfg["Q"], cfg["N"]
    w = {}
    def attn(pfx):
        for nm in ("QueryProjection", "KeyProjection", "ValueProjection", "OutputProjection"):
            w[f"{pfx}/AttentionLayer/{nm}/kernel"] = glorot_normal(rng, D, D); w[f"{pfx}/AttentionLayer/{nm}/bias"] = np.zeros(D, np.float32)
        w[f"{pfx}/LayerNorm/gamma"] = np.ones(D, np.float32); w[f"{pfx}/LayerNorm/beta"] = np.zeros(D, np.float32)
    def ffn(pfx):
        for nm in ("DenseRelu", "DenseLinear"):
            w[f"{pfx}/{nm}/kernel"] = glorot_normal(rng, D, D); w[f"{pfx}/{nm}/bias"] = np.zeros(D, np.float32)
        w[f"{pfx}/LayerNorm/gamma"] = np.ones(D, np.float32); w[f"{pfx}/LayerNorm/beta"] = np.zeros(D, np.float32)
    def head(pfx, d1, d2, nout):
        w[f"{pfx}/{d1}/kernel"] = he_normal(rng, D, D); w[f"{pfx}/{d1}/bias"] = np.zeros(D, np.float32)
        w[f"{pfx}/BatchNorm/gamma"] = np.ones(D, np.float32); w[f"{pfx}/BatchNorm/beta"] = np.zeros(D, np.float32)
        w[f"{pfx}/BatchNorm/moving_mean"] = np.zeros(D, np.float32); w[f"{pfx}/BatchNorm/moving_variance"] = np.ones(D, np.float32)
        w[f"{pfx}/{d2}/kernel"] = glorot_normal(rng, D, nout); w[f"{pfx}/{d2}/bias"] = np.zeros(nout, np.float32)
    for i in range(N):
        w[f"ImageEncoderAttention_{i}/positional_encoding"] = R.positional_table(cfg["rows"], cfg["cols"], D, np.float32)
        attn(f"ImageEncoderAttention_{i}/EncoderBlock_0/SelfAttentionBlock"); ffn(f"ImageEncoderAttention_{i}/EncoderBlock_0/FeedForwardBlock")
        if i >= 1:
            attn(f"DecoderBlock_{i}/SelfAttentionBlock")
        attn(f"DecoderBlock_{i}/JointAttentionBlock"); ffn(f"DecoderBlock_{i}/FeedForwardBlock")
        head(f"CategoryPredictionHead_{i}", "DenseCateg", "DenseLogits", C)
        head(f"AttributePredictionHead_{i}", "Dense", "DenseLinear", A)
        head(f"BoxPredictionHead_{i}", "Dense", "BoxCoords", 4)
    w["DecoderPrep/init_decoder_features"] = rng.normal(0, 0.02, (Q, D)).astype(np.float32)
    batch = synth_batch(0, B, C, A, cfg)
    tg = (batch["category"], batch["attribute"], batch["bbox"], batch["num_objects"])
    times = []
    acc = {}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, grads, stats = R.train_step_reference(w, batch["features"], tg, N, cfg["H"], torch.float32, dropout_seed=2024,
                                                 weights=R.model_weights(1.0))
        if not acc:
            acc = {k: np.zeros_like(v, dtype=np.float32) for k, v in grads.items()}
        lr = R.cosine_decay_restarts(it, 1e-3, 4000, 2.0, .95, .1)
        new_w, acc = R.sgd_step_reference({k: w[k] for k in grads}, grads, acc, lr, .9, True, .1, dtype=np.float32)
        w.update(new_w)
        w.update({k: np.asarray(v, np.float32) for k, v in stats.items()})
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("BDETR_MODE", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one gradient all-reduce after the backward instead of per-block buckets under it")
    ap.add_argument("--no-conc", action="store_true", help="disable multi-stream concurrency inside the step (A/B timing)")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch (A/B timing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-seconds", type=float, default=420.0, help="hard wall-clock limit: a wedged run exits 3 instead of hanging")
    args = ap.parse_args()

    def _watchdog():
        time.sleep(args.max_seconds)
        sys.stderr.write(f"bench.py: exceeded --max-seconds {args.max_seconds:.0f}, aborting\n")
        sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=_watchdog, daemon=True).start()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    C, A = 82, 3
    cfg, B = CFG, CFG["B"]
    workload = (f"BASELINE config 2: Boosted DETR default, {cfg['N']} enc/dec pairs, d_model {cfg['D']}, {cfg['H']} heads, "
                f"{cfg['Q']} queries, {cfg['rows']}x{cfg['cols']} features (640x640/32), batch {B}/GPU, T={cfg['T']}, C={C}, A={A}, "
                "training step = fwd + Hungarian loss at every block + bwd (+ grad all-reduce when N>1) + SGD-Nesterov update "
                "(per-variable clipnorm .1, CosineDecayRestarts), dropout .1 on")
    config = {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
              "l2": "flushed between timed steps (256 MiB write outside the event brackets); per-step CUDA events"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        sec = cpu_reference_steps(cfg, B, C, A, steps, warm, cores)
        val = B / sec
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
                "data": "synthetic", "config": config, "gpu_launches": 0,
                "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                                 "sample": f"{steps} full training steps of {B} images (oracle/reference_path.py, torch CPU fp32, {cores} threads; TF unavailable)"},
                "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from boosted_detr_b200 import _lib
    from boosted_detr_b200.graph import GraphedTrainStep
    from boosted_detr_b200.parallel import DataParallel, init_from_env
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    lib = _lib.load()
    lib.bdetr_set_mode(_lib.MODE_TF32 if args.mode == "tf32" else _lib.MODE_FP32)
    lib.bdetr_set_pdl(0 if args.no_pdl else 1)
    lib.bdetr_set_concurrency(0 if args.no_conc else 1)
    model = make_model(cfg)
    DataParallel(model, overlap=not args.no_overlap)
    batch = synth_batch(rank, B, C, A, cfg)
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs already in HBM, CUDA-graph replay ---------------------------------------
    if args.no_graph:
        dev_batch = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
        step = lambda: model.train_step(dev_batch, return_host=False)
    else:
        gs = GraphedTrainStep(model, batch)
        step = gs.replay
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:                                                # one nvidia-smi poller per job, not one per GPU
        sampler.start()
    lib.bdetr_reset_launch_count()
    if args.no_graph:
        step(); torch.cuda.synchronize()
        launches_per_step = lib.bdetr_launch_count()
    else:
        launches_per_step = gs.launches_per_step if hasattr(gs, "launches_per_step") else None
    times = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    barrier()
    t_dev = torch.tensor([sum(times) / 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    sec_per_step = float(t_dev.item()) / args.steps

    # ---- end-to-end arm: public API with HOST buffers (pinned H2D in, loss D2H out) every step -----------------
    # the step's inputs wait in pinned host memory (what a prefetching input pipeline hands over); every step copies
    # them host->device and reads the metric means + matcher status back
    host_batch = {k: torch.from_numpy(np.ascontiguousarray(v, np.int32 if k == "num_objects" else np.float32)).pin_memory()
                  for k, v in batch.items()}
    # BDETR_E2E_PREFETCH=1: the next step's H2D copy is queued under the running step (GraphedTrainStep.prefetch);
    # off by default until it has been validated on the GPU
    if args.no_graph:
        e2e_fn = lambda: model.train_step(batch)
    elif os.environ.get("BDETR_E2E_PREFETCH") == "1":
        e2e_fn = lambda: gs(host_batch, prefetch=host_batch)
    else:
        e2e_fn = lambda: gs(host_batch)
    for _ in range(3):
        e2e_fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        logs = e2e_fn()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_sec = float(t_e2e.item()) / args.steps
    h2d = sum(v.nbytes for v in batch.values())
    d2h = 4 * 6 + 4 * cfg["N"] * B                         # six metric means + per-block, per-image matcher status flags
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    if rank != 0:
        return                                                   # (no collective below this line: the other ranks are gone)
    if launches_per_step is None:
        # count by running one eager (non-graph) step
        dev_batch = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
        lib.bdetr_reset_launch_count()
        model.grad_allreduce = model.grad_bucket_hook = None     # rank 0 only from here on: no collectives
        model.train_step(dev_batch, return_host=False)
        torch.cuda.synchronize()
        launches_per_step = lib.bdetr_launch_count()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = roofline_block(cfg, B, flush, peaks)
    flops = algorithmic_flops_per_step(cfg, B, C, A)
    line = {"metric": METRIC, "value": B * world / sec_per_step, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32" if args.mode == "tf32" else "fp32", "data": "synthetic (random-init weights)",
            "config": config, "clocks": sampler.summary(),
            "e2e": {"value": B * world / e2e_sec, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_sec * 1e3, "input_prefetch": os.environ.get("BDETR_E2E_PREFETCH") == "1" and not args.no_graph},
            "gpu_launches": int(launches_per_step) * args.steps, "gpu_launches_per_step": int(launches_per_step),
            "cuda_graph": not args.no_graph, "pdl": not args.no_pdl, "allreduce": ("none" if world == 1 else "one call after backward" if args.no_overlap else "per-block buckets overlapped with backward, inside the CUDA graph"), "concurrent_streams": not args.no_conc, "loss": logs.get("loss") if isinstance(logs, dict) else None,
            "step_algorithmic_tflops": flops / sec_per_step / 1e12, "roofline": roof}
    if not args.no_cpu_baseline and world == 1:
        sec = cpu_reference_steps(cfg, B, C, A, 1, 1, cores)
        line["cpu_baseline"] = {"value": B / sec, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"1 full training step of {B} images after 1 warm-up (oracle/reference_path.py, torch CPU fp32, {cores} threads)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
