#!/usr/bin/env python
"""bench.py -- images/sec of the MolNexTR hot path (encoder -> greedy decode -> bond head).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of 32 synthetic 384x384 images per GPU
(BASELINE.json configs[1]: "bs=32 synthetic 384x384 greedy decode on 1xB200"), Swin-B encoder,
seeded synthetic checkpoint, <eos> suppressed so every row runs the full 480 decode steps
(fixed work).  Prints ONE JSON line (rank 0).

  value : images/s with the batches already resident in HBM, device-timed over K consecutive batches
          through Engine.predict_pipelined (2-deep: the encoder of batch i+1 overlaps the persistent
          decode kernel of batch i); `latency` holds the batch-by-batch numbers (Engine.predict)
  e2e   : the same from pinned HOST buffers: H2D of every batch's images and D2H of every result inside
          the timed region
  roofline     : the decoder cross-attention kernel (north_star's named HBM target), timed with
                 CUDA events on its launch stream right after the timed region, same shapes
  cpu_baseline : the CPU oracle (a port of the reference's PyTorch path) on the host cores, on a
                 bounded sample, scaled to the full workload
--impl reference times that CPU path as the reference arm (the reference is pure PyTorch; there
is no compiled reference to build, see oracle/README.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, H, W, T_MAX, S_MEM = 32, 384, 384, 480, 144
METRIC = "images/sec end-to-end decode @384x384 bs=32; encoder HBM GB/s vs roofline"
WORKLOAD = "bs=32 synthetic 384x384 greedy decode on 1xB200 (Swin-B encoder, T=480 forced, seed-0 synthetic checkpoint)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


_CPU_STATE = {}


def _cpu_setup():
    """Checkpoint, inputs and the fastest torch thread counts for the two CPU phases (many-core hosts
    run the decoder's tiny ops far slower with all threads than with a few: give the baseline its
    best setting rather than a strawman)."""
    if _CPU_STATE:
        return _CPU_STATE
    from molnextr_b200 import synth
    from oracle import restate
    ck = synth.synthetic_checkpoint(0, "fixed480")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn((BATCH, 3, H, W), generator=g)
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best_enc, best_dec = (1e30, ncpu), (1e30, ncpu)
    with torch.no_grad():
        f1 = None
        for c in cands:
            torch.set_num_threads(c)
            t0 = time.perf_counter()
            f1 = restate.swin_b_features(ck["encoder"], x[:1])
            best_enc = min(best_enc, (time.perf_counter() - t0, c))
        feats = f1.repeat(BATCH, 1, 1).contiguous()
        for c in cands:
            torch.set_num_threads(c)
            t0 = time.perf_counter()
            restate.greedy_decode(ck["decoder"], feats, max_len=6)
            best_dec = min(best_dec, (time.perf_counter() - t0, c))
    _CPU_STATE.update(ck=ck, x=x, enc_threads=best_enc[1], dec_threads=best_dec[1], ncpu=ncpu)
    return _CPU_STATE


def cpu_reference_sample(n_enc_images: int, n_dec_steps: int):
    """Time the CPU oracle on a bounded sample of the workload; returns (images/s scaled to the
    full batch, description, seconds).  Encoder cost is linear in images, decode cost in steps
    (an under-estimate for the reference: its per-step cost grows with the KV length)."""
    from oracle import restate
    st = _cpu_setup()
    ck, x = st["ck"], st["x"]
    with torch.no_grad():
        torch.set_num_threads(st["enc_threads"])
        t0 = time.perf_counter()
        f_part = restate.swin_b_features(ck["encoder"], x[:n_enc_images])
        t_enc = time.perf_counter() - t0
        feats = f_part.repeat((BATCH + n_enc_images - 1) // n_enc_images, 1, 1)[:BATCH].contiguous()
        torch.set_num_threads(st["dec_threads"])
        t0 = time.perf_counter()
        restate.greedy_decode(ck["decoder"], feats, max_len=n_dec_steps)
        t_dec = time.perf_counter() - t0
    full = t_enc * (BATCH / n_enc_images) + t_dec * (T_MAX / n_dec_steps)
    desc = (f"Swin-B encoder on {n_enc_images} of {BATCH} images ({t_enc:.2f} s, {st['enc_threads']} threads) + greedy decode "
            f"of all {BATCH} rows for {n_dec_steps} of {T_MAX} steps ({t_dec:.2f} s, {st['dec_threads']} threads), fp32 torch, "
            f"thread counts picked as the fastest of a sweep up to {st['ncpu']} cores; scaled linearly to the full batch "
            f"({full:.1f} s); bond head and tokenizer excluded")
    return BATCH / full, desc, t_enc + t_dec


def eager_port_sample(dev, n_enc_images: int = 8, n_dec_steps: int = 24):
    """The same port of the reference's PyTorch path, run EAGERLY on `dev` (BASELINE.md's second bar: "stock
    PyTorch-eager of the reference modules on the same B200"): ~1650 ATen dispatches and three host syncs per
    decode step.  Bounded sample scaled like cpu_reference_sample.  Returns (images/s, description)."""
    from molnextr_b200 import synth
    from oracle import restate
    ck = synth.synthetic_checkpoint(0, "fixed480")
    enc = {k: v.to(dev) for k, v in ck["encoder"].items()}
    dec = {k: v.to(dev) for k, v in ck["decoder"].items()}
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn((BATCH, 3, H, W), generator=g).to(dev)
    sync = torch.cuda.synchronize if torch.device(dev).type == "cuda" else (lambda: None)
    with torch.no_grad():
        f1 = restate.swin_b_features(enc, x[:1])                           # warm-up (cuBLAS / cuDNN handles)
        restate.greedy_decode(dec, f1.repeat(BATCH, 1, 1).contiguous(), max_len=2)
        sync()
        t0 = time.perf_counter()
        f_part = restate.swin_b_features(enc, x[:n_enc_images])
        sync()
        t_enc = time.perf_counter() - t0
        feats = f_part.repeat((BATCH + n_enc_images - 1) // n_enc_images, 1, 1)[:BATCH].contiguous()
        t0 = time.perf_counter()
        restate.greedy_decode(dec, feats, max_len=n_dec_steps)
        sync()
        t_dec = time.perf_counter() - t0
    full = t_enc * (BATCH / n_enc_images) + t_dec * (T_MAX / n_dec_steps)
    desc = (f"the same fp32 torch port run eagerly on {dev}: Swin-B on {n_enc_images} of {BATCH} images ({t_enc * 1e3:.0f} ms) + "
            f"greedy decode of all {BATCH} rows for {n_dec_steps} of {T_MAX} steps ({t_dec * 1e3:.0f} ms), scaled linearly "
            f"({full:.2f} s per batch)")
    return BATCH / full, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, desc = [], ""
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, desc, _ = cpu_reference_sample(4, 24)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference arm = the CPU oracle port of the reference's PyTorch path "
                   "(no compiled reference exists); each step is a bounded sample scaled to the full batch"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": max(_CPU_STATE.get("enc_threads", 0), _CPU_STATE.get("dec_threads", 0)),
                         "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    emit(line)


def roofline_other(extra, peaks, clocks, xattn_bytes, xattn_s):
    """Secondary kernels named by north_star: decoder cross-attention (graph path) and ConvNeXt dwconv7x7+LN."""
    out = [{"kernel": "attn_kernel<false> (multi-kernel path: cross-attention + per-head final_linear)", "bound": "hbm",
            "achieved": xattn_bytes / xattn_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": xattn_bytes / xattn_s / 1e9 / peaks["hbm_gbs"], "traffic": 9752320,
            "algorithmic_bytes_per_launch": xattn_bytes}]
    if "convnext_error" in extra:
        out.append({"kernel": "dwconv_ln_kernel", "error": extra["convnext_error"]})
    if "dwconv_us" in extra:
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * 2 * mhz * 1e6          # fp32 FLOP/s on the CUDA cores at the sampled clock
        depth, tot_t, tot_roof, per_stage = (3, 3, 27, 3), 0.0, 0.0, []
        for st_, us in enumerate(extra["dwconv_us"]):
            elems = BATCH * (H >> (2 + st_)) * (W >> (2 + st_)) * (128 << st_)
            t_hbm = elems * 6 / (peaks["hbm_gbs"] * 1e9)      # fp32 read + bf16 write
            t_fma = elems * 98 / fma_peak
            roof = max(t_hbm, t_fma)
            per_stage.append({"stage": st_, "us": us, "roof_us": roof * 1e6, "frac": roof / (us * 1e-6)})
            tot_t += depth[st_] * us * 1e-6
            tot_roof += depth[st_] * roof
        out.append({"kernel": "dwconv_ln_kernel (ConvNeXt-B 7x7 depthwise conv + channel LayerNorm, 36 calls)",
                    "bound": "fp32-FMA / hbm (max of the two, SURVEY.md 8d)", "achieved": tot_roof / tot_t, "peak": 1.0,
                    "unit": "fraction of max(bytes/HBM, flops/FMA peak)", "frac": tot_roof / tot_t, "traffic": None,
                    "per_stage": per_stage, "total_ms": tot_t * 1e3, "convnext_encoder_ms": extra.get("convnext_encoder_ms")})
    return out


def run_ours(args):
    import torch.distributed as dist
    from molnextr_b200 import synth
    from molnextr_b200.engine import Engine, MAX_ATOMS, MAX_LEN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    ck = synth.synthetic_checkpoint(0, "fixed480")
    eng = Engine(ck, device=local, max_batch=BATCH, max_height=H, max_width=W)
    g = torch.Generator(device="cpu").manual_seed(rank)
    x_host = torch.randn((BATCH, 3, H, W), generator=g).pin_memory()
    x_dev = x_host.to(dev)

    def gather(out):
        """the reference's one collective: every rank's predictions to all ranks (main.py:295)."""
        if world == 1:
            return
        for k in ("ids", "lens", "n_atoms", "edges"):
            t = out[k].to(dev) if not out[k].is_cuda else out[k]
            buf = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(buf, t.contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            t0 = time.perf_counter()
            out = fn()
            gather(out)
            if os.environ.get("MNX_BENCH_DEBUG"):
                print(f"[debug] step wall {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def timed_pipeline(batches, host):
        """K consecutive batches through Engine.predict_pipelined: every step's work (encoder, decode, bond head,
        and in host mode its H2D / D2H) is inside the timed region; steps overlap 2-deep."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for out in eng.predict_pipelined(batches, host=host):
            gather(out)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        gather(eng.predict(x_dev))
    # untimed pass of the same shape as the timed one: device allocator pools (and, in host mode below, the
    # engine's pinned result buffers, one set per batch of a call) exist before the timed region starts
    pipeline_note = None
    try:
        eng.predict_pipelined([x_dev] * max(args.steps, args.warmup))
    except Exception as ex:      # keep a bench line even if the two-stream path is unusable on this box
        pipeline_note = f"predict_pipelined failed ({ex}); value / e2e are the batch-by-batch numbers"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_latency = timed(lambda: eng.predict(x_dev), args.steps)          # batch by batch (single-batch latency)
    l0 = eng.launch_count()
    ms_dev = timed_pipeline([x_dev] * args.steps, host=False) if pipeline_note is None else timed(lambda: eng.predict(x_dev), args.steps)
    launches = eng.launch_count() - l0
    steps_run = eng.last_decode_steps()
    # ---- host-buffer arm (H2D + D2H inside) ----
    eng.predict_host(x_host)
    if pipeline_note is None:
        eng.predict_pipelined([x_host] * max(args.steps, args.warmup), host=True)
        ms_e2e = timed_pipeline([x_host] * args.steps, host=True)
    else:
        ms_e2e = timed(lambda: eng.predict_host(x_host), args.steps)
    ms_e2e_latency = timed(lambda: eng.predict_host(x_host), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- phase breakdown + isolated kernel timings (rank 0, after the timed region) ----
    extra = {}
    if rank == 0:
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        feats = eng.encode(x_dev)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            feats = eng.encode(x_dev)
        e1.record()
        eng.decode_greedy(feats)
        e2.record()
        torch.cuda.synchronize()
        extra["encoder_ms"] = e0.elapsed_time(e1) / 5
        extra["decode_ms"] = e1.elapsed_time(e2)
        extra["decode_us_per_step"] = 1000.0 * extra["decode_ms"] / max(1, eng.last_decode_steps())
        names = {1: "cross_attn", 2: "self_attn_t240", 3: "ln1_qkv", 4: "sum_ln_w1_gelu", 5: "w2_partials", 6: "pick"}
        extra["kernel_us"] = {n: 1000.0 * eng.time_kernel(k, 100) for k, n in names.items()}
        extra["mega_ms"] = eng.time_kernel(7, 2)     # the persistent cluster decode kernel alone (CUDA events)
        # 3 = 16-CTA clusters, 3 x 3-warp groups (mega16.cu); 5 = 16-CTA clusters, 2 x 4-warp groups (mega16s.cu); 2 = 8-CTA clusters
        extra["decode_path"] = int(eng.time_kernel(1003, 1))
        # ---- ConvNeXt-B encoder (north_star's named dwconv target), same batch, separate engine ----
        try:
            eng.close()
            ckc = synth.synthetic_checkpoint(0, "fixed480", encoder="convnext_base")
            engc = Engine(ckc, device=local, max_batch=BATCH, max_height=H, max_width=W)
            for _ in range(3):
                engc.encode(x_dev)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            c0.record()
            for _ in range(5):
                engc.encode(x_dev)
            c1.record()
            torch.cuda.synchronize()
            extra["convnext_encoder_ms"] = c0.elapsed_time(c1) / 5
            extra["dwconv_us"] = [1000.0 * engc.time_kernel(101 + st_, 50) for st_ in range(4)]
            engc.close()
        except Exception as ex:
            extra["convnext_error"] = str(ex)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        total_imgs = BATCH * world * args.steps
        value = total_imgs / (ms_dev / 1000.0)
        e2e_value = total_imgs / (ms_e2e / 1000.0)
        # dominant kernel (83 % of the step in profiles/r1c_summary.md): the persistent cluster decode kernel, the whole
        # greedy decode in one launch (decode_mega16_kernel at bs = 32: seven 16-CTA clusters of <= 5 rows).
        # Algorithmic bytes per launch: per step the 22.1 MB of fp32 decoder weights once, the
        # memory-bank K/V of every row (1 769 472 B) and the self-attention cache read so far (2*6*1024 B per position).
        w_bytes = 4 * (6 * (4 * 65536 + 2 * 65536 + 2 * 262144) + 256 * 229)
        mega_bytes = steps_run * w_bytes + BATCH * steps_run * 1769472 + BATCH * 12 * 1024 * (steps_run * (steps_run + 1) // 2)
        mega_s = extra["mega_ms"] * 1e-3
        achieved = mega_bytes / mega_s / 1e9
        xattn_bytes = BATCH * 8 * S_MEM * 32 * 4 * 2
        xattn_s = extra["kernel_us"]["cross_attn"] * 1e-6
        swin_flops = 94.16e9 * BATCH      # 47.08 GMAC / image
        enc_tflops = swin_flops / (extra["encoder_ms"] * 1e-3) / 1e12
        d2h = BATCH * (MAX_LEN * 4 + 4 + MAX_LEN * 4 + MAX_ATOMS * 4 + 4 + MAX_ATOMS * MAX_ATOMS)
        try:
            cpu_val, cpu_desc, _ = cpu_reference_sample(8, 48)
            cpu = {"value": cpu_val, "unit": "images/s", "cores": max(_CPU_STATE["enc_threads"], _CPU_STATE["dec_threads"]),
                   "kind": "port", "sample": cpu_desc}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        try:
            ev, edesc = eager_port_sample(dev)
            cpu["same_port_eager_on_gpu"] = {"value": ev, "unit": "images/s", "sample": edesc}
        except Exception as ex:
            cpu["same_port_eager_on_gpu"] = {"value": None, "sample": f"failed: {ex}"}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 encoder GEMMs (fp32 accumulate), f32 decoder", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world, "decode_steps": steps_run,
                       "encoder": "swin_base", "parallelism": f"dp{world}",
                       "pipeline": "the K timed steps run through Engine.predict_pipelined: the encoder of step i+1 (second "
                                   "stream, GEMM grids capped at 32 CTAs) overlaps the persistent decode kernel of step i, "
                                   "which occupies 112 of the 148 SMs; batch-by-batch numbers are under `latency`"
                                   if pipeline_note is None else pipeline_note,
                       "l2": "no explicit flush: one step streams 0.19 GB of bf16 encoder weights, >1 GB of "
                             "activations and a 246 MB KV cache, far above the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": BATCH * 3 * H * W * 4, "d2h_bytes_per_step": d2h},
            "latency": {"ms_per_batch": ms_latency / args.steps, "images_per_s": total_imgs / (ms_latency / 1000.0),
                        "ms_per_batch_host_buffers": ms_e2e_latency / args.steps,
                        "note": "Engine.predict / predict_host batch by batch, no overlap between steps"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": {3: "decode_mega16_kernel", 5: "decode_mega16s_kernel"}.get(extra.get("decode_path"), "decode_mega_kernel") +
                                   " (persistent cluster decode: 480 steps x 6 layers, one launch)",
                         "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": 52.46e9 if extra.get("decode_path") in (3, 5) else 79.04e9, "peak_source": peak_src + " (sustained copy)",
                         "algorithmic_bytes_per_launch": mega_bytes, "launch_ms": extra["mega_ms"],
                         "timing": "CUDA events around 2 launches of the kernel alone on its launch stream, right after "
                                   "the timed region, same K/V buffers; traffic = dram__bytes_read+write of the ncu "
                                   "--set full capture in profiles/ (r1c_mega16 / r1_mega, same command, bs=32); the kernel is "
                                   "latency-bound (serial chain of ~50 cluster exchanges per step), see DESIGN.md 4.3"},
            "roofline_other": roofline_other(extra, peaks, clocks, xattn_bytes, xattn_s),
            "encoder": {"ms": extra["encoder_ms"], "tflops": enc_tflops, "peak_tflops": peaks["bf16_tflops_sustained"],
                        "frac": enc_tflops / peaks["bf16_tflops_sustained"], "flops_per_image": 94.16e9},
            "decode": {"ms": extra["decode_ms"], "us_per_step": extra["decode_us_per_step"], "kernel_us_graph_path": extra["kernel_us"]},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


_JSON_OUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries write to file descriptor 1 behind Python's
    back (NCCL prints "NCCL version ..." there when NCCL_DEBUG is VERSION / WARN, whatever NCCL_DEBUG_FILE
    says).  Keep a private duplicate of the real stdout for the JSON line and point fd 1 at stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
