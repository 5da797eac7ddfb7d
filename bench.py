#!/usr/bin/env python
"""bench.py — frames/sec encoded+scored (EVA-CLIP-g/14, 224 px) on N B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic input per rank:
  encode_image on 1024 frames (32 videos x 32 frames, BASELINE.json configs[1]) -> per-video mean-pool + L2 norm ->
  [N > 1: one NCCL all-gather of the normalised video embeddings] -> cosine scores against 512 text queries
  (4 of them re-encoded by the text tower each step: BASELINE configs[2]'s 512 queries : 131072 frames ratio).

`value`  : whole-job frames/s, inputs already resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the public Python API with HOST (pinned) buffers — H2D of the frames and D2H of the
           score matrix inside the timed region.
`roofline`: aggregate of this library's tcgen05 GEMM launches (97 % of the FLOPs), per-launch CUDA events.
`cpu_baseline`: the CPU oracle port (oracle/eva_oracle.py, torch fp32) on the box's host cores, bounded sample.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec encoded+scored (EVA-CLIP-g/14 224px)"


def gemm_traffic():
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (mean over the four GEMMs of one ViT layer at
    1024 frames) READ from the committed summary of the `ncu --set full` capture (profiles/gemm_traffic.json, written by
    tools/profile_layer.sh -> tools/ncu_summary.py --json); null if the file is missing — never a literal in this script."""
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["gemm_dram_bytes_per_launch_mean"]), f"profiles/gemm_traffic.json ({d.get('source', 'ncu')})"
    except Exception:
        return None, "profiles/gemm_traffic.json missing"


UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per rank per step")
    ap.add_argument("--frames-per-video", type=int, default=32)
    ap.add_argument("--queries", type=int, default=512)
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="debug: tiny config instead of EVA-CLIP-g/14")
    ap.add_argument("--config", default="step", choices=["step", "retrieval", "moment", "e2e"],
                    help="step (default, the driver's line): BASELINE configs[1]-shaped step per GPU; retrieval / moment / e2e: "
                         "BASELINE configs[2] / [3] / [4] as whole jobs (tools/bench_extra.py)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the sharded-vs-single-GPU equality check after the timed region")
    ap.add_argument("--videos-per-gpu", type=int, default=512, help="--config retrieval: videos per rank (x 32 frames)")
    ap.add_argument("--clips-per-gpu", type=int, default=64, help="--config moment: clips per rank per step")
    ap.add_argument("--clip-frames", type=int, default=300, help="--config moment: frames per clip")
    ap.add_argument("--videos", type=int, default=256, help="--config e2e: videos in the whole job")
    ap.add_argument("--beam", type=int, default=3)
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0)), "hbm_gbs": p.get("hbm_gbs", 6650.0),
                "source": "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md, sustained)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self) -> int:
        """Index of the next sample: brackets a timed region without restarting nvidia-smi (its start-up takes NVML / driver
        locks for up to a second, which stalls a host-synchronised loop that happens to run at the same time)."""
        return len(self.lines)

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.proc = None
        self.stopped = True

    def summary(self, i0: int = 0, i1: int = None):
        if self.proc is None and not getattr(self, "stopped", False):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines[i0:i1]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_mhz_min": sm_sorted[0], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(cfg, n_frames, steps, warmup, frames_per_video, n_queries):
    """Times the CPU oracle port of the path (torch fp32, all host threads) on `n_frames` frames per step."""
    import torch
    from hirest_b200 import synthetic
    from oracle import eva_oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synthetic.make_eva_state_dict(cfg, seed=0)
    frames = synthetic.make_frames(n_frames, cfg["vision_cfg"]["image_size"], seed=1)
    fpv = min(frames_per_video, n_frames)
    text_hat = torch.nn.functional.normalize(torch.randn(n_queries, cfg["embed_dim"]), dim=-1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            emb = eva_oracle.encode_image(sd, frames, cfg)
            v_hat = eva_oracle.pool_normalize_video(emb[: (n_frames // fpv) * fpv], fpv)
            _ = eva_oracle.similarity(text_hat, v_hat)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return {"frames_per_s": n_frames * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": threads,
            "sample": f"{n_frames} frames/step x {len(times)} steps (+{warmup} warm-up), encode_image + pool + scores, torch fp32 CPU"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from hirest_b200 import synthetic

    cfg = synthetic.EVA_TINY if args.tiny else synthetic.EVA_G14
    model_name = "EVA-TINY(debug)" if args.tiny else "EVA-CLIP-g/14"
    if args.config != "step":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_extra

        return bench_extra.main(args, cfg, model_name, ClockSampler, load_peaks)

    # ------------------------------------------------------------------ reference arm: CPU oracle port
    if args.impl == "reference":
        if rank != 0:
            return 0
        warm = args.warmup
        r = cpu_reference_run(cfg, args.cpu_frames, args.steps, warm, args.frames_per_video, args.queries)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["frames_per_s"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{model_name} frame encoder + retrieval scoring; CPU sample of {args.cpu_frames} frames/step",
                       "frames_per_step": args.cpu_frames, "frames_per_video": args.frames_per_video, "queries": args.queries},
            "cpu_baseline": {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["frames_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    from hirest_b200 import _lib, eva_clip, retrieval

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU oracle timing")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created; stdout carries exactly one JSON
        # line (bench contract), so the banner is sent to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.init(local_rank)
    # (kernel-variant A/B switches: HB_DEBUG_* environment variables, include/hirest_b200_debug.h)

    sd = synthetic.make_eva_state_dict(cfg, seed=0, device=dev)
    model = eva_clip.EVA_CLIP(**cfg, max_image_batch=args.frames, max_text_batch=max(args.queries, 8))
    model.load_state_dict(sd, strict=True)
    del sd
    model = model.to(dev).eval()

    S = cfg["vision_cfg"]["image_size"]
    B, Fv, Q, E = args.frames, args.frames_per_video, args.queries, cfg["embed_dim"]
    assert B % Fv == 0
    new_q = max(1, round(Q * B / (4096 * 32)))  # queries re-encoded per step, keeping cfg3's 512 : 131072 ratio
    frames_dev = synthetic.make_frames(B, S, seed=100 + rank, device=dev)
    tokens_all = synthetic.make_tokens(Q, cfg, seed=2).to(dev)
    with torch.no_grad():
        text_hat = retrieval.normalize(model.encode_text(tokens_all))
    tokens_step = tokens_all[:new_q].contiguous()

    V_total = world * (B // Fv)
    pending = [None]   # all-gather of the previous step's video embeddings, still in flight

    def score_pending():
        """Scores of the step whose all-gather is in flight (None if there is none)."""
        if pending[0] is None:
            return None
        v_all = pending[0].wait()
        pending[0] = None
        return retrieval.similarity(text_hat, v_all, exact=True)

    def step_device(frames=None):
        """One step: re-encode `new_q` queries, encode + pool this rank's frames, START the all-gather of the normalised video
        embeddings, and score the PREVIOUS step's gathered embeddings.  Deferring the scoring by one step takes the collective
        off the critical path: with an in-line gather every step ended in a rank barrier, so the step time was the max over
        (power-capped, jittering) ranks (r01: 0.983 efficiency at 8 GPUs); now a rank only ever waits for a gather its peers
        started a whole step earlier.  `drain()` scores the last step; both are inside every timed region."""
        t_new = retrieval.normalize(model.encode_text(tokens_step))
        text_hat[:new_q].copy_(t_new)
        v_hat = retrieval.encode_videos(model, frames_dev if frames is None else frames, Fv)
        nxt = retrieval.all_gather_embeddings(v_hat, n_total=V_total, async_op=True)
        scores = score_pending()
        pending[0] = nxt
        return scores

    def drain():
        return score_pending()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ONE nvidia-smi sampler for the whole run, started before the warm-up: its start-up (NVML / driver locks, up to ~1 s)
        # must not coincide with a timed region — it stalled the first step of the host-synchronised e2e loop by ~0.5 s.
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step_device()
        drain()
        barrier()
        mark0 = sampler.mark()
        launches0 = lib.hb_launch_count()
        lib.hb_profile_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        marks = []
        for _ in range(args.steps):
            step_device()
            mk = torch.cuda.Event(enable_timing=True)
            mk.record()
            marks.append(mk)
        scores = drain()
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        ms_steps = [round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 1) for i in range(len(marks))]
        prof = _lib.HbProfileSummary()
        _lib.check(lib.hb_profile_stop(prof), "hb_profile_stop")
        launches = lib.hb_launch_count() - launches0
        clocks = sampler.summary(mark0, sampler.mark()) if rank == 0 else None
        value = world * B * args.steps / (ms_total * 1e-3)

        # ---------------- e2e through the public API with host buffers
        e2e = None
        if args.no_e2e:
            sampler.stop()
        if not args.no_e2e:
            # raw uint8 frames in pinned host memory: ToTensor + Normalize run on the GPU inside the patch gather
            # (encode_image accepts uint8), so a step moves 154 MB over PCIe instead of 617 MB of fp32 frames
            gen = torch.Generator().manual_seed(200 + rank)
            frames_host = torch.randint(0, 256, (B, 3, S, S), generator=gen, dtype=torch.uint8).pin_memory()
            tokens_host = tokens_step.cpu().pin_memory()
            scores_host = [torch.empty((Q, world * (B // Fv)), dtype=torch.float32).pin_memory() for _ in range(2)]

            # Streaming pipeline a user would write: the H2D copy of step i+1 (pinned host -> device, side stream) overlaps
            # the compute of step i, and the host picks up the scores of step i-1 (pinned, double-buffered) after it has
            # enqueued step i, so the GPU always has the next step queued.  Every step's H2D copy and D2H result read are
            # inside the timed region.  (Synchronising the stream after EVERY step leaves the GPU idle for the few ms the host
            # needs to enqueue the next one; on some boxes it then takes ~0.4 s to get back to speed — isolated 900-1100 ms
            # steps at < 800 W with unchanged SM clocks, which is what `e2e.ms_steps` is recorded for.)
            copy_stream = torch.cuda.Stream(device=dev)
            dev_bufs = [torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev) for _ in range(2)]
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [None, None]

            copy_marks = []   # (start, end) events of every H2D copy, host enqueue / sync times: diagnostics only

            def issue_copy(i):
                with torch.cuda.stream(copy_stream):
                    if consumed[i % 2] is not None:
                        copy_stream.wait_event(consumed[i % 2])
                    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    c0.record(copy_stream)
                    dev_bufs[i % 2].copy_(frames_host, non_blocking=True)
                    c1.record(copy_stream)
                    copy_marks.append((c0, c1))
                    copied[i % 2].record(copy_stream)

            step_events = []
            host_ms = []
            done = [None, None]

            def run_e2e(n):
                for k in range(2):
                    consumed[k] = None
                step_events.clear()
                copy_marks.clear()
                host_ms.clear()
                issue_copy(0)
                for i in range(n):
                    ev0 = torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    step_events.append(ev0)
                    th0 = time.perf_counter()
                    if i + 1 < n:
                        issue_copy(i + 1)
                    cur = torch.cuda.current_stream()
                    cur.wait_event(copied[i % 2])
                    tk = tokens_host.to(dev, non_blocking=True)
                    t_new = retrieval.normalize(model.encode_text(tk))
                    text_hat[:new_q].copy_(t_new)
                    v_hat = retrieval.encode_videos(model, dev_bufs[i % 2], Fv)
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    consumed[i % 2] = ev
                    nxt = retrieval.all_gather_embeddings(v_hat, n_total=V_total, async_op=True)
                    sc = score_pending()          # scores of step i-1 (its gather has had a whole step to complete)
                    pending[0] = nxt
                    if sc is not None:
                        scores_host[(i - 1) % 2].copy_(sc, non_blocking=True)
                        done[(i - 1) % 2] = torch.cuda.Event()
                        done[(i - 1) % 2].record(cur)
                    th1 = time.perf_counter()
                    if i > 1:
                        done[(i - 2) % 2].synchronize()   # the caller now holds the scores of step i-2 on the host
                    host_ms.append((round((th1 - th0) * 1e3, 1), round((time.perf_counter() - th0) * 1e3, 1)))
                sc = drain()
                scores_host[(n - 1) % 2].copy_(sc, non_blocking=True)
                done[(n - 1) % 2] = torch.cuda.Event()
                done[(n - 1) % 2].record(torch.cuda.current_stream())
                evl = torch.cuda.Event(enable_timing=True)
                evl.record()
                step_events.append(evl)
                if n > 1:
                    done[(n - 2) % 2].synchronize()
                done[(n - 1) % 2].synchronize()

            run_e2e(max(2, min(args.warmup, 3)))
            barrier()
            mark2 = sampler.mark()
            e0.record()
            run_e2e(args.steps)
            barrier()
            clocks_e2e = sampler.summary(mark2, sampler.mark()) if rank == 0 else None
            sampler.stop()
            # e0 -> the event recorded on the stream right after the last step's D2H copy (stream order: the last scores are
            # on the host when it fires).  Taking the end stamp on the device keeps a descheduled host thread (observed: 0.3-1 s
            # inside cudaEventSynchronize on busy boxes, GPU already idle) out of a device-side throughput number.
            ms_e2e = max_over_ranks(e0.elapsed_time(step_events[-1]))
            ms_steps_e2e = [round(step_events[i].elapsed_time(step_events[i + 1]), 1) for i in range(len(step_events) - 1)]
            # diagnostic: the bare pinned-host -> device copy of one step's frames, nothing else running
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            dev_bufs[0].copy_(frames_host, non_blocking=True)
            h1.record()
            torch.cuda.synchronize()
            h2d_ms = h0.elapsed_time(h1)
            e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": frames_host.numel() * frames_host.element_size() + tokens_host.numel() * 8,
                   "h2d_ms_alone": h2d_ms, "clocks": clocks_e2e,
                   "ms_steps": ms_steps_e2e, "h2d_ms_steps": [round(a.elapsed_time(b), 1) for a, b in copy_marks],
                   "host_enqueue_and_total_ms": list(host_ms),
                   "input": "uint8 frames [B,3,224,224] in pinned host memory, normalised on the GPU; H2D double-buffered on a side stream",
                   "d2h_bytes_per_step": scores_host[0].numel() * 4, "ms_per_step": ms_e2e / args.steps}

    # ------------------------------------------------------------------ multi-GPU correctness on the hardware (not timed)
    check = None
    if world > 1 and not args.no_check:
        with torch.no_grad():
            v_all = retrieval.all_gather_embeddings(retrieval.encode_videos(model, frames_dev, Fv), n_total=V_total)
            sc_sharded = retrieval.similarity(text_hat, v_all, exact=True)
            parts = []
            for r in range(world):   # this rank alone over EVERY rank's frames (regenerated from their seeds)
                fr = frames_dev if r == rank else synthetic.make_frames(B, S, seed=100 + r, device=dev)
                parts.append(retrieval.encode_videos(model, fr, Fv))
                del fr
            sc_single = retrieval.similarity(text_hat, torch.cat(parts), exact=True)
            ks = [k for k in (1, 5, 10, 50) if k <= V_total]
            top_equal = all(torch.equal(sc_sharded.topk(k, dim=1).indices, sc_single.topk(k, dim=1).indices) for k in ks)
            flags = torch.tensor([int(torch.equal(sc_sharded, sc_single)), int(torch.equal(v_all, torch.cat(parts))), int(top_equal)],
                                 dtype=torch.int64, device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            check = {"what": f"scores [Q={Q}, V={V_total}] from {world} frame shards + NCCL all-gather vs the same videos encoded by one GPU "
                             "alone (every rank checks; min over ranks)",
                     "scores_bit_equal": bool(flags[0]), "gathered_embeddings_bit_equal": bool(flags[1]),
                     "topk_equal": {"ks": ks, "equal": bool(flags[2])},
                     "max_abs_diff": float((sc_sharded - sc_single).abs().max())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    cats = _lib.PROF_CATEGORIES
    per = {c: {"ms_per_step": prof.ms[i] / args.steps, "launches_per_step": prof.launches[i] / args.steps,
               "tflops": (prof.flops[i] / (prof.ms[i] * 1e-3) / 1e12) if prof.ms[i] > 0 and prof.flops[i] > 0 else None}
           for i, c in enumerate(cats)}
    gemm_ms = sum(prof.ms[i] for i in range(3))
    gemm_flops = sum(prof.flops[i] for i in range(3))
    gemm_launches = sum(prof.launches[i] for i in range(3))
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    kernel_ms_total = sum(prof.ms[i] for i in range(len(cats)))
    flops_frame = synthetic.encode_image_flops(cfg)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "ms_steps": ms_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{model_name} frame encoder, {B}-frame batch per GPU ({B // Fv} videos x {Fv} frames) + mean-pool/L2-norm"
                               f" + {'NCCL all-gather (overlapped with the next step) + ' if world > 1 else ''}cosine scores vs {Q} text queries ({new_q} re-encoded per step)",
                   "frames_per_gpu_per_step": B, "frames_per_video": Fv, "queries": Q, "weights": "seeded random init (synthetic.py)",
                   "residual_stream": "fp32", "gemm_operands": "bf16, fp32 accumulate", "l2": "inputs_larger_than_l2",
                   "parallelism": f"frame-shard dp{world}" if world > 1 else "single GPU"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops"], "traffic": gemm_traffic()[0], "traffic_source": gemm_traffic()[1],
                     "peak_source": peaks["source"],
                     "kernel": "hb::gemm_kernel<CG=2,*> (all tcgen05 GEMM launches)", "launches_per_step": gemm_launches / args.steps,
                     "gemm_share_of_kernel_time": gemm_ms / kernel_ms_total if kernel_ms_total else None,
                     "whole_step_tflops": flops_frame * B / (ms_total / args.steps * 1e-3) / 1e12,
                     "whole_step_frac": flops_frame * B / (ms_total / args.steps * 1e-3) / 1e12 / peaks["bf16_tflops"]},
        "kernels": per,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "check": check,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(cfg, args.cpu_frames, 2, 1, Fv, Q)
        line["cpu_baseline"] = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
