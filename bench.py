#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (BASELINE.json):

    audio-seconds / second (RTFx), tdt-ctc-110m, 10 s clips, 1/2/4/8 x B200

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (PCM -> log-mel -> FastConformer -> TDT greedy) over
one batch of 64 synthetic 10 s clips per GPU (BASELINE.json configs[1]); weak scaling: every
rank owns its own 64 clips, the only exchange is one all-gather of the token buffers.

  value  : whole-job audio-seconds per second with the PCM already resident in HBM
           (pk_stage_pcm once, then pk_run_staged per step on the engine stream).
  e2e    : same metric through the public C-ABI with HOST buffers: every step copies its 41 MB of PCM from
           page-locked memory and reads its tokens back inside the timed region, driven as a serving loop
           (pk_stage_pcm + pk_run_staged + pk_prefetch_pcm(next batch) + pk_fetch_tokens: the H2D copy of
           batch i+1 runs under the kernels of batch i); e2e.sync_call = the single blocking call
           pk_transcribe_batch per batch.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section "Measurement".

--impl reference times the reference's own CPU implementation (oracle/_ref/libpkref.so,
the unmodified reference compiled by oracle/Makefile) on this box's host cores, one 10 s
clip per step (a bounded sample of the same workload).
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
import __graft_entry__ as ge  # noqa: E402

CLIP_SAMPLES = 160000
CLIP_SECONDS = 10.0
BATCH = 64
ENC_GFLOP_PER_CLIP = 28.23      # SURVEY.md section 8d, excl. the input-independent pos_proj
METRIC = "audio-seconds/sec (RTFx) tdt-ctc-110m 10s clips"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.rows, self.proc = dev, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = sorted(sm)[len(sm) // 2:] if sm else [0.0]     # upper half ~ samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_checkpoint(tmpdir):
    pkg = ge.load_package()
    from parakeet_cpp_b200 import synth
    cfg = pkg.make_110m_config(max_batch=BATCH, max_samples=CLIP_SAMPLES)
    wp = os.path.join(tmpdir, "pk110m_seed0.safetensors")
    if not os.path.exists(wp):
        W = synth.make_weights(cfg, seed=0)
        synth.save_safetensors(wp + ".tmp", W)
        os.replace(wp + ".tmp", wp)
    return pkg, synth, cfg, wp


def omp_threads():
    """CPUs the reference's OpenMP team can really run on: the affinity mask capped by the cgroup CPU quota
    (the GPU boxes show 128 CPUs but grant 16 CPUs of time; 128 spinning threads made the reference 4.6x slower).
    The team size is applied with omp_set_num_threads on the OpenMP runtime libpkref.so uses."""
    import ctypes
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def run_reference(args, rank, world):
    """The reference's own CPU path (Transcriber::transcribe, transcribe.hpp:99-179)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refbind as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpkref.so not built"}))
        return
    cores = omp_threads()
    pkg, synth, cfg, wp = make_checkpoint(args.tmp)
    m = R.RefModel(wp, "", 0)
    clips = [synth.make_audio(CLIP_SAMPLES, 1000 + i) for i in range(2)]
    for i in range(args.warmup):
        m.transcribe(clips[i % 2], "tdt")
        if i == 0 and args.warmup > 1:
            break                                  # one warm-up pass is enough for a 10 s CPU step
    t0 = time.perf_counter()
    stage = np.zeros(3)
    for i in range(args.steps):
        _, ms = m.transcribe(clips[i % 2], "tdt")
        stage += ms
    dt = time.perf_counter() - t0
    val = args.steps * CLIP_SECONDS / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "x real-time", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "tdt-ctc-110m TDT decode, 10 s 16 kHz synthetic clips", "clips_per_step": 1,
                       "note": "bounded sample: 1 clip per step of the 64-clip batch"},
            "cpu_baseline": {"value": val, "unit": "x real-time", "cores": cores, "kind": "reference",
                             "sample": f"{args.steps} x one 10 s clip, TDT, OpenMP team = {cores} threads (affinity mask capped by the cgroup CPU quota)",
                             "stage_ms_per_clip": {"mel": stage[0] / args.steps, "encoder": stage[1] / args.steps,
                                                   "decode": stage[2] / args.steps}},
            "e2e": {"value": val, "unit": "x real-time", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--decoder", default="tdt", choices=["tdt", "ctc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tmp", default=os.environ.get("PK_BENCH_TMP", "/tmp/pk_bench"))
    args = ap.parse_args()
    os.makedirs(args.tmp, exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        pkg, synth, cfg, wp = make_checkpoint(args.tmp)
    if world > 1:
        dist.barrier()
    pkg, synth, cfg, wp = make_checkpoint(args.tmp)
    eng = pkg.Engine(cfg, wp, local)
    dec = pkg.Decoder.TDT if args.decoder == "tdt" else pkg.Decoder.CTC

    # this rank's 64 clips (weak scaling): seeds 1000 + global clip index
    pcms = [synth.make_audio(CLIP_SAMPLES, 1000 + rank * BATCH + i) for i in range(BATCH)]
    # the step's host input: one packed fp32 buffer in PAGE-LOCKED host memory
    buf = torch.from_numpy(np.concatenate(pcms)).pin_memory().numpy()
    off = np.arange(BATCH + 1, dtype=np.int64) * CLIP_SAMPLES

    # single cross-GPU exchange: all-gather of the int32 token buffers
    gather = None
    if world > 1:
        ptr, rows, ints = eng.token_buffer()

        class _Buf:
            __cuda_array_interface__ = {"shape": (BATCH, ints), "typestr": "<i4", "data": (ptr, False), "version": 3}
        tok_local = torch.as_tensor(_Buf(), device=f"cuda:{local}")
        tok_all = torch.empty((world * BATCH, ints), dtype=torch.int32, device=f"cuda:{local}")

        def gather():
            eng.sync()
            dist.all_gather_into_tensor(tok_all, tok_local)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(eng.stream(), device=local)   # events must be recorded on the engine stream

    def timed(fn, steps):
        """K steps bracketed by barrier+sync; device time via events on the ENGINE stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        if world > 1:      # max over ranks, measured on the device
            t = torch.tensor([ms, wall * 1e3], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    # ---- device-resident throughput ("value")
    eng.stage(buf, off)

    def step_resident():
        eng.flush_l2()
        eng.run_staged(dec)
        if gather:
            gather()

    for _ in range(args.warmup):
        step_resident()
    l0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms, wall = timed(step_resident, args.steps)
    clocks = sampler.stop()
    launches = eng.launch_count() - l0
    ref_tokens = eng.fetch(BATCH)
    audio_s = args.steps * BATCH * CLIP_SECONDS * world
    value = audio_s / (max(ms, 1e-9) / 1e3)

    # ---- end-to-end through the public API with host buffers ("e2e"): every step copies its 41 MB of PCM
    # from page-locked host memory and reads its tokens back.  Two ways a caller can drive it:
    #   sync     : pk_transcribe_batch (one blocking call per batch, like the reference's transcribe())
    #   pipelined: pk_stage_pcm + pk_run_staged + pk_prefetch_pcm(next batch) + pk_fetch_tokens -- the H2D
    #              copy of batch i+1 runs under the kernels of batch i (double-buffered PCM on the device)
    tok_out = eng._tokens(BATCH)

    def step_sync():
        eng.flush_l2()
        arrs = eng.transcribe_packed(buf, off, dec, tok_out)     # H2D of buf + D2H of the token arrays inside
        if gather:
            gather()
        return arrs

    def step_pipelined():
        eng.flush_l2()
        eng.stage(buf, off)              # adopts the copy started by the previous step's prefetch
        eng.run_staged(dec)
        eng.prefetch(buf, off)           # this is the NEXT step's input: its H2D is inside the timed region too
        arrs = eng.fetch_into(tok_out)   # D2H of this step's tokens
        if gather:
            gather()
        return arrs

    def time_e2e(step_fn):
        for _ in range(2):
            o = step_fn()
        assert [o["ids"][b, :o["len"][b]].tolist() for b in range(BATCH)] == [[t.token_id for t in u] for u in ref_tokens]
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            o = step_fn()
        barrier()
        wall_ = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([wall_], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall_ = float(t[0])
        return wall_, o

    sync_wall, out = time_e2e(step_sync)
    eng.prefetch(buf, off)               # prime the pipeline (outside the timed region; every timed step issues its own)
    e2e_wall, out = time_e2e(step_pipelined)
    eng.stage(buf, off)                  # drain the last prefetch
    eng.sync()
    e2e_value = audio_s / e2e_wall
    n_tok = int(out["len"].sum())
    d2h = BATCH * (1 + eng.cap) * 4 + 3 * BATCH * eng.cap * 4      # token rows + start/end/conf as copied by fetch

    # ---- per-kernel-class device time (separate profiled pass; not the timed value)
    eng.profile_begin()
    PSTEPS = 3
    for _ in range(PSTEPS):
        eng.flush_l2()
        eng.run_staged(dec)
    prof = eng.profile_end()
    pk = peaks()
    gemm_ms, gemm_n, gemm_fl = prof["gemm"]
    gemm_tflops = gemm_fl / max(gemm_ms, 1e-9) / 1e9
    enc_ms = sum(prof[k][0] for k in ("subsample", "gemm", "layernorm", "attention", "dwconv")) / PSTEPS
    math_name = {0: "bf16x3", 1: "bf16", 2: "f32"}[int(cfg.math)]
    # fp32 CUDA-core GEMM is bounded by the fp32 FMA pipe, not by tcgen05: report against the
    # tensor roofline anyway (the bar the north star sets) and say so.
    # DRAM bytes of one launch of the dominant GEMM, from the committed ncu --set full capture
    # (bench.py cannot run ncu on itself; profiles/r01_traffic.json says which launch and how it was taken)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_note = tj["dram_bytes_per_launch"], f'{tj["kernel"]}; {tj["source"]}'
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05; all GEMM launches of one step)",
                "achieved": gemm_tflops, "peak": pk["bf16_sus"], "unit": "TFLOP/s",
                "frac": gemm_tflops / pk["bf16_sus"], "traffic": traffic, "traffic_of": traffic_note,
                "mma_frac": (3.0 if int(cfg.math) == 0 else 1.0) * gemm_tflops / pk["bf16_sus"] if int(cfg.math) != 2 else None,
                "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                "algorithmic_gflop_per_launch": gemm_fl / max(gemm_n, 1) / 1e9, "launches_per_step": gemm_n // PSTEPS,
                "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                "encoder": {"ms_per_clip": enc_ms / BATCH, "ms_per_batch": enc_ms,
                            "algorithmic_tflops": ENC_GFLOP_PER_CLIP * BATCH / max(enc_ms, 1e-9),
                            "frac_of_bf16_peak": ENC_GFLOP_PER_CLIP * BATCH / max(enc_ms, 1e-9) / pk["bf16_sus"]},
                "per_class_ms_per_step": {k: v[0] / PSTEPS for k, v in prof.items()}}

    line = {"metric": METRIC, "value": value, "unit": "x real-time", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": math_name, "data": "synthetic",
            "config": {"workload": f"tdt-ctc-110m {args.decoder.upper()} decode, batch=64x10s synthetic clips per GPU",
                       "clips_per_gpu_per_step": BATCH, "global_clips_per_step": BATCH * world,
                       "parallelism": f"utterance-sharded dp{world}, one all-gather of token buffers",
                       "l2": "256 MiB scratch written between steps (inside the timed region)",
                       "tokens_per_step_rank0": n_tok},
            "e2e": {"value": e2e_value, "unit": "x real-time", "h2d_bytes_per_step": int(buf.nbytes) * 1,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_wall / args.steps,
                    "api": "pk_stage_pcm + pk_run_staged + pk_prefetch_pcm(next batch) + pk_fetch_tokens: pinned host PCM in, "
                           "host token arrays out, H2D of batch i+1 under the kernels of batch i",
                    "sync_call": {"value": audio_s / sync_wall, "ms_per_step": 1e3 * sync_wall / args.steps,
                                  "api": "pk_transcribe_batch (one blocking call per batch)"}},
            "gpu_launches": int(launches), "wall_s": wall, "clocks": clocks, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refbind as R
        if R.available():
            cores = omp_threads()
            m = R.RefModel(wp, "", 0)
            m.transcribe(pcms[0][:32000], args.decoder)                 # touch the weights
            t0 = time.perf_counter()
            ids, stage = m.transcribe(pcms[0], args.decoder)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": CLIP_SECONDS / dt, "unit": "x real-time", "cores": cores, "kind": "reference",
                                    "sample": f"1 x 10 s clip of the batch (clip 0), OpenMP team = {cores} threads (cgroup CPU quota)",
                                    "stage_ms": {"mel": stage[0], "encoder": stage[1], "decode": stage[2]},
                                    "tokens_match_gpu": ids == [t.token_id for t in ref_tokens[0]]}
            m.close()
        else:
            line["cpu_baseline"] = {"value": None, "unit": "x real-time", "cores": 0, "kind": "reference",
                                    "sample": "oracle/_ref/libpkref.so not present"}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
