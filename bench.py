#!/usr/bin/env python
"""bench.py -- headline benchmark of the flake_b200 FLAC encoding hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  For N > 1 the driver launches it under
torch.distributed.run, one rank per GPU.

Workload (BASELINE.json configs[1], "C2"): flake -8 on a 1 h synthetic 16-bit
stereo 44.1 kHz stream = 158,760,000 inter-channel samples, 38,760 blocks of 4096.
A *step* is one pass of the hot path over that whole stream on every rank (each
rank encodes its own stream: frames shard with no collective, weak scaling).

  value   MSamples/s, KERNELS ONLY: PCM already resident in HBM (packed s16le, the
          WAV data layout), device-resident outputs, CUDA events on the launching
          stream, max over ranks.  No H2D/D2H, no MD5.
  e2e     the like-for-like figure against the reference arm, through the host-buffer
          C ABI with BOTH arms free to use every host core: a FIXED corpus of independent
          tracks (C5's shape: E2E_TRACKS x E2E_TRACK_SECONDS of flake -8 16-bit stereo,
          packed s16le in page-locked host memory = the WAV data layout) through
          flake_b200_encode_corpus -- H2D, kernels, D2H of frames + lengths, the MD5 of
          every track's PCM (multi-buffer SIMD on the host cores) and the final stream
          headers all inside the timed region.  The corpus is split over the ranks
          (strong scaling); value = corpus samples / max-over-ranks wall time.
  e2e_single_stream  ONE 1-hour stream through flake_b200_encode_stream (int32 in, the
          flake_encode_frame convention): bounded by the stream's serial MD5 chain on one
          host core, whatever the GPU does.
  parity  the single-stream output byte-compared with the compiled reference over the
          WHOLE stream, plus a sample of corpus tracks: frames compared / mismatching,
          compressed size delta.
  other_configs  C1 / C3 / C4 of BASELINE.json: device-resident rate, stage times,
          roofline of their dominant kernel, parity against the reference.

All arms and the full-size parity tests draw their PCM from ONE generator
(flake_b200.synth.long_pcm, CPU, deterministic).

`--impl reference` times the reference's own CPU implementation (oracle/_ref
compiled from the reference sources and driven from C by oracle/ref_shim.c, else
the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0         # B200_PROFILING.md fallback

CONFIGS = {
    # name: level, channels, bits, rate, samples, block (SURVEY.md section 8 shorthand)
    "C1": dict(level=5, ch=2, bps=16, rate=44100, samples=26_460_000, block=4096,
               what="flake -5 (LPC<=8, order estimate), 10 min 16-bit stereo 44.1 kHz"),
    "C2": dict(level=8, ch=2, bps=16, rate=44100, samples=158_760_000, block=4096,
               what="flake -8 (LPC<=12 log search, Rice partition order 0..6, mid/side estimate), "
                    "1 h 16-bit stereo 44.1 kHz"),
    "C3": dict(level=12, ch=2, bps=24, rate=96000, samples=57_600_000, block=8192,
               what="flake -12 (LPC<=32 exhaustive order search, partition order 0..8, VBS), "
                    "10 min 24-bit stereo 96 kHz"),
    "C4": dict(level=9, ch=8, bps=24, rate=48000, samples=28_800_000, block=4096,
               what="flake -9 (LPC<=12 log search, partition order 0..8, VBS), 10 min 8-channel 24-bit 48 kHz"),
}
C2 = CONFIGS["C2"]
RATE, CHANNELS, BPS, LEVEL, BLOCK = C2["rate"], C2["ch"], C2["bps"], C2["level"], C2["block"]
C2_SAMPLES = C2["samples"]
# the e2e corpus (same format and level as C2): E2E_TRACKS tracks of E2E_TRACK_SAMPLES samples
E2E_TRACKS = 128
E2E_TRACK_SAMPLES = 225 * RATE            # 3 min 45 s: 2423 blocks of 4096 (the last one short)


def _jsonable(o):
    """numpy scalars / arrays that found their way into the line"""
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError("not JSON serializable: %r" % type(o))


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_pcm(cfg, seed, nsamples=None):
    """int32 (n, ch) PCM of a BASELINE.json configuration: the one generator of every arm."""
    from flake_b200 import synth
    return synth.long_pcm(nsamples or cfg["samples"], cfg["ch"], cfg["bps"], cfg["rate"], seed=seed)


def csrc_fingerprint():
    """sha1 over the kernel sources: profiles/traffic.json is only quoted for the kernels it
    was captured from."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "flake_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 100` in the background during the timed regions (the profiling
    recipe's clocks line); median SM clock and any throttle reason that was ever active."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        mhz, mx, reasons, watts = [], None, set(), []
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                mhz.append(float(f[0])); mx = float(f[1]); watts.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(mhz),
                "power_w_max": max(watts) if watts else None}


# --------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# --------------------------------------------------------------------------
def cpu_reference_time(pcm_i32: np.ndarray, cfg, threads: int, steps: int, warmup: int):
    """Time the reference over pcm_i32: `threads` contiguous block ranges, one reference
    context and one thread each, every block through flake_encode_frame in a C loop
    (oracle/ref_shim.c; no per-block Python).  Returns (samples, [seconds], kind, bytes)."""
    from oracle import pyoracle as po
    n = pcm_i32.shape[0]
    times, nbytes = [], 0
    if po.have_ref():
        kind = "reference"
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out, _, _ = po.ref_encode_parallel(pcm_i32, cfg["rate"], cfg["bps"], cfg["level"], threads=threads)
            dt = time.perf_counter() - t0
            nbytes = len(out)
            if it >= warmup:
                times.append(dt)
    else:
        kind = "port"
        seg = (n // threads) // cfg["block"] * cfg["block"]
        res = [0] * threads

        def work(t):
            res[t] = len(po.encode_stream(pcm_i32[t * seg:(t + 1) * seg], cfg["rate"], cfg["bps"], cfg["level"])[0])
        for it in range(warmup + steps):
            ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            dt = time.perf_counter() - t0
            nbytes = sum(res)
            if it >= warmup:
                times.append(dt)
        n = seg * threads
    return n, times, kind, nbytes


def reference_bytes(pcm_i32: np.ndarray, cfg):
    """(frame bytes, bytes per block) of the whole stream from the reference on all host
    cores (untimed; the checker of the parity legs)."""
    from oracle import pyoracle as po
    if po.have_ref():
        out, per_block, _ = po.ref_encode_parallel(pcm_i32, cfg["rate"], cfg["bps"], cfg["level"], threads=0)
        return out, per_block.astype(np.int64), "reference"
    data, flen, fbs, _ = po.encode_stream(pcm_i32, cfg["rate"], cfg["bps"], cfg["level"])
    # the port reports frames; fold them into blocks
    blk = (np.cumsum(fbs) - fbs) // cfg["block"]
    per_block = np.bincount(blk, weights=flen).astype(np.int64)
    return np.frombuffer(data, dtype=np.uint8), per_block, "port"


def parity_record(got: np.ndarray, flen: np.ndarray, fbs: np.ndarray, want: np.ndarray, per_block: np.ndarray, cfg, kind):
    """Frames compared / mismatching against the reference's bytes, size delta in percent."""
    nframes = int(len(flen))
    rec = {"frames_compared": nframes, "frames_mismatching": 0,
           "size_delta_pct": round(100.0 * (len(got) - len(want)) / max(1, len(want)), 6),
           "bytes": int(len(got)), "reference_bytes": int(len(want)), "against": kind,
           "scope": "whole stream, every frame byte"}
    if len(got) == len(want) and np.array_equal(got, want):
        return rec
    blk = (np.cumsum(fbs) - fbs) // cfg["block"]
    mine = np.bincount(blk, weights=flen, minlength=len(per_block)).astype(np.int64)
    mo = np.concatenate([[0], np.cumsum(mine)])
    ro = np.concatenate([[0], np.cumsum(per_block)])
    fpb = np.bincount(blk, minlength=len(per_block))
    bad = 0
    for b in range(min(len(mine), len(per_block))):
        x, y = got[mo[b]:mo[b + 1]], want[ro[b]:ro[b + 1]]
        if len(x) != len(y) or not np.array_equal(x, y):
            bad += int(fpb[b])
    rec["frames_mismatching"] = bad + abs(len(mine) - len(per_block))
    return rec


def corpus_track_i32(index, nsamples):
    from flake_b200 import synth
    return synth.corpus_track(index, nsamples, CHANNELS, BPS, RATE)


def run_reference_arm(args):
    """The reference's own CPU implementation on this box's host cores -- all of them.

    libflake is single threaded and its API is serial per stream (frame numbers and the MD5
    chain through every block), so the way the reference uses a whole box is one file per
    core: exactly the shape of the GPU arm's e2e workload (a corpus of independent tracks).
    Each step encodes a bounded sample of that corpus: the first `cores` tracks, a prefix of
    each (same generator, same track indices as the GPU arm), one thread and one reference
    context per track, every block through flake_encode_frame in a C loop.  `single_thread`
    is one track on one core, for orientation.
    """
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = env_int("FLAKE_BENCH_REF_THREADS", cores)
    per_thread = env_int("FLAKE_BENCH_REF_SAMPLES_PER_THREAD", BLOCK * 700)      # 2.9 M samples, ~0.25 s of CPU
    per_thread = min(per_thread, E2E_TRACK_SAMPLES // BLOCK * BLOCK)
    ntracks = env_int("FLAKE_BENCH_E2E_TRACKS", E2E_TRACKS)
    pcm = np.concatenate([corpus_track_i32(t % ntracks, per_thread) for t in range(threads)])
    total, times, kind, _ = cpu_reference_time(pcm, C2, threads, args.steps, max(1, min(args.warmup, 1)))
    ms = 1e3 * float(np.mean(times))
    val = total / (ms * 1e-3) / 1e6
    sample = "the first %d samples (%.0f s of audio) of corpus tracks 0..%d per step, one thread + one reference " \
             "context per track, C loop over flake_encode_frame" % (per_thread, per_thread / RATE, threads - 1)
    one = None
    if not os.environ.get("FLAKE_BENCH_SKIP_SINGLE_THREAD"):
        t1, times1, _, _ = cpu_reference_time(pcm[:per_thread * 2], C2, 1, 1, 0)
        one = {"value": round(t1 / times1[0] / 1e6, 3), "unit": "MSamples/s", "cores": 1,
               "sample": "%d samples on one thread" % t1}
    line = {
        "impl": "reference", "metric": "MSamples/s encoded, flake -8", "value": round(val, 3),
        "unit": "MSamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(val, 3), "unit": "MSamples/s", "cores": threads, "kind": kind,
                         "sample": sample, "host_cores_available": cores},
        "e2e": {"value": round(val, 3), "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "audio_seconds_per_s": round(val * 1e6 / RATE, 1),
        "single_thread": one,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    ntracks = env_int("FLAKE_BENCH_E2E_TRACKS", E2E_TRACKS)
    return {"workload": "C2: " + C2["what"] + " = 158760000 samples, 38760 blocks of 4096, per GPU",
            "e2e_workload": "C5 shape, same format and level: a fixed corpus of %d independent tracks x %d samples "
                            "(%.1f h of audio in all), split over the ranks; both arms may use every host core" % (
                                ntracks, E2E_TRACK_SAMPLES, ntracks * E2E_TRACK_SAMPLES / RATE / 3600.0),
            "level": LEVEL, "block_size": BLOCK, "channels": CHANNELS, "bits_per_sample": BPS,
            "sample_rate": RATE, "samples_per_gpu": C2_SAMPLES,
            "generator": "flake_b200.synth.long_pcm(seed = rank) for the 1 h stream, synth.corpus_track(i) for "
                         "the corpus; the same in both arms",
            "l2": "inputs (635 MB packed PCM per pass) larger than the 126 MB L2; no explicit flush",
            "sharding": "one stream per rank, no collective" if n_gpus > 1 else "single stream"}


# --------------------------------------------------------------------------
# GPU arm helpers
# --------------------------------------------------------------------------
def to_device_packed(pcm_i32: np.ndarray, cfg, dev):
    """The WAV data-chunk layout in HBM: packed little-endian, ceil(bps/8) bytes per sample."""
    import torch
    from flake_b200 import api, synth
    if cfg["bps"] == 16:
        return torch.from_numpy(pcm_i32.astype(np.int16)).to(dev), api.PCM_S16LE
    if cfg["bps"] == 24:
        raw = np.frombuffer(synth.pack_pcm(pcm_i32, 24), dtype=np.uint8)
        return torch.from_numpy(raw.copy()).to(dev), api.PCM_S24LE
    return torch.from_numpy(pcm_i32).to(dev), api.PCM_S32


class DevicePass:
    """One configuration resident in HBM: the device-only pass, its timing and roofline."""

    def __init__(self, lib, cfg, pcm_i32, dev, stream):
        import torch
        from flake_b200 import api
        self.lib, self.cfg, self.dev, self.stream = lib, cfg, dev, stream
        self.n = pcm_i32.shape[0]
        self.enc = api.Encoder(lib, cfg["ch"], cfg["rate"], cfg["bps"], self.n, cfg["level"])
        self.enc.init()
        self.ctx = C.byref(self.enc.ctx)
        self.d_pcm, self.fmt = to_device_packed(pcm_i32, cfg, dev)
        self.bytes_per_sample = cfg["ch"] * ((cfg["bps"] + 7) // 8)
        cap_s, cap_b, cap_f = C.c_ulonglong(), C.c_ulonglong(), C.c_uint()
        assert lib.flake_b200_device_capacity(self.ctx, C.byref(cap_s), C.byref(cap_b), C.byref(cap_f)) == 0
        self.chunk = int(cap_s.value)
        self.nchunks = (self.n + self.chunk - 1) // self.chunk
        self.d_out = torch.empty(int(cap_b.value), dtype=torch.uint8, device=dev)
        self.d_flen = torch.empty(int(cap_f.value), dtype=torch.int32, device=dev)
        self.d_sum = torch.zeros((self.nchunks, 3), dtype=torch.int64, device=dev)   # 24-byte FbSummary per chunk
        self.base = self.d_pcm.data_ptr()

    def run(self):
        allow_vbs = bool(self.enc.ctx.params.allow_vbs)
        for k in range(self.nchunks):
            s0 = k * self.chunk
            ns = min(self.chunk, self.n - s0)
            first = s0 if allow_vbs else s0 // self.cfg["block"]
            rc = self.lib.flake_b200_encode_device(
                self.ctx, self.base + s0 * self.bytes_per_sample, self.fmt, ns, first, self.d_out.data_ptr(),
                self.d_flen.data_ptr(), None, self.d_sum[k].data_ptr(), self.stream.cuda_stream)
            if rc:
                raise RuntimeError("flake_b200_encode_device failed: %d" % rc)

    def totals(self):
        summ = self.d_sum.cpu().numpy().view(np.uint8).reshape(self.nchunks, 24)
        frames = int(sum(int(np.frombuffer(summ[k, 0:4].tobytes(), np.uint32)[0]) for k in range(self.nchunks)))
        out_bytes = int(sum(int(np.frombuffer(summ[k, 8:16].tobytes(), np.uint64)[0]) for k in range(self.nchunks)))
        return frames, out_bytes

    def stage_profile(self, passes):
        import torch
        self.enc.set_profiling(True)
        for _ in range(passes):
            self.run()
        torch.cuda.synchronize()
        stages = self.enc.stage_times()
        self.enc.set_profiling(False)
        return stages

    def close(self):
        self.enc.close()


def roofline_record(stages, passes, nsamples, nchunks, in_bytes_per_sample, out_bytes, peak_gbs, peak_src, traffic):
    stage_ms = {k: v[0] for k, v in stages.items()}
    stage_n = {k: v[1] for k, v in stages.items()}
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    dom_avg_ms = stage_ms[dom] / max(1, stage_n[dom])
    # algorithmic bytes per launch: packed PCM in + FLAC frame bytes out,
    # for the samples one launch (= one chunk) processes
    algo_bytes = (in_bytes_per_sample + out_bytes / float(nsamples)) * nsamples / float(nchunks)
    achieved = algo_bytes / (dom_avg_ms * 1e-3) / 1e9
    step_ms = sum(stage_ms.values()) / passes
    rec = {"bound": "hbm", "kernel": "k_" + dom, "achieved": round(achieved, 2), "peak": peak_gbs,
           "unit": "GB/s", "frac": round(achieved / peak_gbs, 5),
           "traffic": (traffic or {}).get(dom) if traffic else None,
           "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes),
           "kernel_ms_per_launch": round(dom_avg_ms, 4),
           "whole_step_frac": round(algo_bytes * nchunks / (step_ms * 1e-3) / 1e9 / peak_gbs, 5),
           "stage_ms_per_step": {k: round(v / passes, 3) for k, v in stage_ms.items()},
           "stage_share": {k: round(v / max(1e-9, sum(stage_ms.values())), 4) for k, v in stage_ms.items()},
           "note": "path is instruction-bound (integer/FP64 pipes), not HBM-bound; see DESIGN.md 6"}
    if traffic:
        rec["traffic_all_kernels"] = {k: traffic.get(k) for k in stage_ms if traffic.get(k) is not None}
    return rec


def host_encode(lib, cfg, pcm_i32):
    """Whole stream through the host-buffer C ABI (int32 in, frames + lengths out)."""
    from flake_b200 import api
    enc = api.Encoder(lib, cfg["ch"], cfg["rate"], cfg["bps"], pcm_i32.shape[0], cfg["level"])
    enc.init()
    try:
        data, flen, fbs = enc.encode_stream(pcm_i32, api.PCM_S32, pcm_i32.shape[0])
    finally:
        enc.close()
    return data, flen, fbs


def load_traffic():
    """profiles/traffic.json (ncu dram bytes per launch) -- only if it was captured from the
    kernel sources of this tree."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "profiles/traffic.json missing"
    if t.get("csrc_sha1_16") != csrc_fingerprint():
        return None, "profiles/traffic.json is from other kernel sources (%s): not quoted" % t.get("csrc_sha1_16")
    return t, t.get("source")


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_gpu_arm(args):
    # stdout carries exactly one JSON line: whatever native libraries print there (NCCL's
    # "NCCL version ..." banner under torchrun) is sent to stderr; the line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from flake_b200 import api

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flake_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")     # host-side waits that must not occupy SMs

    nsamples = env_int("FLAKE_BENCH_SAMPLES", C2_SAMPLES)
    lib = api.load_library()
    lib.flake_b200_set_device(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # a real (non-default) torch stream: the library treats a NULL handle as "use the
    # context's own stream", and torch events must see the stream the kernels run on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    # ---- inputs: rank r encodes stream r of the corpus ------------------------------
    pcm_np = workload_pcm(C2, rank, nsamples)
    dp = DevicePass(lib, C2, pcm_np, dev, stream)

    # ---- device-resident timing ------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        dp.run()
    barrier()
    st0 = dp.enc.stats()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        dp.run()
        ev[i + 1].record(stream)
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    st1 = dp.enc.stats()
    launches = int(st1.kernel_launches - st0.kernel_launches)
    # per-kernel CUDA events (on the launching stream) in passes of their own, so that the
    # event records do not sit inside the timed region above
    prof_passes = max(1, min(args.steps, 5))
    stages = dp.stage_profile(prof_passes)
    frames, out_bytes = dp.totals()
    ms_per_step = allmax(total_ms) / args.steps
    value = world * nsamples / (ms_per_step * 1e-3) / 1e6
    rank_ms = None
    if world > 1:
        g = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.tensor([total_ms / args.steps], dtype=torch.float64, device=dev))
        rank_ms = [round(float(x.item()), 3) for x in g]

    # ---- end-to-end through the host-buffer C ABI -----------------------------------
    e2e_steps = max(1, min(args.steps, env_int("FLAKE_BENCH_SINGLE_STEPS", 2)))
    h_pcm32 = torch.from_numpy(pcm_np).pin_memory()
    pcm_pinned = h_pcm32.numpy()
    ctx0 = C.byref(dp.enc.ctx)
    cap = int(lib.flake_b200_max_encoded_size(ctx0, nsamples))
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    nblocks = (nsamples + BLOCK - 1) // BLOCK
    h_flen = np.zeros(nblocks + 1, dtype=np.uint32)
    h_fbs = np.zeros(nblocks + 1, dtype=np.uint32)
    nf = C.c_uint(0)
    e2e_ms, e2e_bytes, md5_hex = [], 0, None
    h2d = d2h = 0
    enc2 = api.Encoder(lib, CHANNELS, RATE, BPS, nsamples, LEVEL)
    enc2.init()
    for it in range(1 + e2e_steps):
        lib.flake_b200_reset_stream(C.byref(enc2.ctx))
        barrier()
        t0 = time.perf_counter()
        rc = lib.flake_b200_encode_stream(C.byref(enc2.ctx), pcm_pinned.ctypes.data, api.PCM_S32, nsamples,
                                          h_out.data_ptr(), cap, h_flen.ctypes.data, h_fbs.ctypes.data,
                                          nblocks + 1, C.byref(nf))
        si, si_bytes = enc2.streaminfo()          # final STREAMINFO incl. MD5: the stream is complete
        dt = time.perf_counter() - t0
        if rc < 0:
            raise RuntimeError("flake_b200_encode_stream failed: %d" % rc)
        st = enc2.stats()
        if it >= 1:
            e2e_ms.append(dt * 1e3)
            e2e_bytes, h2d, d2h = int(rc), int(st.h2d_bytes), int(st.d2h_bytes)
            md5_hex = bytes(si.md5sum).hex()
    e2e_stats = enc2.stats()
    enc2.close()
    e2e_ms_max = allmax(float(np.mean(e2e_ms)))
    e2e_value = world * nsamples / (e2e_ms_max * 1e-3) / 1e6
    single_out = h_out[:e2e_bytes].numpy().copy() if rank == 0 else None     # for the parity leg below

    # ---- the per-block API (what the unmodified CLI calls): one synchronous call per block ---
    per_block_rec = None
    if rank == 0 and not os.environ.get("FLAKE_BENCH_SKIP_PER_BLOCK"):
        try:
            per_block_rec = per_block_leg(lib, api, pcm_np)
        except Exception as exc:
            per_block_rec = {"value": None, "error": str(exc)[:200]}

    # ---- e2e: the fixed corpus through flake_b200_encode_corpus, split over the ranks ------
    del h_out, h_pcm32, pcm_pinned
    corpus_rec = run_corpus_leg(lib, api, torch, dist, rank, world, local, host_group, allmax, barrier,
                                max(1, min(args.steps, env_int("FLAKE_BENCH_E2E_STEPS", 3))))
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        dp.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events between kernels) ---------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    traffic, traffic_src = load_traffic()
    roofline = roofline_record(stages, prof_passes, nsamples, dp.nchunks, CHANNELS * 2, out_bytes, peak_gbs, peak_src,
                               (traffic or {}).get("C2") if traffic else None)
    roofline["traffic_source"] = traffic_src

    # ---- parity of the e2e output against the reference, whole stream (untimed) ----------
    parity = None
    cpu = None
    try:
        got = single_out
        want, per_block, kind = reference_bytes(pcm_np, C2)
        parity = parity_record(got, h_flen[:nf.value].astype(np.int64), h_fbs[:nf.value].astype(np.int64),
                               want, per_block, C2, kind)
        parity["md5_matches_pcm"] = (md5_hex == hashlib.md5(pcm_np.astype("<i2").tobytes()).hexdigest())
    except Exception as exc:
        parity = {"frames_compared": 0, "error": str(exc)[:300]}

    # ---- CPU baseline: the compiled reference on ONE host core, bounded sample --------
    try:
        budget = env_int("FLAKE_BENCH_CPU_SAMPLES", BLOCK * 14000)       # 57 M samples ~ 5 s of CPU
        budget = min(budget, nsamples // BLOCK * BLOCK)
        total, times, kind, _ = cpu_reference_time(pcm_np[:budget], C2, 1, 1, 0)
        cpu = {"value": round(total / times[0] / 1e6, 3), "unit": "MSamples/s", "cores": 1, "kind": kind,
               "sample": "first %d samples (%.0f s of audio) of rank 0's stream, one flake_encode_frame "
                         "loop in C, 1 thread" % (total, total / RATE),
               "host_cores_available": os.cpu_count()}
    except Exception as exc:       # the baseline is informative; never fail the bench on it
        cpu = {"value": None, "unit": "MSamples/s", "cores": 1, "kind": "unavailable", "sample": str(exc)}

    # ---- the other BASELINE.json configurations -----------------------------------------
    others = {}
    dp.close()
    del dp
    if not os.environ.get("FLAKE_BENCH_SKIP_OTHERS"):
        for name in ("C1", "C3", "C4"):
            try:
                others[name] = run_other_config(lib, name, dev, stream, peak_gbs, peak_src, traffic, torch)
            except Exception as exc:
                others[name] = {"error": str(exc)[:300]}

    line = {
        "metric": "MSamples/s encoded, flake -8", "value": round(value, 2), "unit": "MSamples/s",
        "value_scope": "kernels only: PCM and outputs resident in HBM, no copies, no MD5 (e2e is the like-for-like figure)",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": workload_config(world),
        "audio_seconds_per_s": round(value * 1e6 / RATE, 1),
        "e2e": corpus_rec,
        "e2e_single_stream": {
            "value": round(e2e_value, 2), "unit": "MSamples/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms_max, 2),
            "workload": "one 1 h stream per rank through flake_b200_encode_stream",
            "input": "int32 interleaved host buffer (flake_encode_frame convention)",
            "includes": "H2D, all kernels, D2H of frames+lengths, MD5 of the PCM, final STREAMINFO",
            "md5": md5_hex,
            "md5_thread_ms": round(e2e_stats.md5_ms, 1),
            "gpu_ms": round(e2e_stats.gpu_ms, 1),
            "wall": "one stream's MD5 is a serial chain on one host core (md5_thread_ms of ms_per_step)"},
        "per_block_api": per_block_rec,
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "other_configs": others,
        "frames_per_step": frames, "compressed_bytes_per_step": out_bytes,
        "compression_ratio": round(out_bytes / float(nsamples * CHANNELS * 2), 4),
        "step_ms": [round(x, 2) for x in step_ms],
        "rank_ms_per_step": rank_ms,
    }
    os.write(real_stdout, (json.dumps(line, default=_jsonable) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_other_config(lib, name, dev, stream, peak_gbs, peak_src, traffic, torch):
    """Device-resident rate, stage times and roofline of C1 / C3 / C4, and whole-stream parity of
    the host-buffer path against the reference."""
    cfg = CONFIGS[name]
    n = env_int("FLAKE_BENCH_%s_SAMPLES" % name, cfg["samples"])
    pcm = workload_pcm(cfg, 0, n)
    dp = DevicePass(lib, cfg, pcm, dev, stream)
    steps = 5
    for _ in range(3):
        dp.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        dp.run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    stages = dp.stage_profile(3)
    frames, out_bytes = dp.totals()
    in_bps = cfg["ch"] * ((cfg["bps"] + 7) // 8)
    roof = roofline_record(stages, 3, n, dp.nchunks, in_bps, out_bytes, peak_gbs, peak_src,
                           (traffic or {}).get(name) if traffic else None)
    dp.close()
    del dp
    rec = {"workload": name + ": " + cfg["what"], "samples": n, "value": round(n / (ms * 1e-3) / 1e6, 2),
           "unit": "MSamples/s", "channel_msamples_per_s": round(cfg["ch"] * n / (ms * 1e-3) / 1e6, 2),
           "ms_per_step": round(ms, 3), "steps": steps, "frames": frames, "compressed_bytes": out_bytes,
           "compression_ratio": round(out_bytes / float(n * in_bps), 4), "roofline": roof}
    try:
        got, flen, fbs = host_encode(lib, cfg, pcm)
        want, per_block, kind = reference_bytes(pcm, cfg)
        rec["parity"] = parity_record(got, flen.astype(np.int64), fbs.astype(np.int64), want, per_block, cfg, kind)
    except Exception as exc:
        rec["parity"] = {"frames_compared": 0, "error": str(exc)[:300]}
    return rec


def per_block_leg(lib, api, pcm_i32, nblocks=1500):
    """flake_encode_frame in a loop, one 4096-sample block per synchronous call (encode.c:979-1008,
    flake/flake.c:624-663): what a caller that links the library without the batch calls gets.
    Latency bound: upload, five launches, two waits and the download of one block per call."""
    enc = api.Encoder(lib, CHANNELS, RATE, BPS, nblocks * BLOCK, LEVEL)
    enc.init()
    ctx = C.byref(enc.ctx)
    pcm = np.ascontiguousarray(pcm_i32[:nblocks * BLOCK], dtype=np.int32)
    base, step = pcm.ctypes.data, BLOCK * CHANNELS * 4
    for b in range(50):
        lib.flake_encode_frame(ctx, base + b * step, BLOCK)
    lib.flake_b200_reset_stream(ctx)
    t0 = time.perf_counter()
    nbytes = 0
    for b in range(nblocks):
        fs = lib.flake_encode_frame(ctx, base + b * step, BLOCK)
        if fs <= 0:
            raise RuntimeError("flake_encode_frame returned %d" % fs)
        nbytes += fs
    dt = time.perf_counter() - t0
    enc.close()
    return {"value": round(nblocks * BLOCK / dt / 1e6, 2), "unit": "MSamples/s", "us_per_call": round(dt / nblocks * 1e6, 1),
            "blocks": nblocks, "bytes": nbytes,
            "note": "one block per synchronous flake_encode_frame call; the batch calls exist because of this"}


def pcie_probe(torch, dev, nbytes=1 << 30):
    """Pinned host <-> device copy rate of this GPU's link, GB/s (h2d, d2h): what bounds the e2e leg."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = []
    for src, dst in ((h, d), (d, h)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out.append(round(3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1))
    return out


def host_read_probe(buf_u8: np.ndarray, threads: int, seconds: float = 0.25):
    """Aggregate read rate of the host cores over `buf_u8` (GB/s): every thread sums its own slice
    (numpy releases the GIL).  The e2e leg moves ~10 bytes of host DRAM traffic per sample (MD5
    read + DMA read + DMA write), so this is the wall the box runs into as GPUs are added."""
    view = buf_u8[:buf_u8.size // 8 * 8].view(np.uint64)
    per = view.size // max(1, threads)
    done = [0] * threads
    stop = time.perf_counter() + seconds

    def work(t):
        sl = view[t * per:(t + 1) * per]
        while time.perf_counter() < stop:
            int(sl.sum())
            done[t] += sl.nbytes
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return round(sum(done) / (time.perf_counter() - t0) / 1e9, 1)


def run_corpus_leg(lib, api, torch, dist, rank, world, local, host_group, allmax, barrier, steps):
    """e2e: the fixed corpus (E2E_TRACKS tracks, packed s16le in page-locked host memory) through
    flake_b200_encode_corpus.  Rank r takes tracks r, r + world, ...; one library call per step
    encodes them on this rank's GPU (two worker threads, DMA straight from / to the caller's
    buffers) while cores/world MD5 workers hash them in SIMD lanes.  Timed with the host clock
    around the call (its inputs and outputs are host buffers), barrier before, max over ranks."""
    from flake_b200 import corpus as fc, synth
    ntracks = env_int("FLAKE_BENCH_E2E_TRACKS", E2E_TRACKS)
    n = env_int("FLAKE_BENCH_E2E_TRACK_SAMPLES", E2E_TRACK_SAMPLES)
    mine = list(range(rank, ntracks, world))
    cores = os.cpu_count() or 1
    md5_threads = env_int("FLAKE_BENCH_MD5_THREADS", max(1, cores // world))
    dev = torch.device("cuda", local)
    co = fc.Corpus(lib, CHANNELS, RATE, BPS, LEVEL, api.PCM_S16LE, devices=[local], longest=n,
                   md5_threads=md5_threads, threads_per_device=env_int("FLAKE_BENCH_GPU_THREADS", 2))
    cap = co.max_encoded_size(n)
    fcap = co.frame_cap(n)
    h_in = torch.empty((len(mine), n, CHANNELS), dtype=torch.int16).pin_memory()
    h_out = torch.empty((len(mine), cap), dtype=torch.uint8).pin_memory()
    flen = np.zeros((len(mine), fcap), dtype=np.uint32)
    fbs = np.zeros((len(mine), fcap), dtype=np.uint32)
    in_np, out_np = h_in.numpy(), h_out.numpy()
    for j, t in enumerate(mine):
        synth.corpus_track(t, n, CHANNELS, BPS, RATE, out=in_np[j])
    items = (api.FlakeB200CorpusStream * max(1, len(mine)))()
    for j in range(len(mine)):
        it = items[j]
        it.pcm = in_np[j].ctypes.data; it.nsamples = n
        it.out = out_np[j].ctypes.data; it.out_cap = cap
        it.frame_len = flen[j].ctypes.data; it.frame_bs = fbs[j].ctypes.data; it.frame_cap = fcap
    stats = api.FlakeB200CorpusStats()
    hdr = (C.c_ubyte * 16384)()
    times = []
    for itn in range(1 + steps):
        barrier()
        t0 = time.perf_counter()
        rc = lib.flake_b200_corpus_encode(co.handle, items, len(mine), C.byref(stats))
        for j in range(len(mine)):                       # the final stream headers: the streams are complete files
            lib.flake_b200_corpus_stream_header(C.byref(co.ctx), C.byref(items[j]), hdr, len(hdr))
        dt = time.perf_counter() - t0
        if rc < 0:
            raise RuntimeError("flake_b200_corpus_encode failed: %d %s" % (rc, stats.error.decode(errors="replace")))
        if itn >= 1:
            times.append(dt * 1e3)
    my_ms = float(np.mean(times))
    ms = allmax(my_ms)
    total_samples = ntracks * n
    rank_ms = None
    if world > 1:
        g = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.tensor([my_ms], dtype=torch.float64, device=dev))
        rank_ms = [round(float(x.item()), 2) for x in g]
    h2d = torch.tensor([float(stats.h2d_bytes), float(stats.d2h_bytes), float(stats.kernel_launches)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d)
    h2d_b, d2h_b, launches = (int(x) for x in h2d.tolist())
    rec = None
    if rank == 0:
        # parity of a sample of this rank's tracks against the reference (untimed)
        par = {"tracks_compared": 0, "frames_compared": 0, "frames_mismatching": 0, "md5_mismatching": 0}
        try:
            for j in range(min(len(mine), env_int("FLAKE_BENCH_E2E_PARITY_TRACKS", 4))):
                pcm32 = in_np[j].astype(np.int32)
                want, per_block, kind = reference_bytes(pcm32, C2)
                r = parity_record(out_np[j][:items[j].bytes], flen[j][:items[j].nframes].astype(np.int64),
                                  fbs[j][:items[j].nframes].astype(np.int64), want, per_block, C2, kind)
                par["tracks_compared"] += 1
                par["frames_compared"] += r["frames_compared"]
                par["frames_mismatching"] += r["frames_mismatching"]
                par["md5_mismatching"] += int(bytes(items[j].md5sum) != hashlib.md5(in_np[j].tobytes()).digest())
                par["against"] = kind
        except Exception as exc:
            par["error"] = str(exc)[:200]
        pcie = None
        try:
            pcie = pcie_probe(torch, dev)
        except Exception:
            pass
        host_gbs = None
        try:
            host_gbs = host_read_probe(in_np.reshape(-1).view(np.uint8), cores)
        except Exception:
            pass
        rec = {"value": round(total_samples / (ms * 1e-3) / 1e6, 2), "unit": "MSamples/s",
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b, "ms_per_step": round(ms, 2),
               "scaling": "strong", "workload": "fixed corpus: %d tracks x %d samples (%.0f s each, %.2f h in all), "
               "flake -8, 16-bit stereo 44.1 kHz, packed s16le in page-locked host memory; tracks r, r+N, ... on rank r"
               % (ntracks, n, n / RATE, total_samples / RATE / 3600.0),
               "includes": "H2D from the caller's buffers, all kernels, D2H of frames + lengths into the caller's buffers, "
                           "MD5 of every track, final stream headers",
               "api": "flake_b200_corpus_encode (one call per rank and step)",
               "audio_seconds_per_s": round(total_samples / (ms * 1e-3) / RATE, 1),
               "tracks": ntracks, "tracks_per_rank": len(mine), "rank_ms_per_step": rank_ms,
               "gpu_launches_per_step": launches,
               "rank0": {"md5_ms": round(stats.md5_ms, 2), "md5_threads": int(stats.md5_threads),
                         "md5_lanes_per_thread": int(stats.md5_lanes), "gpu_threads": int(stats.gpu_threads),
                         "chunks": int(stats.chunks), "gpu_worker_ms": round(stats.device_ms[0], 2),
                         "h2d_gbs": round(stats.h2d_bytes / (my_ms * 1e-3) / 1e9, 1),
                         "d2h_gbs": round(stats.d2h_bytes / (my_ms * 1e-3) / 1e9, 1),
                         "host_cores": cores},
               "pcie_probe_gbs": {"h2d": pcie[0], "d2h": pcie[1]} if pcie else None,
               "host_read_probe_gbs": host_gbs,
               "host_bytes_per_sample": round((2 * CHANNELS * 2 * total_samples + d2h_b) / float(total_samples), 2),
               "bound": "N = 1: the GPU's PCIe link (pcie_probe_gbs); more GPUs: host DRAM traffic (MD5 read + DMA read + "
                        "DMA write per sample against host_read_probe_gbs)",
               "parity": par}
    co.close()
    del h_in, h_out
    if world > 1:
        dist.barrier(group=host_group)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
