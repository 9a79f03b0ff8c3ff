#!/usr/bin/env python
"""bench.py -- SEA hot-path throughput on B200 (BASELINE.json metric: decode & encode Msamples/s, fraction of roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference [--gpus N --steps K --warmup W] the reference's own CPU decoder on the host cores

A "step" is one pass of the hot path over one batch of synthetic streams.  Headline = batch decode of BASELINE config 4
(4096 independent stereo 60 s 44.1 kHz CBR-3 streams per GPU, weak scaling: every rank owns its own 4096 streams, nothing is
exchanged on the data path).  `value` is measured with the .sea bytes already resident in HBM; `e2e` goes through the
host-buffer C-ABI call (pinned host memory, H2D and D2H inside the timed region).  The same JSON line also carries the encode
throughput (config 5 shape: 1024 stereo 60 s streams, CBR 3 and VBR 3.0) with its INT32 roofline, the HBM roofline of the
decode kernel and a CPU baseline timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE, CHANNELS, SECONDS = 44100, 2, 60
ALG_OPS_PER_CAND_SAMPLE = 49  # SURVEY.md 8d op count of encoder_base.rs:64-89 + lms.rs


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(index: int) -> None:
    """Pins this rank's host threads (and so its first-touch pinned buffers) to the CPUs NVML reports as local to the GPU: the
    host-buffer (e2e) path moves ~6.5 GB per step over PCIe, and with several ranks per box remote-node buffers share one
    inter-socket link."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def ncu_traffic(streams: int, seconds: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel per launch, from the committed `ncu --set full` capture
    (profiles/r01_decode_traffic.json: bytes per stream of the 60 s config-4 shape); None when the shape differs."""
    p = os.path.join(ROOT, "profiles", "r01_decode_traffic.json")
    try:
        t = json.load(open(p))
        if seconds != t["seconds"]:
            return None
        return float(t["dram_bytes_per_launch"]) / t["streams"] * streams
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm

def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_decode_baseline(threads: int, reps: int, seconds: int = SECONDS, prefer_ref: bool = True):
    """Times the reference's CPU decode of one config-4 stream per core.  kind 'reference' = the reference's own C decoder
    (c/sea.h compiled into oracle/_ref, forked workers); 'port' = the oracle restatement (pthreads)."""
    from oracle import sea_oracle as O
    from sea_codec_b200 import synth

    pcm = synth.gen_stream(0, seconds * RATE, CHANNELS, RATE)
    st = O.make_settings(3.0)
    t0 = time.perf_counter()
    sea = O.sea_encode(pcm, RATE, CHANNELS, st)
    enc_single_s = time.perf_counter() - t0
    if prefer_ref and os.path.exists(O.REF_BENCH_PATH):
        with tempfile.NamedTemporaryFile(suffix=".sea", delete=False) as f:
            f.write(sea)
            path = f.name
        try:
            secs, samples = O.ref_c_bench(path, threads, reps)
        finally:
            os.unlink(path)
        kind = "reference"
    else:
        secs, samples = O.bench("decode", threads, reps, pcm, RATE, CHANNELS, st, sea)
        kind = "port"
    return dict(kind=kind, secs=secs, samples=samples, value=samples / secs / 1e6, pcm=pcm, sea=sea, settings=st,
                enc_single_msamples=pcm.size / enc_single_s / 1e6)


def reference_arm(args, info):
    if info.rank != 0:
        return
    cores = host_cores()
    # a bounded sample: every core decodes one 60 s stereo stream `reps` times per step
    reps = 4
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_decode_baseline(cores, reps) if i == 0 else r_again(r, cores, reps)
        if i >= args.warmup:
            vals.append(r["value"])
    value = float(np.mean(vals))
    ms = 1e3 * (cores * reps * SECONDS * RATE * CHANNELS) / (value * 1e6)
    sample = f"{cores} host cores x {reps} decodes of one 60 s 44.1 kHz stereo CBR-3 stream per step (config-4 stream shape)"
    print(json.dumps({
        "impl": "reference", "metric": "decode_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": workload_config(args, 0),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def r_again(prev, cores, reps):
    from oracle import sea_oracle as O

    if prev["kind"] == "reference":
        with tempfile.NamedTemporaryFile(suffix=".sea", delete=False) as f:
            f.write(prev["sea"])
            path = f.name
        try:
            secs, samples = O.ref_c_bench(path, cores, reps)
        finally:
            os.unlink(path)
    else:
        secs, samples = O.bench("decode", cores, reps, prev["pcm"], RATE, CHANNELS, prev["settings"], prev["sea"])
    out = dict(prev)
    out.update(secs=secs, samples=samples, value=samples / secs / 1e6)
    return out


def workload_config(args, unique):
    return {"workload": f"config4: batch decode of {args.streams} independent stereo {args.seconds} s 44.1 kHz CBR-3 .sea streams "
                        f"per GPU (chunk 5120, sf bits 4, sf frames 20); weak scaling, streams sharded by rank, no collectives",
            "streams_per_gpu": args.streams, "seconds": args.seconds, "sample_rate": RATE, "channels": CHANNELS,
            "unique_streams": unique, "l2": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------ our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # ~0.6 s timed region: enough nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="decode streams per GPU (config 4)")
    ap.add_argument("--enc-streams", type=int, default=1024, help="encode streams per GPU (config 5)")
    ap.add_argument("--e2e-streams", type=int, default=512, help="streams per step of the host-buffer (e2e) measurement")
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--unique", type=int, default=32, help="distinct synthetic streams generated per GPU (then replicated)")
    ap.add_argument("--skip-encode", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()

    from sea_codec_b200 import dist

    info = dist.init("gloo" if args.impl == "reference" else None)
    if args.impl == "reference":
        reference_arm(args, info)
        dist.shutdown()
        return

    import torch

    import sea_codec_b200 as S
    from sea_codec_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsea_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(info.local_rank)
    bind_to_gpu_numa_node(info.local_rank)
    dev = torch.device("cuda", info.local_rank)
    ctx = S.Context(info.local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    settings = S.EncoderSettings()  # CBR 3, chunk 5120, sf bits 4, sf frames 20
    frames = args.seconds * RATE
    unique = min(args.unique, args.streams)
    assert args.streams % unique == 0 and args.enc_streams % unique == 0 and args.e2e_streams % unique == 0

    # ---- build the inputs on the device: synthetic PCM -> (our encoder) -> .sea streams, replicated to the batch size
    first, _ = dist.weak_streams(unique, info.rank)
    pcm_u = synth.gen_batch_torch(unique, frames, CHANNELS, RATE, dev, first_stream=first)
    bound = ctx.encode_bound(frames, CHANNELS, settings)
    stride = (bound + 15) // 16 * 16
    sea_u = torch.zeros(unique * stride, dtype=torch.uint8, device=dev)
    lens_u = ctx.encode_batch_device(pcm_u.data_ptr(), np.arange(unique) * frames * CHANNELS, np.full(unique, frames), RATE, CHANNELS,
                                     settings, sea_u.data_ptr(), np.arange(unique) * stride)
    assert np.all(lens_u == bound)
    n = args.streams
    sea = sea_u.view(unique, stride).repeat(n // unique, 1).contiguous().view(-1)
    headers = np.tile(sea_u.view(unique, stride)[:, :22].cpu().numpy(), (n // unique, 1))
    sea_off = np.arange(n, dtype=np.uint64) * stride
    sea_len = np.full(n, bound, dtype=np.uint64)
    spp = frames * CHANNELS  # samples per stream
    pcm_out = torch.empty(n * spp, dtype=torch.int16, device=dev)
    pcm_off = np.arange(n, dtype=np.uint64) * spp
    samples_per_step = n * spp
    alg_bytes = float(n * bound + 2 * samples_per_step)

    kernel_ms = []

    def decode_step():
        got = ctx.decode_batch_device(sea.data_ptr(), sea_off, sea_len, headers, pcm_out.data_ptr(), pcm_off)
        kernel_ms.append(ctx.last_kernel_ms)
        return got

    for _ in range(args.warmup):
        decode_step()
    kernel_ms.clear()
    sampler = ClockSampler(info.local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        got = decode_step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dist.barrier()
    launches = ctx.launch_count - launches0
    ms_total = dist.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = info.world * samples_per_step / (ms_per_step * 1e-3) / 1e6
    assert np.all(got == spp)

    # ---- parity spot checks on the benchmarked buffers (replicas identical; one stream vs the CPU oracle below)
    view = pcm_out.view(n, spp)
    probe = [unique, n - 1] if n > unique else []
    for i in probe:
        assert torch.equal(view[i], view[i % unique]), "replicated streams decoded differently"
    dec0 = view[0].cpu().numpy()
    sea0 = sea_u[:bound].cpu().numpy().tobytes()

    # ---- roofline of the decode kernel (HBM): algorithmic bytes / average kernel duration
    peaks, peak_src = measured_peaks()
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "decode_unrolled_kernel<2,3,pair-repl> (+ decode_staged_kernel<2,0> for the partial last chunks, "
                                          "side stream)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(n, args.seconds), "peak_source": peak_src,
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_msamples_per_s": samples_per_step / (k_ms * 1e-3) / 1e6}

    # ---- the VBR twin of the headline (decode_vbr_kernel): same stream shape, VBR 3.0, a quarter of the streams
    vbr_dec = None
    if not args.skip_encode:
        st_v = S.EncoderSettings(residual_bits=3.0, vbr=True)
        nv = max(unique, n // 4)
        bound_v = ctx.encode_bound(frames, CHANNELS, st_v)
        stride_v = (bound_v + 15) // 16 * 16
        sea_vu = torch.zeros(unique * stride_v, dtype=torch.uint8, device=dev)
        lens_vu = ctx.encode_batch_device(pcm_u.data_ptr(), np.arange(unique) * frames * CHANNELS, np.full(unique, frames), RATE, CHANNELS,
                                          st_v, sea_vu.data_ptr(), np.arange(unique) * stride_v)
        sea_v = sea_vu.view(unique, stride_v).repeat(nv // unique, 1).contiguous().view(-1)
        hdr_v = np.tile(sea_vu.view(unique, stride_v)[:, :22].cpu().numpy(), (nv // unique, 1))
        len_v = np.tile(lens_vu, nv // unique)
        ks = []
        for _ in range(2 + max(3, min(args.steps, 10))):
            got_v = ctx.decode_batch_device(sea_v.data_ptr(), np.arange(nv, dtype=np.uint64) * stride_v, len_v, hdr_v, pcm_out.data_ptr(),
                                            pcm_off[:nv])
            ks.append(ctx.last_kernel_ms)
        assert np.all(got_v == spp)
        ms_v = dist.max_over_ranks(float(np.mean(ks[2:])))
        bytes_v = float(len_v.sum() + 2 * nv * spp)
        vbr_dec = {"value": info.world * nv * spp / (ms_v * 1e-3) / 1e6, "unit": "Msamples/s", "streams_per_gpu": nv, "ms_per_step": ms_v,
                   "roofline": {"bound": "hbm", "kernel": "decode_vbr_kernel<2>", "achieved": bytes_v / (ms_v * 1e-3) / 1e9,
                                "peak": measured_peaks()[0]["hbm_gbs"], "unit": "GB/s",
                                "frac": bytes_v / (ms_v * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"]}}
        del sea_v, sea_vu
        torch.cuda.empty_cache()

    # ---- BASELINE config 3 shape (8 channels, 48 kHz, CBR 4; multichannel lane mapping): 256 streams of 60 s through decode_mc_kernel
    mc_dec = None
    if not args.skip_encode:
        ch8, rate8, fr8, n8, u8 = 8, 48000, 60 * 48000, 256, 4
        st8 = S.EncoderSettings(residual_bits=4.0)
        pcm8 = synth.gen_batch_torch(u8, fr8, ch8, rate8, dev, first_stream=first)
        b8 = ctx.encode_bound(fr8, ch8, st8)
        s8 = (b8 + 15) // 16 * 16
        sea8u = torch.zeros(u8 * s8, dtype=torch.uint8, device=dev)
        ctx.encode_batch_device(pcm8.data_ptr(), np.arange(u8) * fr8 * ch8, np.full(u8, fr8), rate8, ch8, st8, sea8u.data_ptr(),
                                np.arange(u8) * s8)
        sea8 = sea8u.view(u8, s8).repeat(n8 // u8, 1).contiguous().view(-1)
        hdr8 = np.tile(sea8u.view(u8, s8)[:, :22].cpu().numpy(), (n8 // u8, 1))
        spp8 = fr8 * ch8
        ks = []
        for _ in range(2 + max(3, min(args.steps, 10))):
            got8 = ctx.decode_batch_device(sea8.data_ptr(), np.arange(n8, dtype=np.uint64) * s8, np.full(n8, b8, dtype=np.uint64), hdr8,
                                           pcm_out.data_ptr(), np.arange(n8, dtype=np.uint64) * spp8)
            ks.append(ctx.last_kernel_ms)
        assert np.all(got8 == spp8)
        ms8 = dist.max_over_ranks(float(np.mean(ks[2:])))
        bytes8 = float(n8 * b8 + 2 * n8 * spp8)
        mc_dec = {"value": info.world * n8 * spp8 / (ms8 * 1e-3) / 1e6, "unit": "Msamples/s", "streams_per_gpu": n8, "channels": ch8,
                  "seconds": 60, "ms_per_step": ms8,
                  "roofline": {"bound": "hbm", "kernel": "decode_mc_kernel<8,4>", "achieved": bytes8 / (ms8 * 1e-3) / 1e9,
                               "peak": measured_peaks()[0]["hbm_gbs"], "unit": "GB/s",
                               "frac": bytes8 / (ms8 * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"]}}
        del pcm8, sea8, sea8u
        torch.cuda.empty_cache()

    # ---- e2e: the same decode through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    ne = min(args.e2e_streams, n)
    while True:  # pinned host memory is a shared resource on a multi-GPU box: shrink the sample rather than fail
        try:
            h_sea = torch.empty(ne * stride, dtype=torch.uint8).pin_memory()
            h_pcm = torch.empty(ne * spp, dtype=torch.int16).pin_memory()
            break
        except RuntimeError:
            if ne <= unique:
                raise
            ne = max(unique, ne // 2 // unique * unique)
    h_sea.copy_(sea[: ne * stride])
    torch.cuda.synchronize()

    def e2e_step():
        return ctx.decode_batch_host(h_sea.data_ptr(), sea_off[:ne], sea_len[:ne], h_pcm.data_ptr(), pcm_off[:ne])

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0) / e2e_steps
    assert np.array_equal(h_pcm[:spp].numpy(), dec0), "host-buffer decode differs from the device-resident decode"
    e2e = {"value": info.world * ne * spp / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(ne * bound),
           "d2h_bytes_per_step": int(ne * spp * 2), "ms_per_step": e2e_s * 1e3,
           "sample": f"{ne} of the {n} streams per step (bounded pinned-host footprint); PCIe-bound"}
    del h_sea, h_pcm, pcm_out, view
    torch.cuda.empty_cache()

    # ---- encode (config 5 shape): CBR 3 and VBR 3.0, INT32 roofline
    encode = None
    if not args.skip_encode:
        ops_peak, _ = max((ctx.int32_peak(m) for m in (0, 1, 2)), key=lambda t: t[0])
        peaks_int = {m: ctx.int32_peak(m)[0] for m in (0, 1, 2)}
        ns = args.enc_streams
        pcm_e = pcm_u.repeat(ns // unique, 1).contiguous().view(-1)
        out_e = torch.zeros(ns * ((max(bound, ctx.encode_bound(frames, CHANNELS, S.EncoderSettings(residual_bits=3.0, vbr=True))) + 15) // 16 * 16),
                            dtype=torch.uint8, device=dev)
        estride = out_e.numel() // ns
        encode = {"streams_per_gpu": ns, "int32_peak_lane_ops_per_s": {"imad": peaks_int[0], "lop3": peaks_int[1], "mixed": peaks_int[2]}}
        for name, st_e, passes in (("cbr3", settings, 1), ("vbr3", S.EncoderSettings(residual_bits=3.0, vbr=True), 2)):
            def enc_step():
                return ctx.encode_batch_device(pcm_e.data_ptr(), np.arange(ns) * spp, np.full(ns, frames), RATE, CHANNELS, st_e,
                                               out_e.data_ptr(), np.arange(ns) * estride)
            enc_step()
            dist.barrier()
            torch.cuda.synchronize()
            ks = []
            steps_e = max(1, min(args.steps, 3))
            for _ in range(steps_e):
                elens = enc_step()
                ks.append(ctx.last_kernel_ms)
            ms_e = dist.max_over_ranks(float(np.mean(ks)))
            sps = info.world * ns * spp / (ms_e * 1e-3)
            ops = ALG_OPS_PER_CAND_SAMPLE * 16 * passes  # 49 * 2^sf_bits * passes
            encode[name] = {"value": sps / 1e6, "unit": "Msamples/s", "ms_per_step": ms_e, "bytes_per_stream": int(elens[0]),
                            "vbr_ties": ctx.last_vbr_ties,
                            "roofline": {"bound": "int32", "achieved": sps / info.world * ops / 1e12, "peak": ops_peak / 1e12,
                                         "unit": "Tops/s", "frac": sps / info.world * ops / ops_peak, "ops_per_sample": ops}}
        del pcm_e, out_e
        torch.cuda.empty_cache()
        # the same kernel with the machine full: 4096 streams of 10 s (one warp per stream: 1024 streams leave the SMs latency-bound)
        nl, fl_ = 4 * ns, 10 * RATE
        pcm_l = pcm_u.view(unique, frames, CHANNELS)[:, :fl_, :].repeat(nl // unique, 1, 1).contiguous().view(-1)
        bl = ctx.encode_bound(fl_, CHANNELS, settings)
        sl = (bl + 15) // 16 * 16
        out_l = torch.zeros(nl * sl, dtype=torch.uint8, device=dev)
        ks = []
        for _ in range(3):
            ctx.encode_batch_device(pcm_l.data_ptr(), np.arange(nl) * fl_ * CHANNELS, np.full(nl, fl_), RATE, CHANNELS, settings,
                                    out_l.data_ptr(), np.arange(nl) * sl)
            ks.append(ctx.last_kernel_ms)
        ms_l = dist.max_over_ranks(float(np.mean(ks[1:])))
        sps_l = info.world * nl * fl_ * CHANNELS / (ms_l * 1e-3)
        encode["cbr3_4096_streams"] = {"value": sps_l / 1e6, "unit": "Msamples/s", "ms_per_step": ms_l, "streams_per_gpu": nl, "seconds": 10,
                                       "roofline": {"bound": "int32", "achieved": sps_l / info.world * 784 / 1e12, "peak": ops_peak / 1e12,
                                                    "unit": "Tops/s", "frac": sps_l / info.world * 784 / ops_peak, "ops_per_sample": 784}}
        del pcm_l, out_l
        torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): bounded sample of the same stream shape
    cpu = None
    if info.rank == 0 and info.world == 1 and not args.skip_cpu:
        from oracle import sea_oracle as O

        cores = host_cores()
        reps = 4
        r = cpu_decode_baseline(cores, reps, args.seconds)
        if args.seconds == SECONDS:  # the oracle checks stream 0 of the benchmarked batch, bit for bit
            assert r["sea"] == sea0, "GPU-encoded stream 0 differs from the oracle's encode"
        assert np.array_equal(O.sea_decode(sea0).samples, dec0), "GPU-decoded stream 0 differs from the oracle's decode"
        e_secs, e_samples = O.bench("encode", cores, 1, r["pcm"], RATE, CHANNELS, r["settings"])
        cpu = {"value": r["value"], "unit": "Msamples/s", "cores": cores, "kind": r["kind"],
               "sample": f"{cores} cores x {reps} decodes of one {args.seconds} s stereo CBR-3 stream (config-4 stream shape)",
               "encode_cbr3_msamples_per_s": e_samples / e_secs / 1e6, "encode_kind": "port",
               "encode_sample": f"{cores} cores x 1 encode of the same stream (oracle restatement; no Rust toolchain)",
               "parity_checked": "stream 0 of the benchmarked batch: GPU encode == oracle encode, GPU decode == oracle decode"}

    if info.rank == 0:
        print(json.dumps({
            "metric": "decode_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": info.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32", "data": "synthetic", "config": workload_config(args, unique), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "decode_vbr3": vbr_dec, "decode_8ch_cbr4": mc_dec, "encode": encode,
        }))
    ctx.close()
    dist.shutdown()


def _main_with_clean_stdout():
    """stdout carries the one JSON line and nothing else: libraries that chat on fd 1 (NCCL prints its version banner there at
    the first collective) are sent to stderr for the duration of the run."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = sys.stdout
    sys.stdout = os.fdopen(real, "w", buffering=1)
    try:
        main()
    finally:
        sys.stdout.flush()
        sys.stdout = out


if __name__ == "__main__":
    _main_with_clean_stdout()
