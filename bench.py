#!/usr/bin/env python
"""Benchmark of the one hot path: FullSubNet+ inference forward on batches of 3 s / 16 kHz synthetic clips.

    python bench.py --gpus N --steps K --warmup W            # product arm (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's CPU PyTorch path

Metric (BASELINE.json): frames/sec (and real-time factor) on 16 kHz 3 s clips.  A "step" = one forward of the
model over one batch of 64 clips per GPU (BASELINE configs[1]; weak scaling: 64 clips per rank).
  value     model forward with the STFT inputs already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C ABI's HOST-buffer entry point: pinned host inputs -> H2D -> forward ->
            D2H of the mask, every step
  pipeline  (extra) waveform on device -> torch.stft -> model -> decompress -> torch.istft -> one NCCL
            all-gather of the enhanced waveforms
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fullsubnet-plus_b200")]

FRAMES_PER_CLIP = 188            # 48000 samples, n_fft 512, hop 256, center=True
CLIP_SECONDS = 3.0
SR, NSAMP = 16000, 48000


def default_cfg():
    """config/inference.toml:30-44 of the reference."""
    return dict(sb_num_neighbors=15, fb_num_neighbors=0, num_freqs=257, look_ahead=2, sequence_model="LSTM",
                fb_output_activate_function="ReLU", sb_output_activate_function=False, channel_attention_model="TSSE",
                fb_model_hidden_size=512, sb_model_hidden_size=384, weight_init=False,
                norm_type="offline_laplace_norm", num_groups_in_drop_band=2, kersize=[3, 5, 10], subband_num=1)


def flops_per_clip(cfg, T_in=FRAMES_PER_CLIP, L=2):
    """Algorithmic FLOPs.  Returns (sub-band LSTM + Linear, total)."""
    F, Tp, H = cfg["num_freqs"], T_in + cfg["look_ahead"], cfg["sb_model_hidden_size"]
    I = (2 * cfg["sb_num_neighbors"] + 1) + 3 * (2 * cfg["fb_num_neighbors"] + 1)
    sb = Tp * F * (2 * 4 * H * (I + H) + (L - 1) * 2 * 4 * H * 2 * H + 2 * H * 2)
    tcn = 3 * Tp * (8 * (2 * 2 * F * 512 + 2 * 3 * 512) + 2 * F * F)
    ts = 3 * (2 * F * sum(cfg["kersize"]) * Tp + 4 * F * (F // 2))
    return sb, sb + tcn + ts


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


class CpuReference:
    """CPU path of the reference (torch port issuing the same ATen ops, oracle/torch_port.py), one clip per call
    like the reference inferencer.  The thread count is the best of an ascending sweep up to all host cores (torch's
    default of one thread per core is ~80x slower than 16 threads on the 128-core GPU hosts for these small GEMMs;
    the baseline is reported at its best setting, not its default)."""

    def __init__(self, state, cfg, spec, threads=None):
        import torch
        from oracle.torch_port import TorchPort
        self.torch = torch
        self.port = TorchPort({k: v.detach().cpu().numpy() for k, v in state.items()}, cfg, "plus")
        self.sweep = {}
        if threads is None:
            ncpu = os.cpu_count() or 1
            cands = sorted({c for c in (1, 4, 8, 16, 32, 64, ncpu) if c <= ncpu})   # 1 thread: the scalar figure SURVEY.md 8d asks for
            best, threads = None, cands[0]
            for c in cands:
                torch.set_num_threads(c)
                self.run(spec, 1)                                  # warm-up at this setting
                t = min(self.run(spec, 2))
                self.sweep[c] = round(t, 3)
                if best is None or t < best:
                    best, threads = t, c
                elif t > 1.8 * best:
                    break                                          # past the knee: more threads only get slower
        torch.set_num_threads(threads)
        self.threads = torch.get_num_threads()

    def run(self, spec, n_clips, start=0):
        """Forward n_clips clips (cycling through spec); returns the list of per-clip seconds."""
        mag, real, imag = spec
        times = []
        for i in range(n_clips):
            j = (start + i) % mag.shape[0]
            t0 = time.perf_counter()
            self.port.forward(mag[j:j + 1], real[j:j + 1], imag[j:j + 1])
            times.append(time.perf_counter() - t0)
        return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--lstm-impl", default="auto", choices=["auto", "mma", "tcgen05"])
    ap.add_argument("--accurate-math", action="store_true", help="ex2/rcp gate math instead of the default tanh.approx path")
    ap.add_argument("--ref-clips", type=int, default=4, help="reference arm: clips per step (bounded sample)")
    ap.add_argument("--cpu-baseline-clips", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = default_cfg()
    B, K, W = args.batch, args.steps, max(args.warmup, 3 if args.impl == "b200" else args.warmup)
    workload = f"FullSubNet+ default config/inference.toml, batch={B} synthetic 3 s 16 kHz clips per GPU (BASELINE configs[1])"

    from fsnplus_b200.model import FullSubNet_Plus
    from fsnplus_b200.synth import synth_clips
    from fsnplus_b200 import inference as inf

    torch.manual_seed(0)
    model = FullSubNet_Plus(**cfg, lstm_impl=args.lstm_impl, fast_math=not args.accurate_math).eval()    # random init, seed 0
    state = model.state_dict()

    # ------------------------------------------------------------------ reference arm (CPU) -----------------
    if args.impl == "reference":
        if rank != 0:
            return
        clips = synth_clips(args.ref_clips, NSAMP, SR, seed=1000)
        X = inf.stft(clips)
        spec = (X.abs().unsqueeze(1), X.real.unsqueeze(1).contiguous(), X.imag.unsqueeze(1).contiguous())
        ref = CpuReference(state, cfg, spec)
        th = ref.threads
        step_times = []
        for s in range(W + K):
            t0 = time.perf_counter()
            ref.run(spec, args.ref_clips)
            if s >= W:
                step_times.append(time.perf_counter() - t0)
        per_step = statistics.median(step_times)
        fps = args.ref_clips * FRAMES_PER_CLIP / per_step
        line = {
            "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf": per_step / (args.ref_clips * CLIP_SECONDS),
            "config": {"workload": workload, "sample": f"{args.ref_clips} clips per step, one clip per call (the reference "
                       "inference batch size), model forward only"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": th, "kind": "port",
                             "sample": f"{args.ref_clips} clips/step x {K} steps, torch {torch.__version__} CPU, B=1 per call",
                             "host_cores": os.cpu_count(), "thread_sweep_s_per_clip": ref.sweep},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ product arm ---------------------------
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = model.to(dev)

    # inputs: NSETS distinct batches rotated so consecutive steps never reuse L2-resident inputs
    NSETS = 4
    clips = synth_clips(NSETS * B, NSAMP, SR, seed=1000 + 7919 * rank).to(dev)
    X = inf.stft(clips)
    mags = X.abs().unsqueeze(1).contiguous().view(NSETS, B, 1, 257, FRAMES_PER_CLIP)
    reals = X.real.unsqueeze(1).contiguous().view(NSETS, B, 1, 257, FRAMES_PER_CLIP)
    imags = X.imag.unsqueeze(1).contiguous().view(NSETS, B, 1, 257, FRAMES_PER_CLIP)
    in_bytes = 3 * B * 257 * FRAMES_PER_CLIP * 4
    out_bytes = 2 * B * 257 * FRAMES_PER_CLIP * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        with torch.no_grad():
            for i in range(warm):
                fn(i)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            e0.record()
            for i in range(steps):
                fn(warm + i)
            e1.record()
            barrier()
            t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, t0, t1

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms_step, t0, t1 = timed(lambda i: model(mags[i % NSETS], reals[i % NSETS], imags[i % NSETS]), K, W)
    clocks = sampler.stop(t0, t1) if sampler else None
    lstm_ms = model.lstm_ms_history(min(K, 32))
    launches = model.last_launch_count() * K
    lstm_impl = model.last_lstm_impl()
    fps = world * B * FRAMES_PER_CLIP / (ms_step * 1e-3)

    # e2e: host buffers through the C ABI
    pin = lambda x: x.cpu().pin_memory()
    hm, hr, hi = [pin(mags[i]) for i in range(NSETS)], [pin(reals[i]) for i in range(NSETS)], [pin(imags[i]) for i in range(NSETS)]
    houts = [torch.empty((B, 2, 257, FRAMES_PER_CLIP), dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_step(i):
        model.forward_host(hm[i % NSETS], hr[i % NSETS], hi[i % NSETS], out=houts[i % 2], device=dev, pipelined=True)
    # timed(): warm-up, then K pipelined host-buffer forwards; the closing barrier() only synchronises torch's streams, so
    # the copy streams are drained explicitly INSIDE the timed region (sync_host before the closing event).
    def e2e_timed():
        with torch.no_grad():
            for i in range(2):
                e2e_step(i)
            model.sync_host(); barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                e2e_step(2 + i)
            model.sync_host()                      # every D2H of the K steps has landed in pinned host memory
            e1.record()
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / K
    ms_e2e = e2e_timed()
    fps_e2e = world * B * FRAMES_PER_CLIP / (ms_e2e * 1e-3)

    # pipeline: waveform -> STFT -> model -> decompress -> iSTFT (+ all-gather of enhanced waveforms)
    wav = clips.view(NSETS, B, NSAMP)

    def pipe(i):
        enh = inf.enhance_batch(model, wav[i % NSETS])
        return inf.all_gather_enhanced(enh, world * B)
    ms_pipe, _, _ = timed(pipe, K, 2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (sub-band LSTM): tensor-bound
    sb_flops, tot_flops = flops_per_clip(cfg)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_burst, peak_src = pk["bf16_tflops_sustained"], pk["bf16_tflops"], "measured (MEASURED_PEAKS.json, sustained)"
    else:
        peak, peak_burst, peak_src = 1400.0, 1590.0, "fallback (B200_PROFILING.md)"
    traffic = None                                    # dram read+write bytes per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "lstm_traffic.json")
    if os.path.exists(tpath) and lstm_impl == "tcgen05" and B == 64:
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    k_ms = statistics.mean([x for x in lstm_ms if x > 0]) if lstm_ms else float("nan")
    achieved = B * sb_flops / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": f"sub-band LSTM ({lstm_impl})", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "frac_of_burst_peak": achieved / peak_burst, "peak_source": peak_src,
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step, "traffic": traffic,
                "algorithmic_flops_per_launch": B * sb_flops}

    line = {
        "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "rtf": ms_step * 1e-3 / (B * CLIP_SECONDS),
        "config": {"workload": workload, "clips_per_gpu": B, "frames_per_clip": FRAMES_PER_CLIP, "lstm_impl": lstm_impl,
                   "gate_math": "ex2+rcp" if args.accurate_math else "tanh.approx (default; parity identical to 3 digits, profiles/r01_fast_math_accuracy.txt)", "weights": "random init (torch default, seed 0)",
                   "l2": f"inputs rotated over {NSETS} batches ({NSETS * in_bytes / 1e6:.0f} MB > L2); per-step intermediates "
                         "(392 MB of LSTM input tiles) exceed L2"},
        "model_tflops": world * B * tot_flops / (ms_step * 1e-3) / 1e12,
        "roofline": roofline,
        "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_e2e, "rtf": ms_e2e * 1e-3 / (B * CLIP_SECONDS),
                "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "path": "fsn_model_forward_host_async (C ABI, pinned host buffers; H2D / forward / D2H of consecutive steps overlap, all drained inside the timed region)"},
        "pipeline": {"value": world * B * FRAMES_PER_CLIP / (ms_pipe * 1e-3), "unit": "frames/s", "ms_per_step": ms_pipe,
                     "path": "wave(device) -> torch.stft -> model -> decompress_cIRM -> torch.istft"
                             + (" -> NCCL all_gather_into_tensor of enhanced waveforms" if world > 1 else "")},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        n = args.cpu_baseline_clips
        spec = (mags[0][:n].cpu(), reals[0][:n].cpu(), imags[0][:n].cpu())
        ref = CpuReference(state, cfg, spec)
        th = ref.threads
        ref.run(spec, 1)
        times = ref.run(spec, n)
        cfps = FRAMES_PER_CLIP / statistics.median(times)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": th, "kind": "port",
                                "rtf": statistics.median(times) / CLIP_SECONDS,
                                "sample": f"{n} of the {B} clips, one clip per call (reference inference batch size), model forward "
                                          f"only, torch {torch.__version__} CPU fp32, median of {n} after 1 warm-up",
                                "host_cores": os.cpu_count(), "thread_sweep_s_per_clip": ref.sweep}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
