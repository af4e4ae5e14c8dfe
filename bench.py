#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: utterances/s (2 s @ 16 kHz) embedding extraction.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path (fbank -> CMN -> model forward) over one batch of synthetic utterances
per GPU.  Default workload = BASELINE.json configs[1]: ECAPA-TDNN-1024, bf16 tensor-core path, batch 256 of
2.02 s utterances (32320 samples -> 200 frames, the "256 x 200-frame" case).  `value` is measured with the
waveforms already resident in HBM; `e2e` is the same metric through the public host-buffer API
(B200SpeakerModel.extract_from_wav on pinned host int16 PCM: H2D + kernels + D2H inside the timed region).
Multi-GPU: one rank per GPU, utterances sharded with no data-path collective (weak scaling), ONE NCCL all-gather
of the embeddings at the end of the job, inside the timed region.  `--impl reference` times the oracle port of
the reference's CPU PyTorch path (torch fp32 on all host cores) on the same config, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, precision, batch per GPU, samples per utterance, GFLOP/utt of conv+linear layers at that T)
    "ecapa1024_bf16_b256": ("ECAPA_TDNN_c1024", "bf16", 256, 32320, 5.141),
    "ecapa512_bf16_b256": ("ECAPA_TDNN_c512", "bf16", 256, 32320, 1.917),
    "ecapa1024_bf16_b128": ("ECAPA_TDNN_c1024", "bf16", 128, 32320, 5.141),
    "ecapa1024_bf16_b64": ("ECAPA_TDNN_c1024", "bf16", 64, 32320, 5.141),
    "ecapa1024_bf16_b512": ("ECAPA_TDNN_c1024", "bf16", 512, 32320, 5.141),
    "ecapa512_fp32_b16": ("ECAPA_TDNN_c512", "fp32", 16, 32000, 1.898),
    "resnet34_fp16_b64": ("ResNet34", "fp16", 64, 32320, 9.056),
    "campplus_bf16_b64": ("CAMPPlus", "bf16", 64, 32320, 2.252),
    "ecapa512_tf32x3_b256": ("ECAPA_TDNN_c512", "tf32x3", 256, 32320, 1.917),
    "ecapa512_fp32_b256": ("ECAPA_TDNN_c512", "fp32", 256, 32320, 1.917),
    "ecapa1024_tf32x3_b256": ("ECAPA_TDNN_c1024", "tf32x3", 256, 32320, 5.141),
}
DEFAULT_WORKLOAD = "ecapa1024_bf16_b256"
METRIC = "utterances/s (2s@16kHz) embedding extraction"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML polled every 5 ms from a
    thread (the timed region is only tens of ms long); falls back to `nvidia-smi -lms 100` if NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.reasons, self.maxclk, self._stop = None, [], set(), None, False

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                bus = torch.cuda.get_device_properties(self.index).pci_bus_id
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(f"00000000:{bus:02x}:00.0".encode()) if isinstance(bus, int) \
                    else pynvml.nvmlDeviceGetHandleByIndex(self.index)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                    "sm_min_mhz": float(min(self.samples)) if self.samples else None,
                    "sm_max_mhz": float(self.maxclk) if self.maxclk else None, "reasons": sorted(self.reasons),
                    "samples": len(self.samples), "source": "nvml, 5 ms polling during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def _cpu_path(model_name, nsamples, batch):
    """Oracle port of the reference CPU path: numpy fbank+CMN, torch-CPU fp32 forward."""
    from oracle import fbank_np, models_torch
    from wespeaker_b200 import synthetic as syn
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in syn.make_state_dict(model_name, 0).items()}
    wavs = syn.make_wavs(batch, nsamples, seed=0)

    def one():
        feats = np.stack([fbank_np.cmn(fbank_np.fbank(w)) for w in wavs])
        return models_torch.forward(model_name, sd, torch.from_numpy(feats))
    return one


def _best_threads(one):
    """torch's CPU convs collapse when oversubscribed (128 threads: ~1 utt/s on the 128-core bench host), so give the
    reference arm the thread count at which it is fastest."""
    cores = os.cpu_count() or 1
    best, best_t = None, None
    for nt in sorted({t for t in (8, 16, 32, 64, cores) if t <= cores}):
        torch.set_num_threads(nt)
        one()
        t0 = time.perf_counter(); one(); dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
        if dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best, cores


def cpu_reference_run(model_name, nsamples, batch, budget_s, max_batches):
    one = _cpu_path(model_name, nsamples, batch)
    threads, cores = _best_threads(one)
    t0, nb = time.perf_counter(), 0
    while nb < max_batches and (time.perf_counter() - t0 < budget_s or nb == 0):
        one(); nb += 1
    dt = time.perf_counter() - t0
    return nb * batch / dt, threads, cores, nb, dt


def run_reference(args, wl):
    model, _, _, nsamples, _ = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    batch = 16  # each "step" = one bounded sample (16 utterances) of the same workload
    one = _cpu_path(model, nsamples, batch)
    threads, cores = _best_threads(one)
    for _ in range(max(1, min(args.warmup, 3))):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    v = args.steps * batch / dt
    sample = (f"{args.steps} steps x {batch} utts of {nsamples} samples, oracle port of the reference CPU path (numpy fbank+CMN, "
              f"torch-CPU fp32 forward), {threads} torch threads (fastest of 8..{cores} on this {cores}-core host)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "model": model, "batch_per_step": batch, "samples_per_utt": nsamples},
        "cpu_baseline": {"value": v, "unit": "utt/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def plda_bench(dev, n_enroll=32768, n_test=100000, dim=256, iters=3):
    """Secondary metric of BASELINE.json: PLDA scores/s (configs[4]: 1M x 100k, D=256).  The all-pairs job is tiled over
    enroll rows into a reused fp32 score buffer (10^11 scores do not fit in HBM); one tile = n_enroll x n_test."""
    from wespeaker_b200 import synthetic as syn
    from wespeaker_b200.plda import TwoCovPLDA
    from oracle import plda_np
    pm = syn.make_plda(dim, seed=3, normalize_length=True)
    p = TwoCovPLDA.from_arrays(**pm, device=dev.index)
    e = torch.from_numpy(syn.make_embeddings(n_enroll, dim, seed=3)).to(dev)
    t = torch.from_numpy(syn.make_embeddings(n_test, dim, seed=4)).to(dev)
    e_t, t_t = p.transform_batch(e), p.transform_batch(t)
    out = torch.empty((n_enroll, n_test), dtype=torch.float32, device=dev)
    p.score_matrix(e_t, t_t, 1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        p.score_matrix(e_t, t_t, 1, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    scores = float(n_enroll) * n_test
    # CPU: the reference's per-trial Python loop (two_cov_plda.py:248-256), single thread, on a 20k-trial sample
    et, tt = e_t[:200].cpu().numpy(), t_t[:100].cpu().numpy()
    t0 = time.perf_counter()
    for i in range(200):
        for j in range(100):
            plda_np.log_likelihood_ratio(pm, et[i], tt[j], 1)
    cpu_loop = 20000 / (time.perf_counter() - t0)
    ref = plda_np.llr_matrix(pm, et, tt, 1)
    err = float(np.abs(out[:200, :100].double().cpu().numpy() - ref).max())
    return {"value": scores / (ms * 1e-3), "unit": "scores/s", "dtype": "f64", "enroll_tile": n_enroll, "test": n_test, "dim": dim,
            "ms_per_tile": ms, "tflops_f64": scores * 2 * dim / (ms * 1e-3) / 1e12, "out_gbs": scores * 4 / (ms * 1e-3) / 1e9,
            "max_abs_err_vs_fp64_oracle": err,
            "cpu_baseline": {"value": cpu_loop, "unit": "scores/s", "cores": 1, "kind": "port",
                             "sample": "20000 trials, per-trial log_likelihood_ratio loop as in eval_sv"}}


def time_dominant_kernel(model, prec, B, T, iters=10, tc_version=3):
    """Roofline leg: the dominant launch of the step — ECAPA's 1x1 conv 3C->1536 over B*T positions
    (ecapa_tdnn.py:200,218; 47-49% of the model's MACs) — timed live with CUDA events on the launching stream
    through ws_conv.  Algorithmic FLOPs per launch = 2 * B*T * 3C * 1536."""
    import ctypes as C
    from wespeaker_b200 import lib
    if not model.startswith("ECAPA") or prec not in ("fp32", "tf32", "bf16", "fp16"):
        return None
    Cc = 1024 if "1024" in model else 512
    cin, cout = 3 * Cc, 1536
    code, tdt = {"fp32": (0, torch.float32), "tf32": (0, torch.float32), "bf16": (1, torch.bfloat16),
                 "fp16": (2, torch.float16)}[prec]
    dev = torch.device("cuda", torch.cuda.current_device())
    # rotate over enough (x, out) pairs that every launch reads operands not left in L2 by the previous one
    per = B * T * (cin + cout) * (2 if code else 4)
    nbuf = max(2, int(2 * 126e6 // per) + 1)
    xs = [torch.randn(B, 1, T, cin, device=dev).to(tdt) for _ in range(nbuf)]
    outs = [torch.empty(B, 1, T, cout, device=dev, dtype=tdt) for _ in range(nbuf)]
    w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).to(tdt)
    bias = torch.zeros(cout, device=dev)
    L = lib.load()
    descs = []
    for x, o in zip(xs, outs):
        d = lib.ConvDesc()
        d.x, d.B, d.F, d.T, d.Cin, d.x_ld = x.data_ptr(), B, 1, T, cin, cin
        d.w, d.Cout, d.kf, d.kt = w.data_ptr(), cout, 1, 1
        d.dil_f = d.dil_t = d.stride_f = d.stride_t = 1
        d.bias, d.act1, d.out, d.out_ld, d.dtype, d.use_tc = bias.data_ptr(), 1, o.data_ptr(), cout, code, (0 if prec == "fp32" else tc_version)
        descs.append(d)
    st = lib.cur_stream_ptr()
    for i in range(3):
        lib.check(L.ws_conv(C.byref(descs[i % nbuf]), st), "ws_conv")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        lib.check(L.ws_conv(C.byref(descs[i % nbuf]), st), "ws_conv")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * B * T * cin * cout
    return dict(ms=ms, tflops=flops / (ms * 1e-3) / 1e12, flops=flops, nbuf=nbuf,
                kernel=f"ws_conv_gemm_tc{tc_version}_kernel 1x1 {cin}->{cout} over {B * T} positions")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plda", action="store_true")
    ap.add_argument("--tc-version", type=int, default=0, help="override the engine's tcgen05 kernel generation (1,2,3)")
    args = ap.parse_args()
    wl = args.workload
    if args.impl == "reference":
        run_reference(args, wl)
        return
    model_name, prec, B, nsamples, gflop_utt = WORKLOADS[wl]
    torch.set_num_threads(2)  # the GPU arm needs no host parallelism; idle-spinning worker pools only add scheduler noise

    from wespeaker_b200 import parallel
    from wespeaker_b200.models import from_synthetic
    from wespeaker_b200 import synthetic as syn
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    peaks = measured_peaks()

    model = from_synthetic(model_name, 0, precision=prec).to(dev)
    if args.tc_version:
        model.set_option("tc_version", args.tc_version)
    L_frames = 1 + (nsamples - 400) // 160
    # rotate over distinct input batches; the per-step working set (activations) is >> the 126 MB L2 anyway
    nrot = 4
    base = syn.make_wavs(B, nsamples, seed=100 + rank)
    wav_dev = [torch.from_numpy(np.roll(base, i, axis=0)).to(dev) for i in range(nrot)]
    wav_pin = [torch.from_numpy(np.roll(base, i, axis=0).astype(np.int16)).pin_memory() for i in range(nrot)]
    emb = None
    for i in range(max(3, args.warmup)):
        emb = model.extract_from_wav(wav_dev[i % nrot])
    if world > 1:  # warm the communicator: the first NCCL collective pays lazy init
        parallel.gather_embeddings(torch.cat([emb, emb], 0), 2 * B * world)
    torch.cuda.synchronize()
    launches_per_step = model.last_launches()

    # ---------------- device-resident throughput (`value`)
    sampler = ClockSampler(local)
    parallel.barrier(); torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    local_embs = []
    for i in range(args.steps):
        local_embs.append(model.extract_from_wav(wav_dev[i % nrot]))
    allemb = torch.cat(local_embs, 0)
    gathered = parallel.gather_embeddings(allemb, allemb.shape[0] * world)  # the one NCCL all-gather of the job
    e1.record()
    torch.cuda.synchronize(); parallel.barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    assert gathered.shape[0] == args.steps * B * world and torch.isfinite(gathered).all()
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---------------- end-to-end through the public host-buffer API (`e2e`)
    # B200SpeakerModel.extract_stream: every step copies that step's pinned int16 PCM batch H2D, runs fbank+CMN+forward
    # and copies the embeddings back D2H; the copy of batch i+1 overlaps the kernels of batch i (2 staging slots).
    # Warm-up stream first (allocations, pinned pools), then time K complete steps: the clock starts before the first
    # timed batch is submitted (its H2D copy is NOT hidden) and stops when the K-th batch's embeddings are in host memory.
    for out_h in model.extract_stream(wav_pin[i % nrot] for i in range(max(4, args.warmup))):
        pass
    best_dt = None
    for rep in range(3):   # best of 3: the host is a shared 128-core box, a descheduled host thread stalls collect()
        parallel.barrier(); torch.cuda.synchronize()
        t0, nout = time.perf_counter(), 0
        for out_h in model.extract_stream(wav_pin[i % nrot] for i in range(args.steps)):
            nout += out_h.shape[0]
        dt_rep = parallel.max_over_ranks(time.perf_counter() - t0, dev)
        best_dt = dt_rep if best_dt is None else min(best_dt, dt_rep)
    dt = best_dt
    torch.cuda.synchronize()
    parallel.barrier()
    e2e_value = world * B * args.steps / dt
    assert nout == B * args.steps and torch.isfinite(out_h).all()
    assert torch.equal(out_h.to(dev), model.extract_from_wav(wav_dev[(args.steps - 1) % nrot]))

    if rank != 0:
        return
    # ---------------- roofline for the dominant kernel + whole-step tensor utilisation
    dom = time_dominant_kernel(model_name, prec, B, L_frames, tc_version=args.tc_version or 3)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath) and wl == DEFAULT_WORKLOAD:   # DRAM bytes per launch of this kernel from the committed ncu capture
        traffic = json.load(open(tpath))["traffic_bytes_per_launch"]
    step_tf = value / world * gflop_utt / 1e3  # TFLOP/s per GPU, algorithmic
    if dom is not None:
        roof = {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": dom["tflops"] / peaks["tf_burst"], "traffic": traffic, "kernel": dom["kernel"],
                "kernel_ms": dom["ms"], "flops_per_launch": dom["flops"], "peak_source": peaks["src"] + " (burst, kernel timed alone)",
                "step_tflops_per_gpu": step_tf, "step_frac_of_sustained": step_tf / peaks["tf_sustained"]}
    else:
        roof = {"bound": "tensor", "achieved": step_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": step_tf / peaks["tf_sustained"], "traffic": None, "kernel": "whole step (conv GEMMs)",
                "peak_source": peaks["src"] + " (sustained, whole step)"}
    cpu = None
    if not args.no_cpu_baseline:
        v, threads, cores, nb, cdt = cpu_reference_run(model_name, nsamples, 16, budget_s=10.0, max_batches=200)
        cpu = {"value": v, "unit": "utt/s", "cores": threads, "kind": "port",
               "sample": f"{nb} batches x 16 utts of {nsamples} samples in {cdt:.1f}s: numpy fbank+CMN + torch-CPU fp32 forward "
                         f"(oracle port of the reference path), {threads} torch threads = fastest of 8..{cores} on this {cores}-core host"}
    plda = None
    if not args.no_plda:
        plda = plda_bench(dev)
    act_mb = B * L_frames * (1536 * 2 + 128 + (1024 if '1024' in model_name else 512) * 7) * (4 if prec in ("fp32", "tf32", "tf32x3") else 2) / 1e6
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3", "bf16": "bf16", "fp16": "f16"}[prec], "data": "synthetic",
        "config": {"workload": wl, "model": model_name, "batch_per_gpu": B, "samples_per_utt": nsamples,
                   "frames": L_frames, "precision": prec, "gflop_per_utt": gflop_utt,
                   "l2": f"inputs rotate over {nrot} batches; per-step activation working set ~{act_mb:.0f} MB > 126 MB L2",
                   "collective": "one all_gather_into_tensor of embeddings at job end (inside timed region)" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": B * nsamples * 2, "d2h_bytes_per_step": B * model.embed_dim * 4,
                "api": "B200SpeakerModel.extract_stream(pinned int16 PCM host batches): H2D + fbank + CMN + forward + D2H per step, copy/compute overlapped over 4 slots; best of 3 runs of K steps"},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "plda": plda,
    }))


if __name__ == "__main__":
    main()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
