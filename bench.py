#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the VAENAR-TTS mel-synthesis hot path on B200.

Headline workload (BASELINE.json configs[1], "C2"): LJSpeech hparams, batch 16 per GPU, T_text 148, T_mel 870,
inference only (text encoder -> prior flow sample -> decoder), reduction factor 2, synthetic batch and
random-init weights.  One "step" = one VAENAR.inference call over one batch of 16 utterances.

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                      # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0):
  value        frames/s over exactly K steps with inputs resident in HBM (CUDA-graph replay, CUDA events, L2 flushed
               before the timed region, inputs + weights + workspaces > L2).  `--inflight` independent batches (default
               4, each with its own graph / workspace / stream) are in flight at a time -- the serving pipeline
               H2D / compute / D2H of consecutive batches; one inference occupies 64 of the 148 SMs for most of its launch
               chain, so the batches also overlap on the device (measured: 2 -> 1.23, 3 -> 1.05, 4 -> 0.985, 5 -> 0.993,
               6 -> 1.007, 8 -> 1.024 ms per step).  `serial` carries the strictly
               one-step-after-the-other number (latency per step).
  e2e          the same K steps through the public API with pinned-host inputs and pinned-host results inside the timed
               region (H2D + graph + D2H per step), same number of batches in flight; `e2e_with_alignments` also copies
               the decoder alignments VAENAR.inference returns.
  roofline     dominant kernel class timed per launch with CUDA events (exclusive: one launch chain, one stream).
  train        K_train full train steps (BASELINE.json configs[2] at N = 1, the per-GPU share of configs[3] with the
               fused peer-memory gradient exchange + Adam at N > 1), timed the same way.
  cpu_baseline the oracle on the host cores (bounded sample), the only use of oracle/ besides the input generator.
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
sys.path.insert(0, ROOT)

B_PER_GPU, T_TEXT, T_MEL, RF = 16, 148, 870, 2
WORKLOAD = "C2: LJSpeech hparams, batch=16/GPU, T_text=148, T_mel=870, inference (encoder + prior sample + decoder), rf=2"
CONFIG = {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "T_text": T_TEXT, "T_mel": T_MEL, "rf": RF}
FLOPS_PER_FRAME = 27.53e6   # SURVEY.md §8d algorithmic FLOPs per mel frame at C2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d["bf16_tflops"]), tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cores():
    """Host cores this process may really use: min(affinity, cgroup CPU quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def cpu_oracle_time(steps, warmup, batch):
    """Times the reference's CPU implementation of the path (the oracle port, PyTorch eager fp32) on the host
    cores.  The ONLY place bench.py executes oracle/ (besides the synthetic input generator)."""
    import torch
    from oracle import vaenar_oracle as O
    from oracle.hparams import LJHPS as OH
    cores = usable_cores()
    torch.set_num_threads(cores)      # explicit: torchrun exports OMP_NUM_THREADS=1
    P = O.init_params(OH, seed=OH.Train.random_seed)
    texts, mels, t_len, m_len = O.synthetic_batch(OH, batch, T_TEXT, T_MEL)
    Tz = int(((m_len + RF - 1) // RF).max())
    eps = torch.randn(batch, Tz, 128, generator=torch.Generator().manual_seed(0))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.vaenar_inference(P, OH, texts, m_len, t_len, RF, eps)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return sum(times) / len(times), cores, torch.get_num_threads()


def gpu_eager_standin(device, batch, steps=3):
    """the BASELINE target "10x the reference TF2 single-GPU" has no measurable denominator here (TF 2.2 is
    not installable), so the oracle restatement run as PyTorch eager CUDA fp32 on the same B200 is reported as the
    stand-in single-GPU reference.  Library kernels (cuBLAS / ATen), nothing of this repo's CUDA path."""
    import torch
    from oracle import vaenar_oracle as O
    from oracle.hparams import LJHPS as OH
    try:
        P = {k: v.to(device) for k, v in O.init_params(OH, seed=OH.Train.random_seed).items()}
        texts, mels, t_len, m_len = O.synthetic_batch(OH, batch, T_TEXT, T_MEL)
        Tz = int(((m_len + RF - 1) // RF).max())
        eps = torch.randn(batch, Tz, 128, generator=torch.Generator().manual_seed(0)).to(device)
        texts, t_len, m_len = texts.to(device), t_len.to(device), m_len.to(device)
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        times = []
        with torch.no_grad(), torch.device(device):
            for i in range(steps + 2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                O.vaenar_inference(P, OH, texts, m_len, t_len, RF, eps)
                torch.cuda.synchronize()
                if i >= 2:
                    times.append(time.perf_counter() - t0)
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        sec = sum(times) / len(times)
        return {"value": batch * T_MEL / sec, "unit": "frames/s", "ms_per_step": sec * 1e3,
                "what": "oracle restatement, PyTorch eager CUDA fp32 (cuBLAS/ATen) on this GPU, same C2 batch; stand-in for the "
                        "uninstallable TF 2.2 single-GPU reference"}
    except Exception as e:   # reported, never fatal
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, cores, threads = cpu_oracle_time(args.steps, args.warmup, B_PER_GPU)
    frames = B_PER_GPU * T_MEL
    v = frames / sec
    line = {
        "impl": "reference", "metric": "mel-frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "note": "reference CPU path = oracle restatement (PyTorch eager fp32); TensorFlow 2.2 is not installable in this image",
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"full C2 batch (16 x 870 frames) per step, {args.steps} steps"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    return world, rank, local


def train_leg(args, world, rank, local, workload, steps, warmup, profile=True):
    """K full train steps (forward with tape + hand-written backward + gradient exchange when N > 1 + Keras Adam + operand
    re-pack), train.py:120-138.  c3: LJSpeech B32/GPU (BASELINE.json configs[2]); c4: DataBaker B16/GPU (the per-GPU share
    of configs[3]; BASELINE.json names no sequence shape, SURVEY.md 8d proposes T_text 152 / T_mel 640)."""
    import torch
    import torch.distributed as dist
    from vaenar_tts_b200 import VAENAR, LJHPS, DataBakerHPS, _lib
    lib = _lib.load()
    from oracle.vaenar_oracle import synthetic_batch
    from oracle.hparams import LJHPS as OLJ, DataBakerHPS as ODB
    c4 = workload == "c4"
    OH, HPS = (ODB, DataBakerHPS) if c4 else (OLJ, LJHPS)
    B, Tt, Tm, rf = (16 if c4 else args.train_batch), (152 if c4 else T_TEXT), (640 if c4 else T_MEL), RF
    dev = f"cuda:{local}"
    texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm, seed=OH.Train.random_seed + rank)
    h_texts, h_mels = texts.pin_memory(), mels.pin_memory()
    d_texts, d_mels, d_t, d_m = (x.to(dev) for x in (texts, mels, t_len, m_len))
    model = VAENAR(HPS, device=dev, seed=OH.Train.random_seed, noise_seed_offset=rank)
    model.init(d_texts, d_m, d_t)                       # init_step of train.py:172-179 (data-dependent ActNorm)
    if world > 1:
        model.broadcast_parameters(0)                   # all replicas start from rank 0's initialisation
    peer = world > 1 and not args.nccl_allreduce
    if peer:
        model.enable_peer_optimizer()                   # gradient exchange + Adam as one kernel over NVLink peer memory
    klw = float(OH.Train.kl_weight_init) if hasattr(OH.Train, "kl_weight_init") else 1e-5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_dev():
        return model.train_step(d_texts, d_mels, d_t, d_m, klw, rf)

    def step_e2e():
        out = model.train_step(h_texts.to(dev, non_blocking=True), h_mels.to(dev, non_blocking=True), d_t, d_m, klw, rf)
        return torch.stack(out).cpu()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    n0 = lib.vaenar_launch_count()
    step_dev()
    torch.cuda.synchronize()
    launches = lib.vaenar_launch_count() - n0
    for _ in range(max(warmup, 3) - 1):
        step_dev()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        step_dev()
        b.record()
    torch.cuda.synchronize()
    ms_dev = sum(a.elapsed_time(b) for a, b in evs)
    barrier()
    ms_e2e = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        losses = step_e2e()
        ms_e2e += (time.perf_counter() - t0) * 1e3
    barrier()
    # the exchange + optimiser half alone (train.py:136-137): N = 1 fused Adam; N > 1 barrier + reduce-scatter/Adam/all-gather
    # kernel over peer memory + barrier (or NCCL all-reduce + Adam)
    ex = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ex:
        a.record()
        model.exchange_and_apply()
        b.record()
    torch.cuda.synchronize()
    ms_ex = statistics.median(a.elapsed_time(b) for a, b in ex)
    barrier()
    t = torch.tensor([ms_dev, ms_e2e, ms_ex], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_ex = float(t[0]), float(t[1]), float(t[2])
    frames_total = world * B * Tm * steps
    out = None
    if rank == 0:
        pk = peaks()
        Tz = (Tm + rf - 1) // rf
        xblk = Tz * (1048576 + 512 * (Tz + Tt)) + Tt * 262144
        mac = Tt * (11534336 + 2048 * Tt) + 16 * xblk + Tz * (151552 + 6 * 65536 + 135168) + Tz * rf * 1433600
        flops_per_frame = 6.0 * mac / Tm            # train step = 3 x forward
        step_tflops = flops_per_frame * B * Tm / (ms_dev / steps * 1e-3) / 1e12
        out = {
            "workload": (f"{'C4 share: DataBaker' if c4 else 'C3: LJSpeech'} hparams, batch={B}/GPU, T_text={Tt}, T_mel={Tm}, full "
                         "train_step (encoder + posterior + prior flow + decoder + KL, backward, gradient exchange, Adam), rf=2"),
            "value": frames_total / (ms_dev / 1e3), "unit": "frames/s", "steps": steps, "ms_per_step": ms_dev / steps,
            "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(h_texts.numel() * 4 + h_mels.numel() * 4), "d2h_bytes_per_step": 16},
            "exchange_ms": ms_ex,
            "exchange": ("fused Keras Adam, one kernel (no exchange at N = 1)" if world == 1 else
                         ("2 stream barriers + ONE kernel: reduce-scatter over NVLink peer loads -> Adam on the shard -> all-gather "
                          "by peer stores (138.9 MB fp32 gradients)" if peer else "NCCL all-reduce of the flat gradient buffer + fused Adam")),
            "parallelism": f"dp{world}" if world > 1 else "single GPU",
            "launches_per_step": int(launches), "losses_last_step": [float(x) for x in losses],
            "loss_scale": float(model.loss_scale), "skipped_steps": int(model.skipped_steps),
            "whole_step": {"tflops": step_tflops * world, "tflops_per_gpu": step_tflops,
                           "frac_of_sustained_per_gpu": step_tflops / pk["tflops_sustained"]},
        }
        if profile:
            # per-class timing pass: one stream (the weight-gradient stream folded into the main one): exclusive times
            os.environ["VAENAR_NO_WGRAD_STREAM"] = "1"
            lib.vaenar_profile_enable(1)
            model.train_step_grads(d_texts, d_mels, d_t, d_m, klw, rf)
            os.environ.pop("VAENAR_NO_WGRAD_STREAM", None)
            rep = json.loads(lib.vaenar_profile_report().decode())
            lib.vaenar_profile_enable(0)
            tot_ms = sum(v["ms"] for v in rep.values()) or 1.0
            out["classes"] = {k: {"launches_per_step": v["launches"], "ms_per_step": v["ms"], "share": v["ms"] / tot_ms,
                                  "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12, "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9}
                              for k, v in rep.items() if v["ms"] > 0}
    del model
    torch.cuda.empty_cache()
    return out


def mel_inversion_leg(dev, with_cpu):
    """SURVEY.md 8f rank 4 (audio/utils.py:24-40): the mels of one C2 batch (16 x 870 frames) -> 60 Griffin-Lim
    iterations -> inverse pre-emphasis -> int16, all on the device (vaenar_tts_b200.audio).  Device time by CUDA events;
    the CPU figure is the numpy oracle (oracle/audio_oracle.py, the checker) on a bounded sample."""
    import numpy as np
    import torch
    from vaenar_tts_b200 import LJHPS
    from vaenar_tts_b200.audio import Audio
    a = Audio(LJHPS.Audio, device=dev)
    B, T = B_PER_GPU, T_MEL
    rng = np.random.default_rng(0)
    tt = np.arange(T)[:, None] / (T - 1)
    mm = np.arange(80)[None, :] / 80
    env = np.sin(np.pi * tt) ** 0.5
    mel = np.clip(0.55 * env * (1 - 0.6 * mm) + 0.15 * np.sin(2 * np.pi * (3 * tt + 2 * mm)) * env
                  + 0.05 * rng.standard_normal((B, T, 80)), 0, 1).astype(np.float32)
    d_mel = torch.from_numpy(mel).to(dev)
    lens = [T] * B

    def run():
        wav = a.inv_mel_spectrogram_batch(d_mel, lens, seed=1)
        return a.to_int16_batch(a.inv_preemphasize_batch(wav, lens), lens)
    run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = statistics.median(ms) * 1e-3
    hop, sr, iters = LJHPS.Audio.frame_shift_sample, LJHPS.Audio.sample_rate, LJHPS.Audio.griffin_lim_iters
    out = {"workload": f"mel inversion of one C2 batch: {B} x {T} frames, {iters} Griffin-Lim iterations (fp64 FFTs), "
                       "inverse pre-emphasis, int16", "ms_per_batch": t * 1e3,
           "frames_per_s": B * T / t, "audio_seconds_per_second": B * hop * (T - 1) / sr / t,
           "launches_per_batch": iters + 7}
    if with_cpu:
        from oracle import audio_oracle as AO
        o = AO.Audio(AO.LJAudio)
        Tc = 150
        S = o.linear_magnitudes(mel[0, :Tc].T)
        r = rng.random(S.shape)
        t0 = time.perf_counter()
        o._griffin_lim(S, rand=r, iters=iters)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": Tc / dt, "unit": "frames/s", "kind": "port", "cores": os.cpu_count(),
                               "sample": f"numpy restatement of audio.py:93-102, 1 utterance x {Tc} frames, {iters} iterations"}
    return out


def run_ours(args):
    import torch
    # ---------------- CPU baseline first (rank 0, before the process group exists: the other ranks simply wait in the
    # rendezvous; explicit thread count because torchrun exports OMP_NUM_THREADS=1)
    rank_env = int(os.environ.get("RANK", "0"))
    cpu = None
    if rank_env == 0 and not args.skip_cpu:
        sec, cores, threads = cpu_oracle_time(3, 1, B_PER_GPU)
        cpu = {"value": B_PER_GPU * T_MEL / sec, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "oracle (PyTorch eager fp32) on one full C2 batch (16 x 870 frames), mean of 3 runs"}
    import torch.distributed as dist
    world, rank, local = _dist_setup()
    from vaenar_tts_b200 import VAENAR, LJHPS, InferenceSession, _lib
    lib = _lib.load()
    from oracle.vaenar_oracle import synthetic_batch   # input generator only (shapes/lengths per SURVEY.md §8d)
    from oracle.hparams import LJHPS as OH

    B, Tt, Tm = B_PER_GPU, T_TEXT, T_MEL
    Tz = (Tm + RF - 1) // RF
    dev = f"cuda:{local}"
    model = VAENAR(LJHPS, device=dev, seed=OH.Train.random_seed)
    # zero-init projections would switch the coupling nets off numerically (not in cost); keep Keras defaults.
    nfl = max(1, args.inflight)

    def make_sessions(n, with_ali):
        ss = []
        for i in range(n):
            texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm, seed=OH.Train.random_seed + rank * 16 + i)
            s = InferenceSession(model, B, Tt, Tz, rf=RF, return_alignments=with_ali, seed=rank * 16 + i)
            s.set_inputs(texts, t_len, m_len)
            s.run_e2e()                      # eager pass (also packs weights)
            torch.cuda.synchronize()
            ss.append(s)
        return ss

    sessions = make_sessions(nfl, False)
    n1 = lib.vaenar_launch_count()
    sessions[0]._launch()
    torch.cuda.synchronize()
    launches_per_call = lib.vaenar_launch_count() - n1 + 1       # + the noise kernel of run_device
    for s in sessions:
        s.capture()
    streams = [torch.cuda.Stream(device=dev) for _ in range(nfl)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_serial(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()               # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)            # ms over `steps`

    def run_inflight(ss, steps, e2e):
        """exactly `steps` steps, round-robin over the sessions, each session on its own stream"""
        cur = torch.cuda.current_stream()
        for st in streams[:len(ss)]:
            st.wait_stream(cur)
        for k in range(steps):
            i = k % len(ss)
            with torch.cuda.stream(streams[i]):
                if e2e:
                    ss[i].run_e2e()
                else:
                    ss[i].run_device()
        for st in streams[:len(ss)]:
            cur.wait_stream(st)

    def timed_inflight(ss, steps):
        flush.zero_()                   # L2 flushed before the timed region; the working set of the in-flight batches
        torch.cuda.synchronize()        # (weights 64 MB + 2 x workspace) exceeds L2 on its own
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_inflight(ss, steps, False)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    def timed_e2e(ss, steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_inflight(ss, steps, True)   # H2D (pinned) + graph + D2H (pinned) per step, inside the timed region
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    # ---------------- device-resident numbers: serial (latency) and in flight (value)
    for _ in range(max(args.warmup, 3)):
        for s in sessions:
            s.run_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_serial = timed_serial(sessions[0].run_device, args.steps)
    barrier()
    run_inflight(sessions, 2 * nfl, False)
    barrier()
    ms_dev = timed_inflight(sessions, args.steps)
    barrier()
    # ---------------- end-to-end numbers (pinned host in, pinned host out, inside the timed region)
    run_inflight(sessions, 2 * nfl, True)
    barrier()
    ms_e2e = timed_e2e(sessions, args.steps)
    barrier()
    ms_e2e_serial = 0.0
    for _ in range(args.steps):          # wall clock per call incl. launch overhead, H2D, graph, D2H and the sync
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sessions[0].run_e2e()
        torch.cuda.current_stream().synchronize()
        ms_e2e_serial += (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # with the alignments VAENAR.inference also returns (models/models.py:199-210): 2 x 16.5 MB fp32 per step to the host
    ali_sessions = make_sessions(nfl, True)
    for s in ali_sessions:
        s.capture()
    run_inflight(ali_sessions, 2 * nfl, True)
    barrier()
    ms_e2e_ali = timed_e2e(ali_sessions, args.steps)
    ms_dev_ali = timed_inflight(ali_sessions, args.steps)
    barrier()
    d2h_ali = ali_sessions[0].d2h_bytes
    del ali_sessions
    t = torch.tensor([ms_dev, ms_e2e, ms_serial, ms_e2e_serial, ms_e2e_ali, ms_dev_ali], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_serial, ms_e2e_serial, ms_e2e_ali, ms_dev_ali = (float(x) for x in t)
    frames_total = world * B * Tm * args.steps

    line = None
    if rank == 0:
        # ---------------- roofline: per-launch CUDA-event timing of the kernel classes (eager, one stream: exclusive)
        pk = peaks()
        s0 = sessions[0]
        lib.vaenar_profile_enable(1)
        reps = 3
        for _ in range(reps):
            s0._launch()
        rep = json.loads(lib.vaenar_profile_report().decode())
        lib.vaenar_profile_enable(0)
        tot_ms = sum(v["ms"] for v in rep.values()) or 1.0
        dom = max(rep, key=lambda k: rep[k]["ms"])
        classes = {k: {"launches_per_step": v["launches"] // reps, "ms_per_step": v["ms"] / reps,
                       "share": v["ms"] / tot_ms, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                       "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} for k, v in rep.items() if v["ms"] > 0}
        d = rep[dom]
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        step_tflops = FLOPS_PER_FRAME * B * Tm / (ms_dev / args.steps * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tflops"], "peak_source": pk["source"] + " cuBLAS bf16 burst",
                    "traffic": None, "classes": classes,
                    "classes_note": "eager launches on ONE stream, each bracketed by CUDA events: exclusive times of the instrumented "
                                    "tensor-core kernels; SIMT glue kernels are not instrumented",
                    "classes_sum_ms": tot_ms / reps, "serial_step_ms": ms_serial / args.steps,
                    "whole_step": {"tflops": step_tflops, "frac_of_sustained": step_tflops / pk["tflops_sustained"],
                                   "serial_tflops": FLOPS_PER_FRAME * B * Tm / (ms_serial / args.steps * 1e-3) / 1e12}}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roofline["traffic"] = json.load(open(tr)).get(dom)
            except Exception:
                pass
        standin = None if args.skip_cpu else gpu_eager_standin(dev, B)
        line = {
            "metric": "mel-frames/sec", "value": frames_total / (ms_dev / 1e3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (flow + softmax + LN in f32)",
            "data": "synthetic", "config": dict(CONFIG),
            "execution": {"parallelism": f"replicas x{world} (no data-path collective)", "batches_in_flight": nfl,
                          "l2": "flushed before the timed region (serial numbers: between steps); weights + workspaces exceed L2",
                          "how": "CUDA graph replay of the C-ABI launch sequence, one graph / workspace / stream per in-flight batch"},
            "serial": {"value": frames_total / (ms_serial / 1e3), "unit": "frames/s", "ms_per_step": ms_serial / args.steps,
                       "e2e_ms_per_step": ms_e2e_serial / args.steps, "e2e_value": frames_total / (ms_e2e_serial / 1e3),
                       "what": "one batch at a time (latency per step), L2 flushed between steps"},
            "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": sessions[0].h2d_bytes, "d2h_bytes_per_step": sessions[0].d2h_bytes},
            "e2e_with_alignments": {"value": frames_total / (ms_e2e_ali / 1e3), "unit": "frames/s",
                                    "ms_per_step": ms_e2e_ali / args.steps, "device_ms_per_step": ms_dev_ali / args.steps,
                                    "h2d_bytes_per_step": sessions[0].h2d_bytes, "d2h_bytes_per_step": d2h_ali},
            "gpu_launches": int(launches_per_call * args.steps),
            "launches_per_step": int(launches_per_call),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_standin": standin,
        }
    del sessions
    torch.cuda.empty_cache()
    # ---------------- training leg (BASELINE.json configs[2] at N = 1; the per-GPU share of configs[3] at N > 1)
    if not args.no_train:
        tsteps = max(3, min(args.steps, args.train_steps))
        tr = train_leg(args, world, rank, local, "c3" if world == 1 else "c4", tsteps, 3, profile=False)
        if rank == 0:
            line["train"] = tr
    if rank == 0 and world == 1 and not args.no_audio:
        line["mel_inversion"] = mel_inversion_leg(dev, not args.skip_cpu)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """Training as the primary workload (--workload c3 | c4): same JSON contract, `value` = mel frames per second of
    training."""
    import torch.distributed as dist
    world, rank, local = _dist_setup()
    tr = train_leg(args, world, rank, local, args.workload, args.steps, args.warmup, profile=True)
    if rank == 0:
        pk = peaks()
        classes = tr.pop("classes", {})
        dom = max(classes, key=lambda k: classes[k]["ms_per_step"]) if classes else None
        line = {
            "metric": "mel-frames/sec", "value": tr["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (residual streams, LN/BN/softmax statistics, flow, Adam in f32)",
            "data": "synthetic", "config": {"workload": tr["workload"], "parallelism": tr["parallelism"]},
            "e2e": tr["e2e"], "gpu_launches": tr["launches_per_step"] * args.steps, "launches_per_step": tr["launches_per_step"],
            "train": tr,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": classes[dom]["tflops"] if dom else None,
                         "peak": pk["tflops"], "unit": "TFLOP/s", "frac": classes[dom]["tflops"] / pk["tflops"] if dom else None,
                         "peak_source": pk["source"] + " cuBLAS bf16 burst", "traffic": None, "classes": classes,
                         "whole_step": tr["whole_step"]},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline / eager stand-in legs (profiling runs only)")
    ap.add_argument("--no-train", action="store_true", help="omit the training leg of the default (c2) run")
    ap.add_argument("--no-audio", action="store_true", help="omit the mel-inversion leg (Griffin-Lim) of the default run at N = 1")
    ap.add_argument("--inflight", type=int, default=4,
                    help="independent batches in flight (own graph, workspace, stream): H2D / compute / D2H of consecutive "
                         "batches overlap")
    ap.add_argument("--train-steps", type=int, default=20, help="timed train steps of the default run's training leg")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default, the BASELINE.json metric): inference (+ a training leg); c3: full train_step (LJSpeech, "
                         "B32/GPU); c4: full train_step, DataBaker hparams, B16/GPU (run with --gpus 4 for the named config)")
    ap.add_argument("--train-batch", type=int, default=32, help="per-GPU batch of the c3 workload")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="c3/c4, N > 1: NCCL all-reduce + Adam instead of the fused peer-memory optimiser kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c3", "c4"):
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
