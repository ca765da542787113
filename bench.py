#!/usr/bin/env python
"""bench.py -- mel-frames/sec of the VAENAR-TTS mel-synthesis hot path on B200.

Workload (BASELINE.json configs[1], "C2"): LJSpeech hparams, batch 16 per GPU, T_text 148, T_mel 870,
inference only (text encoder -> prior flow sample -> decoder), reduction factor 2, synthetic batch and
random-init weights.  One "step" = one VAENAR.inference call over one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                      # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  value = frames/s with inputs resident in HBM (CUDA-graph replay, CUDA events,
L2 flushed between steps); e2e = the same through the public API with pinned-host inputs / output copies inside
the timed region; roofline = dominant tensor-core kernel class timed per launch with CUDA events; cpu_baseline =
the oracle on the host cores (bounded sample), the only use of oracle/ here.
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
    cores.  The ONLY place bench.py executes oracle/."""
    import torch
    from oracle import vaenar_oracle as O
    from oracle.hparams import LJHPS as OH
    cores = usable_cores()
    torch.set_num_threads(cores)
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
        "config": {"workload": WORKLOAD, "note": "reference CPU path = oracle restatement (PyTorch eager fp32); "
                   "TensorFlow 2.2 is not installable in this image"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"full C2 batch (16 x 870 frames) per step, {args.steps} steps"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
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
    from vaenar_tts_b200 import VAENAR, LJHPS, InferenceSession, _lib
    lib = _lib.load()
    from oracle.vaenar_oracle import synthetic_batch   # input generator only (shapes/lengths per SURVEY.md §8d)
    from oracle.hparams import LJHPS as OH

    B, Tt, Tm = B_PER_GPU, T_TEXT, T_MEL
    texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm, seed=OH.Train.random_seed + rank)
    Tz = (Tm + RF - 1) // RF
    model = VAENAR(LJHPS, device=f"cuda:{local}", seed=OH.Train.random_seed)
    # zero-init projections would switch the coupling nets off numerically (not in cost); keep Keras defaults.
    sess = InferenceSession(model, B, Tt, Tz, rf=RF, return_alignments=False, seed=rank)
    sess.set_inputs(texts, t_len, m_len)
    n0 = lib.vaenar_launch_count()
    sess.run_e2e()                      # eager pass (also sizes workspace, packs weights)
    torch.cuda.synchronize()
    launches_per_call = None
    n1 = lib.vaenar_launch_count()
    sess._launch()
    torch.cuda.synchronize()
    launches_per_call = lib.vaenar_launch_count() - n1 + 1       # + the noise kernel of run_device
    sess.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()               # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)            # ms over `steps`

    # ---------------- device-resident number (value)
    for _ in range(max(args.warmup, 3)):
        sess.run_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(sess.run_device, args.steps)
    barrier()
    # ---------------- end-to-end number (pinned host in, pinned host out, inside the timed region)
    for _ in range(3):
        sess.run_e2e()
    barrier()
    ms_e2e = 0.0
    for _ in range(args.steps):          # wall clock per call incl. launch overhead, H2D, graph, D2H and the sync
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sess.run_e2e()
        torch.cuda.current_stream().synchronize()
        ms_e2e += (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    frames_total = world * B * Tm * args.steps
    value = frames_total / (ms_dev / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)

    if rank == 0:
        # ---------------- roofline: per-launch CUDA-event timing of the tensor-core kernel classes (eager, N=1 rank)
        pk = peaks()
        lib.vaenar_profile_enable(1)
        reps = 3
        for _ in range(reps):
            sess._launch()
        rep = json.loads(lib.vaenar_profile_report().decode())
        lib.vaenar_profile_enable(0)
        tot_ms = sum(v["ms"] for v in rep.values()) or 1.0
        dom = max(rep, key=lambda k: rep[k]["ms"])
        classes = {k: {"launches_per_step": v["launches"] // reps, "ms_per_step": v["ms"] / reps,
                       "share": v["ms"] / tot_ms, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                       "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} for k, v in rep.items() if v["ms"] > 0}
        d = rep[dom]
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tflops"], "peak_source": pk["source"] + " cuBLAS bf16 burst",
                    "traffic": None, "classes": classes,
                    "whole_step": {"tflops": FLOPS_PER_FRAME * B * Tm / (ms_dev / args.steps * 1e-3) / 1e12,
                                   "frac_of_sustained": FLOPS_PER_FRAME * B * Tm / (ms_dev / args.steps * 1e-3) / 1e12 /
                                   pk["tflops_sustained"]}}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roofline["traffic"] = json.load(open(tr)).get(dom)
            except Exception:
                pass
        # ---------------- CPU baseline (bounded sample: one full C2 batch, 1 warm-up + 3 runs)
        if args.skip_cpu:
            cpu = None
        else:
            sec, cores, threads = cpu_oracle_time(3, 1, B)
            cpu = {"value": B * Tm / sec, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "oracle (PyTorch eager fp32) on one full C2 batch (16 x 870 frames), mean of 3 runs"}
        line = {
            "metric": "mel-frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (flow + softmax + LN in f32)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "T_text": Tt, "T_mel": Tm, "rf": RF,
                       "parallelism": f"replicas x{world} (no data-path collective)", "l2": "flushed between timed steps",
                       "execution": "CUDA graph replay of the C-ABI launch sequence"},
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": sess.h2d_bytes, "d2h_bytes_per_step": sess.d2h_bytes},
            "gpu_launches": int(launches_per_call * args.steps),
            "launches_per_step": int(launches_per_call),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """Secondary workload (BASELINE.json configs[2], "C3"): LJSpeech hparams, batch 32 per GPU, full train_step
    (forward + ELBO + hand-written backward + one NCCL all-reduce of the flat gradients when N > 1 + Keras Adam + operand
    re-pack).  Same JSON contract; `value` = mel frames per second of training."""
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
    from vaenar_tts_b200 import VAENAR, LJHPS, DataBakerHPS, _lib
    lib = _lib.load()
    from oracle.vaenar_oracle import synthetic_batch
    from oracle.hparams import LJHPS as OLJ, DataBakerHPS as ODB
    c4 = args.workload == "c4"
    # C4 (BASELINE.json configs[3]): DataBaker hparams, global batch 64 = 16 per GPU on 4 GPUs; BASELINE.json names no
    # sequence shape, SURVEY.md 8d proposes T_text 152 / T_mel 640.
    OH, HPS = (ODB, DataBakerHPS) if c4 else (OLJ, LJHPS)
    B, Tt, Tm, rf = (args.train_batch if not c4 else 16), (152 if c4 else T_TEXT), (640 if c4 else T_MEL), RF
    dev = f"cuda:{local}"
    texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm, seed=OH.Train.random_seed + rank)
    h_texts, h_mels = texts.pin_memory(), mels.pin_memory()
    d_texts, d_mels, d_t, d_m = (x.to(dev) for x in (texts, mels, t_len, m_len))
    model = VAENAR(HPS, device=dev, seed=OH.Train.random_seed)
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
    for _ in range(max(args.warmup, 3) - 1):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        step_dev()
        b.record()
    torch.cuda.synchronize()
    ms_dev = sum(a.elapsed_time(b) for a, b in evs)
    barrier()
    ms_e2e = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        losses = step_e2e()
        ms_e2e += (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    frames_total = world * B * Tm * args.steps
    if rank == 0:
        pk = peaks()
        # per-class timing pass: local (no collective on this rank alone) and with the weight-gradient stream folded into
        # the main stream, so that every class is timed as its own execution time, not as an overlapped interval
        os.environ["VAENAR_NO_WGRAD_STREAM"] = "1"
        lib.vaenar_profile_enable(1)
        model.train_step_grads(d_texts, d_mels, d_t, d_m, klw, rf)
        os.environ.pop("VAENAR_NO_WGRAD_STREAM", None)
        rep = json.loads(lib.vaenar_profile_report().decode())
        lib.vaenar_profile_enable(0)
        tot_ms = sum(v["ms"] for v in rep.values()) or 1.0
        classes = {k: {"launches_per_step": v["launches"], "ms_per_step": v["ms"], "share": v["ms"] / tot_ms,
                       "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12, "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9}
                   for k, v in rep.items() if v["ms"] > 0}
        dom = max(rep, key=lambda k: rep[k]["ms"])
        achieved = classes[dom]["tflops"]
        # SURVEY.md 8d algorithmic FLOPs of one train step (fwd + bwd = 3 x fwd, rf = 2), per mel frame
        Tz = (Tm + rf - 1) // rf
        xblk = Tz * (1048576 + 512 * (Tz + Tt)) + Tt * 262144
        mac = Tt * (11534336 + 2048 * Tt) + 16 * xblk + Tz * (151552 + 6 * 65536 + 135168) + Tz * rf * 1433600
        flops_per_frame = 6.0 * mac / Tm
        step_tflops = flops_per_frame * B * Tm / (ms_dev / args.steps * 1e-3) / 1e12
        line = {
            "metric": "mel-frames/sec", "value": frames_total / (ms_dev / 1e3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (residual streams, LN/BN/softmax statistics, flow, Adam in f32)",
            "data": "synthetic",
            "config": {"workload": f"{'C4: DataBaker' if c4 else 'C3: LJSpeech'} hparams, batch={B}/GPU, T_text={Tt}, T_mel={Tm}, "
                       "full train_step (encoder + posterior + prior flow + decoder + KL, backward, Adam), rf=2", "batch_per_gpu": B,
                       "parallelism": (f"dp{world} (" + ("reduce-scatter + Adam + all-gather fused in one kernel over NVLink peer "
                                       "memory" if peer else "one NCCL all-reduce of the flat gradient buffer") + ")")
                       if world > 1 else "single GPU",
                       "l2": "flushed between timed steps", "execution": "eager C-ABI launch sequence"},
            "e2e": {"value": frames_total / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h_texts.numel() * 4 + h_mels.numel() * 4), "d2h_bytes_per_step": 16},
            "gpu_launches": int(launches * args.steps), "launches_per_step": int(launches), "clocks": clocks,
            "losses_last_step": [float(x) for x in losses],
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "peak_source": pk["source"] + " cuBLAS bf16 burst", "traffic": None,
                         "classes": classes,
                         "whole_step": {"tflops": step_tflops, "frac_of_sustained": step_tflops / pk["tflops_sustained"]}},
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
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs only)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default, the BASELINE.json metric): inference; c3: full train_step (LJSpeech, B32/GPU); "
                         "c4: full train_step, DataBaker hparams, B16/GPU (run with --gpus 4 for the named config)")
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
