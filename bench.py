#!/usr/bin/env python
"""bench.py — audio-seconds separated per second on B200 for the Conv-TasNet hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode train|fwd]

Workload (BASELINE.json configs[1]): paper config N=256 L=20 B=256 H=512 P=3 X=8 R=4 C=2 gLN non-causal, batch 3 per
GPU x 4 s segments @ 8 kHz, one training step = forward + PIT SI-SNR loss + backward (+ gradient all-reduce when N>1)
+ clip_grad_norm_(5) + Adam — the body of the reference's hot loop (src/solver.py:188-196).  Synthetic 8 kHz mixtures,
reference-default random init.  One JSON line on stdout (rank 0).

`value`  : whole-job audio-s/s with inputs resident in HBM, timed with CUDA events over exactly K steps.
`e2e`    : the same through the public API with HOST (pinned) inputs: H2D of mixture/source/lengths and a D2H read of
           the loss inside the timed region, every step.
`roofline`: the dominant kernel (the 1x1-conv GEMM) timed live on its own launches.
`cpu_baseline`: the CPU oracle port of the reference step on the box's host cores (rank 0, N=1, bounded sample).
`--impl reference`: times that CPU port alone, same metric/config (the reference has no GPU-specific code and is
           not pip-installable).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PAPER = dict(N=256, L=20, B=256, H=512, P=3, X=8, R=4, C=2, norm_type="gLN", causal=False, mask_nonlinear="relu")
SR = 8000
PER_GPU_BATCH = 3
SEG_SECONDS = 4
WORKLOAD = ("configs[1]: paper config N=256 L=20 B=256 H=512 P=3 X=8 R=4 C=2 gLN non-causal, batch 3 per GPU x 4 s "
            "@ 8 kHz, training step = fwd + PIT SI-SNR + bwd (+ NCCL grad all-reduce) + clip(5) + Adam")


def synthetic(M, T, C, L, seed):
    """SURVEY §8d inputs: sources ~ N(0, 0.05^2), mixture = clamp(sum), last item 3*S+7 samples short."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(M, C, T, generator=g) * 0.05
    lengths = torch.full((M,), T, dtype=torch.long)
    short = T - 3 * (L // 2) - 7
    lengths[-1] = short
    src[-1, :, short:] = 0
    mix = src.sum(1).clamp_(-0.9, 0.9)
    return mix, src, lengths


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def reference_steps(n_steps, warmup, M, T, budget_s=150.0):
    """The reference training step (solver.py:188-196) restated on the CPU oracle, all host threads.
    Returns (audio_seconds_per_second, seconds_per_step, sample description, cores)."""
    from oracle import conv_tasnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Config(**PAPER)
    sd = O.init_state_dict(cfg, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)

    def one(mix, src, lens):
        est = O.forward(cfg, params, mix)
        loss, *_ = O.cal_loss(src, est, lens)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 5)
        opt.step()
        return loss.item()

    mix, src, lens = synthetic(M, T, cfg.C, cfg.L, 1234 + 2)
    t0 = time.perf_counter()
    one(mix, src, lens)  # first warm-up, also the probe that sizes the sample
    probe = time.perf_counter() - t0
    m_used = M
    if probe * (n_steps + max(0, warmup - 1)) > budget_s and M > 1:
        m_used = 1  # bounded sample: one utterance of the batch
        mix, src, lens = mix[:1], src[:1], torch.full((1,), T, dtype=torch.long)
    for _ in range(max(0, warmup - 1)):
        one(mix, src, lens)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        one(mix, src, lens)
    dt = (time.perf_counter() - t0) / n_steps
    sample = (f"{n_steps} steps of the CPU oracle port (torch {torch.__version__}, {cores} threads) on "
              f"{m_used} x {T / SR:.0f} s of the {M} x {T / SR:.0f} s per-GPU batch, fwd+PIT+bwd+clip+Adam")
    return m_used * T / SR / dt, dt, sample, cores


def run_reference(args, rank):
    if rank != 0:
        return
    T = SEG_SECONDS * SR
    val, dt, sample, cores = reference_steps(args.steps, args.warmup, PER_GPU_BATCH, T)
    line = {"impl": "reference", "metric": "audio-seconds separated per second (training step)", "value": val,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="run only warm-up + K device-resident steps (for ncu)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying "
                    "one captured CUDA graph per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from conv_tasnet_b200 import ConvTasNet, cal_loss, _lib
    from conv_tasnet_b200.data_parallel import ShardedDataParallel
    from conv_tasnet_b200.optim import FusedAdam

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT at VERSION/INFO level; stdout carries exactly one JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("CTN_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    T = SEG_SECONDS * SR
    M = PER_GPU_BATCH
    torch.manual_seed(0)
    model = ConvTasNet(**PAPER).cuda()
    model.train()
    dp = ShardedDataParallel(model) if world > 1 else model
    opt = FusedAdam(model, lr=1e-3, max_grad_norm=5.0)
    mix_h, src_h, len_h = synthetic(M, T, PAPER["C"], PAPER["L"], 1234 + 2 + rank)
    mix_h, src_h, len_h = mix_h.pin_memory(), src_h.pin_memory(), len_h.pin_memory()
    mix_d, src_d, len_d = mix_h.to(dev), src_h.to(dev), len_h.to(dev)

    def step(mix, src, lens):
        est = dp(mix)
        loss, _, _, _ = cal_loss(src, est, lens)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3

    # ---- one training step captured as a CUDA graph (conv_tasnet_b200.graph.GraphedTrainStep: the launch sequence is
    # static — no host sync, no allocation in the C ABI, device-side step counter) and replayed per step ------------
    from conv_tasnet_b200.graph import GraphedInference, GraphedTrainStep
    use_graph = not args.no_graph and not args.profile_only
    gstep = GraphedTrainStep(dp, opt) if use_graph else None
    launches_per_step = 0
    if gstep is not None:
        n_before = L.ctn_launch_count()
        gstep(mix_d, src_d, len_d)  # warm-up + capture
        launches_per_step = (L.ctn_launch_count() - n_before) // (gstep.warmup + 1)
        if not gstep.captured:
            gstep = None

    def run_step():
        if gstep is not None:
            return gstep(mix_d, src_d, len_d)
        return step(mix_d, src_d, len_d)

    # ---- device-resident throughput -------------------------------------------------------------------
    for _ in range(args.warmup):
        run_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = L.ctn_launch_count()
    secs = timed(run_step, args.steps)
    launches = (launches_per_step * args.steps) if gstep is not None else (L.ctn_launch_count() - n0)
    audio = world * M * T / SR * args.steps
    value = audio / secs
    if args.profile_only:
        sampler.stop()
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": secs / args.steps * 1e3, "gpu_launches": int(launches)}))
        return

    # ---- end to end through the public API with host inputs -------------------------------------------
    losses = []

    def e2e_step():
        if gstep is not None:  # H2D from pinned memory into the graph's static inputs, replay, D2H of the loss
            losses.append(gstep(mix_h, src_h, len_h).item())
            return
        mix = mix_h.to(dev, non_blocking=True)
        src = src_h.to(dev, non_blocking=True)
        lens = len_h.to(dev, non_blocking=True)
        losses.append(step(mix, src, lens).item())  # D2H read of the loss, every step

    for _ in range(2):
        e2e_step()
    secs_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop()
    h2d = mix_h.numel() * 4 + src_h.numel() * 4 + len_h.numel() * 8

    # ---- forward-only throughput (the metric's other half) --------------------------------------------
    model.eval()
    with torch.no_grad():
        infer = GraphedInference(dp) if use_graph else dp
        for _ in range(3):
            infer(mix_d)
        secs_fwd = timed(lambda: infer(mix_d), args.steps)
    model.train()

    # ---- roofline of the dominant kernel by time share (profiles/r1_final_summary.md): the weight-gradient GEMM
    # dW[H,B] = dz1[F,H]^T x[F,B] on tcgen05 (bf16x3 split, MN-major operands), timed live on its own launches --------
    pk, pk_src = peaks()
    K = L.ctn_num_frames(ctypes.byref(model._cfg), T)
    F = M * K
    Gm = torch.randn(F, PAPER["H"], device=dev)
    Xm = torch.randn(F, PAPER["B"], device=dev)
    dWm = torch.zeros(PAPER["H"], PAPER["B"], device=dev)
    flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)  # >> L2 (126 MB); also keeps the GPU busy while the
    # host enqueues the timed launch, so the event pair brackets the kernel alone and not host launch latency
    st = _lib.stream()

    def wgrad():
        _lib.check(L.ctn_wgrad(Gm.data_ptr(), Xm.data_ptr(), dWm.data_ptr(), F, PAPER["H"], PAPER["B"], K, None, None,
                               None, None, None, st))
    for _ in range(3):
        wgrad()
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wgrad()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) * 1e-3
    t_k = tot / reps
    flops = 2.0 * F * PAPER["B"] * PAPER["H"]
    ach = flops / t_k / 1e12
    roofline = {"kernel": "tc_wgrad_kernel<256>: dW[512,256] = dz1[F,512]^T x[F,256] (tcgen05, bf16x3 split, MN-major "
                "operands, split-K + coalesced red.v4), the largest single kernel of the step (2 per TemporalBlock + 2)", "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                "traffic": 30042112, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per "
                "launch (profiles/r1b_wgrad_full_raw.csv); algorithmic input 29.5 MB",
                "peak_source": pk_src + ", bf16 burst (the bf16x3 split issues 3 MMAs per algorithmic MAC, so the "
                "ceiling of this fraction is 1/3)", "launch_us": t_k * 1e6, "alg_flops_per_launch": flops,
                "launches_per_step": 2 * PAPER["R"] * PAPER["X"] + 2,
                "share_of_step": (2 * PAPER["R"] * PAPER["X"] + 2) * t_k / (secs / args.steps)}

    # whole-step algorithmic rates (SURVEY §8d per-frame figures x frames)
    frames = world * F
    step_flops = frames * (17.30e6 * 3 - 2 * PAPER["N"] * PAPER["L"])
    step_bytes = frames * (91422 + 241950 + 5 * PAPER["C"] * 10) * 4 + 10 * 4 * 8710720 * world
    t_step = secs / args.steps

    line = {"metric": "audio-seconds separated per second (training step)", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": M, "segment_s": SEG_SECONDS, "global_batch": M * world,
                       "parallelism": f"dp{world}", "l2": "per-step working set (1.7 GB activation stash) exceeds L2 "
                       "(126 MB); roofline kernel timed with an explicit 1 GB L2 flush between launches"},
            "e2e": {"value": audio / secs_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": secs_e2e / args.steps * 1e3, "last_loss": losses[-1]},
            "fwd": {"value": audio / secs_fwd, "unit": "audio-s/s", "ms_per_step": secs_fwd / args.steps * 1e3},
            "gpu_launches": int(launches), "cuda_graph": gstep is not None, "clocks": clocks, "roofline": roofline,
            "step_algorithmic": {"tflops": step_flops / t_step / 1e12, "gbs": step_bytes / t_step / 1e9,
                                 "hbm_bound_ms": step_bytes / world / (pk["hbm_gbs"] * 1e9) * 1e3,
                                 "frac_of_hbm_bound": (step_bytes / world / (pk["hbm_gbs"] * 1e9)) / t_step}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, dt, sample, cores = reference_steps(2, 1, M, T, budget_s=30.0)
        line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
