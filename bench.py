#!/usr/bin/env python
"""bench.py — audio-seconds separated per second on B200 for the Conv-TasNet hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 0..4] [--mode train|fwd]

Workloads = BASELINE.json `configs` (paper config N=256 L=20 B=256 H=512 P=3 X=8 R=4, 8 kHz, synthetic mixtures,
reference-default random init):
    --config 0  one 4 s mixture, forward + cal_loss (the reference's CPU-runnable case)
    --config 1  (default) training step, batch 3 per GPU x 4 s: fwd + PIT SI-SNR + bwd (+ NCCL all-reduce) + clip(5) + Adam
                — the body of the reference's hot loop (src/solver.py:188-196)
    --config 2  causal cLN variant, batch 32 x 4 s, forward
    --config 3  3-speaker (C=3, 6-permutation PIT) training step, batch 16 per GPU x 4 s
    --config 4  long-utterance inference, 8 x 60 s per GPU (64 x 60 s sharded by utterance over 8 GPUs, no collective)
`--mode` overrides the config's own mode (e.g. `--config 1 --mode fwd`: forward throughput of the headline batch).
One JSON line on stdout (rank 0).

`value`  : whole-job audio-s/s with inputs resident in HBM, timed with CUDA events over exactly K steps.
`e2e`    : the same through the public API with HOST (pinned) inputs: H2D of the step's inputs and a D2H read of its
           result (the loss; for forward modes a checksum of the estimate) inside the timed region, every step.
`roofline`: the dominant kernel timed live on its own launches (L2 flushed between them).
`cpu_baseline`: the CPU oracle port of the reference step on the box's host cores (rank 0, N=1, bounded sample).
`--impl reference`: times that CPU port alone, same metric/config (the reference has no GPU-specific code and is
           not pip-installable).
"""
import argparse
import csv
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
PAPER_STR = "N=256 L=20 B=256 H=512 P=3 X=8 R=4"
TRAIN_STR = "training step = fwd + PIT SI-SNR + bwd (+ gradient all-reduce over NVLink) + clip(5) + Adam"
CONFIGS = {
    0: dict(model={}, M=1, seconds=4, mode="fwd_loss",
            workload=f"configs[0]: paper config {PAPER_STR} C=2 gLN non-causal, one 4 s @ 8 kHz mixture, fp32 forward + cal_loss"),
    1: dict(model={}, M=3, seconds=4, mode="train",
            workload=f"configs[1]: paper config {PAPER_STR} C=2 gLN non-causal, batch 3 per GPU x 4 s @ 8 kHz, {TRAIN_STR}"),
    2: dict(model=dict(norm_type="cLN", causal=True), M=32, seconds=4, mode="fwd", dtype="bf16",
            workload=f"configs[2]: causal cLN variant (causal=1, norm_type=cLN, Chomp1d) of the paper config {PAPER_STR} C=2, "
                     "batch 32 per GPU x 4 s @ 8 kHz, bf16 forward (H-wide activations stored as bf16, single-bf16 MMAs, "
                     "fp32 accumulation / residual stream / statistics / I-O)"),
    3: dict(model=dict(C=3), M=16, seconds=4, mode="train",
            workload=f"configs[3]: 3-speaker (C=3, 6-permutation PIT) paper config {PAPER_STR} gLN non-causal, batch 16 per GPU x 4 s "
                     f"@ 8 kHz, {TRAIN_STR}"),
    4: dict(model={}, M=8, seconds=60, mode="fwd",
            workload=f"configs[4]: long-utterance inference, paper config {PAPER_STR} C=2 gLN non-causal, 8 x 60 s @ 8 kHz per GPU "
                     "(64 x 60 s sharded by utterance over 8 GPUs, no collective; gLN reduction over ~48k frames per utterance)"),
}
METRIC = {"train": "audio-seconds separated per second (training step)",
          "fwd": "audio-seconds separated per second (forward)",
          "fwd_loss": "audio-seconds separated per second (forward + cal_loss)"}


def model_kwargs(cfg_id):
    kw = dict(PAPER)
    kw.update(CONFIGS[cfg_id]["model"])
    return kw


def config_dict(cfg_id, mode, world):
    """The `config` object of the JSON line — identical for the B200 arm and the reference arm."""
    c = CONFIGS[cfg_id]
    return {"workload": c["workload"] + ("" if mode == c["mode"] else f" [--mode {mode}]"), "config_id": cfg_id,
            "mode": mode, "per_gpu_batch": c["M"], "segment_s": c["seconds"], "global_batch": c["M"] * world,
            "parallelism": f"dp{world}" if mode == "train" else f"utterance-sharded x{world} (no collective)",
            "l2": "per-step working set exceeds L2 (126 MB); roofline kernel timed with an explicit L2 flush between launches"}


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


def profiled_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the newest committed `ncu --set full` raw page under
    profiles/ that holds a kernel whose name contains `kernel_substr`; (None, why) when there is none."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        pages = sorted((f for f in os.listdir(pdir) if f.endswith("_full_raw.csv")), reverse=True)
    except OSError:
        return None, "profiles/ is missing"
    for fn in pages:
        try:
            with open(os.path.join(pdir, fn), newline="") as fh:
                rows = list(csv.reader(fh))
            hdr, units = rows[0], rows[1]
            ik = hdr.index("Kernel Name")
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    tot = float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
                    return int(tot), f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/{fn})"
        except Exception:
            continue
    return None, "no ncu --set full capture of this kernel under profiles/"


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
def reference_steps(cfg_id, mode, n_steps, warmup, budget_s=150.0):
    """The reference step of this config restated on the CPU oracle, all host threads.  The sample is bounded: when the
    probe step says the full per-GPU batch would not finish in `budget_s`, the batch is cut to one utterance (and a 60 s
    utterance to 8 s) and the throughput is that of the sample — labelled in `sample`.
    Returns (audio_seconds_per_second, seconds_per_step, sample description, cores)."""
    from oracle import conv_tasnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = model_kwargs(cfg_id)
    cfg = O.Config(**kw)
    sd = O.init_state_dict(cfg, seed=0)
    M, T = CONFIGS[cfg_id]["M"], CONFIGS[cfg_id]["seconds"] * SR
    train = mode == "train"
    params = {k: v.clone().requires_grad_(train) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3) if train else None

    def one(mix, src, lens):
        if train:
            est = O.forward(cfg, params, mix)
            loss, *_ = O.cal_loss(src, est, lens)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(params.values()), 5)
            opt.step()
            return loss.item()
        with torch.no_grad():
            est = O.forward(cfg, params, mix, training=False)
            if mode == "fwd_loss":
                return O.cal_loss(src, est, lens)[0].item()
            return est.abs().sum().item()

    # probe on one short utterance to size the sample
    m_used, t_used = M, T
    pm, ps, pl = synthetic(1, min(T, 4 * SR), cfg.C, cfg.L, 99)
    t0 = time.perf_counter()
    one(pm, ps, pl)
    probe = (time.perf_counter() - t0) / (min(T, 4 * SR) / SR)  # seconds of CPU per audio-second (cold)
    total_steps = n_steps + max(0, warmup)
    if probe * M * (T / SR) * total_steps > budget_s:
        m_used = 1
        if probe * (T / SR) * total_steps > budget_s:
            t_used = min(T, 8 * SR)
    mix, src, lens = synthetic(m_used, t_used, cfg.C, cfg.L, 1234 + cfg_id)
    for _ in range(max(0, warmup - 1)):
        one(mix, src, lens)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        one(mix, src, lens)
    dt = (time.perf_counter() - t0) / n_steps
    what = {"train": "fwd+PIT+bwd+clip+Adam", "fwd": "forward (no_grad)", "fwd_loss": "forward + cal_loss (no_grad)"}[mode]
    sample = (f"{n_steps} steps of the CPU oracle port (torch {torch.__version__}, {cores} threads) on "
              f"{m_used} x {t_used / SR:.0f} s of the {M} x {T / SR:.0f} s per-GPU batch, {what}"
              + ("" if (m_used, t_used) == (M, T) else " — bounded sample, throughput of the sample (BASELINE.md §3)"))
    return m_used * t_used / SR / dt, dt, sample, cores


def run_reference(args, rank, mode):
    if rank != 0:
        return
    val, dt, sample, cores = reference_steps(args.config, mode, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC[mode], "value": val,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.config, mode, args.gpus),
            "reference_batch": "the CPU arm runs ONE per-GPU batch on rank 0 whatever --gpus says (it has no multi-GPU "
                               "path): only the N=1 ratio is like for like",
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
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--mode", default=None, choices=["train", "fwd", "fwd_loss"])
    ap.add_argument("--dtype", default=None, choices=["f32", "bf16"], help="forward modes only: bf16 = the reduced-precision "
                    "inference path (default: the config's own, bf16 for --config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default=None, choices=["peer", "nccl"],
                    help="N > 1 training: gradient exchange = the one-kernel peer-memory all-reduce inside the step's graph "
                         "(csrc/peer_reduce.cu) or bucketed NCCL all-reduces between per-stage graphs; default: "
                         "CTN_EXCHANGE or peer")
    ap.add_argument("--profile-only", action="store_true", help="run only warm-up + K device-resident steps (for ncu)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying "
                    "one captured CUDA graph per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    mode = args.mode or CONFIGS[args.config]["mode"]
    dtype = args.dtype or CONFIGS[args.config].get("dtype", "f32")
    if mode == "train":
        dtype = "f32"  # the reference trains in fp32 and so does this path

    if os.environ.get("BENCH_WATCHDOG"):  # debug: dump every thread's Python stack if the run is still going after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=False)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, mode)
        return

    import torch.distributed as dist
    from conv_tasnet_b200 import ConvTasNet, cal_loss, _lib
    from conv_tasnet_b200.data_parallel import ShardedDataParallel
    from conv_tasnet_b200.graph import GraphedInference, GraphedTrainStep
    from conv_tasnet_b200.optim import FusedAdam

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # NCCL_DEBUG is left as the caller set it
    L = _lib.lib()

    kw = model_kwargs(args.config)
    T = CONFIGS[args.config]["seconds"] * SR
    M = CONFIGS[args.config]["M"]
    train = mode == "train"
    torch.manual_seed(0)
    model = ConvTasNet(**kw).cuda()
    model.train() if train else model.eval()
    if dtype == "bf16":
        model.half_inference(True)
    # data parallel training shards the batch and all-reduces gradients; inference shards by utterance with no collective
    exchange = args.exchange or os.environ.get("CTN_EXCHANGE", "peer")
    if world > 1 and train:
        try:
            dp = ShardedDataParallel(model, peer_reduce=(exchange == "peer"))
        except RuntimeError as e:  # (collective: the set-up fails on every rank or on none) fall back to NCCL
            if rank == 0:
                sys.stderr.write(f"bench: peer-memory exchange unavailable ({e}); using NCCL\n")
            exchange = "nccl"
            dp = ShardedDataParallel(model, peer_reduce=False)
    else:
        dp = model
    opt = FusedAdam(model, lr=1e-3, max_grad_norm=5.0) if train else None
    mix_h, src_h, len_h = synthetic(M, T, kw["C"], kw["L"], 1234 + args.config + rank)
    mix_h, src_h, len_h = mix_h.pin_memory(), src_h.pin_memory(), len_h.pin_memory()
    mix_d, src_d, len_d = mix_h.to(dev), src_h.to(dev), len_h.to(dev)
    use_graph = not args.no_graph and not args.profile_only

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

    launches_per_step = 0
    graphed = False
    if train:
        def eager_step(mix, src, lens):
            est = dp(mix)
            loss, _, _, _ = cal_loss(src, est, lens)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss

        # one training step captured as a CUDA graph (conv_tasnet_b200.graph.GraphedTrainStep: the launch sequence is
        # static — no host sync, no allocation in the C ABI, device-side step counter) and replayed per step
        gstep = GraphedTrainStep(dp, opt) if use_graph else None
        if gstep is not None:
            n_before = L.ctn_launch_count()
            gstep(mix_d, src_d, len_d)  # warm-up + capture
            launches_per_step = (L.ctn_launch_count() - n_before) // (gstep.warmup + 1)
            if not gstep.captured:
                gstep = None
        graphed = gstep is not None

        def run_step():
            return gstep(mix_d, src_d, len_d) if gstep is not None else eager_step(mix_d, src_d, len_d)

        def e2e_step():  # H2D from pinned memory (into the graph's static inputs), the step, D2H of the loss
            if gstep is not None:
                return gstep(mix_h, src_h, len_h).item()
            return eager_step(mix_h.to(dev, non_blocking=True), src_h.to(dev, non_blocking=True),
                              len_h.to(dev, non_blocking=True)).item()
        h2d, d2h = mix_h.numel() * 4 + src_h.numel() * 4 + len_h.numel() * 8, 4
    else:
        infer = GraphedInference(model) if use_graph else model
        with_loss = mode == "fwd_loss"

        def fwd(mix, src, lens):
            with torch.no_grad():
                est = infer(mix)
                if with_loss:
                    return cal_loss(src, est, lens)[0]
                return est

        if use_graph:  # kernels per step, counted on one eager pass (the graph replays the same sequence)
            n_before = L.ctn_launch_count()
            with torch.no_grad():
                est0 = model(mix_d)
                if with_loss:
                    cal_loss(src_d, est0, len_d)
            launches_per_step = L.ctn_launch_count() - n_before
            del est0
            fwd(mix_d, src_d, len_d)
            graphed = True

        def run_step():
            return fwd(mix_d, src_d, len_d)

        if with_loss:
            def e2e_step():
                return fwd(mix_h.to(dev, non_blocking=True) if not use_graph else mix_h, src_h.to(dev, non_blocking=True),
                           len_h.to(dev, non_blocking=True)).item()
            h2d, d2h = mix_h.numel() * 4 + src_h.numel() * 4 + len_h.numel() * 8, 4
        else:
            est_h = torch.empty(M, kw["C"], T, dtype=torch.float32).pin_memory()

            def e2e_step():  # separate.py's use: mixture up, separated sources down
                est = fwd(mix_h if use_graph else mix_h.to(dev, non_blocking=True), None, None)
                est_h.copy_(est, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return float(est_h[0, 0, 0])
            h2d, d2h = mix_h.numel() * 4, est_h.numel() * 4

    # ---- device-resident throughput -------------------------------------------------------------------
    for _ in range(args.warmup):
        run_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = L.ctn_launch_count()
    secs = timed(run_step, args.steps)
    launches = (launches_per_step * args.steps) if graphed else (L.ctn_launch_count() - n0)
    audio = world * M * T / SR * args.steps
    value = audio / secs
    if args.profile_only:
        sampler.stop()
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": secs / args.steps * 1e3, "gpu_launches": int(launches)}))
        return

    # ---- end to end through the public API with host inputs -------------------------------------------
    last = None
    for _ in range(2):
        last = e2e_step()

    def e2e_timed():
        nonlocal last
        last = e2e_step()
    secs_e2e = timed(e2e_timed, args.steps)
    clocks = sampler.stop()

    # ---- the metric's other half for the training configs: forward-only throughput of the same batch ---
    fwd_line = None
    if train:
        model.eval()
        with torch.no_grad():
            inf = GraphedInference(model) if use_graph else model
            for _ in range(3):
                inf(mix_d)
            secs_fwd = timed(lambda: inf(mix_d), args.steps)
        model.train()
        fwd_line = {"value": audio / secs_fwd, "unit": "audio-s/s", "ms_per_step": secs_fwd / args.steps * 1e3}

    # ---- roofline of the dominant kernel, timed live on its own launches with the L2 flushed in between ----------
    pk, pk_src = peaks()
    K = L.ctn_num_frames(ctypes.byref(model._cfg), T)
    F = M * K
    Bc, Hc = kw["B"], kw["H"]
    flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)  # >> L2 (126 MB); also keeps the GPU busy while the
    # host enqueues the timed launch, so the event pair brackets the kernel alone and not host launch latency
    st = _lib.stream()
    if train:  # weight gradient dW[H,B] = dz1[F,H]^T x[F,B] (2 per TemporalBlock + 2), the largest kernel of the step
        Gm = torch.randn(F, Hc, device=dev)
        Xm = torch.randn(F, Bc, device=dev)
        dWm = torch.zeros(Hc, Bc, device=dev)

        def kern():
            _lib.check(L.ctn_wgrad(Gm.data_ptr(), Xm.data_ptr(), dWm.data_ptr(), F, Hc, Bc, K, None, None, None, None,
                                   None, st))
        kname, ksub = ("tc_wgrad_kernel<256>: dW[512,256] = dz1[F,512]^T x[F,256] (tcgen05, bf16x3 split, MN-major "
                       "operands, split-K + coalesced red.v4)"), "tc_wgrad_kernel"
        n_launch = 2 * kw["R"] * kw["X"] + 2
    else:      # the 1x1 bottleneck conv of every TemporalBlock: z1[F,H] = x[F,B] W1^T with the gLN sums in the epilogue
        Am = torch.randn(F, Bc, device=dev)
        Wm = torch.randn(Hc, Bc, device=dev) / 16
        hi = Wm.to(torch.bfloat16)  # pre-split bf16 hi/lo planes, as the model holds them for inference
        lo = (Wm - hi.float()).to(torch.bfloat16)
        Dm = torch.empty(F, Hc, device=dev)

        def kern():
            _lib.check(L.ctn_conv1x1_planes(Am.data_ptr(), hi.data_ptr(), lo.data_ptr(), 0, Dm.data_ptr(), F, Hc, Bc, K,
                                            st))
        # (timed in the bf16x3 flavour for every forward config: the bf16 path's single-plane conv is its cheaper sibling)
        kname, ksub = ("1x1-conv GEMM z1[F,512] = x[F,256] W1^T (tcgen05, bf16x3 split of the inference forward, "
                       "pre-split weight planes)"), "gemm_kernel"
        n_launch = 2 * kw["R"] * kw["X"] + 2
    for _ in range(3):
        kern()
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kern()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) * 1e-3
    t_k = tot / reps
    flops = 2.0 * F * Bc * Hc
    # algorithmic bytes of one launch: both operands read once, the output written once (weight gradient: fp32 [H, B];
    # forward conv: fp32 [F, H] out, [F, B] in)
    alg_bytes = 4.0 * F * (Bc + Hc) + (4.0 * Hc * Bc if train else 0.0)
    traffic, traffic_src = profiled_traffic(ksub)
    # which roofline binds this kernel: its arithmetic intensity (B H / (2 (B + H)) = 85 flop/B at the paper widths)
    # against the machine's ridge (measured bf16 peak / measured copy bandwidth = ~250 flop/B) -> the HBM roof.  The
    # tensor-pipe figure is kept next to it (the x3 operand split issues 3 MMAs per algorithmic MAC: its ceiling is 1/3).
    t_hbm, t_tensor = alg_bytes / (pk["hbm_gbs"] * 1e9), flops / (pk["bf16_tflops"] * 1e12)
    tensor_view = {"bound": "tensor", "achieved": flops / t_k / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                   "frac": flops / t_k / 1e12 / pk["bf16_tflops"],
                   "note": "bf16 burst peak; the bf16x3 split issues 3 MMAs per algorithmic MAC, so 1/3 is this fraction's ceiling"}
    hbm_view = {"bound": "hbm", "achieved": alg_bytes / t_k / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / t_k / 1e9 / pk["hbm_gbs"]}
    main, other = (hbm_view, tensor_view) if t_hbm >= t_tensor else (tensor_view, hbm_view)
    roofline = {"kernel": kname, **main, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk_src + (", copy bandwidth (burst)" if main["bound"] == "hbm" else ", bf16 burst"),
                "binding": f"arithmetic intensity {flops / alg_bytes:.0f} flop/B vs ridge "
                           f"{pk['bf16_tflops'] * 1e3 / pk['hbm_gbs']:.0f} flop/B: bound time hbm {t_hbm * 1e6:.2f} us, "
                           f"tensor {t_tensor * 1e6:.2f} us",
                "other_roof": other, "launch_us": t_k * 1e6, "alg_flops_per_launch": flops,
                "alg_bytes_per_launch": alg_bytes, "launches_per_step": n_launch,
                "share_of_step": n_launch * t_k / (secs / args.steps)}

    # whole-step algorithmic rates (SURVEY §8d per-frame figures x frames)
    frames = world * F
    N_, L_, C_, RX = kw["N"], kw["L"], kw["C"], kw["R"] * kw["X"]
    S_ = L_ // 2
    n_params = 8710720 + (C_ - 2) * N_ * Bc
    mac_fwd = N_ * L_ + N_ * Bc + RX * (Bc * Hc + Hc * kw["P"] + Hc * Bc) + Bc * C_ * N_ + C_ * N_ * L_
    e_fwd = S_ + 3 * N_ + 2 * Bc + C_ * S_ + RX * (3 * Bc + 4 * Hc)
    e_bwd = S_ + C_ * S_ + 5 * Bc + 8 * N_ + 2 * C_ * N_ + RX * (5 * Bc + 12 * Hc)
    if train:
        step_flops = frames * (2 * mac_fwd * 3 - 2 * N_ * L_)
        step_bytes = frames * (e_fwd + e_bwd + 5 * C_ * S_) * 4 + 10 * 4 * n_params * world
    else:
        step_flops = frames * 2 * mac_fwd
        esz = 2 if dtype == "bf16" else 4  # SURVEY §8d counts the bf16 config at 2 bytes per element (a LOWER bound on the
        # bytes this path moves: it keeps the residual stream, the encoder / decoder and all I/O in fp32)
        step_bytes = frames * (e_fwd + (2 * C_ * S_ if mode == "fwd_loss" else 0)) * esz + 4 * n_params * world
    t_step = secs / args.steps

    line = {"metric": METRIC[mode], "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config_dict(args.config, mode, world),
            "e2e": {"value": audio / secs_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": secs_e2e / args.steps * 1e3, "last_result": last},
            "gpu_launches": int(launches), "cuda_graph": graphed, "clocks": clocks, "roofline": roofline,
            "step_algorithmic": {"tflops": step_flops / t_step / 1e12, "gbs": step_bytes / t_step / 1e9,
                                 "hbm_bound_ms": step_bytes / world / (pk["hbm_gbs"] * 1e9) * 1e3,
                                 "frac_of_hbm_bound": (step_bytes / world / (pk["hbm_gbs"] * 1e9)) / t_step}}
    if fwd_line is not None:
        line["fwd"] = fwd_line
    if world > 1 and train:
        peer = getattr(dp, "peer_active", lambda: False)()
        line["exchange"] = ("one-kernel peer-memory all-reduce over NVLink inside the step's CUDA graph (csrc/peer_reduce.cu)"
                            if peer else "bucketed NCCL all-reduce between per-stage CUDA graphs")
        if peer and dp._peer.error():
            line["exchange_error"] = "a rank did not arrive within the kernel's timeout"

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, dt, sample, cores = reference_steps(args.config, mode, 2, 1, budget_s=30.0)
        line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
