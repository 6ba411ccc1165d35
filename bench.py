#!/usr/bin/env python
"""bench.py — denoising steps/sec on the 2-D smoke 64x64x32-frame configuration at batch 64 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]
                    [--config metric|smoke16+rollout|jellyfish128|smoke256x8|smoke128x64-ddim|burgers] [--precision tf32|3xtf32]

A "step" is one iteration of GaussianDiffusion.p_sample_loop over the batch: joint U-Net forward, prior U-Net forward,
guidance + prior re-weighting + posterior update + re-imposed initial condition, including the torch.randn noise draw
(SURVEY.md 8(d)).  Trajectories never interact, so ranks own disjoint slices of the batch:

  --scaling strong (default; BASELINE.json: "batch 64 at 1/2/4/8 B200"): the GLOBAL batch is 64, each rank
      samples 64/N trajectories through diffphycon_b200.distributed.sample_sharded on a K-step schedule — global noise stream
      sliced per rank (an N-rank run reproduces the 1-rank trajectories), the single NCCL all-gather of the sampled controls
      INSIDE the timed region.  `value` = batch-64 steps/s.  At N > 1 the line also carries `weak` (64 per rank, no gather).
  --scaling weak: every rank owns 64 trajectories; `value` = batch-64 steps/s summed over ranks.

Timed on the device (CUDA events), max over ranks.  Synthetic inputs and seeded default-init weights (no checkpoints/datasets
are reachable).  --config selects the other BASELINE.json configurations (their own metric names; see CONFIGS).

--impl reference times the reference's algorithm on the host CPU cores: the oracle port (oracle/*.py — plain PyTorch fp32
restatement pinned to the unmodified reference by tests/golden) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoising steps/sec (2D smoke 64x64x32, batch 64)"
UNIT = "steps/s"
FRAMES, SIZE, CH = 32, 64, 6
FLOPS_PER_SAMPLE_STEP = 1.7945e12     # two U-Net forwards (908.8 + 885.7 GFLOP) per trajectory
ALGO_BYTES_PER_SAMPLE_STEP = 7.93e9   # fused fp32 algorithmic minimum per trajectory

CONFIGS = {
    # name: (metric, BASELINE.json config it restates)
    "metric": (METRIC, "metric row M"),
    "smoke16+rollout": ("denoising steps/sec (2D smoke 64x64x32, batch 16) + phi rollout of the sampled controls", "configs[1]"),
    "jellyfish128": ("denoising steps/sec (2D jellyfish 128x128x20, batch 8, joint + prior + surrogate-net guidance)", "configs[2]"),
    "smoke256x8": ("denoising steps/sec (2D smoke 64x64x32, batch 256 over 8 GPUs = 32 per GPU)", "configs[3]"),
    "smoke128x64-ddim": ("DDIM steps/sec (2D smoke 128x128x64, batch 64 over 8 GPUs = 8 per GPU, eta 1, guided)", "configs[4]"),
    "burgers": ("denoising steps/sec (1D Burgers FOPC 16x128, two-model DDPM-200, batch 4)", "configs[0]"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_steps_per_s(steps, warmup, batch_cpu=1):
    """The reference algorithm on the host cores (oracle port), bounded sample: `batch_cpu` trajectories of the metric
    shape; returns (batch-64 steps/s extrapolated linearly, measured seconds per step at batch_cpu, cores)."""
    from oracle import smoke_sampler_oracle as so
    from oracle import unet3d_oracle as uo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cj = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=6)
    cw = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=2)
    pj, pw = uo.make_params(cj, 0), uo.make_params(cw, 1)
    sched = so.make_schedule(1000, "sigmoid")
    R = torch.tensor(so.SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch_cpu, FRAMES, CH, SIZE, SIZE, generator=g)
    init = torch.rand(batch_cpu, SIZE, SIZE, generator=g) / 2
    x[:, 0, 0] = init
    times = []
    t = 999
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tt = torch.full((batch_cpu,), t, dtype=torch.long)
        ej = uo.forward(pj, cj, x, tt)
        ew = uo.forward(pw, cw, x[:, :, 3:5], tt)
        z = torch.randn(x.shape, generator=g)
        x, _ = so.p_sample_step(sched, x, t, ej, ew, z, init, lambda v: so.guidance_fn(v, R, 0.0),
                                design_guidance="standard", standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        t -= 1
    sec = sum(times) / len(times)
    return (batch_cpu / sec) / 64.0, sec, cores


def run_reference(args, rank):
    if rank != 0:
        return
    v, sec, cores = cpu_reference_steps_per_s(args.steps, args.warmup, batch_cpu=4)
    sample = f"4 of 64 trajectories at the metric shape, {args.warmup} warm-up + {args.steps} timed steps, linear extrapolation to batch 64"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 16.0 * sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "smoke 64x64x32 frames, batch 64, DDPM p_sample step (2 U-Nets + guidance + posterior)",
                   "note": "CPU reference arm does not use the GPUs; value is the host-core rate"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Everything any library prints on fd 1 during the run (e.g. NCCL's version banner) goes to stderr; the one JSON line is
    written to the original stdout at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def measure_tf32_peak(dev):
    """cuBLAS TF32 GEMM 8192^3 (2*N^3 FLOP), best of 10 — the recipe MEASURED_PEAKS.json uses for bf16, in the arithmetic type the
    convolutions compute in.  Library call, used only as the roofline denominator."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def time_region(fn, barrier, dist, dev, world):
    """CUDA-event time of fn() in ms, bracketed by barrier + synchronize, max over ranks."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


def smoke_nets(dpc, args, dev):
    torch.manual_seed(0)
    mj = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=2)
    for m in (mj, mw):
        m.precision = args.precision
        m.use_tcgen05 = not args.no_tcgen05
        m.micro_batch = args.micro_batch or None
    return mj.to(dev), mw.to(dev)


def smoke_diffusion(dpc, nets, dev, frames, size, timesteps=1000, sampling_timesteps=None, eta=0.0):
    return dpc.GaussianDiffusion(list(nets), image_size=size, frames=frames, timesteps=timesteps,
                                 sampling_timesteps=timesteps if sampling_timesteps is None else sampling_timesteps,
                                 loss_type='l2', objective='pred_noise', standard_fixed_ratio=1e5, coeff_ratio=0,
                                 eval_2ddpm=True, w_prob_exp=0.97, ddim_sampling_eta=eta).to(dev)


def blob_init(B, size, dev):
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    blob = torch.exp(-((xx - 0.1) ** 2 + (yy + 0.2) ** 2) / 0.1)
    return (blob[None].repeat(B, 1, 1) / 2.0).to(dev).contiguous()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="metric", choices=list(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--batch", type=int, default=0, help="override the batch (metric: global batch for strong, per GPU for weak)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "3xtf32"])
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-tcgen05", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rollout", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="strong mode replays one captured CUDA graph per denoising step (GaussianDiffusion.use_cuda_graph, captured "
                         "before the timed region; bit-identical to the eager loop, tests/test_sampler_gpu.py); this flag uses the "
                         "eager per-kernel launch loop instead")
    ap.add_argument("--profile", action="store_true", help="after the timed region, print a per-launch-category CUDA-event breakdown "
                                                           "of one more step to stderr (development aid)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 2:
            args.steps = 2        # bounded sample: ~8-10 s of CPU work per 4-trajectory step
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank)
        return

    import diffphycon_b200 as dpc
    from diffphycon_b200 import _lib
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config != "metric":
        run_other_config(args, dpc, _lib, dist, dev, rank, world, local_rank, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    from diffphycon_b200.distributed import sample_sharded, shard_bounds
    K, W = args.steps, max(args.warmup, 0)
    strong = args.scaling == "strong"
    Bg = args.batch or 64                                   # strong: global batch; weak: per-rank batch
    lo, hi = shard_bounds(Bg, world, rank) if strong else (0, Bg)
    B = hi - lo                                             # trajectories this rank owns
    nets = smoke_nets(dpc, args, dev)
    diff = smoke_diffusion(dpc, nets, dev, FRAMES, SIZE)
    design_fn = dpc.StockSmokeGuidance(dpc.SMOKE_RESCALER, w_energy=0.0)
    torch.manual_seed(1234 + rank)
    init_g = blob_init(Bg if strong else B, SIZE, dev)      # strong: the GLOBAL init, identical on every rank
    init = init_g[lo:hi].contiguous() if strong else init_g
    shape = (B, FRAMES, CH, SIZE, SIZE)
    x = torch.randn(shape, device=dev)
    x[:, 0, 0] = init

    def step(xc, t, ini=init, shp=shape):
        xn, _ = diff.p_sample(shp, xc, t, None, design_fn=design_fn, design_guidance="standard", init=ini, _impose_init=True)
        return xn

    t_cur = 999
    for _ in range(W):
        x = step(x, t_cur)
        t_cur -= 1
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LaunchCounter.count
    gather = None
    g0 = 0
    if strong:
        # the public multi-GPU call: a K-step schedule through sample_sharded, global noise stream, one all-gather of the controls
        diff_k = smoke_diffusion(dpc, nets, dev, FRAMES, SIZE, timesteps=K)
        if world > 1:                                       # NCCL communicator / gather buffers are set up outside the timed region
            warm = torch.zeros(shard_bounds(Bg, world, 0)[1], FRAMES, 2, SIZE, SIZE, device=dev)
            bucket = torch.empty(world * warm.shape[0], *warm.shape[1:], device=dev)
            dist.all_gather_into_tensor(bucket, warm)
            del warm, bucket
        if args.cuda_graph:                                 # capture (warm-up + one captured step) happens outside the timed region
            diff_k.use_cuda_graph = True
            if os.environ.get("DPC_TWO_STREAMS"):            # development A/B: force the two U-Nets onto two captured streams (1) or one (0)
                diff_k.two_streams = os.environ["DPC_TWO_STREAMS"] == "1"
            sample_sharded(diff_k, Bg, design_fn=design_fn, design_guidance="standard", init=init_g, global_noise=True,
                           gather_channels=slice(3, 5))
            launches0 = _lib.LaunchCounter.count
        g0 = _lib.LaunchCounter.graph_launches
        torch.manual_seed(4321)                             # same generator state on every rank: global noise, sliced per rank
        ms, ctrl = time_region(lambda: sample_sharded(diff_k, Bg, design_fn=design_fn, design_guidance="standard", init=init_g,
                                                      global_noise=True, gather_channels=slice(3, 5)),
                               barrier, dist, dev, world)
        assert ctrl.shape[0] == Bg and torch.isfinite(ctrl).all(), "non-finite sampled controls"
        gather = {"collective": "all_gather_into_tensor(sampled controls), inside the timed region" if world > 1 else "none (1 rank)",
                  "bytes_per_rank": int(ctrl.numel() // max(world, 1)) * 4, "once_per_sampling_run": True}
        ms_per_step = ms / K
        value = (Bg / 64.0) / (ms_per_step / 1e3)
    else:
        def loop():
            nonlocal x, t_cur
            for _ in range(K):
                x = step(x, t_cur)
                t_cur -= 1
        ms, _ = time_region(loop, barrier, dist, dev, world)
        assert torch.isfinite(x).all(), "non-finite state after the timed steps"
        ms_per_step = ms / K
        value = world * (B / 64.0) / (ms_per_step / 1e3)
    launches = _lib.LaunchCounter.count - launches0
    sampler.stop_flag = True

    # ---- secondary: weak scaling (64 trajectories per rank, no collective) when the headline is strong and N > 1 ----
    weak = None
    if strong and world > 1:
        Bw = 64
        xw = torch.randn(Bw, FRAMES, CH, SIZE, SIZE, device=dev)
        iw = blob_init(Bw, SIZE, dev)
        xw[:, 0, 0] = iw
        shw = tuple(xw.shape)
        xw = step(xw, 500, iw, shw)

        def loop_w():
            nonlocal xw
            for i in range(K):
                xw = step(xw, 499 - i, iw, shw)
        ms_w, _ = time_region(loop_w, barrier, dist, dev, world)
        weak = {"per_gpu_batch": Bw, "ms_per_step": ms_w / K, "value": world * (Bw / 64.0) / (ms_w / K / 1e3), "unit": UNIT}
        del xw

    # ---- end-to-end through the public API with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hout = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        hinit = torch.empty(init.shape, dtype=torch.float32, pin_memory=True)
        hinit.copy_(init)
        n_e2e = max(10, min(K, 20))
        state = {"hx": hx, "hout": hout}

        def loop_e2e():
            for i in range(n_e2e):
                xd = state["hx"].to(dev, non_blocking=True)
                idv = hinit.to(dev, non_blocking=True)
                xn, _ = diff.p_sample(shape, xd, 900 - i, None, design_fn=design_fn, design_guidance="standard", init=idv,
                                      _impose_init=True)
                state["hout"].copy_(xn, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                state["hx"], state["hout"] = state["hout"], state["hx"]
        ms_e, _ = time_region(loop_e2e, barrier, dist, dev, world)
        nbytes = x.numel() * 4
        units = (Bg / 64.0) if strong else world * (B / 64.0)
        e2e = {"value": units / (ms_e / n_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": nbytes + init.numel() * 4,
               "d2h_bytes_per_step": nbytes, "steps": n_e2e,
               "note": "per-rank pinned host state copied in and out around every p_sample call"}

    roof = None
    if rank == 0 and not args.no_roofline:
        roof = dominant_kernel_roofline(args, _lib, dev, ms_per_step, Bg if strong else B * world)

    # ---- post-sampling rollout of the sampled controls, informational ----
    rollout = None
    if rank == 0 and not args.no_rollout:
        rollout = run_rollout(dpc, x, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU path's throughput still grows with the batch at 1-2 trajectories (threads idle in the small layers), so the
        # baseline is quoted on 4 trajectories and the 1- and 2-trajectory rates are kept beside it
        v4, sec4, cores = cpu_reference_steps_per_s(1, 1, batch_cpu=4)
        v1, sec1, _ = cpu_reference_steps_per_s(1, 1, batch_cpu=1)
        v2, sec2, _ = cpu_reference_steps_per_s(1, 1, batch_cpu=2)
        cpu = {"value": v4, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "4 of 64 trajectories at the metric shape, 1 warm-up + 1 timed step (%.1f s/step), linear extrapolation to "
                         "batch 64; 1 trajectory: %.1f s/step, 2 trajectories: %.1f s/step" % (sec4, sec1, sec2),
               "value_from_batch1": v1, "value_from_batch2": v2}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32(3xtf32)", "data": "synthetic",
            "config": {"workload": "smoke 64x64x32 frames, DDPM p_sample step (joint+prior Unet3D dim64 (1,2,4), stock guidance, posterior)",
                       "per_gpu_batch": B, "global_batch": Bg if strong else B * world,
                       "sharding": ("global batch 64 sliced over ranks (sample_sharded), one all-gather of the controls per sampling run"
                                    if strong else "independent trajectories per rank, no collective"),
                       "l2": "inputs larger than L2 (activations are GBs per layer)", "precision": args.precision,
                       "tcgen05": not args.no_tcgen05, "micro_batch": args.micro_batch or None,
                       "cuda_graph": bool(args.cuda_graph and strong),
                       "graph_launches": (_lib.LaunchCounter.graph_launches - g0) if strong else 0},
            "clocks": sampler.summary(), "gpu_launches": launches, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
            "rollout": rollout, "gather": gather, "weak": weak,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(args, _lib, dev, ms_per_step, B_global):
    """The dominant kernel (3x3x3 conv 64->64 at 32x64x64, the most frequent layer) timed alone on its stream."""
    from diffphycon_b200 import packing
    pk = peaks()
    Bc = 16
    xa = torch.randn(Bc, FRAMES, SIZE, SIZE, 64, device=dev)
    w = torch.randn(64, 64, 3, 3, 3, device=dev) / (27 * 64) ** 0.5
    wp, _, _ = packing.pack_conv3d(w)
    bias = torch.zeros(64, device=dev)
    taps = packing.tap_table(3, 3, 3, SIZE, SIZE, dev)
    y = torch.empty_like(xa)
    stats = torch.zeros(Bc, 8, 2, dtype=torch.float64, device=dev)
    p = _lib.ConvParams()
    p.x1, p.C1, p.C2 = xa.data_ptr(), 64, 0
    p.w, p.bias, p.y, p.taps, p.ntaps = wp.data_ptr(), bias.data_ptr(), y.data_ptr(), taps.data_ptr(), 27
    p.gn_stats, p.gn_groups = stats.data_ptr(), 8
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = Bc, FRAMES, SIZE, SIZE, FRAMES, SIZE, SIZE
    p.st = p.sh = p.sw = 1
    p.pt = p.ph = p.pw = 1
    p.oh_mul = p.ow_mul = 1
    p.Hfull, p.Wfull = SIZE, SIZE
    p.Cout, p.Npad, p.Kpad = 64, wp.shape[0], wp.shape[1]
    used_tc = False
    for _ in range(3):
        used_tc = _lib.conv(p, tcgen05=not args.no_tcgen05)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.conv(p, tcgen05=not args.no_tcgen05)
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / reps
    flops = 2.0 * Bc * FRAMES * SIZE * SIZE * 64 * 64 * 27
    ach = flops / (kms / 1e3) / 1e12
    del xa, y
    tf32_peak = measure_tf32_peak(dev)
    traffic, traffic_src = None, None   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture
    for name in ("r2_dominant_kernel_traffic.json", "r1_dominant_kernel_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if used_tc and tj.get("kernel_batch") == Bc:
                traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], name
                break
        except (OSError, ValueError, KeyError):
            continue
    step_tflops = FLOPS_PER_SAMPLE_STEP * B_global / (ms_per_step / 1e3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": pk["tensor_burst"], "unit": "TFLOP/s",
            "frac": ach / pk["tensor_burst"], "tf32_peak_measured": tf32_peak, "frac_of_tf32_peak_measured": ach / tf32_peak,
            "traffic": traffic,
            "traffic_note": "bytes per launch (ncu --set full, profiles/%s); algorithmic = %d" % (traffic_src, 2 * Bc * FRAMES * SIZE * SIZE * 64 * 4),
            "kernel": "conv3d_tcgen05 3x3x3 64->64" if used_tc else "conv_igemm (mma.sync) 3x3x3 64->64",
            "kernel_ms": kms, "kernel_batch": Bc,
            "peak_source": pk["source"] + " bf16 burst (kernel timed alone); tf32_peak_measured = cuBLAS TF32 GEMM 8192^3, best of 10, this run",
            "step_tensor_tflops": step_tflops,
            "step_tensor_frac_of_sustained_bf16": step_tflops / pk["tensor_sustained"],
            "step_tensor_frac_of_tf32_peak_measured": step_tflops / tf32_peak,
            "step_hbm_algorithmic_gbs": ALGO_BYTES_PER_SAMPLE_STEP * B_global / (ms_per_step / 1e3) / 1e9,
            "step_hbm_frac": ALGO_BYTES_PER_SAMPLE_STEP * B_global / (ms_per_step / 1e3) / 1e9 / pk["hbm"]}


def run_rollout(dpc, x, dev, frames=256):
    from diffphycon_b200 import smoke_rollout as sr
    B = x.shape[0]
    sim = sr.init_sim_128()
    ctrl = (x[:, :, 3:5] * torch.tensor([16.0, 20.0], device=dev).view(1, 1, 2, 1, 1)).contiguous()
    c1, c2 = ctrl[:, :, 0].contiguous(), ctrl[:, :, 1].contiguous()
    dens = (x[:, 0, 0] * 2.0).clamp(min=0).contiguous()
    sr.solver_batch(sim, sr.init_velocity_(), dens[:1], c1[:1, :4].contiguous(), c2[:1, :4].contiguous(), 8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro = sr.solver_batch(sim, sr.init_velocity_(), dens, c1, c2, frames)
    e1.record()
    torch.cuda.synchronize()
    rms = e0.elapsed_time(e1)
    return {"trajectories": B, "frames": frames, "ms_total": rms, "trajectories_per_s": B / (rms / 1e3),
            "mean_cg_iterations": float(ro["iterations"][:, 1:].float().mean()),
            "note": "a thread-block cluster of 2 or 4 CTAs per trajectory, fp64 CG with the reference's 500-iteration cap; the "
                    "reference needs ~48 s per trajectory on one CPU core (SURVEY.md section 6)"}


# ------------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations
# ------------------------------------------------------------------------------------------------------------------------
def run_other_config(args, dpc, _lib, dist, dev, rank, world, local_rank, barrier):
    K, W = args.steps, max(args.warmup, 0)
    metric, which = CONFIGS[args.config]
    sampler = ClockSampler(local_rank)
    extra = {}
    design_fn = dpc.StockSmokeGuidance(dpc.SMOKE_RESCALER, w_energy=0.0)
    torch.manual_seed(1234 + rank)
    if args.config in ("smoke16+rollout", "smoke256x8", "smoke128x64-ddim"):
        frames, size = (64, 128) if args.config == "smoke128x64-ddim" else (FRAMES, SIZE)
        B = args.batch or {"smoke16+rollout": 16, "smoke256x8": 32, "smoke128x64-ddim": 8}[args.config]
        nets = smoke_nets(dpc, args, dev)
        ddim = args.config == "smoke128x64-ddim"
        diff = smoke_diffusion(dpc, nets, dev, frames, size, sampling_timesteps=250 if ddim else None, eta=1.0 if ddim else 0.0)
        init = blob_init(B, size, dev)
        shape = (B, frames, CH, size, size)
        x = torch.randn(shape, device=dev)
        x[:, 0, 0] = init
        if ddim:
            times = list(reversed(torch.linspace(-1, 999, steps=251).int().tolist()))
            pairs = list(zip(times[:-1], times[1:]))

            def step(xc, i):
                return diff.ddim_step(xc, pairs[i][0], pairs[i][1], design_fn=design_fn, design_guidance="standard", init=init)
        else:
            def step(xc, i):
                return diff.p_sample(shape, xc, 999 - i, None, design_fn=design_fn, design_guidance="standard", init=init,
                                     _impose_init=True)[0]
        units_per_step = world * B      # trajectories advanced per step over all ranks
        flops_step = (FLOPS_PER_SAMPLE_STEP if size == 64 else (7359.4e9 + 7174.7e9)) * B
        workload = f"smoke {size}x{size}x{frames} frames, {'DDIM-250 eta 1' if ddim else 'DDPM'} step, {B} trajectories per GPU"
    elif args.config == "jellyfish128":
        from diffphycon_b200 import diffusion_2d_jellyfish as dj
        B, frames, size = args.batch or 8, 20, 128
        torch.manual_seed(0)
        mj = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=7, out_dim=4)
        mw = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=7, out_dim=1)
        fm = dpc.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
        bd = dpc.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
        for m in (mj, mw, fm, bd):
            m.precision = args.precision
            m.use_tcgen05 = not args.no_tcgen05
            m.to(dev)
        diff = dj.GaussianDiffusion([mj, mw], image_size=size, frames=frames, cond_steps=1, timesteps=1000, sampling_timesteps=1000,
                                    loss_type='l2', objective='pred_noise', coeff_ratio_J=0.3, coeff_ratio_w=0.3, eval_2ddpm=True,
                                    w_prob_exp=0.7, use_guidance_in_model_predictions=False).to(dev)
        guidance = dpc.JellyfishGuidance(fm, bd, p_min=-1.5, p_max=2.5, reg_ratio=1000.0)
        torch.manual_seed(1234 + rank)
        state_0 = torch.rand(B, 3, size, size, device=dev) * 2 - 1
        bd_0 = torch.cat([(torch.rand(B, 1, size, size, device=dev) > 0.7).float(), torch.rand(B, 2, size, size, device=dev) - 0.5], 1)
        thetas_0 = torch.rand(B, device=dev) * 0.7 + 0.2
        st = diff._begin((B, frames, 7, size, size), [state_0, bd_0], thetas_0, bd)

        def step(_x, i):
            diff._ddpm_step(st, 999 - i, guidance, "standard-alpha")
            return st.x
        x = st.x
        units_per_step = world * B
        # per sample-step at 64^2 2 x 570 + 292 + ~3 x (292 + 83) GF, x 4 at 128^2
        flops_step = 4 * (2 * 570e9 + 292e9 + 3 * (292e9 + 83e9)) * B
        workload = (f"jellyfish {size}x{size}x{frames}, DDPM step: joint + prior Unet3D (7->4, 7->1), force_fn guidance = ForceUnet + "
                    f"boundary-updater Unet forward AND backward on the engine, update_bd forward; {B} trajectories per GPU")
    elif args.config == "burgers":
        from diffphycon_b200 import diffusion_1d_burgers as db
        from diffphycon_b200.burgers_unet import Unet2D
        from diffphycon_b200.burgers import burgers_numeric_solve_free
        B, T = args.batch or 4, 200
        torch.manual_seed(0)
        uw = Unet2D(dim=64, out_dim=2, dim_mults=(1, 2, 4), channels=2, resnet_block_groups=1).to(dev)
        w_ = Unet2D(dim=32, out_dim=2, dim_mults=(1, 2, 4, 8), channels=2, resnet_block_groups=1).to(dev)
        uw.precision = w_.precision = args.precision
        diff = db.GaussianDiffusion((uw, w_), seq_length=(16, 128), timesteps=T, auto_normalize=False, use_conv2d=True, temporal=True,
                                    is_condition_u0=True, is_condition_uT=True, eval_two_models=True, prior_beta=1.5).to(dev)
        diff.use_cuda_graph = bool(args.cuda_graph)     # network forwards of a step replayed from one captured graph
        xs = torch.linspace(0, 1, 130, device=dev)[1:-1]
        u0 = torch.exp(-((xs[None] - torch.rand(B, 1, device=dev)) ** 2) * 50) - 0.5 * torch.exp(-((xs[None] - torch.rand(B, 1, device=dev)) ** 2) * 80)
        f0 = torch.zeros(B, 10, 128, device=dev)
        target = burgers_numeric_solve_free(u0, f0, visc=0.01, T=1.0, dt=1e-4, num_t=10)
        kw = dict(nablaJ=db.get_nablaJ(lambda xx: torch.zeros(xx.shape[0], device=xx.device) + 0.0 * xx.sum((1, 2, 3))),
                  J_scheduler=db.cosine_beta_J_schedule, w_scheduler=db.sigmoid_schedule_flip, u_init=target[:, 0] / 10,
                  u_final=target[:, 10] / 10)
        x = None

        def run_all(_x, i):        # one "step" call = one full 200-step sampling run + the finite-difference rollout of its controls
            y = diff.sample(batch_size=B, clip_denoised=True, guidance_u0=True, **kw)
            extra["rollout_traj"] = burgers_numeric_solve_free(u0, (y[:, 1, :10] * 10).contiguous(), visc=0.01, T=1.0, dt=1e-4, num_t=10)
            return y
        step = run_all
        units_per_step = None
        flops_step = 32.1e9 * B / 4        # 24.05 + 8.04 GFLOP per B = 4 step
        workload = f"Burgers FOPC two-model DDPM, {T} steps, [{B},2,16,128], + 10 000-step finite-difference rollout per run"
    else:
        raise AssertionError(args.config)

    for i in range(W):
        x = step(x, i)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LaunchCounter.count

    def loop():
        nonlocal x
        for i in range(K):
            x = step(x, W + i)
    ms, _ = time_region(loop, barrier, dist, dev, world)
    launches = _lib.LaunchCounter.count - launches0
    sampler.stop_flag = True
    assert torch.isfinite(x).all()
    if args.profile and rank == 0:
        with _lib.Profiler() as prof:
            x = step(x, W + K)
        summ = prof.summary()
        tot = sum(t for _, t in summ.values())
        for k, (n, t) in sorted(summ.items(), key=lambda kv: -kv[1][1])[:45]:
            sys.stderr.write(f"{t:9.2f} ms {100 * t / tot:5.1f}% n={n:4d} {k}\n")
        sys.stderr.write(f"profiled step total {tot:.2f} ms\n")
    if args.config == "burgers":
        ms_per_step = ms / (K * 200)
        value = 1e3 / ms_per_step
    else:
        ms_per_step = ms / K
        value = 1e3 / ms_per_step          # steps/s of the config's whole (global) batch
    rollout = None
    if args.config == "smoke16+rollout" and rank == 0 and not args.no_rollout:
        rollout = run_rollout(dpc, x, dev)
    pk = peaks()
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32(3xtf32)", "data": "synthetic",
                "config": {"workload": workload, "baseline_config": which, "global_batch": units_per_step, "precision": args.precision,
                           "tcgen05": not args.no_tcgen05},
                "clocks": sampler.summary(), "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": flops_step / (ms_per_step / 1e3) / 1e12, "peak": pk["tensor_sustained"],
                             "unit": "TFLOP/s", "frac": flops_step / (ms_per_step / 1e3) / 1e12 / pk["tensor_sustained"], "traffic": None,
                             "kernel": "whole step of ONE GPU (SURVEY.md 8(d) FLOP counts per trajectory x per-GPU batch)",
                             "peak_source": pk["source"] + " bf16 sustained"},
                "rollout": rollout, "e2e": None, "cpu_baseline": None}
        _emit(line)


if __name__ == "__main__":
    main()
