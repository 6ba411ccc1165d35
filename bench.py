#!/usr/bin/env python
"""bench.py — denoising steps/sec on the 2-D smoke 64x64x32-frame configuration at batch 64 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one iteration of GaussianDiffusion.p_sample_loop over a batch of 64 trajectories: joint U-Net forward,
prior U-Net forward, guidance + prior re-weighting + posterior update + re-imposed initial condition, including the
torch.randn noise draw.  Weak scaling: every rank owns its own batch of 64 trajectories (they never
interact); `value` = batch-64 steps/s summed over ranks, timed on the device, max over ranks.
Synthetic inputs and seeded default-init weights (no checkpoints/datasets are reachable).

--impl reference times the reference's algorithm on the host CPU cores: the oracle port (oracle/*.py — plain PyTorch
fp32 restatement pinned to the unmodified reference by tests/golden) on a bounded sample of the same workload.
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
    v, sec, cores = cpu_reference_steps_per_s(args.steps, args.warmup)
    sample = f"1 of 64 trajectories at the metric shape, {args.warmup} warm-up + {args.steps} timed steps, linear extrapolation to batch 64"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 64.0 * sec * 1e3, "higher_is_better": True, "scaling": "weak",
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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="trajectories per GPU (the metric is quoted at 64)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "3xtf32"])
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-tcgen05", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rollout", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3        # bounded sample: ~8 s of CPU work per step per trajectory
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
    B = args.batch
    torch.manual_seed(0)
    mj = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=2)
    for m in (mj, mw):
        m.precision = args.precision
        m.use_tcgen05 = not args.no_tcgen05
        m.micro_batch = args.micro_batch or None
    diff = dpc.GaussianDiffusion([mj, mw], image_size=SIZE, frames=FRAMES, timesteps=1000, sampling_timesteps=1000,
                                 loss_type='l2', objective='pred_noise', standard_fixed_ratio=1e5, coeff_ratio=0,
                                 eval_2ddpm=True, w_prob_exp=0.97).to(dev)
    design_fn = dpc.StockSmokeGuidance(dpc.SMOKE_RESCALER, w_energy=0.0)
    torch.manual_seed(1234 + rank)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, SIZE), torch.linspace(-1, 1, SIZE), indexing="ij")
    blob = torch.exp(-((xx - 0.1) ** 2 + (yy + 0.2) ** 2) / 0.1)
    init = (blob[None].repeat(B, 1, 1) / 2.0).to(dev).contiguous()
    shape = (B, FRAMES, CH, SIZE, SIZE)
    x = torch.randn(shape, device=dev)
    x[:, 0, 0] = init

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(xc, t):
        xn, _ = diff.p_sample(shape, xc, t, None, design_fn=design_fn, design_guidance="standard", init=init,
                              _impose_init=True)
        return xn

    t_cur = 999
    for _ in range(args.warmup):
        x = step(x, t_cur)
        t_cur -= 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LaunchCounter.count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x = step(x, t_cur)
        t_cur -= 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.LaunchCounter.count - launches0
    sampler.stop_flag = True
    barrier()
    ms_t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    assert torch.isfinite(x).all(), "non-finite state after the timed steps"
    ms_per_step = ms_max / args.steps
    value = world * (B / 64.0) / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hout = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        hinit = torch.empty(init.shape, dtype=torch.float32, pin_memory=True)
        hinit.copy_(init)
        n_e2e = max(2, min(args.steps, 3))
        barrier()
        e0.record()
        for i in range(n_e2e):
            xd = hx.to(dev, non_blocking=True)
            idv = hinit.to(dev, non_blocking=True)
            xn, _ = diff.p_sample(shape, xd, t_cur - i, None, design_fn=design_fn, design_guidance="standard", init=idv,
                                  _impose_init=True)
            hout.copy_(xn, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hx, hout = hout, hx
        e1.record()
        torch.cuda.synchronize()
        ms_e = torch.tensor([e0.elapsed_time(e1) / n_e2e], device=dev)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        nbytes = x.numel() * 4
        e2e = {"value": world * (B / 64.0) / (float(ms_e.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": nbytes + init.numel() * 4, "d2h_bytes_per_step": nbytes, "steps": n_e2e}

    # ---- dominant kernel (3x3x3 conv 64->64 at 32x64x64, the most frequent layer) timed alone on its stream ----
    roof = None
    if rank == 0:
        from diffphycon_b200 import packing
        pk = peaks()
        Bc = min(B, 16)
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
        e0.record()
        for _ in range(reps):
            _lib.conv(p, tcgen05=not args.no_tcgen05)
        e1.record()
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        flops = 2.0 * Bc * FRAMES * SIZE * SIZE * 64 * 64 * 27
        ach = flops / (kms / 1e3) / 1e12
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_dominant_kernel_traffic.json")))
            if used_tc and tj.get("kernel_batch") == Bc:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except (OSError, ValueError, KeyError):
            traffic = None
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tensor_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor_burst"], "frac_of_tf32_roof": ach / (pk["tensor_burst"] / 2), "traffic": traffic,
                "traffic_note": "bytes per launch (ncu, profiles/r1_ncu_full_v12.txt); algorithmic = %d" % (2 * Bc * FRAMES * SIZE * SIZE * 64 * 4),
                "kernel": "conv3d_tcgen05 3x3x3 64->64" if used_tc else "conv_igemm (mma.sync) 3x3x3 64->64",
                "kernel_ms": kms, "kernel_batch": Bc, "peak_source": pk["source"] + " bf16 burst (kernel timed alone); TF32 nominal peak is half of bf16",
                "step_tensor_tflops": FLOPS_PER_SAMPLE_STEP * B / (ms_per_step / 1e3) / 1e12,
                "step_tensor_frac_of_sustained_bf16": FLOPS_PER_SAMPLE_STEP * B / (ms_per_step / 1e3) / 1e12 / pk["tensor_sustained"],
                "step_hbm_algorithmic_gbs": ALGO_BYTES_PER_SAMPLE_STEP * B / (ms_per_step / 1e3) / 1e9,
                "step_hbm_frac": ALGO_BYTES_PER_SAMPLE_STEP * B / (ms_per_step / 1e3) / 1e9 / pk["hbm"]}

    # ---- the single collective of the path: all-gather of the sampled controls [B,32,2,64,64] per rank (untimed step) ----
    gather = None
    if world > 1:
        ctrl_local = x[:, :, 3:5].contiguous()
        bucket = torch.empty(world * ctrl_local.shape[0], *ctrl_local.shape[1:], dtype=ctrl_local.dtype, device=dev)
        dist.all_gather_into_tensor(bucket, ctrl_local)       # warm-up (NCCL communicator setup)
        barrier()
        e0.record()
        dist.all_gather_into_tensor(bucket, ctrl_local)
        e1.record()
        torch.cuda.synchronize()
        g_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(g_ms, op=dist.ReduceOp.MAX)
        gather = {"collective": "all_gather(sampled controls)", "bytes_per_rank": ctrl_local.numel() * 4,
                  "ms": float(g_ms.item()), "once_per_sampling_run": True}

    # ---- post-sampling rollout of the sampled controls, informational ----
    rollout = None
    if rank == 0 and not args.no_rollout:
        from diffphycon_b200 import smoke_rollout as sr
        sim = sr.init_sim_128()
        ctrl = (x[:, :, 3:5] * torch.tensor([16.0, 20.0], device=dev).view(1, 1, 2, 1, 1)).contiguous()
        c1, c2 = ctrl[:, :, 0].contiguous(), ctrl[:, :, 1].contiguous()
        dens = (x[:, 0, 0] * 2.0).clamp(min=0).contiguous()
        sr.solver_batch(sim, sr.init_velocity_(), dens[:1], c1[:1, :4].contiguous(), c2[:1, :4].contiguous(), 8)
        torch.cuda.synchronize()
        e0.record()
        ro = sr.solver_batch(sim, sr.init_velocity_(), dens, c1, c2, 256)
        e1.record()
        torch.cuda.synchronize()
        rms = e0.elapsed_time(e1)
        rollout = {"trajectories": B, "frames": 256, "ms_total": rms, "trajectories_per_s": B / (rms / 1e3),
                   "mean_cg_iterations": float(ro["iterations"][:, 1:].float().mean()),
                   "note": "one persistent CTA per trajectory, fp64 CG with the reference's 500-iteration cap; the "
                           "reference needs ~48 s per trajectory on one CPU core (SURVEY.md section 6)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_steps_per_s(2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "1 of 64 trajectories at the metric shape, 1 warm-up + 2 timed steps (%.1f s/step), linear extrapolation to batch 64" % sec}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32(3xtf32)", "data": "synthetic",
            "config": {"workload": "smoke 64x64x32 frames, DDPM p_sample step (joint+prior Unet3D dim64 (1,2,4), stock guidance, posterior)",
                       "per_gpu_batch": B, "global_batch": B * world, "sharding": "independent trajectories per rank, no per-step collective",
                       "l2": "inputs larger than L2 (activations are GBs per layer)", "precision": args.precision,
                       "tcgen05": not args.no_tcgen05, "micro_batch": args.micro_batch or None},
            "clocks": sampler.summary(), "gpu_launches": launches, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "rollout": rollout, "gather": gather,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
