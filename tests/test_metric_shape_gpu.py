"""GPU parity of the BENCHMARKED path (precision="tf32", use_tcgen05=True: CTA-pair tcgen05 convolutions with full-halo
boxes, the fused temporal / spatial-linear blocks, the sliding-window stem, S=4 sub-tiling) at the BENCHMARKED shape, against
  (1) strided subsamples of the UNMODIFIED reference's outputs and stage activations (tests/golden/metric_shape.npz), and
  (2) the CPU oracle evaluated here on the same seeded inputs (full tensors, every element).
Also: one teacher-forced p_sample at the metric shape, the other BASELINE.json shapes (jellyfish 20x128x128 7->4 channels,
smoke 64 frames x 128x128) at batch 1 against the oracle, and a teacher-forced TF32-mode DDPM loop on the reference's trace.

Tolerance (contraction class, TF32 operands / fp32 accumulate, the reference's own cuDNN numerics class): 3e-3 of the
tensor's max magnitude.  Measured errors are appended to gpurun_out/parity_r2.json when that directory exists."""
import json
import os

import numpy as np
import pytest
import torch

import diffphycon_b200 as dpc
from oracle import smoke_sampler_oracle as so
from oracle import unet3d_oracle as uo

pytestmark = pytest.mark.gpu

TOL_TF32 = 3e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report(key, value):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "parity_r2.json")
    data = json.load(open(p)) if os.path.exists(p) else {}
    data[key] = value
    json.dump(data, open(p, "w"), indent=1, sort_keys=True)


def metric_input(channels, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 32, channels, 64, 64, generator=g)


def engine_net(cfg_kw, seed, precision="tf32"):
    net = dpc.Unet3D_with_Conv3D(**cfg_kw)
    net.load_state_dict(uo.make_params(uo.UnetCfg(**cfg_kw), seed), strict=True)
    net.precision = precision
    net.use_tcgen05 = True
    return net.cuda()


@pytest.mark.parametrize("tag,channels,seed", [("joint", 6, 31), ("prior", 2, 33)])
def test_metric_shape_forward_vs_reference_and_oracle(tag, channels, seed, golden_dir):
    z = np.load(os.path.join(golden_dir, "metric_shape.npz"))
    kw = dict(dim=64, dim_mults=(1, 2, 4), channels=channels)
    net = engine_net(kw, seed)
    net.taps = {}
    x = metric_input(channels, seed + 1)
    t = torch.tensor([int(z[f"{tag}/t"])])
    y = net(x.cuda(), t.cuda()).cpu()
    errs = {}
    # (1) the unmodified reference, subsampled
    ref = torch.from_numpy(z[f"{tag}/y"])
    errs["y_vs_reference"] = (y[:, ::8, :, ::8, ::8] - ref).abs().max().item() / max(1.0, float(z[f"{tag}/y_absmax"]))
    for k in z.files:
        if k.startswith(f"{tag}/act/"):
            nm = k.split("/", 2)[2]
            a = torch.from_numpy(z[k])
            got = net.taps[nm].cpu()[:, :, ::8, ::8, ::8]
            errs["act_" + nm] = (got - a).abs().max().item() / max(1.0, float(z[f"{tag}/absmax/{nm}"]))
    # (2) the oracle, every element
    otaps = {}
    yo = uo.forward(uo.make_params(uo.UnetCfg(**kw), seed), uo.UnetCfg(**kw), x, t, taps=otaps)
    errs["y_vs_oracle_full"] = (y - yo).abs().max().item() / max(1.0, yo.abs().max().item())
    for nm, a in otaps.items():
        if nm in net.taps and a.dim() == 5:
            errs["full_" + nm] = (net.taps[nm].cpu() - a).abs().max().item() / max(1.0, a.abs().max().item())
    report(f"metric_shape_forward/{tag}", errs)
    bad = {k: v for k, v in errs.items() if v > TOL_TF32}
    assert not bad, bad


def test_metric_shape_p_sample_teacher_forced(golden_dir):
    """One p_sample (smoke.py:671-699) at [1,32,6,64,64], t = 500, on the engine (tf32 / tcgen05) against the unmodified
    reference's recorded result.  x_start = sr*x - srm1*eps: an eps error d moves x_start (before the 1-Lipschitz clamp) by
    srm1*d and x_{t-1} by coef1*srm1*d, so the bounds are TOL_TF32 * eps_scale * srm1 (* coef1) + 1e-5."""
    z = np.load(os.path.join(golden_dir, "metric_shape.npz"))
    t = int(z["p_sample/t"])
    mj = engine_net(dict(dim=64, dim_mults=(1, 2, 4), channels=6), 31)
    mw = engine_net(dict(dim=64, dim_mults=(1, 2, 4), channels=2), 33)
    diff = dpc.GaussianDiffusion([mj, mw], image_size=64, frames=32, timesteps=1000, sampling_timesteps=1000, loss_type="l2",
                                 objective="pred_noise", standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True,
                                 w_prob_exp=0.97).cuda()
    init = torch.from_numpy(z["p_sample/init"])
    x = metric_input(6, 78)
    x[:, 0, 0] = init
    torch.manual_seed(1234 + t)
    noise = torch.randn(1, 32, 6, 64, 64)
    diff.sample_noise = lambda shape, device: noise.to(device)
    pred, x_start = diff.p_sample(x.shape, x.cuda(), t, None, design_fn=dpc.StockSmokeGuidance(), design_guidance="standard",
                                  init=init.cuda(), _impose_init=True)
    sch = diff._sched()
    srm1 = float(sch["sqrt_recipm1_alphas_cumprod"][t])
    c1 = float(sch["posterior_mean_coef1"][t])
    eps_scale = 3.0   # |eps| of the seeded nets at this shape stays below 3 (y_absmax in the golden)
    assert float(z["joint/y_absmax"]) <= eps_scale
    e_xs = (x_start.cpu()[:, ::8, :, ::8, ::8] - torch.from_numpy(z["p_sample/x_start"])).abs().max().item()
    e_pr = (pred.cpu()[:, ::8, :, ::8, ::8] - torch.from_numpy(z["p_sample/pred"])).abs().max().item()
    report("metric_shape_p_sample", dict(t=t, x_start_err=e_xs, pred_err=e_pr, srm1=srm1, coef1=c1,
                                         bound_x_start=TOL_TF32 * eps_scale * srm1 + 1e-5,
                                         bound_pred=TOL_TF32 * eps_scale * srm1 * c1 + 1e-5))
    assert e_xs <= TOL_TF32 * eps_scale * srm1 + 1e-5, e_xs
    assert e_pr <= TOL_TF32 * eps_scale * srm1 * c1 + 1e-5, e_pr
    assert torch.equal(pred[:, 0, 0].cpu(), init)


@pytest.mark.parametrize("name,frames,size,channels,out_dim", [("jellyfish_20x128", 20, 128, 7, 4),
                                                                ("smoke_64x128", 64, 128, 6, None)])
def test_other_baseline_shapes_vs_oracle(name, frames, size, channels, out_dim):
    """BASELINE.json configs 3 and 5 at batch 1: the tf32 / tcgen05 path against the CPU oracle, every element."""
    kw = dict(dim=64, dim_mults=(1, 2, 4), channels=channels)
    if out_dim is not None:
        kw["out_dim"] = out_dim
    cfg = uo.UnetCfg(**kw)
    params = uo.make_params(cfg, 41)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, frames, channels, size, size, generator=g)
    t = torch.tensor([654])
    net = engine_net(kw, 41)
    net.taps = {}
    y = net(x.cuda(), t.cuda()).cpu()
    taps = {k: v.cpu() for k, v in net.taps.items()}
    net.taps = None
    otaps = {}
    yo = uo.forward(params, cfg, x, t, taps=otaps)
    errs = {"y": (y - yo).abs().max().item() / max(1.0, yo.abs().max().item())}
    for nm, a in otaps.items():
        if nm in taps and a.dim() == 5:
            errs[nm] = (taps[nm] - a).abs().max().item() / max(1.0, a.abs().max().item())
    report(f"other_shapes/{name}", errs)
    bad = {k: v for k, v in errs.items() if v > TOL_TF32}
    assert not bad, bad


def test_ddpm_loop_teacher_forced_tf32(golden_dir):
    """The reference's 4-step DDPM loop (tests/golden/sampler_loop_ddpm4_trace.npz: state entering each step, noise drawn
    in it), one engine p_sample per step in the BENCHMARKED precision mode.  Per-step bound as above:
    coef1(t) * srm1(t) * (TOL_TF32 * eps_scale) + 1e-5, with eps_scale the joint net's measured |eps| maximum at that step
    (the clamp of x_start to [-1, 1] is 1-Lipschitz, so it can only shrink the error)."""
    z = np.load(os.path.join(golden_dir, "sampler_loop_ddpm4_trace.npz"))
    cj, cw = dict(dim=32, dim_mults=(1, 2), channels=6), dict(dim=32, dim_mults=(1, 2), channels=2)
    mj, mw = engine_net(cj, 11), engine_net(cw, 12)
    diff = dpc.GaussianDiffusion([mj, mw], image_size=16, frames=4, timesteps=4, sampling_timesteps=4,
                                 standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True, w_prob_exp=0.97).cuda()
    sch = diff._sched()
    init = torch.from_numpy(z["init"]).cuda()
    rows = {}
    for t in (3, 2, 1, 0):
        x = torch.from_numpy(z[f"x{t}"]).cuda()
        if t > 0:
            noise = torch.from_numpy(z[f"z{t}"])
            diff.sample_noise = lambda shape, device, _n=noise: _n.to(device)
        out, _ = diff.p_sample(x.shape, x, t, None, design_fn=dpc.StockSmokeGuidance(), design_guidance="standard", init=init,
                               _impose_init=True)
        ref = torch.from_numpy(z[f"x{t - 1}"] if t > 0 else z["y"])
        tt = torch.full((x.shape[0],), t, device="cuda", dtype=torch.long)
        eps_scale = max(1.0, mj(x, tt).abs().max().item())
        amp = float(sch["posterior_mean_coef1"][t]) * float(sch["sqrt_recipm1_alphas_cumprod"][t])
        err = (out.cpu() - ref).abs().max().item()
        rows[f"t{t}"] = dict(err=err, bound=TOL_TF32 * eps_scale * amp + 1e-5, amp=amp, eps_scale=eps_scale)
        assert err <= TOL_TF32 * eps_scale * amp + 1e-5, (t, err, amp)
    report("ddpm4_teacher_forced_tf32", rows)


def test_graph_replay_equals_eager_at_metric_shape():
    """The benchmark replays one captured CUDA graph per denoising step (bench.py strong mode).  At the benchmarked architecture and
    shape (dim 64, (1,2,4), 32 frames of 64x64, tcgen05 kernels, fused attention blocks) the graph loop must reproduce the eager
    loop bit for bit on the same seed: same kernels, same arithmetic, same noise stream."""
    from diffphycon_b200 import _lib
    torch.manual_seed(0)
    mj = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=6)
    mw = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=2)
    d = dpc.GaussianDiffusion([mj, mw], image_size=64, frames=32, timesteps=3, sampling_timesteps=3, eval_2ddpm=True,
                              standard_fixed_ratio=1e5, coeff_ratio=0, w_prob_exp=0.97).cuda()
    init = torch.rand(2, 64, 64, generator=torch.Generator().manual_seed(5)).cuda() / 2
    fn = dpc.StockSmokeGuidance()
    torch.manual_seed(11)
    ref = d.sample(batch_size=2, design_fn=fn, init=init)
    d.use_cuda_graph = True
    n0 = _lib.LaunchCounter.graph_launches
    torch.manual_seed(11)
    y = d.sample(batch_size=2, design_fn=fn, init=init)
    assert _lib.LaunchCounter.graph_launches - n0 == 3
    assert torch.isfinite(y).all() and torch.equal(y, ref), (y - ref).abs().max().item()
