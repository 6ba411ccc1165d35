"""Golden vectors AT THE BENCHMARKED SHAPE (smoke 64x64, 32 frames, dim 64, mults (1,2,4)) from the UNMODIFIED reference.

Run in the build container only:  python tests/golden/make_golden_metric_shape.py
Inputs are regenerated from seeds (torch CPU generator), so only strided SUBSAMPLES of the outputs / intermediate
activations are committed (a full [1,32,6,64,64] output is 3 MB): every 8th frame and every 8th pixel in both directions,
all channels — 4 x 8 x 8 points per channel from every region of the volume, which is what a layout / halo / tiling bug
at this shape would corrupt.  Stored:
  metric_shape.npz   : joint net (6 ch) and prior net (2 ch) forward at [1,32,C,64,64], t = 321 / 877, with stage taps;
                       one teacher-forced p_sample (smoke.py:671-699) at t = 500 through the reference GaussianDiffusion
  sampler_loop_ddpm4_trace.npz : the 4-step DDPM loop of sampler_loop_ddpm4.npz re-run with the state entering each
                       step and the noise drawn in it recorded, for teacher-forced TF32-mode parity
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_import  # noqa: E402
from oracle import unet3d_oracle as uo  # noqa: E402
from oracle.smoke_sampler_oracle import SMOKE_RESCALER  # noqa: E402
from make_golden import build_ref_unet, ref_guidance_fn  # noqa: E402

SEED_J, SEED_W = 31, 33
TAPS = ("init_conv", "init_temporal_attn", "downs.0.0", "downs.0.2", "downs.0.3", "downs.0.4", "downs.1.3", "mid_spatial_attn",
        "mid_block2", "ups.0.4", "ups.1.3", "ups.2.3")


def sub(a: torch.Tensor) -> np.ndarray:
    """[B,F,C,H,W] state -> every 8th frame / row / column, all channels."""
    return a[:, ::8, :, ::8, ::8].contiguous().numpy()


def sub_act(a: torch.Tensor) -> np.ndarray:
    """NCDHW activation -> every 8th frame / row / column, all channels."""
    return a[:, :, ::8, ::8, ::8].contiguous().numpy()


def metric_inputs(channels, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 32, channels, 64, 64, generator=g)
    return x


def gen_forward(out):
    for tag, ch, seed, t in (("joint", 6, SEED_J, 321), ("prior", 2, SEED_W, 877)):
        cfg = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=ch)
        net, _ = build_ref_unet(cfg, seed)
        x = metric_inputs(ch, seed + 1)
        tt = torch.tensor([t])
        acts, hooks = {}, []

        def mk(nm):
            def hook(_m, _i, o):
                acts[nm] = o.detach().clone()
            return hook
        for nm in TAPS:
            mod = net
            for part in nm.split("."):
                mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
            hooks.append(mod.register_forward_hook(mk(nm)))
        with torch.no_grad():
            y = net(x, tt)
        for h in hooks:
            h.remove()
        out[f"{tag}/t"] = np.int64(t)
        out[f"{tag}/y"] = sub(y)                       # [1,4,C,8,8]
        out[f"{tag}/y_absmax"] = np.float32(y.abs().max())
        for nm, a in acts.items():                     # NCDHW -> [1,C,4,h/8,w/8]
            out[f"{tag}/act/{nm}"] = sub_act(a)
            out[f"{tag}/absmax/{nm}"] = np.float32(a.abs().max())
        print(tag, "forward done", float(y.abs().mean()))


def gen_p_sample(out):
    d = ref_import.smoke_diffusion_module()
    mj, _ = build_ref_unet(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=6), SEED_J)
    mw, _ = build_ref_unet(uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=2), SEED_W)
    R = torch.tensor(SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    diff = d.GaussianDiffusion([mj, mw], image_size=64, frames=32, timesteps=1000, sampling_timesteps=1000, loss_type="l2",
                               objective="pred_noise", standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True,
                               w_prob_exp=0.97)

    def design_fn(x, low=None, init=None, init_u=None):
        return ref_guidance_fn(x, R, w_energy=0.0)
    g = torch.Generator().manual_seed(77)
    init = torch.rand(1, 64, 64, generator=g) / 2.0
    x = metric_inputs(6, 78)
    x[:, 0, 0] = init
    t = 500
    torch.manual_seed(1234 + t)
    pred, x_start = diff.p_sample((1, 32, 6, 64, 64), x.clone(), t, None, design_fn=design_fn, design_guidance="standard")
    pred[:, 0, 0] = init
    out["p_sample/t"] = np.int64(t)
    out["p_sample/init"] = init.numpy()
    out["p_sample/pred"] = sub(pred.detach())
    out["p_sample/x_start"] = sub(x_start.detach())
    print("p_sample done")


def gen_ddpm_trace():
    d = ref_import.smoke_diffusion_module()
    mj, _ = build_ref_unet(uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6), 11)
    mw, _ = build_ref_unet(uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2), 12)
    R = torch.tensor(SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    B, Fr, S = 2, 4, 16
    g = torch.Generator().manual_seed(7)
    init = torch.rand(B, S, S, generator=g) / 2.0

    def design_fn0(x, low=None, init=None, init_u=None):
        return ref_guidance_fn(x, R, w_energy=0.0)
    diff = d.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=4, sampling_timesteps=4,
                               standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True, w_prob_exp=0.97)
    trace = {}
    orig_ps = diff.p_sample

    def rec_ps(shape, x, t, *a, **k):
        trace[f"x{t}"] = x.detach().clone().numpy()
        return orig_ps(shape, x, t, *a, **k)
    diff.p_sample = rec_ps
    orig_randn = torch.randn
    n_draw = [0]

    def rec_randn(*a, **k):
        n = orig_randn(*a, **k)
        trace[f"draw{n_draw[0]}"] = n.detach().clone().numpy()
        n_draw[0] += 1
        return n
    torch.randn = rec_randn
    try:
        torch.manual_seed(42)
        y = diff.sample(batch_size=B, design_fn=design_fn0, design_guidance="standard", init=init)
    finally:
        torch.randn = orig_randn
    # draw0 = the initial state; draw1.. = the noise of steps t = 3, 2, 1 (t = 0 draws none, smoke.py:684)
    out = dict(init=init.numpy(), y=y.numpy())
    for t, i in ((3, 1), (2, 2), (1, 3)):
        out[f"z{t}"] = trace[f"draw{i}"]
    for t in (3, 2, 1, 0):
        out[f"x{t}"] = trace[f"x{t}"]
    prev = np.load(os.path.join(HERE, "sampler_loop_ddpm4.npz"))
    assert np.array_equal(prev["y"], out["y"]), "trace run differs from the committed loop golden"
    np.savez_compressed(os.path.join(HERE, "sampler_loop_ddpm4_trace.npz"), **out)
    print("ddpm trace done", n_draw[0], "draws")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    which = sys.argv[1:] or ["metric", "trace"]
    if "metric" in which:
        out = {}
        gen_forward(out)
        gen_p_sample(out)
        np.savez_compressed(os.path.join(HERE, "metric_shape.npz"), **out)
    if "trace" in which:
        gen_ddpm_trace()
