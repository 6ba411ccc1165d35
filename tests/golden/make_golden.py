"""Generate the golden vectors under tests/golden/ by executing the UNMODIFIED reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference is imported through oracle/shims (five missing third-party packages).
Weights are the deterministic synthetic weights of oracle.unet3d_oracle.make_params (rebuilt anywhere from the key
names), loaded into the reference module with load_state_dict(strict=True) — which also pins the key/shape inventory.
What is stored: inputs, outputs and a few intermediate activations, as float32 .npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import unet3d_oracle as uo  # noqa: E402
from oracle.smoke_sampler_oracle import SMOKE_RESCALER  # noqa: E402

UNET_CASES = {
    # name: (cfg kwargs, batch, frames, size, times, seed)
    "unet_small_c6": (dict(dim=32, dim_mults=(1, 2), channels=6), 2, 4, 16, (3, 900), 1),
    "unet_small_c2": (dict(dim=32, dim_mults=(1, 2), channels=2), 1, 5, 16, (499,), 2),
    "unet_smoke_arch": (dict(dim=64, dim_mults=(1, 2, 4), channels=6), 1, 6, 16, (777,), 3),
    "unet_jelly_arch": (dict(dim=32, dim_mults=(1, 2), channels=7, out_dim=4), 1, 4, 16, (12,), 4),
}


def build_ref_unet(cfg: uo.UnetCfg, seed: int):
    m = ref_import.unet3d_module()
    net = m.Unet3D_with_Conv3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, out_dim=cfg.out_dim)
    params = uo.make_params(cfg, seed)
    ref_keys = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(v.shape) for k, v in params.items()}
    assert list(ref_keys.keys()) == list(mine.keys()), "state_dict key order differs from oracle.param_shapes"
    assert ref_keys == mine, "state_dict shapes differ from oracle.param_shapes"
    net.load_state_dict(params, strict=True)
    net.eval()
    return net, params


def gen_unet():
    for name, (kw, b, f, s, times, seed) in UNET_CASES.items():
        cfg = uo.UnetCfg(**kw)
        net, _ = build_ref_unet(cfg, seed)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(b, f, cfg.channels, s, s, generator=g)
        t = torch.tensor(times, dtype=torch.long)
        acts = {}
        hooks = []

        def mk(nm):
            def hook(_m, _i, o):
                acts[nm] = o.detach().clone()
            return hook
        hooks.append(net.init_conv.register_forward_hook(mk("init_conv")))
        hooks.append(net.init_temporal_attn.register_forward_hook(mk("init_temporal_attn")))
        hooks.append(net.downs[0][0].register_forward_hook(mk("downs.0.0")))
        hooks.append(net.downs[0][2].register_forward_hook(mk("downs.0.2")))
        hooks.append(net.downs[0][3].register_forward_hook(mk("downs.0.3")))
        hooks.append(net.downs[0][4].register_forward_hook(mk("downs.0.4")))
        hooks.append(net.mid_spatial_attn.register_forward_hook(mk("mid_spatial_attn")))
        hooks.append(net.mid_block2.register_forward_hook(mk("mid_block2")))
        hooks.append(net.ups[0][4].register_forward_hook(mk("ups.0.4")))
        with torch.no_grad():
            y = net(x, t)
        for h in hooks:
            h.remove()
        out = dict(x=x.numpy(), t=t.numpy(), y=y.numpy(), seed=np.int64(seed))
        if name == "unet_small_c6":  # intermediate activations for one case only (keeps the fixtures small)
            for k, v in acts.items():
                out["act/" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "y", tuple(y.shape), float(y.abs().mean()))


def ref_guidance_fn(x, RESCALER, w_energy=0.0):
    """The stock guidance of inference/inference_2d_smoke.py:30-44 (that file cannot be imported: it needs
    matplotlib/imageio/accelerate and data at import time), reproduced for the golden run."""
    from torch.autograd import grad
    x = x * RESCALER
    state = x
    guidance_success = state[:, -1, -1].mean((-1, -2)).sum()
    guidance_energy = state[:, :, 3:5].square().mean((1, 2, 3, 4)).sum()
    guidance = -guidance_success + w_energy * guidance_energy
    grad_x = grad(guidance, x, grad_outputs=torch.ones_like(guidance))[0]
    return grad_x


def gen_sampler():
    d = ref_import.smoke_diffusion_module()
    cj = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=6)
    cw = uo.UnetCfg(dim=32, dim_mults=(1, 2), channels=2)
    mj, _ = build_ref_unet(cj, 11)
    mw, _ = build_ref_unet(cw, 12)
    R = torch.tensor(SMOKE_RESCALER).reshape(1, 1, 6, 1, 1)
    B, Fr, S = 2, 4, 16
    g = torch.Generator().manual_seed(7)
    init = torch.rand(B, S, S, generator=g) / 2.0

    # --- teacher-forced single p_sample steps --------------------------------------------------------------
    for tag, guidance, w_energy, sfr, cr, gamma in (
        ("std", "standard", 0.0, 1e5, 0.0, 0.97),
        ("alpha", "standard-alpha", 0.5, 0.01, 0.1, 0.9),
    ):
        diff = d.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=1000, sampling_timesteps=1000,
                                   loss_type="l2", objective="pred_noise", standard_fixed_ratio=sfr, coeff_ratio=cr,
                                   eval_2ddpm=True, w_prob_exp=gamma)

        def design_fn(x, low=None, init=None, init_u=None, _w=w_energy):
            return ref_guidance_fn(x, R, w_energy=_w)

        out = dict(init=init.numpy(), w_energy=np.float64(w_energy), standard_fixed_ratio=np.float64(sfr),
                   coeff_ratio=np.float64(cr), w_prob_exp=np.float64(gamma))
        for t in (999, 500, 1, 0):
            x = torch.randn(B, Fr, 6, S, S, generator=g)
            x[:, 0, 0] = init
            tt = torch.full((B,), t, dtype=torch.long)
            with torch.no_grad():
                ej = mj(x, tt)
                ew = mw(x[:, :, 3:5], tt)
            torch.manual_seed(1234 + t)
            pred, x_start = diff.p_sample((B, Fr, 6, S, S), x.clone(), t, None, design_fn=design_fn,
                                          design_guidance=guidance)
            pred[:, 0, 0] = init  # p_sample_loop re-imposes the condition (smoke.py:720)
            torch.manual_seed(1234 + t)
            z = torch.randn(B, Fr, 6, S, S)
            out.update({f"t{t}/x": x.numpy(), f"t{t}/eps_joint": ej.numpy(), f"t{t}/eps_w": ew.numpy(),
                        f"t{t}/z": z.numpy(), f"t{t}/pred": pred.numpy(), f"t{t}/x_start": x_start.numpy()})
        np.savez_compressed(os.path.join(HERE, f"sampler_step_{tag}.npz"), **out)
        print("sampler_step", tag)

    # --- schedule buffers ----------------------------------------------------------------------------------
    sch = {}
    for name in ("sigmoid", "cosine", "linear"):
        for T in (1000, 200):
            diff = d.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=T, beta_schedule=name,
                                       eval_2ddpm=True)
            for k, v in diff.state_dict().items():
                if not k.startswith("model") and k != "loss_weight":
                    sch[f"{name}{T}/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), **sch)

    # --- whole loops: DDPM with T=4, DDIM 1000 -> 3 steps, eta=1 ---------------------------------------------
    def design_fn0(x, low=None, init=None, init_u=None):
        return ref_guidance_fn(x, R, w_energy=0.0)

    diff = d.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=4, sampling_timesteps=4,
                               standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True, w_prob_exp=0.97)
    torch.manual_seed(42)
    y = diff.sample(batch_size=B, design_fn=design_fn0, design_guidance="standard", init=init)
    np.savez_compressed(os.path.join(HERE, "sampler_loop_ddpm4.npz"), init=init.numpy(), y=y.numpy())
    diff = d.GaussianDiffusion([mj, mw], image_size=S, frames=Fr, timesteps=1000, sampling_timesteps=3,
                               ddim_sampling_eta=1.0, standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True,
                               w_prob_exp=0.97)
    # record the state entering every DDIM step and the noise drawn in it, so that parity tests can teacher-force
    # single steps (whole-loop comparisons compound rounding differences by sqrt(1/abar - 1) ~ 1.8e3)
    trace = {}
    orig_mp = diff.model_predictions

    def rec_mp(shape, x, t, *a, **k):
        trace[f"x{len([q for q in trace if q.startswith('x')])}"] = x.detach().clone().numpy()
        return orig_mp(shape, x, t, *a, **k)
    diff.model_predictions = rec_mp
    orig_rl = torch.randn_like

    def rec_rl(t, **k):
        n = orig_rl(t, **k)
        trace[f"z{len([q for q in trace if q.startswith('z')])}"] = n.detach().clone().numpy()
        return n
    torch.randn_like = rec_rl
    torch.manual_seed(43)
    y = diff.sample(batch_size=B, design_fn=design_fn0, design_guidance="standard", init=init)
    torch.randn_like = orig_rl
    np.savez_compressed(os.path.join(HERE, "sampler_loop_ddim3.npz"), init=init.numpy(), y=y.numpy(), **trace)
    print("sampler loops done")


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["unet", "sampler"]
    if "unet" in which:
        gen_unet()
    if "sampler" in which:
        gen_sampler()
