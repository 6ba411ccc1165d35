"""Golden trace of the Burgers two-model guided DDPM sampler from the UNMODIFIED reference
diffusion/diffusion_1d_burgers.py + model/burgers_1d/unet.py (build container only): per-step inputs of p_sample, the
noise drawn in it and its output, plus the final sample."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import param_gen, ref_import  # noqa: E402

um = ref_import.burgers_unet_module()
dm = ref_import.burgers_diffusion_module()
KW_UW = dict(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=1)
KW_W = dict(dim=32, dim_mults=(1, 2, 4), channels=2, resnet_block_groups=1)


def build(kw, seed):
    net = um.Unet2D(**kw)
    net.load_state_dict(param_gen.make_params({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed), strict=True)
    return net.eval()


def run(name, models, guidance_u0, **dkw):
    T, B = 6, 2
    d = dm.GaussianDiffusion(models, seq_length=(16, 128), timesteps=T, auto_normalize=False, use_conv2d=True, temporal=True,
                             is_condition_u0=True, is_condition_uT=True, **dkw)
    g = torch.Generator().manual_seed(5)
    u_init, u_final = torch.rand(B, 128, generator=g) - 0.5, torch.rand(B, 128, generator=g) - 0.5
    target = torch.rand(B, 128, generator=g) - 0.5

    def loss_fn(x):
        return (x[:, 0, 10, :] - target).square().mean(-1) + 0.05 * x[:, 1, :10, :].square().mean((-1, -2))

    trace, calls = {}, []
    orig_ps = d.p_sample

    def rec_ps(x, t, *a, **k):
        tag = f"{t}" if "pred_noise" not in k else f"{t}b"
        first = f"x{tag}" not in trace                  # with recurrence p_sample runs recurrence_k times per t: keep the first pass
        if first:
            trace[f"x{tag}"] = x.detach().clone().numpy()
        out = orig_ps(x, t, *a, **k)
        if first:
            trace[f"pred{tag}"] = out[0].detach().clone().numpy()
            trace[f"xstart{tag}"] = out[1].detach().clone().numpy()
        return out

    d.p_sample = rec_ps
    orig_rl = torch.randn_like

    def rec_rl(t_, **k):
        n = orig_rl(t_, **k)
        trace[f"z{len([q for q in trace if q.startswith('z')])}"] = n.detach().clone().numpy()
        return n

    torch.randn_like = rec_rl
    torch.manual_seed(77)
    y = d.sample(batch_size=B, clip_denoised=True, nablaJ=dm.get_nablaJ(loss_fn),
                 J_scheduler=lambda t: 0.5 * dm.cosine_beta_J_schedule(t), w_scheduler=dm.sigmoid_schedule_flip,
                 guidance_u0=guidance_u0, u_init=u_init, u_final=u_final)
    torch.randn_like = orig_rl
    np.savez_compressed(os.path.join(HERE, name + ".npz"), u_init=u_init.numpy(), u_final=u_final.numpy(),
                        target=target.numpy(), y=y.detach().numpy(), **trace)
    print(name, sorted(trace), tuple(y.shape), float(y.abs().mean()))


uw, w = build(KW_UW, 31), build(KW_W, 32)
VARIANTS = {
    "burgers_sampler": lambda: run("burgers_sampler", (uw, w), True, eval_two_models=True, prior_beta=1.5),
    "burgers_sampler_single_ut": lambda: run("burgers_sampler_single_ut", uw, False),
    "burgers_sampler_model_w": lambda: run("burgers_sampler_model_w", w, True, is_model_w=True, prior_beta=0.7),
    # self recurrence (burgers.py:472-482, :535-578): two passes per diffusion step, re-noised after each
    "burgers_sampler_recurrent": lambda: run("burgers_sampler_recurrent", (uw, w), True, eval_two_models=True, prior_beta=1.5,
                                             recurrence=True, recurrence_k=2),
}
for name in (sys.argv[1:] or list(VARIANTS)):
    VARIANTS[name]()
