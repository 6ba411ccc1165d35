"""Checkpoint adapter: loads the files the reference's Trainer writes — `model-{n}.pt`,
diffusion/diffusion_2d_smoke.py:942-956: {'step', 'model', 'opt', 'ema', 'scaler'} with 'model' = GaussianDiffusion.state_dict() —
into this package's GaussianDiffusion WITHOUT constructing Trainer / Accelerator / EMA / a dataset (what Trainer.load does at
:958-985, minus the optimizer, EMA and AMP-scaler state that sampling never touches), and rebuilds the two-model smoke sampler of
inference/inference_2d_smoke.py:46-127 (load_ddpm_model).  Pure host logic: tensors are copied into the modules; the CUDA
kernels see them through the normal weight packing on the next forward."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from .diffusion_2d_smoke import GaussianDiffusion
from .unet3d import Unet3D_with_Conv3D


def read_trainer_checkpoint(path: str, map_location="cpu") -> dict:
    """torch.load of a Trainer checkpoint, tensors and plain containers only (no arbitrary unpickling)."""
    data = torch.load(path, map_location=map_location, weights_only=True)
    if not isinstance(data, dict) or "model" not in data:
        raise ValueError(f"{path}: not a Trainer checkpoint (expected a dict with a 'model' entry, diffusion_2d_smoke.py:946-952)")
    return data


def load_trainer_checkpoint(diffusion: GaussianDiffusion, path: str, map_location="cpu", use_ema: bool = False,
                            strict: bool = True) -> int:
    """Loads data['model'] (or, with use_ema, the EMA copy stored under data['ema'] as 'ema_model.*', ema_pytorch layout)
    into a single-model GaussianDiffusion, like Trainer.load (diffusion_2d_smoke.py:975).  Returns the training step."""
    data = read_trainer_checkpoint(path, map_location)
    state = data["model"]
    if use_ema:
        ema = data.get("ema") or {}
        state = {k[len("ema_model."):]: v for k, v in ema.items() if k.startswith("ema_model.")}
        if not state:
            raise ValueError(f"{path}: no 'ema_model.*' entries under 'ema'")
    # accelerate / DataParallel may have wrapped the module when the file was written
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}
    diffusion.load_state_dict(state, strict=strict)
    return int(data.get("step", 0))


def load_ddpm_model(joint_dir: str, joint_milestone, w_dir: str, w_milestone, *, image_size: int = 64, frames: int = 32,
                    using_ddim: bool = False, ddim_sampling_steps: int = 100, ddim_eta: float = 0.0,
                    standard_fixed_ratio: float = 0.01, coeff_ratio: float = 0.1, w_prob_exp: float = 1.0,
                    device: Optional[str] = None, map_location="cpu") -> Tuple[GaussianDiffusion, Tuple[int, int]]:
    """inference/inference_2d_smoke.py:46-127 without the two Trainer objects: joint model (6 channels) and control prior
    (2 channels) are restored from `<dir>/model-<milestone>.pt` and combined into the eval_2ddpm sampler.
    Returns (diffusion, (joint_step, w_step))."""
    kw = dict(image_size=image_size, frames=frames, timesteps=1000,
              sampling_timesteps=ddim_sampling_steps if using_ddim else 1000, ddim_sampling_eta=ddim_eta, loss_type='l2',
              objective='pred_noise', standard_fixed_ratio=standard_fixed_ratio, coeff_ratio=coeff_ratio)
    steps = []
    models = []
    for channels, d, ms in ((6, joint_dir, joint_milestone), (2, w_dir, w_milestone)):
        net = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=channels)
        single = GaussianDiffusion(net, eval_2ddpm=False, **kw)
        steps.append(load_trainer_checkpoint(single, os.path.join(d, f"model-{ms}.pt"), map_location=map_location))
        models.append(single.model)
    diffusion = GaussianDiffusion(models, eval_2ddpm=True, w_prob_exp=w_prob_exp, **kw)
    if device is not None:
        diffusion = diffusion.to(device)
    return diffusion.eval(), (steps[0], steps[1])
