"""Checkpoint adapter (diffphycon_b200/checkpoint.py): files in the layout the reference's Trainer.save writes
(diffusion_2d_smoke.py:942-956) load into this package's classes with strict key checking, bit for bit."""
import os

import pytest
import torch

import diffphycon_b200 as dpc
from diffphycon_b200 import checkpoint as ck
from oracle import unet3d_oracle as uo


def _trainer_file(tmp_path, name, channels, seed, step, wrap_module=False):
    """A model-{n}.pt as Trainer.save builds it: 'model' = GaussianDiffusion.state_dict() (network under 'model.*' plus the
    schedule buffers), an EMA copy in ema_pytorch's layout, optimizer and scaler entries that sampling ignores."""
    cfg = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=channels)
    net = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=channels)
    net.load_state_dict(uo.make_params(cfg, seed), strict=True)
    diff = dpc.GaussianDiffusion(net, image_size=64, frames=32, timesteps=1000, loss_type='l2', objective='pred_noise')
    sd = {k: v.clone() for k, v in diff.state_dict().items()}
    ema = {"initted": torch.tensor(True), "step": torch.tensor(step)}
    ema.update({"ema_model." + k: v * 0.5 for k, v in sd.items()})
    ema.update({"online_model." + k: v for k, v in sd.items()})
    if wrap_module:
        sd = {"module." + k: v for k, v in sd.items()}
    d = tmp_path / name
    d.mkdir()
    torch.save({"step": step, "model": sd, "opt": {"state": {}, "param_groups": [{"lr": 1e-4}]}, "ema": ema, "scaler": None},
               str(d / "model-7.pt"))
    return str(d), diff


def test_single_model_checkpoint_roundtrip(tmp_path):
    d, src = _trainer_file(tmp_path, "joint", 6, 3, 1234, wrap_module=True)
    net = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=6)
    dst = dpc.GaussianDiffusion(net, image_size=64, frames=32, timesteps=1000, loss_type='l2', objective='pred_noise')
    assert ck.load_trainer_checkpoint(dst, os.path.join(d, "model-7.pt")) == 1234
    for (k1, v1), (k2, v2) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    ck.load_trainer_checkpoint(dst, os.path.join(d, "model-7.pt"), use_ema=True)
    k = "model.init_conv.weight"
    assert torch.equal(dst.state_dict()[k], src.state_dict()[k] * 0.5)


def test_two_model_sampler_from_checkpoints(tmp_path):
    dj, sj = _trainer_file(tmp_path, "joint", 6, 5, 10)
    dw, sw = _trainer_file(tmp_path, "w", 2, 6, 20)
    diff, steps = ck.load_ddpm_model(dj, 7, dw, 7, standard_fixed_ratio=1e5, coeff_ratio=0.0, w_prob_exp=0.97)
    assert steps == (10, 20) and diff.eval_2ddpm and diff.w_prob_exp == 0.97
    for name, src, mod in (("joint", sj.model, diff.model_joint), ("w", sw.model, diff.model_thetas)):
        for (k1, v1), (k2, v2) in zip(src.state_dict().items(), mod.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), (name, k1)


def test_rejects_foreign_files_and_missing_keys(tmp_path):
    p = tmp_path / "x.pt"
    torch.save({"weights": torch.zeros(3)}, str(p))
    net = dpc.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=2)
    dst = dpc.GaussianDiffusion(net, image_size=64, frames=32, timesteps=1000, loss_type='l2', objective='pred_noise')
    with pytest.raises(ValueError, match="not a Trainer checkpoint"):
        ck.load_trainer_checkpoint(dst, str(p))
    d, _ = _trainer_file(tmp_path, "joint6", 6, 3, 1)
    with pytest.raises(RuntimeError):                      # a 6-channel file does not fit the 2-channel network: strict load
        ck.load_trainer_checkpoint(dst, os.path.join(d, "model-7.pt"))
