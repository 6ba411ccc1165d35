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


# ---- load-only Trainer stand-ins (diffphycon_b200/trainer_shim.py) and checkpoints in the reference's key inventory -------
def _reference_layout_state(channels, seed, golden_dir):
    """GaussianDiffusion.state_dict() as the REFERENCE writes it: network keys from the reference-pinned inventory
    (oracle.unet3d_oracle.param_shapes is checked key-for-key against the reference module by tests/golden/make_golden.py,
    including the shared rotary_emb.freqs entries), schedule buffers from the reference's own values (tests/golden/schedules.npz)."""
    import numpy as np
    cfg = uo.UnetCfg(dim=64, dim_mults=(1, 2, 4), channels=channels)
    sd = {"model." + k: v for k, v in uo.make_params(cfg, seed).items()}
    z = np.load(os.path.join(golden_dir, "schedules.npz"))
    for k in z.files:
        if k.startswith("sigmoid1000/"):
            sd[k.split("/", 1)[1]] = torch.from_numpy(z[k])
    sd["loss_weight"] = torch.ones(1000)          # objective pred_noise: clipped snr / snr (smoke.py:565)
    return sd


def test_inference_script_flow_with_swapped_imports(tmp_path, golden_dir):
    """inference/inference_2d_smoke.py:46-127 (load_ddpm_model) line for line, with only the import lines changed."""
    from diffphycon_b200.diffusion_2d_smoke import GaussianDiffusion, Trainer
    from diffphycon_b200 import Unet3D_with_Conv3D
    states = {}
    for name, ch, seed in (("joint", 6, 3), ("w", 2, 4)):
        states[name] = _reference_layout_state(ch, seed, golden_dir)
        (tmp_path / name).mkdir()
        torch.save({"step": 500 + ch, "model": states[name], "opt": {}, "ema": {}, "scaler": None}, str(tmp_path / name / "model-9.pt"))
    nets = {}
    for name, ch in (("joint", 6), ("w", 2)):
        model = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=ch)
        diffusion = GaussianDiffusion(model, image_size=64, frames=32, timesteps=1000, sampling_timesteps=1000, ddim_sampling_eta=0.0,
                                      loss_type='l2', objective='pred_noise', standard_fixed_ratio=1e5, coeff_ratio=0.0,
                                      eval_2ddpm=False)
        diffusion.eval()
        trainer = Trainer(diffusion, dataset="Smoke", dataset_path="/nonexistent", results_path=str(tmp_path / name), amp=False)
        trainer.load(9)
        assert trainer.step == 500 + ch and trainer.device == torch.device("cpu")
        nets[name] = diffusion.model
        for k, v in states[name].items():
            assert torch.equal(diffusion.state_dict()[k], v), k
    both = GaussianDiffusion([nets["joint"], nets["w"]], image_size=64, frames=32, timesteps=1000, sampling_timesteps=1000,
                             loss_type='l2', objective='pred_noise', standard_fixed_ratio=1e5, coeff_ratio=0.0, eval_2ddpm=True,
                             w_prob_exp=0.97)
    assert both.model_joint is nets["joint"] and both.model_thetas is nets["w"]
    with pytest.raises(NotImplementedError):
        trainer.train()
    with pytest.raises(NotImplementedError):
        both(torch.zeros(1))


def test_burgers_and_jellyfish_checkpoint_formats(tmp_path):
    from diffphycon_b200 import diffusion_1d_burgers as db
    from diffphycon_b200.burgers_unet import Unet2D
    from diffphycon_b200.trainer_shim import load_jellyfish_surrogates
    from oracle import jellyfish_nets_oracle as jo
    from oracle import param_gen
    # Burgers: results_folder / cos10000-model-{n}.pt (diffusion_1d_burgers.py:949), or a file name
    net = Unet2D(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=1)
    src = db.GaussianDiffusion(net, seq_length=(16, 128), timesteps=50, auto_normalize=False, use_conv2d=True, temporal=True)
    shapes = {k: tuple(v.shape) for k, v in src.state_dict().items() if k.startswith("model.")}
    sd = dict(src.state_dict())
    sd.update(param_gen.make_params(shapes, 9))
    torch.save({"step": 77, "model": sd, "opt": {}, "ema": {}, "scaler": None, "loss": 0.1}, str(tmp_path / "cos10000-model-10.pt"))
    dst = db.GaussianDiffusion(Unet2D(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=1), seq_length=(16, 128), timesteps=50,
                               auto_normalize=False, use_conv2d=True, temporal=True)
    tr = db.Trainer(dst, None, results_folder=str(tmp_path), train_num_steps=1, save_and_sample_every=1)
    tr.load(10)
    assert tr.step == 77
    for k, v in sd.items():
        assert torch.equal(dst.state_dict()[k], v), k
    tr.load("cos10000-model-10.pt")
    # jellyfish surrogates: bare state_dict files (inference_2d_jellyfish.py:263, :273) in the reference's key inventory
    pf = jo.make_params("force", 1, dim=64, dim_mults=(1, 2, 4, 8), channels=4, out_dim=1)
    pu = jo.make_params("unet", 2, dim=64, dim_mults=(1, 2, 4, 8), channels=3, out_dim=3)
    torch.save(pf, str(tmp_path / "force.pt"))
    torch.save(pu, str(tmp_path / "bd.pt"))
    fm, bd = load_jellyfish_surrogates(str(tmp_path / "force.pt"), str(tmp_path / "bd.pt"), image_size=64)
    for k, v in pf.items():
        assert torch.equal(fm.state_dict()[k], v), k
    for k, v in pu.items():
        assert torch.equal(bd.state_dict()[k], v), k
