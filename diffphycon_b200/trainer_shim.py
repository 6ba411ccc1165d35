"""Load-only stand-ins for the reference `Trainer` classes, so that the inference drivers work with only their import lines
swapped (inference/inference_2d_smoke.py:14, :70-77, :102-109; inference_2d_jellyfish.py:23, :159-177, :207-224;
inference_1d_burgers.py:8, :188-195): they build a `Trainer(diffusion, ...)`, call `.load(milestone)` and keep using
`diffusion.model` / `trainer.device`.

The reference Trainer (diffusion_2d_smoke.py:843-985, diffusion_2d_jellyfish.py, diffusion_1d_burgers.py:844-1035) needs a
dataset on disk, an `accelerate.Accelerator`, an Adam optimizer and an EMA copy just to call `load`; sampling touches none of
them.  These shims accept the same constructor arguments (training-only ones are stored and otherwise ignored), restore
`data['model']` into the GaussianDiffusion exactly like `Trainer.load` (`model.load_state_dict(data['model'])`, strict) and
refuse `train()` / `save()` loudly: training is outside this engine's scope."""
from __future__ import annotations

import os
from pathlib import Path

import torch


class _LoadOnlyTrainer:
    FILE_PATTERN = "model-{}.pt"

    def _setup(self, diffusion_model, results_dir, options):
        self.model = diffusion_model
        self.channels = getattr(diffusion_model, "channels", None)
        self.image_size = getattr(diffusion_model, "image_size", None)
        self.results_path = self.results_folder = Path(results_dir)
        self.options = dict(options)
        self.step = 0
        self.ema = None
        self.opt = None

    @property
    def device(self):
        """`accelerator.device` in the reference: the device the diffusion's buffers live on."""
        for b in self.model.buffers():
            return b.device
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def checkpoint_file(self, milestone) -> str:
        if isinstance(milestone, str) and not milestone.isdigit():
            return str(self.results_path / milestone)          # diffusion_1d_burgers.py:958-959: a file name
        return str(self.results_path / self.FILE_PATTERN.format(milestone))

    def load(self, milestone):
        """diffusion_2d_smoke.py:956-985 minus optimizer / EMA / scaler state."""
        path = self.checkpoint_file(milestone)
        data = torch.load(path, map_location=self.device, weights_only=True)
        if not isinstance(data, dict) or "model" not in data:
            raise ValueError(f"{path}: not a Trainer checkpoint (expected a dict with a 'model' entry)")
        state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in data["model"].items()}
        self.model.load_state_dict(state)
        self.step = int(data.get("step", 0))
        return self

    def save(self, milestone):
        raise NotImplementedError("diffphycon_b200 is a sampling engine: Trainer.save / train are not implemented "
                                  "(use the reference Trainer to train; its checkpoints load here)")

    def train(self):
        raise NotImplementedError("diffphycon_b200 is a sampling engine: Trainer.train is not implemented "
                                  "(use the reference Trainer to train; its checkpoints load here)")


class SmokeTrainer(_LoadOnlyTrainer):
    """diffusion_2d_smoke.py:843-866: Trainer(diffusion_model, dataset, dataset_path, *, ..., results_path='./results', ...)."""

    def __init__(self, diffusion_model, dataset=None, dataset_path=None, *, results_path='./results', **training_options):
        self._setup(diffusion_model, results_path, dict(training_options, dataset=dataset, dataset_path=dataset_path))


class JellyfishTrainer(SmokeTrainer):
    """diffusion_2d_jellyfish.py Trainer: same file layout (`results_path / model-{milestone}.pt`); extra keyword arguments
    (frames, traj_len, ts, log_path, calculate_fid, ...) are training-side."""


class BurgersTrainer(_LoadOnlyTrainer):
    """diffusion_1d_burgers.py:844-866: Trainer(diffusion_model, dataset, *, ..., results_folder='./results', ...); files are
    `cos10000-model-{milestone}.pt` (:949), or a file name when `milestone` is a string (:958-959)."""
    FILE_PATTERN = "cos10000-model-{}.pt"

    def __init__(self, diffusion_model, dataset=None, *, results_folder='./results', **training_options):
        self._setup(diffusion_model, results_folder, dict(training_options, dataset=dataset))


def load_jellyfish_surrogates(force_model_checkpoint: str, boundary_updater_model_checkpoint: str, image_size: int = 64,
                              device=None):
    """inference/inference_2d_jellyfish.py:256-274: ForceUnet(dim=image_size, out_dim=1, channels=4) and
    Unet(dim=image_size, out_dim=3, channels=3), restored from BARE state_dict files.  (Quirk 6: ForceUnet.final is
    Linear(512, 1), so `image_size` must be 64 here; larger images are fed to the same dim-64 networks.)"""
    from .jellyfish_nets import ForceUnet, Unet
    force_model = ForceUnet(dim=image_size, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
    force_model.load_state_dict(torch.load(force_model_checkpoint, map_location="cpu", weights_only=True))
    bd_updater = Unet(dim=image_size, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
    bd_updater.load_state_dict(torch.load(boundary_updater_model_checkpoint, map_location="cpu", weights_only=True))
    if device is not None:
        force_model, bd_updater = force_model.to(device), bd_updater.to(device)
    return force_model, bd_updater
