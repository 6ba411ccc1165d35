"""Smoke dataset reader (second half) — same constructor, item layout and RESCALER as the
reference `Smoke` class (dataset/data_2d.py:139-209) so that `inference_2d_smoke.load_data` (:447-463) can take it unchanged:
    train: `<root>/train/sim_{id:06d}/{Density,Velocity,Control,Smoke}.npy`  -> ([32, 6, 64, 64] / RESCALER, sim_id)
    test : `<root>/test/control/sim_{id:06d}/...`                            -> ([256, 6, 64, 64] not rescaled, sim_id)
Arrays are stored (H, W, C, T); channels are concatenated as density(1), velocity(2), control(2), smoke ratio(1) where the
ratio is Smoke[:, 1] / Smoke.sum(-1) broadcast over the frame.  Pure host code (numpy + torch), no kernels involved."""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset


class Smoke(Dataset):
    def __init__(self, dataset_path, time_steps=256, steps=32, all_size=128, size=64, is_train=True):
        super().__init__()
        self.root = dataset_path
        self.steps = steps
        self.time_steps = time_steps
        self.time_interval = int(time_steps / steps)
        self.all_size = all_size
        self.size = size
        self.space_interval = int(all_size / size)
        self.is_train = is_train
        self.dirname = "train" if self.is_train else "test"
        self.sub_dirname = "control"
        self.n_simu = 20000 if self.is_train else 50
        self.RESCALER = torch.tensor([2, 18, 20, 16, 20, 1]).reshape(1, 6, 1, 1)

    def __len__(self):
        return self.n_simu

    def _sim_dir(self, sim_id):
        parts = [self.root, self.dirname] + ([] if self.is_train else [self.sub_dirname]) + ["sim_{:06d}".format(sim_id)]
        return os.path.join(*parts)

    def __getitem__(self, sim_id):
        d = self._sim_dir(sim_id)

        def field(name):                                     # (H, W, C, T) -> (C, T, H, W)
            return torch.tensor(np.load(os.path.join(d, name + ".npy")), dtype=torch.float).permute(2, 3, 0, 1)

        dens, vel, ctrl = field("Density"), field("Velocity"), field("Control")
        s = torch.tensor(np.load(os.path.join(d, "Smoke.npy")), dtype=torch.float)
        s = s[:, 1] / s.sum(-1)
        s = s.reshape(1, s.shape[0], 1, 1).expand(1, s.shape[0], self.size, self.size)
        nkeep = 32 if self.is_train else 256
        state = torch.cat((dens, vel, ctrl, s), dim=0)[:, :nkeep]
        if self.is_train:
            return state.permute(1, 0, 2, 3) / self.RESCALER, sim_id
        return state.permute(1, 0, 2, 3), sim_id
