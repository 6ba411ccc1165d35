"""Jellyfish dataset reader — same constructor, on-disk layout, normalisation and item tuples
as the reference `Jellyfish` class (dataset/data_2d.py:11-140), so `inference_2d_jellyfish.load_data` can take it unchanged:

    <root>/{train_data,test_data}/normalization_max_min.pkl      {vx,vy,p}_{max,min}
    <root>/.../states/sim_{id:06d}.npz['a']                      [T, 3, S, S]   (vx, vy, pressure)
    <root>/.../bdry_merged_mask_offsets/sim_{id:06d}.npz['a']    [T, S', S', 3] (mask + 2 offsets)
    <root>/.../bdry_head_thetas/sim_{id:06d}.npz['thetas']       [T]

States are clamped to [0, 1] after min-max scaling, mapped to [-1, 1]; NaNs become 0.  `normalization()` returns the
{p,vx,vy}_{min,max} dictionary that inference_2d_jellyfish.py:29-37 unpickles at import time (the pressure range feeds
`JellyfishGuidance`).  Pure host code (numpy + torch), no kernels involved."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset


def normalization(dataset_path, split="train_data") -> dict:
    """The dictionary of `<root>/<split>/normalization_max_min.pkl` (inference_2d_jellyfish.py:29-37)."""
    with open(os.path.join(dataset_path, split, "normalization_max_min.pkl"), "rb") as fh:
        return pickle.load(fh)


class Jellyfish(Dataset):
    def __init__(self, dataset, dataset_path, time_steps=40, steps=20, time_interval=1, is_train=True, is_testdata=False,
                 for_pipeline=False, only_vis_pressure=False):
        super().__init__()
        if not dataset.startswith('jellyfish'):
            raise ValueError(f"unknown dataset {dataset!r} (the reference has a bare `raise` here)")
        self.dataset, self.root = dataset, dataset_path
        self.steps, self.time_steps, self.time_interval = steps, time_steps, time_interval
        self.is_train, self.is_testdata = is_train, is_testdata
        self.win_size = steps * time_interval
        self.for_pipeline, self.only_vis_pressure = for_pipeline, only_vis_pressure
        self.dirname = "train_data" if is_train else "test_data"
        if is_testdata:
            self.n_simu = 100 if is_train else 50
        else:
            self.n_simu = 1000 if is_train else 100
        self.time_steps_effective = (time_steps - self.win_size) // time_interval
        fn = os.path.join(self.root, self.dirname, "normalization_max_min.pkl")
        if not os.path.isfile(fn):
            raise FileNotFoundError(fn)                        # bare `raise` in the reference
        for k, v in normalization(self.root, self.dirname).items():
            if k in ("vx_max", "vx_min", "vy_max", "vy_min", "p_max", "p_min"):
                setattr(self, k, v)

    def __len__(self):
        return self.n_simu * self.time_steps_effective if self.is_train else self.n_simu

    def _npz(self, sub, sim_id, key):
        return np.load(os.path.join(self.root, self.dirname, sub, "sim_{:06d}.npz".format(sim_id)))[key]

    def _scaled(self, v, lo, hi):
        return (torch.clamp((v - lo) / (hi - lo), 0, 1) - 0.5).unsqueeze(1) * 2

    def __getitem__(self, idx):
        if self.for_pipeline or self.is_train:
            sim_id, time_id = divmod(idx, self.time_steps_effective)
        else:
            sim_id, time_id = idx, 0
        state_full = torch.FloatTensor(self._npz("states", sim_id, "a"))
        if not self.for_pipeline:
            pressure = self._scaled(state_full[:, 2], self.p_min, self.p_max)
            if not self.only_vis_pressure:
                state_full = torch.cat((self._scaled(state_full[:, 0], self.vx_min, self.vx_max),
                                        self._scaled(state_full[:, 1], self.vy_min, self.vy_max), pressure), 1)
            else:
                state_full = pressure
        state_full[torch.isnan(state_full)] = 0
        win = slice(time_id, time_id + self.win_size)
        state = state_full[win]
        bd_full = self._npz("bdry_merged_mask_offsets", sim_id, "a")
        bd = torch.FloatTensor(np.transpose(bd_full[win], (0, 3, 1, 2)))
        bd[torch.isnan(bd)] = 0
        thetas_full = self._npz("bdry_head_thetas", sim_id, "thetas")
        thetas = torch.FloatTensor(thetas_full[win])
        bd_0 = lambda: torch.FloatTensor(np.transpose(bd_full[0], (2, 0, 1)))
        if self.for_pipeline:
            return state, bd, thetas, bd_0(), sim_id, time_id
        if self.is_train:
            return state, bd, thetas, sim_id, time_id
        return state_full[0], thetas[0], bd_0(), sim_id, torch.FloatTensor(thetas_full[:self.win_size])
