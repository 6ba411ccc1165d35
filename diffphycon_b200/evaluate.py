"""Evaluation stage of the smoke task on the device — mirror of
`InferencePipeline.multi_evaluate` (inference/inference_2d_smoke.py:317-427): re-impose the initial density, drop the sampled
force inside the indirect-control window, roll the controls through the simulator (one launch for the whole batch instead of one
forked process per trajectory), and reduce the objective / error metrics from the rollout outputs with one kernel
(`dpc_smoke_eval_sums`) instead of the GPU -> CPU -> processes -> GPU round trip of the reference."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import smoke_rollout as sr


@torch.no_grad()
@_lib.device_guarded
def multi_evaluate(pred: torch.Tensor, data: torch.Tensor, w_energy: float = 0.0, per_timelength: int = 256,
                   mask_window=(8, 56), sim=None):
    """pred [B,F,6,S,S] sampled (rescaled) trajectories, data [B,T,6,Sd,Sd] the test item the initial density comes from
    (inference_2d_smoke.py:321, :307).  Returns a dict of per-trajectory numpy arrays (J_total, J_target, J_energy, mse,
    mse_wo_smoke, n_l2, n_l2_density, n_l2_v1, n_l2_v2, mae_smoke) plus `means`, the five-tuple the reference returns
    (J_total, J_target, J_energy, mse, n_l2), and the rollout outputs under `rollout`."""
    if pred.device.type != "cuda":
        raise RuntimeError("diffphycon_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    B, F, C, S, _ = pred.shape
    assert C == 6 and 128 % S == 0
    pred = pred.clone().float()
    r = int(data.shape[-1] / S)
    pred[:, 0, 0] = data[:, 0, 0, ::r, ::r].to(pred)                         # initial condition (:321)
    lo, hi = mask_window
    ctrl = pred[:, :, 3:5].clone()
    ctrl[:, :, :, lo:hi, lo:hi] = 0                                          # indirect control (:328)
    sim = sim or sr.init_sim_128()
    ro = sr.solver_batch(sim, sr.init_velocity_(), data[:, 0, 0].to(pred).contiguous(), ctrl[:, :, 0].contiguous(),
                         ctrl[:, :, 1].contiguous(), per_timelength)
    sums = torch.empty(B, 12, dtype=torch.float64, device=pred.device)
    _lib.smoke_eval_sums(pred.contiguous(), ro["densitys"], ro["velocitys"], ro["smoke_out"], sums, B, F, S, per_timelength,
                         lo, hi)
    s = sums.cpu().numpy()
    n = F * S * S                                          # the reference's means run over all F frames (frame 0 contributes 0)
    e2, d2 = s[:, 0:6], s[:, 6:11]
    t_last = (F - 1) * (per_timelength // F)
    smoke_last = ro["smoke_out"][:, t_last].cpu().numpy()
    out = {
        "mse": (e2[:, 0] + e2[:, 1] + e2[:, 2] + e2[:, 5]) / (4 * n),                       # :403
        "mse_wo_smoke": (e2[:, 0] + e2[:, 1] + e2[:, 2]) / (3 * n),                         # :404
        "n_l2": np.sqrt(e2[:, 0:3].sum(1)) / np.sqrt(d2[:, 0:3].sum(1)),                    # :405
        "n_l2_density": np.sqrt(e2[:, 0]) / np.sqrt(d2[:, 0]),                              # :406
        "n_l2_v1": np.sqrt(e2[:, 1]) / np.sqrt(d2[:, 1]),                                   # :407
        "n_l2_v2": np.sqrt(e2[:, 2]) / np.sqrt(d2[:, 2]),                                   # :408
        "mae_smoke": np.abs(s[:, 11] / (S * S) - smoke_last),                               # :409
        "J_target": -smoke_last,                                                            # :411
        "J_energy": (d2[:, 3] + d2[:, 4]) / (2 * n),                                        # :412
    }
    out["J_total"] = out["J_target"] + w_energy * out["J_energy"]                           # :413
    out["means"] = tuple(np.array([out[k].mean()]) for k in ("J_total", "J_target", "J_energy", "mse", "n_l2"))
    out["rollout"] = ro
    return out
