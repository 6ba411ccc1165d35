"""Burgers evaluation metric on the device — mirror of `utils.py::burgers_metric` (utils.py:1203-1284), the function
`inference/inference_1d_burgers.py` scores a sampled control with: roll the force through the finite-difference solver
(`dpc_burgers_rollout`, one launch for the batch) and compare the controlled final state with the target.  Same arguments
and return values; the reductions are [B,128]-sized and stay in torch on the device."""
from __future__ import annotations

import torch

from .burgers import burgers_numeric_solve_free


def _default_solver(u_target, f):
    return burgers_numeric_solve_free(u_target[:, 0, :], f, visc=0.01, T=1.0, dt=1e-4, num_t=10)


@torch.no_grad()
def burgers_metric(u_target: torch.Tensor, f: torch.Tensor, target='final_u', partial_control='full', report_all=False,
                   diffused_u=None, evaluate_u=False, partially_observed=None, solver=_default_solver, **kwargs):
    """u_target [B, Nt, Nx], f [B, Nt-1, Nx], both NOT rescaled.  Returns (J_actual, control_energy) as the reference does:
    J_actual = per-sample MSE at the final time (or the six-tuple mse, mse_median, mae, mae_median, nmse, nmae with
    report_all), control_energy = sum of squares of the (partially zeroed) force."""
    if kwargs != {}:
        print('WARNING: kwargs', [k for k in kwargs.keys()], 'are not used.')
    u_target, f = u_target.clone(), f.clone()
    assert len(u_target.size()) == len(f.size()) == 3
    if partial_control is None or partial_control == 'full':
        pass
    elif partial_control == 'front_rear_quarter':
        Nx = f.size(2)
        f[:, :, Nx // 4: (Nx * 3) // 4] = 0                                  # utils.py:1248-1250
    u_controlled = diffused_u.clone() if evaluate_u else solver(u_target, f)
    if partially_observed is not None:
        Nx = u_controlled.size(-1)
        if partially_observed == 'front_rear_quarter':
            idx = torch.cat((torch.arange(0, Nx // 4), torch.arange((3 * Nx) // 4, Nx))).to(u_controlled.device)
            u_controlled = u_controlled[..., idx]
            u_target = u_target[..., idx]
        else:
            raise NotImplementedError
    if target != 'final_u':
        raise ValueError('Undefined target to evaluate')
    diff = u_controlled[:, -1, :] - u_target[:, -1, :]
    mse = diff.square().mean(-1)
    if not report_all:
        J_actual = mse
    else:
        ep = 1e-5
        mse_median, _ = diff.square().median(-1)
        mae = diff.abs().mean(-1)
        mae_median, _ = diff.abs().median(-1)
        nmse = mse / (u_target[:, -1, :].square().mean() + ep)
        nmae = mae / (u_target[:, -1, :].abs().mean() + ep)
        J_actual = (mse, mse_median, mae, mae_median, nmse, nmae)
    control_energy = f.square().sum((-1, -2))
    return J_actual, control_energy
