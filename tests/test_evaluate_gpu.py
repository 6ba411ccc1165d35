"""Evaluation stage (diffphycon_b200/evaluate.py, dpc_smoke_eval_sums): (1) pinned to the UNMODIFIED reference method
(InferencePipeline.multi_evaluate lifted with ast and run with the reference's own solver, tests/golden/multi_evaluate.npz), and
(2) per-trajectory metrics against a line-by-line torch restatement of the metric block (inference/inference_2d_smoke.py:384-416)
applied to the same rollout outputs."""
import numpy as np
import pytest
import torch

from diffphycon_b200 import evaluate as ev

pytestmark = pytest.mark.gpu


def _reference_metrics(pred, ro, T, w_energy, lo, hi):
    """pred [B,F,6,S,S] (initial density re-imposed, controls NOT masked: the reference masks only the copy it rolls out, :328),
    ro = rollout outputs.  Mirrors :366-371 (solver_out), :384-386 (strided views), :398-413 (metrics)."""
    B, F, _, S, _ = pred.shape
    ctrl = pred[:, :, 3:5].double().clone()
    ctrl[:, :, :, lo:hi, lo:hi] = 0
    ti, si = T // F, 128 // S
    solver_out = torch.zeros(B, T, 6, 128, 128, dtype=torch.float64)
    solver_out[:, :, 0] = ro["densitys"].double().cpu()
    solver_out[:, :, 1] = ro["velocitys"][..., 0].cpu()
    solver_out[:, :, 2] = ro["velocitys"][..., 1].cpu()
    tile = lambda c: c.reshape(B, F, 1, S, 1, S, 1).expand(B, F, ti, S, si, S, si).reshape(B, T, 128, 128)
    solver_out[:, :, 3] = tile(ctrl[:, :, 0].cpu())
    solver_out[:, :, 4] = tile(ctrl[:, :, 1].cpu())
    solver_out[:, :, 5] = ro["smoke_out"].cpu()[:, :, None, None].expand(B, T, 128, 128)
    data_super = solver_out[:, :, :, ::si, ::si]
    data = data_super[:, ::int(data_super.shape[1] / F)]
    p = pred.cpu()
    mask = torch.ones_like(p)
    mask[:, 0] = False
    p = p * mask
    data = data * mask
    diff = p - data
    out = {
        "mse": torch.cat((diff[:, :, :3], diff[:, :, [-1]]), dim=2).square().mean((1, 2, 3, 4)),
        "mse_wo_smoke": diff[:, :, :3].square().mean((1, 2, 3, 4)),
        "n_l2": diff[:, :, :3].square().sum((1, 2, 3, 4)).sqrt() / data[:, :, :3].square().sum((1, 2, 3, 4)).sqrt(),
        "n_l2_density": diff[:, :, 0].square().sum((1, 2, 3)).sqrt() / data[:, :, 0].square().sum((1, 2, 3)).sqrt(),
        "n_l2_v1": diff[:, :, 1].square().sum((1, 2, 3)).sqrt() / data[:, :, 1].square().sum((1, 2, 3)).sqrt(),
        "n_l2_v2": diff[:, :, 2].square().sum((1, 2, 3)).sqrt() / data[:, :, 2].square().sum((1, 2, 3)).sqrt(),
        "mae_smoke": (p[:, -1, 5].mean((1, 2)) - data[:, -1, 5].mean((1, 2))).abs(),
        "J_target": -data[:, -1, -1, 0, 0],
        "J_energy": data[:, :, 3:5].square().mean((1, 2, 3, 4)),
    }
    out["J_total"] = out["J_target"] + w_energy * out["J_energy"]
    return {k: v.numpy() for k, v in out.items()}


def test_multi_evaluate_matches_reference_metric_block():
    g = torch.Generator().manual_seed(3)
    B, F, S, T = 2, 8, 64, 32
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    blob = torch.exp(-((xx - 0.1) ** 2 + (yy + 0.3) ** 2) / 0.05)
    pred = torch.randn(B, F, 6, S, S, generator=g) * 0.5
    pred[:, :, 0] = pred[:, :, 0].abs()
    data = torch.zeros(B, T, 6, S, S)
    data[:, 0, 0] = blob[None] * torch.tensor([1.0, 0.7])[:, None, None]
    res = ev.multi_evaluate(pred.cuda(), data.cuda(), w_energy=0.3, per_timelength=T)
    p = pred.clone()
    p[:, 0, 0] = data[:, 0, 0]
    ref = _reference_metrics(p, res["rollout"], T, 0.3, 8, 56)
    for k, v in ref.items():
        assert np.allclose(res[k], v, rtol=1e-6, atol=1e-9), (k, res[k], v)   # the restatement sums pred in fp32 where the kernel uses fp64
    assert np.allclose(res["means"][0], ref["J_total"].mean()) and np.allclose(res["means"][4], ref["n_l2"].mean())


@pytest.mark.parametrize("B", [1, 2])
def test_multi_evaluate_matches_unmodified_reference(B, golden_dir):
    """The whole evaluation stage — re-imposed initial density, indirect-control window, 256-frame rollout of every trajectory,
    metric block — against the values the UNMODIFIED `InferencePipeline.multi_evaluate` returned for the same seeded inputs
    (tests/golden/make_golden_multi_evaluate.py: the method lifted from inference/inference_2d_smoke.py:299-427 with ast,
    executed with the reference's own solver / phi).  Relative tolerance 2e-5: fp64 rollout (1e-8 class), fp32 densities."""
    import os
    from tests.multi_evaluate_fixture import inputs
    z = np.load(os.path.join(golden_dir, "multi_evaluate.npz"))
    pred, data = inputs(B)
    res = ev.multi_evaluate(pred.cuda(), data.cuda(), w_energy=float(z["w_energy"]), per_timelength=256)
    for i, name in enumerate(("J_total", "J_target", "J_energy", "mse", "n_l2")):
        ref = float(z[f"b{B}/{name}"][0])
        got = float(res["means"][i][0])
        assert abs(got - ref) <= 2e-5 * max(abs(ref), 1e-3), (name, got, ref)
