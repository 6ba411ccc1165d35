r"""Golden vectors of the jellyfish surrogate networks from the UNMODIFIED reference (build container
only):  python tests/golden/make_golden_jellyfish_nets.py

  * `Unet` (boundary updater) and `ForceUnet` of diffusion/diffusion_2d_jellyfish.py:276-481, constructed by the reference
    module itself, loaded with the deterministic synthetic weights of oracle.jellyfish_nets_oracle.make_params
    (load_state_dict(strict=True) pins the key / shape inventory);
  * `force_fn` / `reg_theta` / `unnormalize_state` of inference/inference_2d_jellyfish.py:35-36, :47-60, :85-114.  That module
    cannot be imported (it unpickles a dataset file and imports matplotlib / SAC code at import time), so the three function
    definitions are lifted from the UNMODIFIED source file with `ast` and executed with synthetic p_min / p_max.
Stored: inputs, network outputs, and the guidance gradient design_fn(x, bd_0) = cat(grad_state, grad_theta)."""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import jellyfish_nets_oracle as jo  # noqa: E402
from oracle import ref_import  # noqa: E402

ref_import._prepare()
import importlib  # noqa: E402

jm = importlib.import_module("diffusion.diffusion_2d_jellyfish")
P_MIN, P_MAX, REG = -1.7, 2.9, 1000.0


def lifted_force_fn():
    src = open(os.path.join(ref_import.REFERENCE_ROOT, "inference", "inference_2d_jellyfish.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("unnormalize_state", "reg_theta", "force_fn")]
    assert len(keep) == 3
    ns = {"torch": torch, "grad": torch.autograd.grad, "p_min": P_MIN, "p_max": P_MAX}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "inference_2d_jellyfish.py", "exec"), ns)
    return ns["force_fn"]


def build(kind, seed):
    if kind == "unet":
        net = jm.Unet(dim=64, out_dim=3, dim_mults=(1, 2, 4, 8), channels=3)
        kw = dict(dim=64, dim_mults=(1, 2, 4, 8), channels=3, out_dim=3)
    else:
        net = jm.ForceUnet(dim=64, out_dim=1, dim_mults=(1, 2, 4, 8), channels=4)
        kw = dict(dim=64, dim_mults=(1, 2, 4, 8), channels=4, out_dim=1)
    ref = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(v) for k, v in jo.param_shapes(kind, **kw).items()}
    assert ref == mine, (set(ref) ^ set(mine), [k for k in ref if k in mine and ref[k] != mine[k]])
    net.load_state_dict(jo.make_params(kind, seed, **kw), strict=True)
    return net.eval()


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    bd_updater, force_model = build("unet", 51), build("force", 52)
    force_fn = lifted_force_fn()
    out = {}
    for tag, B, Fr, S in (("s32", 2, 3, 32), ("s64", 1, 2, 64)):
        g = torch.Generator().manual_seed(60 + S)
        x = torch.rand(B, Fr, 4, S, S, generator=g) * 2 - 1
        x[:, :, 3] = 0.5 + 0.3 * x[:, :, 3]          # angle field around 0.5
        bd_0 = torch.cat([(torch.rand(B, Fr, 1, S, S, generator=g) > 0.7).float(), torch.rand(B, Fr, 2, S, S, generator=g) - 0.5], 2)
        args = types.SimpleNamespace(only_vis_pressure=False, device="cpu", reg_ratio=REG)
        gs, gt = force_fn(x.clone(), bd_0, force_model, bd_updater, args)
        grad = torch.cat([gs, gt.unsqueeze(2)], dim=2)
        args0 = types.SimpleNamespace(only_vis_pressure=False, device="cpu", reg_ratio=0.0)   # network part of d/d theta alone
        gs0, gt0 = force_fn(x.clone(), bd_0, force_model, bd_updater, args0)
        out[f"{tag}/grad_noreg"] = torch.cat([gs0, gt0.unsqueeze(2)], dim=2).detach().numpy()
        print(tag, "network-only grad_theta absmax", float(gt0.abs().max()))
        with torch.no_grad():
            theta = x[:, :, 3].mean((-1, -2))
            pred_bd = bd_updater(bd_0.reshape(B * Fr, 3, S, S), theta.reshape(B * Fr))
            pressure = (0.5 * x[:, :, 2] + 0.5) * (P_MAX - P_MIN) + P_MIN
            inp = torch.cat((pressure.reshape(B * Fr, 1, S, S), pred_bd), 1)
            force = force_model(inp)
        out.update({f"{tag}/x": x.numpy(), f"{tag}/bd_0": bd_0.numpy(), f"{tag}/grad": grad.detach().numpy(),
                    f"{tag}/pred_bd": pred_bd.numpy(), f"{tag}/force": force.numpy()})
        print(tag, "grad_state absmax", float(gs.abs().max()), "grad_theta absmax", float(gt.abs().max()), "force", force.flatten()[:3])
    out["p_min"], out["p_max"], out["reg_ratio"] = np.float64(P_MIN), np.float64(P_MAX), np.float64(REG)
    out["seed_unet"], out["seed_force"] = np.int64(51), np.int64(52)
    np.savez_compressed(os.path.join(HERE, "jellyfish_nets.npz"), **out)


if __name__ == "__main__":
    main()
