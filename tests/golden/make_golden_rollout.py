"""Golden vectors for the post-sampling smoke rollout, produced by the UNMODIFIED reference
`dataset/apps/evaluate_solver.py::solver` + vendored `phi/` executed through the AST index-fix import hook of
oracle/ref_import.py (build container only):   python tests/golden/make_golden_rollout.py
Stored: the obstacle / velocity masks of init_sim_128, one pressure solve (divergence -> pressure, iteration count) and a
4-frame rollout (3 simulation steps) with every output array of `solver`."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def inputs(seed, nt, nx):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, nx), np.linspace(-1, 1, nx), indexing="ij")
    # one blob in the free stream, one sitting on a target bucket so that the smoke accounting / zeroing is exercised
    dens = (np.exp(-((xx - 0.1) ** 2 + (yy + 0.3) ** 2) / 0.05) + np.exp(-((xx + 0.52) ** 2 + (yy - 0.81) ** 2) / 0.02)).astype(np.float32)
    # smooth-ish random controls of the magnitude the sampler produces after rescaling (|c| ~ 1)
    c1 = (rng.standard_normal((nt, nx, nx)) * 0.6).astype(np.float32)
    c2 = (rng.standard_normal((nt, nx, nx)) * 0.6 + 0.3).astype(np.float32)
    return dens, c1, c2


def main():
    es = ref_import.evaluate_solver_module()
    from phi.solver.sparse import SparseCGPressureSolver
    from phi.math.nd import StaggeredGrid
    sim = es.init_sim_128()
    out = dict(fluid_mask=sim._fluid_mask[0, :, :, 0].astype(np.int8), active_mask=sim._active_mask[0, :, :, 0].astype(np.int8),
               velocity_mask=sim._velocity_mask.staggered[0].astype(np.float32))
    # ---- one pressure solve on a random staggered field ----
    rng = np.random.default_rng(5)
    v = rng.standard_normal((1, 128, 128, 2))
    vel = sim.with_boundary_conditions(StaggeredGrid(v.copy()))
    div = vel.divergence()
    pressure, iters = SparseCGPressureSolver().solve(div.copy(), sim._active_mask, sim._fluid_mask, sim._boundary, 1e-8,
                                                     return_loop_counter=True)
    out.update(cg_velocity=v[0], cg_divergence=div[0, :, :, 0], cg_pressure=pressure[0, :, :, 0], cg_iterations=np.int64(iters))
    proj = sim.divergence_free(StaggeredGrid(v.copy()), solver=SparseCGPressureSolver(), accuracy=1e-8)
    proj = sim.with_boundary_conditions(proj)
    out.update(cg_projected=proj.staggered[0])
    # ---- rollout ----
    nt, nx, T = 2, 64, 4
    dens, c1, c2 = inputs(11, nt, nx)
    v0 = es.init_velocity_()
    d, zd, vs, c1t, c2t, rec = es.solver(sim, v0, dens, c1, c2, per_timelength=T)
    out.update(init_density=dens, c1=c1, c2=c2, init_velocity=np.asarray(v0)[0], densitys=d, zero_densitys=zd, velocitys=vs,
               smoke_out_record=rec[:, 0, 0])
    np.savez_compressed(os.path.join(HERE, "smoke_rollout.npz"), **out)
    print("cg iterations", iters, "smoke_out", rec[:, 0, 0])
    for k, v_ in out.items():
        print(k, getattr(v_, "shape", None), getattr(v_, "dtype", None))


if __name__ == "__main__":
    main()
