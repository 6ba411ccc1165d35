"""Golden values of the smoke evaluation stage: `InferencePipeline.multi_evaluate` / `per_evaluate` of
inference/inference_2d_smoke.py:299-427, lifted from the UNMODIFIED source file with `ast` (the module itself needs matplotlib,
accelerate, a dataset and checkpoints at import) and executed with the reference's own `solver` / `init_sim_128` /
`init_velocity_` (dataset/apps/evaluate_solver.py + vendored phi through the index-fix import hook of oracle/ref_import.py).
The one substitution: `multiprocess` is replaced by an in-process stand-in (Process.start() runs the target, Queue is a list),
which changes where the rollouts run, not what they compute.  Build container only (about 50 s of CPU per trajectory):
    python tests/golden/make_golden_multi_evaluate.py"""
import ast
import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402
from tests.multi_evaluate_fixture import inputs  # noqa: E402

W_ENERGY = 0.3


class _Queue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)

    def get(self):
        return self.items.pop(0)


class _Process:
    def __init__(self, target, args):
        self.target, self.args = target, args

    def start(self):
        self.target(*self.args)

    def join(self):
        pass


def lifted_pipeline():
    es = ref_import.evaluate_solver_module()
    src = open(os.path.join(ref_import.REFERENCE_ROOT, "inference", "inference_2d_smoke.py")).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "InferencePipeline"][0]
    cls.body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("per_evaluate", "multi_evaluate")]
    assert len(cls.body) == 2
    ns = {"np": np, "torch": torch, "time": time, "mp": types.SimpleNamespace(Queue=_Queue, Process=_Process),
          "init_sim_128": es.init_sim_128, "init_velocity_": es.init_velocity_, "solver": es.solver,
          "gif_density": lambda *a, **k: None}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), "inference_2d_smoke.py", "exec"), ns)
    pipe = ns["InferencePipeline"].__new__(ns["InferencePipeline"])
    pipe.args_general = types.SimpleNamespace(w_energy=W_ENERGY)
    return pipe


def main():
    pipe = lifted_pipeline()
    out = {"w_energy": np.float64(W_ENERGY)}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                  # multi_evaluate writes plot_pred.npy into the working directory
        try:
            for B in (1, 2):
                pred, data = inputs(B)
                res = pipe.multi_evaluate(pred.clone(), data.clone())
                for name, v in zip(("J_total", "J_target", "J_energy", "mse", "n_l2"), res):
                    out[f"b{B}/{name}"] = np.asarray(v, dtype=np.float64)
                print(B, [float(v[0]) for v in res], flush=True)
        finally:
            os.chdir(cwd)
    np.savez(os.path.join(HERE, "multi_evaluate.npz"), **out)


if __name__ == "__main__":
    main()
