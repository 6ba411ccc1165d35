"""diffphycon_b200 — B200-native (sm_100a) guided-diffusion sampling engine behind DiffPhyCon's Python API.

Public surface (mirrors the reference modules named in SURVEY.md section 8(b)):
    Unet3D_with_Conv3D          model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py:356
    GaussianDiffusion           diffusion/diffusion_2d_smoke.py:451
    StockSmokeGuidance          inference/inference_2d_smoke.py:30-44 (closed-form, fused into the step kernel)
    Unet, ForceUnet, JellyfishGuidance   diffusion/diffusion_2d_jellyfish.py:276-481 (boundary updater, force surrogate) and
                                inference/inference_2d_jellyfish.py:85-114, :276-279 (force_fn / design_fn): forward AND backward
Sub-modules mirror the other reference modules of the path:
    diffusion_2d_jellyfish.GaussianDiffusion    diffusion/diffusion_2d_jellyfish.py:529
    diffusion_1d_burgers.GaussianDiffusion, get_nablaJ, cosine_beta_J_schedule, ...   diffusion/diffusion_1d_burgers.py
    burgers_unet.Unet2D                         model/burgers_1d/unet.py:273
    smoke_rollout.{init_sim_128, init_velocity_, solver}   dataset/apps/evaluate_solver.py
    burgers.burgers_numeric_solve_free          dataset/apps/generate_burgers.py:207
    data_smoke.Smoke                            dataset/data_2d.py:139
    burgers_metric.burgers_metric               utils.py:1203
    evaluate.multi_evaluate                     inference/inference_2d_smoke.py:317-427 (InferencePipeline.multi_evaluate)
    checkpoint.{load_trainer_checkpoint, load_ddpm_model}   Trainer.load (diffusion_2d_smoke.py:958-985), inference_2d_smoke.py:46-127
The arithmetic lives in libdpc_b200.so (include/dpc_b200.h); there is no CPU or PyTorch fallback.
"""
from .diffusion_2d_smoke import SMOKE_RESCALER, GaussianDiffusion, StockSmokeGuidance  # noqa: F401
from .jellyfish_nets import ForceUnet, JellyfishGuidance, Unet  # noqa: F401
from .unet3d import Unet3D_with_Conv3D  # noqa: F401

__all__ = ["Unet3D_with_Conv3D", "GaussianDiffusion", "StockSmokeGuidance", "SMOKE_RESCALER", "Unet", "ForceUnet",
           "JellyfishGuidance"]
