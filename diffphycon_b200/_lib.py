"""ctypes binding of libdpc_b200.so (the C-ABI declared in include/dpc_b200.h).

There is NO fallback: if the shared library is missing, or a kernel launch fails, this module raises.  PyTorch is
used only for device memory and streams; every function here takes torch CUDA tensors and passes raw pointers.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import functools
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdpc_b200.so")
ABI_VERSION = 1

c_fp = C.c_void_p


class ConvParams(C.Structure):
    """struct dpc_conv_params (include/dpc_b200.h)."""
    _fields_ = [
        ("x1", c_fp), ("x2", c_fp), ("w", c_fp), ("bias", c_fp), ("residual", c_fp), ("taps", c_fp), ("y", c_fp),
        ("gn_stats", c_fp),
        ("C1", C.c_int32), ("C2", C.c_int32),
        ("B", C.c_int32), ("Fi", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32),
        ("Fo", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("ntaps", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("pt", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32),
        ("Cout", C.c_int32), ("Npad", C.c_int32), ("Kpad", C.c_int32),
        ("Hfull", C.c_int32), ("Wfull", C.c_int32), ("oh_mul", C.c_int32), ("oh_off", C.c_int32),
        ("ow_mul", C.c_int32), ("ow_off", C.c_int32),
        ("out_layout", C.c_int32), ("gn_groups", C.c_int32), ("precise", C.c_int32),
        ("res_scale", c_fp), ("res_shift", c_fp),
    ]


class StepCoefs(C.Structure):
    """struct dpc_step_coefs (include/dpc_b200.h)."""
    _fields_ = [
        ("sqrt_recip_alphas_cumprod", C.c_float), ("sqrt_recipm1_alphas_cumprod", C.c_float),
        ("guidance_coef", C.c_float), ("prior_coef", C.c_float), ("w_energy", C.c_float),
        ("rescaler", C.c_float * 6),
        ("posterior_mean_coef1", C.c_float), ("posterior_mean_coef2", C.c_float), ("sigma", C.c_float),
        ("add_noise", C.c_int32),
        ("sqrt_alpha_next", C.c_float), ("c", C.c_float), ("ddim_sigma", C.c_float),
        ("last", C.c_int32),
    ]


_lib = None

_SIGNATURES = {
    "dpc_abi_version": ([], C.c_int),
    "dpc_last_error": ([], C.c_char_p),
    "dpc_device_is_sm100": ([], C.c_int),
    "dpc_conv_igemm": ([C.POINTER(ConvParams), c_fp], C.c_int),
    "dpc_conv3d_tcgen05": ([C.POINTER(ConvParams), c_fp], C.c_int),
    "dpc_groupnorm_silu": ([c_fp, c_fp, c_fp, c_fp, c_fp, C.c_int64, C.c_int64, c_fp, c_fp, C.c_int32, C.c_int64,
                            C.c_int32, C.c_int32, C.c_float, c_fp], C.c_int),
    "dpc_layernorm_channels": ([c_fp, c_fp, c_fp, c_fp, C.c_int64, C.c_int32, C.c_float, C.c_int32, c_fp], C.c_int),
    "dpc_upsample_nearest2x": ([c_fp, c_fp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_pack_input": ([c_fp, c_fp] + [C.c_int32] * 8 + [c_fp], C.c_int),
    "dpc_temporal_attention": ([c_fp] * 5 + [C.c_int32] * 6 + [c_fp], C.c_int),
    "dpc_temporal_block_fused": ([c_fp] * 7 + [C.c_int32] * 5 + [C.c_float, c_fp], C.c_int),
    "dpc_spatial_attention": ([c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_spatial_linear_block_fused": ([c_fp] * 7 + [C.c_int32] * 4 + [C.c_float, c_fp], C.c_int),
    "dpc_stem_conv_tcgen05": ([c_fp] * 4 + [C.c_int32] * 9 + [c_fp], C.c_int),
    "dpc_final_proj": ([c_fp] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_spatial_attention_mma": ([c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_smoke_eval_sums": ([c_fp] * 5 + [C.c_int32] * 6 + [c_fp], C.c_int),
    "dpc_gn_fold": ([c_fp] * 5 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_float, c_fp], C.c_int),
    "dpc_gn_stats_merge": ([c_fp] * 2 + [C.c_int32] * 3 + [c_fp], C.c_int),
    "dpc_spatial_linear_attention": ([c_fp, c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_time_embed": ([c_fp] * 8 + [C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_time_proj": ([c_fp] * 4 + [C.c_int32] * 3 + [c_fp], C.c_int),
    "dpc_ddpm_guided_step": ([c_fp] * 6 + [C.c_int32, C.POINTER(StepCoefs), c_fp, c_fp] + [C.c_int32] * 4 + [c_fp],
                             C.c_int),
    "dpc_ddim_guided_step": ([c_fp] * 6 + [C.c_int32, C.POINTER(StepCoefs), c_fp, c_fp] + [C.c_int32] * 4 + [c_fp],
                             C.c_int),
    "dpc_sampler_prepare": ([c_fp] * 3 + [C.c_int32, c_fp, C.c_int32, c_fp, c_fp], C.c_int),
    "dpc_guided_step_dev": ([C.c_int32] + [c_fp] * 9 + [C.c_int32] * 4 + [c_fp], C.c_int),
    "dpc_predict_x_start": ([c_fp, c_fp, C.c_float, C.c_float, C.c_int32, c_fp, C.c_int64, c_fp], C.c_int),
    "dpc_renoise": ([c_fp, c_fp, C.c_float, C.c_float, c_fp, C.c_int64, c_fp], C.c_int),
    "dpc_burgers_model_output": ([c_fp] * 5 + [C.c_int32] + [C.c_float] * 4 + [C.c_int32, C.c_int64, C.c_int64, c_fp], C.c_int),
    "dpc_ddpm_posterior_step": ([c_fp] * 7 + [C.c_float] * 3 + [C.c_int32] + [C.c_float] * 3 + [C.c_int64, c_fp], C.c_int),
    "dpc_jelly_x_start": ([c_fp] * 3 + [C.c_float] * 2 + [C.c_int32, C.c_int64, C.c_int64, c_fp], C.c_int),
    "dpc_jelly_step": ([c_fp] * 12 + [C.c_float] * 5 + [C.c_int32] * 4 + [C.c_int64, c_fp], C.c_int),
    "dpc_jelly_write_bd": ([c_fp] * 4 + [C.c_int32] * 3 + [C.c_int64, c_fp], C.c_int),
    "dpc_burgers_rollout": ([c_fp] * 3 + [C.c_int32] * 4 + [C.c_float] * 6 + [c_fp], C.c_int),
    "dpc_spatial_linear_attention_ex": ([c_fp] * 4 + [C.c_int32] * 3 + [C.c_float, c_fp], C.c_int),
    "dpc_linattn2d_bwd": ([c_fp] * 6 + [C.c_int32] * 3 + [C.c_float, C.c_float, c_fp], C.c_int),
    "dpc_attention2d_bwd": ([c_fp] * 4 + [C.c_int32] * 3 + [C.c_float, c_fp], C.c_int),
    "dpc_gn_silu_bwd": ([c_fp] * 5 + [C.c_int64, C.c_int64] + [c_fp] * 4 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                                                         c_fp], C.c_int),
    "dpc_layernorm_channels_bwd": ([c_fp] * 5 + [C.c_int64, C.c_int32, C.c_float, C.c_int32, c_fp], C.c_int),
    "dpc_add": ([c_fp] * 3 + [C.c_int64, c_fp], C.c_int),
    "dpc_sumpool2x2": ([c_fp, c_fp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_mean_head": ([c_fp] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_mean_head_bwd": ([c_fp] * 3 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_time_embed_f32": ([c_fp] * 8 + [C.c_int32, C.c_int32, c_fp], C.c_int),
    "dpc_time_mlp_bwd": ([c_fp] * 9 + [C.c_int32] * 3 + [c_fp], C.c_int),
    "dpc_smoke_rollout": ([c_fp] * 14 + [C.c_int32] * 4 + [C.c_double, C.c_double, C.c_int32, c_fp], C.c_int),
}

EXPORTS = tuple(_SIGNATURES)


def library_path() -> str:
    return _SO


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} not found: build it with `python -m diffphycon_b200.build` (there is no CPU/PyTorch fallback)")
        handle = C.CDLL(_SO)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.dpc_abi_version() != ABI_VERSION:
            raise RuntimeError("libdpc_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().dpc_last_error().decode()} (rc={rc})")


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor that lives on the CURRENT device (kernels are launched on the current
    device's current stream: a tensor of another device would be dereferenced on the wrong GPU)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "call through the module API (it makes the tensors' device current) or use torch.cuda.device(...)")
    return t.data_ptr()


def stream_ptr():
    """The current stream of the current device; the module entry points run under `on_device(tensor)`."""
    return torch.cuda.current_stream().cuda_stream


def on_device(t):
    """Context manager that makes the device of `t` (a tensor or a torch.device) current for the launches inside."""
    dev = t.device if isinstance(t, torch.Tensor) else torch.device(t)
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def device_guarded(fn):
    """Decorator for the public entry points: the device of the first CUDA tensor argument is made current for the call,
    so that `stream_ptr()` / the per-device launch state on the C side belong to the device the data lives on."""
    @functools.wraps(fn)
    def wrapper(*a, **k):
        for v in a:
            if isinstance(v, torch.Tensor) and v.is_cuda:
                with torch.cuda.device(v.device):
                    return fn(*a, **k)
        for v in k.values():
            if isinstance(v, torch.Tensor) and v.is_cuda:
                with torch.cuda.device(v.device):
                    return fn(*a, **k)
        return fn(*a, **k)
    return wrapper


class LaunchCounter:
    """Counts kernel launches issued through this binding (bench.py reports it as gpu_launches)."""
    count = 0
    graph_launches = 0     # CUDA-graph replays (each replays every kernel of one captured denoising step)


class Profiler:
    """Optional per-call CUDA-event timing of the launches issued through this binding (development aid):
        with Profiler() as prof: net(x, t)  ->  prof.summary() = {label: (calls, total ms)}"""
    active = None

    def __init__(self):
        self.events = []

    def __enter__(self):
        Profiler.active = self
        return self

    def __exit__(self, *a):
        Profiler.active = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for label, e0, e1 in self.events:
            n, t = out.get(label, (0, 0.0))
            out[label] = (n + 1, t + e0.elapsed_time(e1))
        return out


def _timed(label):
    def deco(fn):
        def wrapper(*a, **k):
            prof = Profiler.active
            if prof is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            lab = label
            if label == "conv":
                p = a[0]
                lab = f"conv[{'tc' if r else 'igemm'} taps={p.ntaps} cin={p.C1 + p.C2} cout={p.Cout} w={p.Wi}]"
            prof.events.append((lab, e0, e1))
            return r
        return wrapper
    return deco


@_timed("conv")
def conv(params: ConvParams, tcgen05: bool = False, tc_only: bool = False) -> bool:
    """Launch a convolution.  With tcgen05=True tries the TMA/tcgen05 kernel first; returns True if it ran.
    Falls back to the generic tensor-core implicit GEMM only on the documented 'shape not supported' code (-2); with
    tc_only=True nothing is launched in that case and False is returned."""
    L = lib()
    if tcgen05:
        rc = L.dpc_conv3d_tcgen05(C.byref(params), stream_ptr())
        if rc == 0:
            LaunchCounter.count += 1
            return True
        if rc != -2:
            check(rc, "dpc_conv3d_tcgen05")
    if tc_only:
        return False
    check(L.dpc_conv_igemm(C.byref(params), stream_ptr()), "dpc_conv_igemm")
    LaunchCounter.count += 1
    return False


@_timed("groupnorm_silu")
def groupnorm_silu(y, stats, gamma, beta, scale_shift, ss_stride, ss_off, residual, out, B, rows_per_sample, Cn, groups,
                   eps=1e-5):
    check(lib().dpc_groupnorm_silu(ptr(y), ptr(stats), ptr(gamma), ptr(beta), ptr(scale_shift), ss_stride, ss_off,
                                   ptr(residual), ptr(out), B, rows_per_sample, Cn, groups, eps, stream_ptr()),
          "dpc_groupnorm_silu")
    LaunchCounter.count += 1


@_timed("gn_fold")
def gn_fold(stats, gamma, beta, scale, shift, B, rows_per_sample, Cn, groups, eps=1e-5):
    check(lib().dpc_gn_fold(ptr(stats), ptr(gamma), ptr(beta), ptr(scale), ptr(shift), B, rows_per_sample, Cn, groups, eps,
                            stream_ptr()), "dpc_gn_fold")
    LaunchCounter.count += 1


@_timed("gn_stats_merge")
def gn_stats_merge(stats_in, stats_out, B, groups_in, groups_out):
    check(lib().dpc_gn_stats_merge(ptr(stats_in), ptr(stats_out), B, groups_in, groups_out, stream_ptr()), "dpc_gn_stats_merge")
    LaunchCounter.count += 1


@_timed("layernorm_channels")
def layernorm_channels(x, gamma, out, rows, Cn, eps=1e-5, residual=None, use_rsqrt=False):
    check(lib().dpc_layernorm_channels(ptr(x), ptr(gamma), ptr(residual), ptr(out), rows, Cn, eps, 1 if use_rsqrt else 0,
                                       stream_ptr()), "dpc_layernorm_channels")
    LaunchCounter.count += 1


@_timed("upsample_nearest2x")
def upsample_nearest2x(x, out, BF, H, W, Cn):
    check(lib().dpc_upsample_nearest2x(ptr(x), ptr(out), BF, H, W, Cn, stream_ptr()), "dpc_upsample_nearest2x")
    LaunchCounter.count += 1


@_timed("pack_input")
def pack_input(x, out, B, F, Ctot, c0, Cin, H, W, Cpad):
    check(lib().dpc_pack_input(ptr(x), ptr(out), B, F, Ctot, c0, Cin, H, W, Cpad, stream_ptr()), "dpc_pack_input")
    LaunchCounter.count += 1


@_timed("temporal_attention")
def temporal_attention(qkv, rope_cos, rope_sin, pos_bias, out, B, F, HW, heads, use_rope=True, precise=False, relative_bias=False):
    """relative_bias: pos_bias[h][i][j] is a function of j - i only (RelativePositionBias): read through a shared-memory table."""
    check(lib().dpc_temporal_attention(ptr(qkv), ptr(rope_cos), ptr(rope_sin), ptr(pos_bias), ptr(out), B, F, HW, heads,
                                       (1 if use_rope else 0) | (2 if relative_bias else 0), 1 if precise else 0, stream_ptr()),
          "dpc_temporal_attention")
    LaunchCounter.count += 1


@_timed("temporal_block_fused")
def temporal_block_fused(x, w_qkv, w_out, rope_cos, rope_sin, pos_bias, y, B, F, HW, Cn, heads, eps=1e-5) -> bool:
    """LayerNorm + to_qkv + temporal attention + to_out + residual in one launch; False if the shape is not served (-2)."""
    rc = lib().dpc_temporal_block_fused(ptr(x), ptr(w_qkv), ptr(w_out), ptr(rope_cos), ptr(rope_sin), ptr(pos_bias), ptr(y),
                                        B, F, HW, Cn, heads, eps, stream_ptr())
    if rc == -2:
        return False
    check(rc, "dpc_temporal_block_fused")
    LaunchCounter.count += 1
    return True


@_timed("spatial_attention")
def spatial_attention(qkv, out, BF, HW, heads, precise=True):
    """precise=False: q k^T and P v as TF32 tensor-core MMAs (dpc_spatial_attention_mma) when the shape is served."""
    if not precise:
        rc = lib().dpc_spatial_attention_mma(ptr(qkv), ptr(out), BF, HW, heads, stream_ptr())
        if rc != -2:
            check(rc, "dpc_spatial_attention_mma")
            LaunchCounter.count += 1
            return
    check(lib().dpc_spatial_attention(ptr(qkv), ptr(out), BF, HW, heads, stream_ptr()), "dpc_spatial_attention")
    LaunchCounter.count += 1


@_timed("final_proj")
def final_proj(x, w, bias, out, BF, HW, Cn, Cout) -> bool:
    """final 1x1x1 conv to <= 6 channels in the reference layout; False if the shape is not served (-2)."""
    rc = lib().dpc_final_proj(ptr(x), ptr(w), ptr(bias), ptr(out), BF, HW, Cn, Cout, stream_ptr())
    if rc == -2:
        return False
    check(rc, "dpc_final_proj")
    LaunchCounter.count += 1
    return True


@_timed("stem_conv_tcgen05")
def stem_conv(x, w, bias, y, B, F, H, W, Cpad, N, kt, kh, kw) -> bool:
    """init_conv on the tcgen05 sliding-window kernel; False if the shape is not served (-2)."""
    rc = lib().dpc_stem_conv_tcgen05(ptr(x), ptr(w), ptr(bias), ptr(y), B, F, H, W, Cpad, N, kt, kh, kw, stream_ptr())
    if rc == -2:
        return False
    check(rc, "dpc_stem_conv_tcgen05")
    LaunchCounter.count += 1
    return True


@_timed("spatial_linear_block_fused")
def spatial_linear_block_fused(x, w_qkv, w_out, b_out, ctx_ws, mt_ws, y, BF, HW, Cn, heads, eps=1e-5) -> bool:
    """LayerNorm + to_qkv + linear attention + to_out + residual in three launches; False if the shape is not served."""
    rc = lib().dpc_spatial_linear_block_fused(ptr(x), ptr(w_qkv), ptr(w_out), ptr(b_out), ptr(ctx_ws), ptr(mt_ws), ptr(y),
                                              BF, HW, Cn, heads, eps, stream_ptr())
    if rc == -2:
        return False
    check(rc, "dpc_spatial_linear_block_fused")
    LaunchCounter.count += 3
    return True


@_timed("spatial_linear_attention")
def spatial_linear_attention(qkv, ctx_ws, out, BF, HW, heads):
    check(lib().dpc_spatial_linear_attention(ptr(qkv), ptr(ctx_ws), ptr(out), BF, HW, heads, stream_ptr()),
          "dpc_spatial_linear_attention")
    LaunchCounter.count += 2


@_timed("time_embed")
def time_embed(t, freqs, w1, b1, w2, b2, hidden_ws, t_emb, B, dim):
    check(lib().dpc_time_embed(ptr(t), ptr(freqs), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(hidden_ws), ptr(t_emb), B, dim,
                               stream_ptr()), "dpc_time_embed")
    LaunchCounter.count += 2


@_timed("time_proj")
def time_proj(t_emb, W, bias, out, B, tdim, total):
    check(lib().dpc_time_proj(ptr(t_emb), ptr(W), ptr(bias), ptr(out), B, tdim, total, stream_ptr()), "dpc_time_proj")
    LaunchCounter.count += 1


@_timed("guided_step")
def guided_step(ddim, x, eps_joint, eps_w, noise, init, g, coefs: StepCoefs, x_out, x_start_out, B, F, H, W):
    fn = lib().dpc_ddim_guided_step if ddim else lib().dpc_ddpm_guided_step
    check(fn(ptr(x), ptr(eps_joint), ptr(eps_w), ptr(noise), ptr(init), ptr(g), 1 if g is None else 0, C.byref(coefs),
             ptr(x_out), ptr(x_start_out), B, F, H, W, stream_ptr()), "dpc_guided_step")
    LaunchCounter.count += 1


@_timed("sampler_prepare")
def sampler_prepare(t_table, c_table, step_index, nsteps, tt, B, cur):
    check(lib().dpc_sampler_prepare(ptr(t_table), ptr(c_table), ptr(step_index), nsteps, ptr(tt), B, ptr(cur), stream_ptr()),
          "dpc_sampler_prepare")
    LaunchCounter.count += 1


@_timed("guided_step")
def guided_step_dev(ddim, x, eps_joint, eps_w, noise, init, coefs_host: StepCoefs, coefs_dev, x_out, x_start_out, B, F, H, W):
    check(lib().dpc_guided_step_dev(1 if ddim else 0, ptr(x), ptr(eps_joint), ptr(eps_w), ptr(noise), ptr(init), C.byref(coefs_host),
                                    ptr(coefs_dev), ptr(x_out), ptr(x_start_out), B, F, H, W, stream_ptr()), "dpc_guided_step_dev")
    LaunchCounter.count += 1


@_timed("renoise")
def renoise(x, z, a, b, out):
    check(lib().dpc_renoise(ptr(x), ptr(z), a, b, ptr(out), x.numel(), stream_ptr()), "dpc_renoise")
    LaunchCounter.count += 1


@_timed("predict_x_start")
def predict_x_start(x, eps, sr, srm1, clip, out):
    check(lib().dpc_predict_x_start(ptr(x), ptr(eps), sr, srm1, 1 if clip else 0, ptr(out), x.numel(), stream_ptr()),
          "dpc_predict_x_start")
    LaunchCounter.count += 1


@_timed("smoke_rollout")
def smoke_rollout(fluid_mask, velocity_mask, init_velocity, init_density, c1, c2, vel_ws, x_ws, dens_ws, densitys,
                  zero_densitys, velocitys, smoke_out, iterations, B, nt, nx, T, dt, accuracy, max_iterations):
    check(lib().dpc_smoke_rollout(ptr(fluid_mask), ptr(velocity_mask), ptr(init_velocity), ptr(init_density), ptr(c1), ptr(c2),
                                  ptr(vel_ws), ptr(x_ws), ptr(dens_ws), ptr(densitys), ptr(zero_densitys), ptr(velocitys),
                                  ptr(smoke_out), ptr(iterations), B, nt, nx, T, float(dt), float(accuracy),
                                  int(max_iterations), stream_ptr()), "dpc_smoke_rollout")
    LaunchCounter.count += 1


def smoke_eval_sums(pred, densitys, velocitys, smoke_out, sums, B, F, S, T, mask_lo, mask_hi):
    check(lib().dpc_smoke_eval_sums(ptr(pred), ptr(densitys), ptr(velocitys), ptr(smoke_out), ptr(sums), B, F, S, T, mask_lo,
                                    mask_hi, stream_ptr()), "dpc_smoke_eval_sums")
    LaunchCounter.count += 1


@_timed("burgers_rollout")
def burgers_rollout(u0, f, traj, N, s, Nt, steps, t0, t1, d0, d1, d2, dt):
    check(lib().dpc_burgers_rollout(ptr(u0), ptr(f), ptr(traj), N, s, Nt, steps, t0, t1, d0, d1, d2, dt, stream_ptr()),
          "dpc_burgers_rollout")
    LaunchCounter.count += 1


@_timed("burgers_model_output")
def burgers_model_output(x, eps1, eps2, out, x_start, mode, coef, beta, sr, srm1, Cn, plane):
    check(lib().dpc_burgers_model_output(ptr(x), ptr(eps1), ptr(eps2), ptr(out), ptr(x_start), mode, coef, beta, sr, srm1, Cn,
                                         plane, x.numel(), stream_ptr()), "dpc_burgers_model_output")
    LaunchCounter.count += 1


@_timed("ddpm_posterior_step")
def ddpm_posterior_step(x, eps, g, noise, x_out, x_start_out, pred_noise_out, gscale, sr, srm1, clip, c1, c2, sigma):
    check(lib().dpc_ddpm_posterior_step(ptr(x), ptr(eps), ptr(g), ptr(noise), ptr(x_out), ptr(x_start_out), ptr(pred_noise_out),
                                        gscale, sr, srm1, 1 if clip else 0, c1, c2, sigma, x.numel(), stream_ptr()),
          "dpc_ddpm_posterior_step")
    LaunchCounter.count += 1


@_timed("jelly_x_start")
def jelly_x_start(x, eps, x_start, sr, srm1, clip):
    """x [B,F,7,H,W], eps / x_start [B,F,4,H,W]."""
    B, F, _, H, W = x.shape
    check(lib().dpc_jelly_x_start(ptr(x), ptr(eps), ptr(x_start), sr, srm1, 1 if clip else 0, B * F, H * W, stream_ptr()),
          "dpc_jelly_x_start")
    LaunchCounter.count += 1


@_timed("jelly_step")
def jelly_step(x, x_start, eps, eps_w, g, noise, state_0, thetas_0, x_next, x_w, dtheta, theta_mean, ga, gb, c1, c2, sigma,
               ddim, cond_steps):
    B, F, _, H, W = x.shape
    check(lib().dpc_jelly_step(ptr(x), ptr(x_start), ptr(eps), ptr(eps_w), ptr(g), ptr(noise), ptr(state_0), ptr(thetas_0),
                               ptr(x_next), ptr(x_w), ptr(dtheta), ptr(theta_mean), ga, gb, c1, c2, sigma, 1 if ddim else 0,
                               B, F, cond_steps, H * W, stream_ptr()), "dpc_jelly_step")
    LaunchCounter.count += 1


@_timed("jelly_write_bd")
def jelly_write_bd(pred_bd, bd_0, x_next, x_w, cond_steps):
    B, F, _, H, W = x_next.shape
    check(lib().dpc_jelly_write_bd(ptr(pred_bd), ptr(bd_0), ptr(x_next), ptr(x_w), B, F, cond_steps, H * W, stream_ptr()),
          "dpc_jelly_write_bd")
    LaunchCounter.count += 1


# ---- jellyfish surrogate networks (include/dpc_b200.h, "Jellyfish surrogate networks") -------------------------------------
ATT_SCALE = 32 ** -0.5


@_timed("spatial_linear_attention")
def spatial_linear_attention_ex(qkv, ctx_ws, kstat, out, BF, HW, heads, v_scale):
    check(lib().dpc_spatial_linear_attention_ex(ptr(qkv), ptr(ctx_ws), ptr(kstat), ptr(out), BF, HW, heads, v_scale, stream_ptr()),
          "dpc_spatial_linear_attention_ex")
    LaunchCounter.count += 2


@_timed("linattn2d_bwd")
def linattn2d_bwd(qkv, ctx, kstat, dout, dctx_ws, dqkv, BF, HW, heads, v_scale):
    check(lib().dpc_linattn2d_bwd(ptr(qkv), ptr(ctx), ptr(kstat), ptr(dout), ptr(dctx_ws), ptr(dqkv), BF, HW, heads, ATT_SCALE,
                                  v_scale, stream_ptr()), "dpc_linattn2d_bwd")
    LaunchCounter.count += 2


@_timed("attention2d_bwd")
def attention2d_bwd(qkv, out, dout, dqkv, BF, HW, heads):
    check(lib().dpc_attention2d_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(dqkv), BF, HW, heads, ATT_SCALE, stream_ptr()),
          "dpc_attention2d_bwd")
    LaunchCounter.count += 1


@_timed("gn_silu_bwd")
def gn_silu_bwd(y, stats, gamma, beta, scale_shift, ss_stride, ss_off, dout, dy, sums_ws, dss, B, rows_per_sample, Cn, groups,
                eps=1e-5):
    check(lib().dpc_gn_silu_bwd(ptr(y), ptr(stats), ptr(gamma), ptr(beta), ptr(scale_shift), ss_stride, ss_off, ptr(dout), ptr(dy),
                                ptr(sums_ws), ptr(dss), B, rows_per_sample, Cn, groups, eps, stream_ptr()), "dpc_gn_silu_bwd")
    LaunchCounter.count += 2


@_timed("layernorm_channels_bwd")
def layernorm_channels_bwd(x, gamma, dy, add, dx, rows, Cn, eps=1e-5, use_rsqrt=True):
    check(lib().dpc_layernorm_channels_bwd(ptr(x), ptr(gamma), ptr(dy), ptr(add), ptr(dx), rows, Cn, eps, 1 if use_rsqrt else 0,
                                           stream_ptr()), "dpc_layernorm_channels_bwd")
    LaunchCounter.count += 1


@_timed("add")
def add(a, b, out, n):
    check(lib().dpc_add(ptr(a), ptr(b), ptr(out), n, stream_ptr()), "dpc_add")
    LaunchCounter.count += 1


@_timed("sumpool2x2")
def sumpool2x2(dy, dx, N, H, W, Cn):
    check(lib().dpc_sumpool2x2(ptr(dy), ptr(dx), N, H, W, Cn, stream_ptr()), "dpc_sumpool2x2")
    LaunchCounter.count += 1


@_timed("mean_head")
def mean_head(x, W, bias, out, N, HW, Cn, O):
    check(lib().dpc_mean_head(ptr(x), ptr(W), ptr(bias), ptr(out), N, HW, Cn, O, stream_ptr()), "dpc_mean_head")
    LaunchCounter.count += 1


@_timed("mean_head_bwd")
def mean_head_bwd(dout, W, dx, N, HW, Cn, O):
    check(lib().dpc_mean_head_bwd(ptr(dout), ptr(W), ptr(dx), N, HW, Cn, O, stream_ptr()), "dpc_mean_head_bwd")
    LaunchCounter.count += 1


@_timed("time_embed")
def time_embed_f32(t, freqs, w1, b1, w2, b2, hidden_ws, t_emb, B, dim):
    check(lib().dpc_time_embed_f32(ptr(t), ptr(freqs), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(hidden_ws), ptr(t_emb), B, dim,
                                   stream_ptr()), "dpc_time_embed_f32")
    LaunchCounter.count += 2


@_timed("time_mlp_bwd")
def time_mlp_bwd(t, freqs, w1, b1, w2, w_proj, t_emb, dss, dt, B, dim, total):
    check(lib().dpc_time_mlp_bwd(ptr(t), ptr(freqs), ptr(w1), ptr(b1), ptr(w2), ptr(w_proj), ptr(t_emb), ptr(dss), ptr(dt), B, dim,
                                 total, stream_ptr()), "dpc_time_mlp_bwd")
    LaunchCounter.count += 1
