"""Host-side weight repacking for the CUDA kernels (done once per parameter version, on the device with torch ops).

Layouts consumed by csrc/conv_igemm.cu (see include/dpc_b200.h):
    packed weight  [Npad][Kpad] fp32, K-major, k = tap*Cin + ci, tap = (dt*kh + dh)*kw + dw
    tap table      int32 [ntaps][4] = (dt, dh, dw, (dt*Hi + dh)*Wi + dw)
Reference weight layouts: nn.Conv3d [Cout,Cin,kt,kh,kw] (conv3d.py:403, :192), nn.ConvTranspose3d [Cin,Cout,1,4,4]
(conv3d.py:159-160), nn.Linear [N,K] (conv3d.py:288-289), nn.Conv2d 1x1 [N,K,1,1] (conv3d.py:240-241).
"""
from __future__ import annotations

import torch


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def tf32_round(w: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away) to the 10-bit TF32 mantissa, like cvt.rna.tf32.f32; the tensor cores then see
    exactly representable operands on the weight side."""
    bits = w.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def _pad2(mat: torch.Tensor, npad: int, kpad: int) -> torch.Tensor:
    out = torch.zeros(npad, kpad, dtype=torch.float32, device=mat.device)
    out[: mat.shape[0], : mat.shape[1]] = mat
    return out.contiguous()


def npad_of(cout: int) -> int:
    return 64 if cout <= 64 else round_up(cout, 128)


def pack_conv3d(weight: torch.Tensor, cin_pad: int | None = None, tf32: bool = True):
    """[Cout,Cin,kt,kh,kw] -> packed [Npad][Kpad]; returns (packed, ntaps, cin_eff)."""
    cout, cin, kt, kh, kw = weight.shape
    cin_eff = cin if cin_pad is None else cin_pad
    w = weight.detach().float().permute(0, 2, 3, 4, 1)  # [Cout, kt, kh, kw, Cin]
    if cin_eff != cin:
        w = torch.nn.functional.pad(w, (0, cin_eff - cin))
    mat = w.reshape(cout, kt * kh * kw * cin_eff)
    if tf32:
        mat = tf32_round(mat)
    return _pad2(mat, npad_of(cout), round_up(mat.shape[1], 32)), kt * kh * kw, cin_eff


def pack_stem_conv(weight: torch.Tensor, cin_pad: int, tf32: bool = True):
    """[Cout,Cin,kt,kh,7] -> [Cout][kt*kh*(cin_pad/4)*32] for csrc/stem_conv_tcgen05.cu: k = (((dt*kh + dh)*P + plane)*8 + dw)*4
    + c with channel = 4*plane + c; the dw = 7 column and the padded channels are zero."""
    cout, cin, kt, kh, kw = weight.shape
    assert kw == 7 and cin_pad % 4 == 0 and cin_pad >= cin
    w = weight.detach().float().permute(0, 2, 3, 4, 1)                    # [Cout, kt, kh, kw, Cin]
    w = torch.nn.functional.pad(w, (0, cin_pad - cin, 0, 8 - kw))         # [Cout, kt, kh, 8, cin_pad]
    w = w.reshape(cout, kt, kh, 8, cin_pad // 4, 4).permute(0, 1, 2, 4, 3, 5).reshape(cout, -1)
    return (tf32_round(w) if tf32 else w).contiguous()


def pack_linear(weight: torch.Tensor, tf32: bool = True):
    """[N,K] (or [N,K,1,1] / [N,K,1,1,1]) -> packed [Npad][Kpad]."""
    n, k = weight.shape[0], weight.shape[1]
    mat = weight.detach().float().reshape(n, k)
    if tf32:
        mat = tf32_round(mat)
    return _pad2(mat, npad_of(n), round_up(k, 32))


def pack_conv_transpose_1x4x4(weight: torch.Tensor, tf32: bool = True):
    """ConvTranspose3d(C, C, (1,4,4), stride (1,2,2), padding (0,1,1)) as four 1x2x2 convolutions, one per output parity
    class (ph, pw): out[2i+ph, 2j+pw] = sum_{dh,dw} in[i + dh - (1-ph), j + dw - (1-pw)] * w[:, :, 0, kh, kw] with
    kh = 3 - 2*dh (ph = 0) or 2 - 2*dh (ph = 1), same for kw.  Returns {(ph, pw): packed [Npad][Kpad]}."""
    cin, cout = weight.shape[0], weight.shape[1]
    w = weight.detach().float()
    out = {}
    for ph in (0, 1):
        for pw in (0, 1):
            taps = []
            for dh in (0, 1):
                for dw in (0, 1):
                    kh = 3 - 2 * dh if ph == 0 else 2 - 2 * dh
                    kw = 3 - 2 * dw if pw == 0 else 2 - 2 * dw
                    taps.append(w[:, :, 0, kh, kw].t())  # [Cout, Cin]
            mat = torch.stack(taps, dim=1).reshape(cout, 4 * cin)  # k = tap*Cin + ci
            if tf32:
                mat = tf32_round(mat)
            out[(ph, pw)] = _pad2(mat, npad_of(cout), round_up(4 * cin, 32))
    return out


def tap_table(kt: int, kh: int, kw: int, hi: int, wi: int, device) -> torch.Tensor:
    rows = []
    for dt in range(kt):
        for dh in range(kh):
            for dw in range(kw):
                rows.append((dt, dh, dw, (dt * hi + dh) * wi + dw))
    return torch.tensor(rows, dtype=torch.int32, device=device).contiguous()
