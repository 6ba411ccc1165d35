"""Development aid: time the tcgen05 stem conv (dpc_stem_conv_tcgen05) at the metric shape."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Fr, H, W = 32, 64, 64
for C in (6, 2):
    cpad = packing.round_up(C, 4)
    xin = torch.randn(B, Fr, H, W, cpad, device="cuda")
    w = torch.randn(64, C, 7, 7, 7, device="cuda") / (343 * C) ** 0.5
    ws = packing.pack_stem_conv(w, cpad); bias = torch.zeros(64, device="cuda")
    y = torch.empty(B, Fr, H, W, 64, device="cuda")
    for _ in range(2): assert _lib.stem_conv(xin, ws, bias, y, B, Fr, H, W, cpad, 64, 7, 7, 7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.stem_conv(xin, ws, bias, y, B, Fr, H, W, cpad, 64, 7, 7, 7)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"stem C={C} (cpad {cpad}) B={B}: {ms:.3f} ms  issued {2.0*B*Fr*H*70*64*49*cpad*8/ms/1e9:.0f} TFLOP/s, useful {2.0*B*Fr*H*W*64*343*C/ms/1e9:.0f} TFLOP/s")
