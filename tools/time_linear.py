"""Development aid: time the 1x1x1 / Linear mode of the tcgen05 conv kernel (CUDA events) for the layer shapes of the step."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
dev = "cuda"
def run(B, Fr, S, cin, cout, residual=False, tc=True, reps=5):
    xa = torch.randn(B, Fr, S, S, cin, device=dev)
    w = torch.randn(cout, cin, device=dev) / cin ** 0.5
    wp = packing.pack_linear(w); taps = packing.tap_table(1, 1, 1, S, S, dev)
    y = torch.empty(B, Fr, S, S, cout, device=dev)
    res = torch.randn(B, Fr, S, S, cout, device=dev) if residual else None
    p = _lib.ConvParams()
    p.x1, p.C1, p.C2 = xa.data_ptr(), cin, 0
    p.w, p.bias, p.y, p.taps, p.ntaps = wp.data_ptr(), None, y.data_ptr(), taps.data_ptr(), 1
    p.residual = res.data_ptr() if residual else None
    p.gn_stats, p.gn_groups = None, 0
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = B, Fr, S, S, Fr, S, S
    p.st = p.sh = p.sw = 1; p.pt = p.ph = p.pw = 0; p.oh_mul = p.ow_mul = 1; p.Hfull, p.Wfull = S, S
    p.Cout, p.Npad, p.Kpad = cout, wp.shape[0], wp.shape[1]
    for _ in range(2): _lib.conv(p, tcgen05=tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.conv(p, tcgen05=tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows = B * Fr * S * S
    gb = rows * 4.0 * (cin + cout * (2 if residual else 1)) / 1e9
    print(f"linear tc={tc} B={B} S={S} {cin}->{cout} res={residual}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s algorithmic  {2.0*rows*cin*cout/ms/1e9:.0f} TFLOP/s", flush=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for tc in (True, False):
    run(B, 32, 16, 128, 256, True, tc); run(B, 32, 16, 256, 384, False, tc); run(B, 32, 32, 128, 384, False, tc)
    run(B, 32, 32, 128, 128, True, tc); run(B, 32, 64, 128, 64, False, tc); run(B, 32, 16, 128, 128, True, tc)
