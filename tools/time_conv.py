"""Development aid: time the tcgen05 conv alone (CUDA events) for a few layer shapes."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
dev = "cuda"
def run(B, Fr, S, cin, cout, tc=True, reps=5):
    xa = torch.randn(B, Fr, S, S, cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5
    wp, _, _ = packing.pack_conv3d(w); bias = torch.zeros(cout, device=dev); taps = packing.tap_table(3, 3, 3, S, S, dev)
    y = torch.empty(B, Fr, S, S, cout, device=dev); stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
    p = _lib.ConvParams()
    p.x1, p.C1, p.C2 = xa.data_ptr(), cin, 0
    p.w, p.bias, p.y, p.taps, p.ntaps = wp.data_ptr(), bias.data_ptr(), y.data_ptr(), taps.data_ptr(), 27
    p.gn_stats, p.gn_groups = stats.data_ptr(), 8
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = B, Fr, S, S, Fr, S, S
    p.st = p.sh = p.sw = 1; p.pt = p.ph = p.pw = 1; p.oh_mul = p.ow_mul = 1; p.Hfull, p.Wfull = S, S
    p.Cout, p.Npad, p.Kpad = cout, wp.shape[0], wp.shape[1]
    for _ in range(2): _lib.conv(p, tcgen05=tc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.conv(p, tcgen05=tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * Fr * S * S * cin * cout * 27
    print(f"PAIR={os.environ.get('DPC_TC_PAIR','1')} DBG={os.environ.get('DPC_TC_DEBUG','0')} B={B} S={S} {cin}->{cout}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shapes = [(64, 64, 64), (32, 128, 128), (16, 256, 256)]
if len(sys.argv) > 2 and sys.argv[2] == "all":
    shapes += [(64, 128, 64), (16, 512, 128), (16, 128, 128), (32, 256, 64), (32, 64, 64), (32, 64, 128), (16, 128, 256)]
for S_, ci, co in shapes:
    run(B, 32, S_, ci, co)
