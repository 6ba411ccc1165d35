"""Development aid: quad-mode conv with delta weights (which input frame / tap / channel lands where)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffphycon_b200 import _lib, packing
dev = "cuda"
B, Fr, S, cin, cout = 1, 8, 32, 64, 64
def run(w, xa):
    wp, _, _ = packing.pack_conv3d(w); bias = torch.zeros(cout, device=dev); taps = packing.tap_table(3, 3, 3, S, S, dev)
    y = torch.full((B, Fr, S, S, cout), float("nan"), device=dev)
    p = _lib.ConvParams()
    p.x1, p.C1, p.C2 = xa.data_ptr(), cin, 0
    p.w, p.bias, p.y, p.taps, p.ntaps = wp.data_ptr(), bias.data_ptr(), y.data_ptr(), taps.data_ptr(), 27
    p.gn_stats, p.gn_groups = None, 0
    p.B, p.Fi, p.Hi, p.Wi, p.Fo, p.Ho, p.Wo = B, Fr, S, S, Fr, S, S
    p.st = p.sh = p.sw = 1; p.pt = p.ph = p.pw = 1; p.oh_mul = p.ow_mul = 1; p.Hfull, p.Wfull = S, S
    p.Cout, p.Npad, p.Kpad = cout, wp.shape[0], wp.shape[1]
    assert _lib.conv(p, tcgen05=True)
    torch.cuda.synchronize()
    return y
xa = torch.zeros(B, Fr, S, S, cin, device=dev)
for f in range(Fr): xa[0, f] = f + 1
xa[..., 1] *= 10          # channel 1 = 10 x
for dt in range(3):
    w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
    for c in range(cout): w[c, c, dt, 1, 1] = 1.0
    y = run(w, xa)
    print(f"delta at dt={dt}: y[f, 10,10, ch0] =", [round(float(y[0, f, 10, 10, 0]), 2) for f in range(Fr)], " ch1 =", [round(float(y[0, f, 10, 10, 1]), 1) for f in range(Fr)], " ch40 =", [round(float(y[0, f, 10, 10, 40]), 1) for f in range(Fr)])
    print("   expected ch0:", [float(f + dt) if 1 <= f + dt <= Fr else 0.0 for f in range(Fr)])
w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
for c in range(cout): w[c, (c + 1) % cin, 1, 1, 1] = 1.0     # y[c] = x[c+1]
y = run(w, xa)
print("channel shift: y[f=2, ch0..3] =", [round(float(y[0, 2, 10, 10, c]), 1) for c in range(4)], "expected [30, 3, 3, 3]")
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
xa = torch.randn(B, Fr, S, S, cin, device=dev)
w = torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5
y = run(w, xa)
ref = F.conv3d(xa.permute(0, 4, 1, 2, 3), w, padding=1).permute(0, 2, 3, 4, 1)
err = (y - ref).abs().amax(dim=-1)[0]          # [F, H, W]
print("max err per frame:", [round(float(err[f].max()), 3) for f in range(Fr)])
print("max err per row (frame 1):", [round(float(err[1, h].max()), 2) for h in range(S)])
print("max err per col (frame 1):", [round(float(err[1, :, x].max()), 2) for x in range(S)])
for dh, dw in ((0, 1), (1, 0), (2, 2)):
    w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
    for c in range(cout): w[c, c, 1, dh, dw] = 1.0
    y = run(w, xa); ref = F.conv3d(xa.permute(0, 4, 1, 2, 3), w, padding=1).permute(0, 2, 3, 4, 1)
    err = (y - ref).abs().amax(dim=-1)[0]
    print(f"delta (1,{dh},{dw}): max err {float(err.max()):.3f}; rows with err:", [h for h in range(S) if float(err[:, h].max()) > 1e-2][:40])
print("---- spatial identity check (delta at centre tap, random x)")
w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
for c in range(cout): w[c, c, 1, 1, 1] = 1.0
y = run(w, xa)
for f in (0, 1, 5):
    # find for a few output pixels which input pixel of the same frame they equal (channel 0..7 signature)
    for (h, x_) in ((0, 0), (3, 5), (10, 10), (20, 7), (31, 31)):
        sig = y[0, f, h, x_, :8]
        d = (xa[0, :, :, :, :8] - sig).abs().sum(-1)        # [F, H, W]
        idx = int(d.argmin()); ff, hh, ww = idx // (S * S), (idx // S) % S, idx % S
        print(f"  y[f={f},h={h},w={x_}] matches x[f={ff},h={hh},w={ww}] (dist {float(d.min()):.3f})")
