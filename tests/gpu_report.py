"""Prints a per-case error table on the GPU box (development aid; run via gpurun)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffphycon_b200 as dpc
from oracle import unet3d_oracle as uo
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"unet_small_c6": dict(dim=32, dim_mults=(1, 2), channels=6), "unet_small_c2": dict(dim=32, dim_mults=(1, 2), channels=2),
         "unet_smoke_arch": dict(dim=64, dim_mults=(1, 2, 4), channels=6), "unet_jelly_arch": dict(dim=32, dim_mults=(1, 2), channels=7, out_dim=4)}
for name, kw in CASES.items():
    z = np.load(os.path.join(G, name + ".npz"))
    for prec in ("3xtf32", "tf32"):
        for tc in (False, True):
            net = dpc.Unet3D_with_Conv3D(**kw); net.load_state_dict(uo.make_params(uo.UnetCfg(**kw), int(z["seed"])))
            net.precision = prec; net.use_tcgen05 = tc; net.taps = {}; net = net.cuda()
            try:
                y = net(torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["t"]).cuda()).cpu()
            except Exception as e:
                print(name, prec, tc, "EXC", e); continue
            ref = torch.from_numpy(z["y"])
            msg = [f"{name} {prec} tc={tc} out_err={(y-ref).abs().max().item()/max(1,ref.abs().max().item()):.2e}"]
            for k in z.files:
                if k.startswith("act/"):
                    a = torch.from_numpy(z[k]); msg.append(f"{k[4:]}={(net.taps[k[4:]].cpu()-a).abs().max().item()/max(1,a.abs().max().item()):.1e}")
            print(" ".join(msg), flush=True)
