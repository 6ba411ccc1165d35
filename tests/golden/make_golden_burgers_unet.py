"""Golden vectors for the Burgers `Unet2D` from the UNMODIFIED reference
model/burgers_1d/unet.py (build container only).  Weights: oracle.param_gen.make_params over the reference's own
state_dict inventory, loaded with strict=True (which also pins the key/shape inventory of diffphycon_b200.Unet2D)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import param_gen, ref_import  # noqa: E402
from diffphycon_b200.burgers_unet import Unet2D  # noqa: E402

CASES = {
    "burgers_unet_uw": (dict(dim=64, dim_mults=(1, 2, 4), channels=2, resnet_block_groups=1), 21),
    "burgers_unet_w": (dict(dim=32, dim_mults=(1, 2, 4, 8), channels=2, resnet_block_groups=1), 22),
    "burgers_unet_g8": (dict(dim=32, dim_mults=(1, 2), channels=2, resnet_block_groups=8, out_dim=3), 23),
}
m = ref_import.burgers_unet_module()
for name, (kw, seed) in CASES.items():
    net = m.Unet2D(**kw)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(v.shape) for k, v in Unet2D(**kw).state_dict().items()}
    assert list(shapes) == list(mine) and shapes == mine, "state_dict inventory differs from the reference"
    net.load_state_dict(param_gen.make_params(shapes, seed), strict=True)
    net.eval()
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(2, 2, 16, 128, generator=g)
    t = torch.tensor([7, 901])
    with torch.no_grad():
        y = net(x, t)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), t=t.numpy(), y=y.numpy(), seed=np.int64(seed))
    print(name, tuple(y.shape), float(y.abs().mean()))
