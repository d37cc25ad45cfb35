import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, wan_vae
from apex_studio_b200.vae import AutoencoderKLWan, WanVAEConfig
DEV = "cuda"
for base_dim, shape in [(96, (16, 2, 18, 16)), (96, (16, 2, 16, 16)), (96, (16, 1, 18, 16)), (64, (16, 2, 18, 16))]:
    vae = AutoencoderKLWan(WanVAEConfig(base_dim=base_dim))
    vae.load_state_dict(wan_vae.make_weights(base_dim=base_dim, seed=3), device=DEV)
    z = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).to(DEV, torch.bfloat16)
    a1, a2 = vae.decode_tile(z).clone(), vae.decode_tile(z).clone()
    b1, b2 = vae.decode_tile_py(z).clone(), vae.decode_tile_py(z).clone()
    d = lambda x, y: (x.float() - y.float()).abs().max().item()
    print(base_dim, shape, "C-C", d(a1, a2), "Py-Py", d(b1, b2), "C-Py", d(a1, b1), "finite", bool(torch.isfinite(a1.float()).all()), bool(torch.isfinite(b1.float()).all()),
          "frac_diff", (a1 != b1).float().mean().item())
