"""Short driver for ncu: a few launches of the fused ICNN decode kernel at B=65536 (one GPU)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import module, ops, utils as vutils, _C

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32")
ap.add_argument("--H", type=int, default=1024)
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(0)
ic = vutils.trained_like_icnn_(module.ICNN(2, a.H).cuda(), rng)
P = [t.detach() for t in ic._flat_params()]
z = torch.randn(a.B, 2, device="cuda")
v = torch.randn(a.B, 2, device="cuda")
prec = _C.PRECISIONS[a.precision]
ws = ops.icnn_prepare(P, 2, a.H, 0, prec, a.B, a.bwd)
for _ in range(a.iters):
    psi, xhat, m1, m2 = ops.icnn_decode_fwd(z, ws, 2, a.H, 0, 0.1, prec, True, True, a.bwd)
    if a.bwd:
        ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws, 2, a.H, 0, 0.1, prec)
torch.cuda.synchronize()
print("ok", float(xhat.abs().mean()))
