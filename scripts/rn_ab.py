"""A/B: does tcgen05 kind::tf32 ignore the 13 low mantissa bits of K-major shared-memory operands?  Runs the same TF32
forward with the library given by B200VAE_LIB and dumps psi/xhat; compare the two dumps bitwise."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import module, ops, utils as vutils
rng = np.random.default_rng(0)
ic = vutils.trained_like_icnn_(module.ICNN(2, 1024).cuda(), rng)
P = [t.detach() for t in ic._flat_params()]
z = torch.tensor(rng.normal(0, 1, (4096, 2)), dtype=torch.float32, device="cuda")
v = torch.tensor(rng.normal(0, 1, (4096, 2)), dtype=torch.float32, device="cuda")
ws = ops.icnn_prepare(P, 2, 1024, 0, 1, 4096, True)
psi, xhat, m1, m2 = ops.icnn_decode_fwd(z, ws, 2, 1024, 0, 0.1, 1, True, True, True)
dz, g = ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws, 2, 1024, 0, 0.1, 1)
torch.cuda.synchronize()
np.savez(sys.argv[1], psi=psi.cpu().numpy(), xhat=xhat.cpu().numpy(), m1=m1.cpu().numpy(), dz=dz.cpu().numpy(), W1=g[7].cpu().numpy(), W0=g[6].cpu().numpy())
print("saved", sys.argv[1])
