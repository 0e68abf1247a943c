"""Short driver for ncu: one decode + backward of the MNIST-shaped decoder (ICNN(32,512) -> ICNN(784,1024)) through the
wide-input tcgen05 kernels (csrc/icnn_wide_tc.cu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import module

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32x3")
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda")
rng = np.random.default_rng(0)
ics = []
for d, H in ((32, 512), (784, 1024)):
    ic = module.ICNN(d, H, precision=a.precision).to(dev)
    with torch.no_grad():
        ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
        ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
        ic.A[0].bias.copy_(torch.tensor(rng.normal(-0.3, 1.0, (H,)), dtype=torch.float32))
    ics.append(ic)
z = torch.randn(a.B, 32, device=dev, requires_grad=True)
vy = torch.randn(a.B, 784, device=dev)
for _ in range(a.iters):
    _, x1 = ics[0].brenier(z, 0.1)
    _, y = ics[1].brenier(x1, 0.1)
    (y * vy).sum().backward()
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
