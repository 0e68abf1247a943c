"""Short driver for ncu: the tiled all-pairs Lipschitz kernel at N points (d = Dx = 2), a few launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_song_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
torch.manual_seed(0)
X = torch.randn(N, 2, device="cuda")
Y = torch.tanh(X @ torch.randn(2, 2, device="cuda")) * 3
for _ in range(3):
    st, _ = ops.lipschitz_allpairs(X, Y, 1e-3)
torch.cuda.synchronize()
print("ok", st.tolist())
