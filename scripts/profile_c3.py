"""Short driver for ncu: eager train steps of BASELINE configs[2] (LIDVAE, encoder 128-64-64-32-16-8-4-2, batch 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import model, train, utils as vutils
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
m = model.LIDVAE(inverse_lipschitz=0.2, beta=0.001, dataset="pinwheel", hidden_channels=[128, 64, 64, 32, 16, 8, 4, 2], precision=prec).cuda()
rng = np.random.default_rng(1)
for ic in (m.decoder[0], m.decoder[1]):
    vutils.trained_like_icnn_(ic, rng)
tr = train.DataParallelTrainer(m, lr=1e-3)
x = torch.randn(B, 2, device="cuda")
for _ in range(int(os.environ.get("STEPS", "4"))):
    tr.step(x)
torch.cuda.synchronize()
print("ok")
