"""Short driver for an ncu launch list: eager train steps of LIDVAE(dataset='mnist') at batch 256 (BASELINE configs[3],
ICNN decoder) through train.DataParallelTrainer.step; torch.cuda.profiler brackets the LAST step (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import model, train, utils as vutils

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
m = model.LIDVAE(dataset="mnist", inverse_lipschitz=0.2, beta=0.001, precision=prec).cuda().train()
rng = np.random.default_rng(1)
for ic in (m.decoder[0], m.decoder[1]):
    vutils.trained_like_icnn_(ic, rng)
tr = train.DataParallelTrainer(m, lr=1e-3)
x = torch.rand(256, 1, 28, 28, device="cuda")
for _ in range(4):
    tr.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = tr.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out[0]))
