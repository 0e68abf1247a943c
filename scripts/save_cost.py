"""Forward time of the 3xTF32 pair kernel with / without keeping the GEMM2 accumulators (B200VAE_SAVE_GX1=0 disables)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import module, ops, utils as vutils
B = 65536
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for H in (512, 1024):
    rng = np.random.default_rng(0)
    ic = vutils.trained_like_icnn_(module.ICNN(2, H).cuda(), rng)
    P = [t.detach() for t in ic._flat_params()]
    z = torch.randn(B, 2, device="cuda"); v = torch.randn(B, 2, device="cuda")
    ws = ops.icnn_prepare(P, 2, H, 0, 3, B, True)
    def t(fn):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return np.mean(ts)
    out = {}
    def fwd(): out["r"] = ops.icnn_decode_fwd(z, ws, 2, H, 0, 0.1, 3, True, True, True)
    tf = t(fwd)
    _, _, m1, m2 = out["r"]
    tb = t(lambda: ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws, 2, H, 0, 0.1, 3))
    print(f"SAVE={os.environ.get('B200VAE_SAVE_GX1', '1')} H={H}: training forward {tf:.3f} ms, backward {tb:.3f} ms")
