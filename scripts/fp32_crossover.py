"""FP32: fused sample-stationary kernels (one CTA per 128 samples) vs the tiled GEMM chain, decode + backward, d=2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import module, ops, utils as vutils
rng = np.random.default_rng(0)
for H in (512, 1024):
    ic = vutils.trained_like_icnn_(module.ICNN(2, H).cuda(), rng)
    P = list(ic._flat_params())
    for B in (256, 1024, 4096, 8192, 16384, 32768):
        z = torch.randn(B, 2, device="cuda", requires_grad=True); v = torch.randn(B, 2, device="cuda")
        res = {}
        for name, fn in (("fused", ops.IcnnBrenierFn), ("tiled", ops.IcnnBrenierWideFn)):
            def step():
                for p in P: p.grad = None
                z.grad = None
                _, x = fn.apply(z, 0.1, 0, 0, *P)
                (x * v).sum().backward()
            for _ in range(2): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): step()
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 5
        print(f"H={H} B={B}: fused {res['fused']:.3f} ms, tiled {res['tiled']:.3f} ms")
