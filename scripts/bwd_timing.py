"""Per-call timing of the tensor-core backward (rows + dP0 + finalize) at B = 65536: python scripts/bwd_timing.py [precs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import icnn_oracle as io
from vae_song_b200 import ops, _C

precs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 3]
B = 65536
for H in (512, 1024):
    rng = np.random.default_rng(5)
    p = io.random_params(rng, 2, H, np.float64, "mixed")
    P = [torch.tensor(np.asarray(p[k], np.float32), device="cuda") for k in io.PARAM_KEYS]
    z = torch.randn(B, 2, device="cuda")
    for vs in (1.0, 1e-5):
        v = torch.randn(B, 2, device="cuda") * vs
        for prec in precs:
            ws = ops.icnn_prepare(P, 2, H, 0, prec, B, True)
            _, _, m1, m2 = ops.icnn_decode_fwd(z, ws, 2, H, 0, 0.1, prec, True, True, True)
            tf = []
            for it in range(6):          # the training-variant forward (masks + saved accumulators)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.icnn_decode_fwd(z, ws, 2, H, 0, 0.1, prec, True, True, True); e1.record()
                torch.cuda.synchronize()
                tf.append(e0.elapsed_time(e1))
            print(f"H={H} prec={prec} training forward: " + " ".join(f"{t:.3f}" for t in tf) + " ms", flush=True)
            ts = []
            for it in range(6):
                n0 = _C.launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws, 2, H, 0, 0.1, prec); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"H={H} v~{vs:g} prec={prec}: " + " ".join(f"{t:.3f}" for t in ts) + f" ms  ({_C.launch_count() - n0} launches/call)", flush=True)
