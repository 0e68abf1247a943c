"""Timing of the fused wide-input ICNN kernels (csrc/icnn_wide.cu) on the MNIST-shaped decoder of BASELINE configs[3]
(ICNN(32,512) -> eye(784,32) pad -> ICNN(784,1024)) against stock PyTorch eager on the same GPU: the reference's own
formulation (module.py:142-148 + autograd.grad(create_graph=True), model.py:818-830) with cuBLAS FP32 GEMMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import eager_icnn_potential
import numpy as np, torch
from vae_song_b200 import module, ops, utils as vutils

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
PRECS = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "tf32x3", "tf32"]
kappa = 0.1
rng = np.random.default_rng(0)
ics = []
for d, H in ((32, 512), (784, 1024)):
    ic = module.ICNN(d, H).to(dev)
    with torch.no_grad():
        ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
        ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
        ic.A[0].bias.copy_(torch.tensor(rng.normal(-0.3, 1.0, (H,)), dtype=torch.float32))
    ics.append(ic)
z = torch.randn(B, 32, device=dev, requires_grad=True)
vy = torch.randn(B, 784, device=dev)


def fused(train):
    _, x1 = ics[0].brenier(z, kappa)
    _, y = ics[1].brenier(x1, kappa)          # [B,32] into ICNN(784,.): the eye(784,32) pad is implicit
    if train:
        for ic in ics:
            ic.zero_grad(set_to_none=True)
        z.grad = None
        (y * vy).sum().backward()
    return y


def eager(train):
    def brenier(ic, x):
        psi = eager_icnn_potential(x, ic._mode(), *ic._flat_params()) + kappa * x.pow(2).sum(1, keepdim=True)
        return torch.autograd.grad(psi, [x], torch.ones_like(psi), create_graph=True)[0]
    x1 = brenier(ics[0], z)
    y = brenier(ics[1], torch.nn.functional.linear(x1, torch.eye(784, 32, device=dev)))
    if train:
        for ic in ics:
            ic.zero_grad(set_to_none=True)
        z.grad = None
        (y * vy).sum().backward()
    return y


def timeit(fn, train, n=10):
    for _ in range(3):
        fn(train)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn(train)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fl_dec = vutils.flops_decode(32, 512) + vutils.flops_decode(784, 1024)
yb = eager(False).detach()
te = {False: timeit(eager, False), True: timeit(eager, True)}
for prec in PRECS:
    for ic in ics:
        ic.precision = prec
    ya = fused(False).detach()
    diff = (ya - yb).abs()
    print(f"B={B} {prec}: fused vs eager max rel diff {float(diff.max() / yb.abs().max()):.2e}, median {float(diff.median() / yb.abs().max()):.2e}")
    for train in (False, True):
        tf = timeit(fused, train)
        what = "decode+backward" if train else "decode"
        extra = f"  = {fl_dec * B / (tf * 1e-3) / 1e12:.1f} TFLOP/s dense-algorithmic" if not train else ""
        print(f"  {what}: fused {tf:.3f} ms ({B / tf * 1e3 / 1e6:.2f} M samples/s), PyTorch eager cuBLAS-FP32 {te[train]:.3f} ms -> {te[train] / tf:.2f}x{extra}")
