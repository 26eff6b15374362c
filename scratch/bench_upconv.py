"""ConvTranspose k2 s2 + concat (UpConvCat) micro-benchmark at the cfg-3 shapes: forward and backward, CUDA-event timed."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
from ich_b200 import ops, config
config.set(precision='bf16')
for name, n, d, h, w, cin, cout in [('upT0', 8, 8, 16, 16, 256, 128), ('upT1', 8, 16, 32, 32, 128, 64), ('upT2', 8, 32, 64, 64, 64, 32)]:
    x = torch.randn(n, d, h, w, cin, device='cuda', dtype=torch.bfloat16, requires_grad=True)
    res = torch.randn(n, 2 * d, 2 * h, 2 * w, cout, device='cuda', dtype=torch.bfloat16, requires_grad=True)
    wt = (torch.randn(cin, cout, 2, 2, 2, device='cuda') * 0.05).requires_grad_(True)
    b = torch.zeros(cout, device='cuda', requires_grad=True)
    g = torch.randn(n, 2 * d, 2 * h, 2 * w, 2 * cout, device='cuda', dtype=torch.bfloat16)
    def fwd():
        return ops.UpConvCat.apply(x, res, wt, b, 2, None)
    def t(fn, reps=10):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tf = t(fwd)
    out = fwd()
    def bwd():
        torch.autograd.grad(out, [x, res, wt, b], g, retain_graph=True)
    tb = t(bwd)
    print(f'{name} {cin}->{cout}: fwd (slab copy + convT) {tf:.0f} us | bwd (s2d + dgrad + wgrad) {tb:.0f} us', flush=True)
