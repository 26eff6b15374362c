"""Plane-streaming kernel forced onto the mid-resolution layers (ICH_TC_STREAM=2), with / without thread-block clusters + multicast weights
(ICH_TC_STREAM_CLUSTER=2|4): correctness against torch's fp32 conv (checker) and timing.  Env is read once per process."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
import torch.nn.functional as F
from ich_b200 import ops, config
config.set(precision='bf16')
torch.backends.cudnn.allow_tf32 = False
print('env', {k: v for k, v in os.environ.items() if k.startswith('ICH_TC')})
SHAPES = [(1, 4, 12, 64, 64, 64), (2, 6, 24, 64, 128, 64), (1, 5, 8, 64, 64, 32), (8, 32, 64, 64, 64, 64), (8, 32, 64, 64, 128, 64), (8, 32, 64, 64, 32, 32),
          (8, 32, 64, 64, 64, 32), (8, 16, 32, 32, 128, 128), (1, 20, 16, 128, 32, 16), (2, 9, 12, 128, 16, 32)]
for n, d, h, w, cin, cout in SHAPES:
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(n, d, h, w, cin, device='cuda', generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, 3, device='cuda', generator=g) * 0.05).bfloat16().float()
    var = ops._tc_variant(x, cin, cout, (3, 3, 3))
    y = ops.conv_forward(x, wt, None)
    torch.cuda.synchronize()
    big = n * d * h * w * cin > 2e8
    xr = x.float().permute(0, 4, 1, 2, 3).contiguous()
    yr = F.conv3d(xr, wt, None, padding=1)
    err = ((y.permute(0, 4, 1, 2, 3).float() - yr).norm() / yr.norm()).item()
    for _ in range(2):
        ops.conv_forward(x, wt, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_forward(x, wt, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * d * h * w * cin * cout * 27
    print(f'{n}x{d}x{h}x{w} {cin}->{cout} variant {var} rel err {err:.2e} {"OK" if err < 4e-3 else "FAIL"} | {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s', flush=True)
