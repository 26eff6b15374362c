"""First-layer (Cin = 1) tcgen05 kernels: parity vs torch conv (fp32, on the GPU) + timing at the cfg-3 / cfg-2 sizes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
import torch.nn.functional as F
from ich_b200 import ops, config
from ich_b200._lib import call

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
config.set(precision='bf16')
S = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def run(n, d, h, w, cout, k3d, timing=False):
    g = torch.Generator(device='cuda').manual_seed(n + d + h + w + cout)
    kd = 3 if k3d else 1
    x = torch.randn(n, 1, d, h, w, device='cuda', generator=g).bfloat16()
    wt = (torch.randn(cout, 1, kd, 3, 3, device='cuda', generator=g) * 0.2).bfloat16().float()
    dy = torch.randn(n, cout, d, h, w, device='cuda', generator=g).bfloat16()
    xc = x.permute(0, 2, 3, 4, 1).contiguous()
    dyc = dy.permute(0, 2, 3, 4, 1).contiguous()
    wr = wt.clone().requires_grad_(True)
    yr = F.conv3d(x.float(), wr, None, padding=(kd // 2, 1, 1))
    yr.backward(dy.float())
    y = ops.conv_forward(xc, wt, None)
    dw = ops.conv_wgrad(xc, dyc, wt)
    # fused statistics
    sums = torch.empty(2, cout, device='cuda', dtype=torch.float64)
    y2 = torch.empty_like(y)
    pk = ops._pack(wt, 'conv_fwd')
    call('ich_conv_cin1_tc_fwd', xc.data_ptr(), 1, pk.data_ptr(), None, y2.data_ptr(), cout, sums[0].data_ptr(), sums[1].data_ptr(), n, d, h, w, cout, kd, 0, S())
    yf = y2.double().reshape(-1, cout)
    e_y = rel(y.permute(0, 4, 1, 2, 3), yr)
    e_w = rel(dw, wr.grad)
    e_s = max(rel(sums[0], yf.sum(0)), rel(sums[1], (yf * yf).sum(0)))
    ok = e_y < 1e-2 and e_w < 2e-3 and e_s < 1e-5 and torch.equal(y, y2)
    msg = f'{n}x{d}x{h}x{w} cout {cout} kd {kd}: y {e_y:.2e} dw {e_w:.2e} stats {e_s:.2e} {"OK" if ok else "FAIL"}'
    if timing:
        for kind, fn in (('fwd', lambda: ops.conv_forward(xc, wt, None)),
                         ('fwd+stats', lambda: call('ich_conv_cin1_tc_fwd', xc.data_ptr(), 1, pk.data_ptr(), None, y2.data_ptr(), cout, sums[0].data_ptr(), sums[1].data_ptr(), n, d, h, w, cout, kd, 0, S())),
                         ('wgrad', lambda: ops.conv_wgrad(xc, dyc, wt))):
            fn(); fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            byts = n * d * h * w * (2 + 2 * cout)
            msg += f' | {kind} {ms * 1e3:.0f} us {byts / ms / 1e6:.0f} GB/s'
    print(msg, flush=True)
    return ok


ok = True
for case in [(2, 4, 8, 8, 8, True), (1, 1, 16, 24, 16, False), (1, 3, 6, 140, 32, True), (2, 2, 5, 16, 16, True), (1, 5, 7, 300, 24, True),
             (2, 1, 33, 257, 32, False)]:
    ok &= run(*case)
ok &= run(8, 64, 128, 128, 16, True, timing=True)
ok &= run(32, 1, 512, 512, 32, False, timing=True)
ok &= run(2, 64, 128, 128, 8, True, timing=True)
print('ALL OK' if ok else 'SOME FAILED')
sys.exit(0 if ok else 1)
