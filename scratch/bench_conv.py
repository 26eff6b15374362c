"""Per-layer conv micro-benchmark (cfg-3 shapes): CUDA-event timing of fwd / dgrad / wgrad through the C-ABI."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
from ich_b200 import ops, config

LAYERS = [  # name, N, D, H, W, Cin, Cout
    ('d0.c2', 8, 64, 128, 128, 16, 32), ('u2.c1', 8, 64, 128, 128, 64, 32), ('u2.c2', 8, 64, 128, 128, 32, 32),
    ('d1.c1', 8, 32, 64, 64, 32, 32), ('d1.c2', 8, 32, 64, 64, 32, 64), ('u1.c1', 8, 32, 64, 64, 128, 64), ('u1.c2', 8, 32, 64, 64, 64, 64),
    ('d2.c2', 8, 16, 32, 32, 64, 128), ('u0.c1', 8, 16, 32, 32, 256, 128), ('u0.c2', 8, 16, 32, 32, 128, 128),
    ('bt.c2', 8, 8, 16, 16, 128, 256),
]
only = sys.argv[1].split(',') if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
config.set(precision='bf16')
tot = {'fwd': [0, 0], 'dgrad': [0, 0], 'wgrad': [0, 0]}
for name, n, d, h, w, cin, cout in LAYERS:
    if only and name not in only:
        continue
    x = torch.randn(n, d, h, w, cin, device='cuda', dtype=torch.bfloat16)
    dy = torch.randn(n, d, h, w, cout, device='cuda', dtype=torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, 3, device='cuda') * 0.05
    flops = 2.0 * n * d * h * w * cin * cout * 27
    res = []
    for kind, fn in (('fwd', lambda: ops.conv_forward(x, wt, None)), ('dgrad', lambda: ops.conv_dgrad(dy, wt)),
                     ('wgrad', lambda: ops.conv_wgrad(x, dy, wt))):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append(f'{kind} {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF/s')
        tot[kind][0] += flops; tot[kind][1] += ms
    print(f'{name} {n}x{d}x{h}x{w} {cin:3d}->{cout:3d} | ' + ' | '.join(res), flush=True)
for k, (f, ms) in tot.items():
    if ms:
        print(f'total {k}: {ms:.2f} ms, {f / ms / 1e9:.1f} TF/s')

# fused-statistics variant of the forward (what ConvBnRelu uses in training)
if os.environ.get('BENCH_STATS'):
    from ich_b200._lib import call
    tot = [0.0, 0.0]
    for name, n, d, h, w, cin, cout in LAYERS:
        x = torch.randn(n, d, h, w, cin, device='cuda', dtype=torch.bfloat16)
        wt = torch.randn(cout, cin, 3, 3, 3, device='cuda') * 0.05
        y = torch.empty(n, d, h, w, cout, device='cuda', dtype=torch.bfloat16)
        sums = torch.empty(2, cout, device='cuda', dtype=torch.float64)
        var = ops._tc_variant(x, cin, cout, (3, 3, 3))
        pk = ops._pack(wt, 'conv_fwd_tc_s' if var == 2 else 'conv_fwd_tc')
        fn = lambda: call('ich_conv_tc_fwd_stats', x.data_ptr(), cin, pk.data_ptr(), y.data_ptr(), cout, sums[0].data_ptr(), sums[1].data_ptr(),
                          n, d, h, w, cin, cout, 3, 3, 3, torch.cuda.current_stream().cuda_stream)
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * n * d * h * w * cin * cout * 27
        tot[0] += fl; tot[1] += ms
        print(f'{name} fwd+stats {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s (variant {var})')
    print(f'total fwd+stats: {tot[1]:.2f} ms, {tot[0] / tot[1] / 1e9:.1f} TF/s')
