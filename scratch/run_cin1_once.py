import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
from ich_b200 import ops, config
from ich_b200._lib import call
config.set(precision='bf16')
n, d, h, w, cout, kd = 8, 64, 128, 128, 16, 3
x = torch.randn(n, d, h, w, 1, device='cuda').bfloat16()
dy = torch.randn(n, d, h, w, cout, device='cuda').bfloat16()
wt = torch.randn(cout, 1, kd, 3, 3, device='cuda') * 0.2
sums = torch.empty(2, cout, device='cuda', dtype=torch.float64)
y = torch.empty(n, d, h, w, cout, device='cuda', dtype=torch.bfloat16)
pk = ops._pack(wt, 'conv_fwd')
S = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    call('ich_conv_cin1_tc_fwd', x.data_ptr(), 1, pk.data_ptr(), None, y.data_ptr(), cout, sums[0].data_ptr(), sums[1].data_ptr(), n, d, h, w, cout, kd, 0, S)
    ops.conv_wgrad(x, dy, wt)
torch.cuda.synchronize()
print('done')
