import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
from ich_b200 import ops, config
config.set(precision='bf16')
n, d, h, w, c = 8, 64, 128, 128, 32
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = torch.randn(n, d, h, w, c, device='cuda', dtype=torch.bfloat16)
dy = torch.randn(n, d, h, w, c, device='cuda', dtype=torch.bfloat16)
wt = torch.randn(c, c, 3, 3, 3, device='cuda') * 0.05
print('ptr x %x dy %x' % (x.data_ptr(), dy.data_ptr()))
print('fwd(x)   ', t(lambda: ops.conv_forward(x, wt, None)))
print('fwd(dy)  ', t(lambda: ops.conv_forward(dy, wt, None)))
print('dgrad(dy)', t(lambda: ops.conv_dgrad(dy, wt)))
print('dgrad(x) ', t(lambda: ops.conv_dgrad(x, wt)))
out = torch.empty_like(x)
from ich_b200._lib import call
pk = ops._pack(wt, 'conv_fwd_tc_s'); pd = ops._pack(wt, 'conv_dgrad_tc_s')
s = torch.cuda.current_stream().cuda_stream
print('fixed out, fwd pack, x ', t(lambda: call('ich_conv_tc_fwd', x.data_ptr(), c, pk.data_ptr(), None, out.data_ptr(), c, n, d, h, w, c, c, 3, 3, 3, 0, s)))
print('fixed out, dgrad pack, x', t(lambda: call('ich_conv_tc_fwd', x.data_ptr(), c, pd.data_ptr(), None, out.data_ptr(), c, n, d, h, w, c, c, 3, 3, 3, 0, s)))
print('fixed out, fwd pack, dy', t(lambda: call('ich_conv_tc_fwd', dy.data_ptr(), c, pk.data_ptr(), None, out.data_ptr(), c, n, d, h, w, c, c, 3, 3, 3, 0, s)))
out2 = torch.empty(n * d * h * w * c + 4096, device='cuda', dtype=torch.bfloat16)[2048:2048 + n * d * h * w * c].view(n, d, h, w, c)
print('out shifted by 4 KB     ', t(lambda: call('ich_conv_tc_fwd', x.data_ptr(), c, pk.data_ptr(), None, out2.data_ptr(), c, n, d, h, w, c, c, 3, 3, 3, 0, s)))
print('--- pack address / value experiments')
def run(pack):
    return t(lambda: call('ich_conv_tc_fwd', x.data_ptr(), c, pack.data_ptr(), None, out.data_ptr(), c, n, d, h, w, c, c, 3, 3, 3, 0, s))
print('pk ptr %x pd ptr %x' % (pk.data_ptr(), pd.data_ptr()))
print('pk.clone()', run(pk.clone()), ' pd.clone()', run(pd.clone()))
big = torch.empty(1 << 22, device='cuda', dtype=torch.bfloat16)
for off in (0, 64, 128, 512, 4096, 27648, 65536, 1 << 20):
    a = big[off:off + pk.numel()].view_as(pk); a.copy_(pk)
    b = big[off:off + pd.numel()].view_as(pd)
    ta = run(a); b.copy_(pd); tb = run(b)
    print('offset %8d: fwd-pack values %.3f  dgrad-pack values %.3f' % (off * 2, ta, tb))
z = torch.zeros_like(pk); print('zeros pack', run(z))
r = torch.randn_like(pk.float()).mul_(0.05).bfloat16(); print('fresh random pack', run(r))
print('pk abs mean', pk.float().abs().mean().item(), 'pd abs mean', pd.float().abs().mean().item(), 'equal sets', torch.equal(pk.float().sort()[0] if False else pk.float().flatten().sort()[0], pd.float().flatten().sort()[0]))
