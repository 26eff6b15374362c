"""Does a concurrent H2D copy (side stream) slow the training step?  And where in the step is it cheapest?"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from ich_b200 import config
from src.models.networks.UNet import UNet
from src.models.optim.LossFunctions import ComboLoss
dev = torch.device('cuda', 0)
config.set(precision='bf16')
net = UNet(**bench.NET_KW).to(dev).train()
lossf = ComboLoss(**bench.LOSS_KW)
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
shape = (8, 1) + bench.PATCH
xh = torch.rand(*shape).pin_memory(); mh = (torch.rand(*shape) > 0.98).float().pin_memory()
xd, md = xh.to(dev), mh.to(dev)
xb, mb = torch.empty_like(xd), torch.empty_like(md)
side = torch.cuda.Stream(dev)
def step(x, m):
    opt.zero_grad(); out = net(x); loss = lossf(out, m); loss.backward(); opt.step(); return loss
def run(mode, n=10):
    for _ in range(3): step(xd, md)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        if mode == 'side':
            with torch.cuda.stream(side):
                if i == n - 1: c0.record()
                xb.copy_(xh, non_blocking=True); mb.copy_(mh, non_blocking=True)
                if i == n - 1: c1.record()
        elif mode == 'side_dep':      # copy ordered after the previous step, like the prefetcher
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                xb.copy_(xh, non_blocking=True); mb.copy_(mh, non_blocking=True)
        elif mode == 'inline':
            xb.copy_(xh, non_blocking=True); mb.copy_(mh, non_blocking=True)
        l = step(xd, md)
        if mode == 'side_dep':
            torch.cuda.current_stream().wait_stream(side)
        l.item()
    e1.record(); torch.cuda.synchronize()
    msg = f'{mode:10s} {e0.elapsed_time(e1) / n:.2f} ms/step'
    if mode == 'side': msg += f'  (copy itself {c0.elapsed_time(c1):.2f} ms)'
    print(msg, flush=True)
for mode in ('none', 'inline', 'side', 'side_dep', 'none'):
    run(mode)
