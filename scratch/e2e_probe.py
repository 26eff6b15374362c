import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from ich_b200 import config
from ich_b200.staging import DevicePrefetcher
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
def step(x, m):
    opt.zero_grad(); out = net(x); loss = lossf(out, m); loss.backward(); opt.step(); return loss
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('resident, no sync      %.2f ms' % timeit(lambda: step(xd, md)))
print('resident, item()       %.2f ms' % timeit(lambda: step(xd, md).item()))
print('simple .to, no sync    %.2f ms' % timeit(lambda: step(xh.to(dev, non_blocking=True), mh.to(dev, non_blocking=True))))
print('simple .to, item()     %.2f ms' % timeit(lambda: step(xh.to(dev, non_blocking=True), mh.to(dev, non_blocking=True)).item()))
def pf(n, sync):
    for x, m in DevicePrefetcher([(xh, mh)] * n, dev):
        l = step(x, m)
        if sync: l.item()
for sync in (False, True):
    pf(3, sync); torch.cuda.synchronize(); t0 = time.perf_counter(); pf(10, sync); torch.cuda.synchronize()
    print('prefetch, item=%s      %.2f ms' % (sync, (time.perf_counter() - t0) / 10 * 1e3))
# CPU-side launch cost of one step (no GPU wait): enqueue only
torch.cuda.synchronize(); t0 = time.perf_counter(); step(xd, md); t1 = time.perf_counter(); torch.cuda.synchronize()
print('host enqueue time of one step %.2f ms' % ((t1 - t0) * 1e3))
