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
side = torch.cuda.Stream(dev)
bufs = [(torch.empty_like(xd), torch.empty_like(md)) for _ in range(3)]
def step(x, m):
    opt.zero_grad(); out = net(x); loss = lossf(out, m); loss.backward(); opt.step(); return loss
def timed(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def resident(n):
    for _ in range(n): step(xd, md).item()
def manual(n):
    cur = torch.cuda.current_stream()
    def stage(i):
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            bufs[i % 3][0].copy_(xh, non_blocking=True); bufs[i % 3][1].copy_(mh, non_blocking=True)
    stage(0)
    for i in range(n):
        cur.wait_stream(side)
        if i + 1 < n: stage(i + 1)
        step(*bufs[i % 3]).item()
def pf(n):
    for x, m in DevicePrefetcher([(xh, mh)] * n, dev): step(x, m).item()
PF = DevicePrefetcher([(xh, mh)] * 10, dev)
def pf_reuse(n):
    for x, m in PF: step(x, m).item()
for _ in range(2):
    for name, fn in (('resident', resident), ('manual', manual), ('prefetcher', pf), ('prefetcher reused', pf_reuse), ('resident', resident)):
        fn(3)
        print(f'{name:18s} {timed(fn):.2f} ms/step', flush=True)
