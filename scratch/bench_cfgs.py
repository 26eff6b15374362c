"""Timing of the other BASELINE configs (parity-test cases, not bench lines): cfg-2 2-D slices, cfg-4 contrastive, cfg-5 inference."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests')); import conftest  # noqa
import numpy as np
import torch
import torch.nn.functional as F
from ich_b200 import config, infer
from src.models.networks.UNet import UNet, UNet_Encoder, Partial_UNet
from src.models.optim.LossFunctions import BinaryDiceLoss, InfoNCELoss, LocalInfoNCELoss
dev = 'cuda'
config.set(precision='bf16')

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

# cfg-2
net = UNet(depth=5, use_3D=False, top_filter=32, midchannels_factor=1, p_dropout=0.0).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
x = torch.rand(32, 1, 512, 512, device=dev); m = (torch.rand(32, 1, 512, 512, device=dev) > 0.98).float()
lf = BinaryDiceLoss(reduction='mean', p=2, alpha=0.2)
def step2():
    opt.zero_grad(); l = lf(net(x), m); l.backward(); opt.step(); return l
ms = timeit(step2)
print(f'cfg-2 2D U-Net d5 tf32 mcf1, batch 32 of 512x512: {ms:.2f} ms/step, {32*512*512/ms/1e3:.1f} M pixels/s, {9241.7/ms:.1f} conv-TFLOP/s', flush=True)
del net, opt, x, m; torch.cuda.empty_cache()

# cfg-4 global
enc = UNet_Encoder(depth=4, use_3D=True, top_filter=32, midchannels_factor=2, MLP_head=[512, 128], p_dropout=0.0).to(dev).train()
opt = torch.optim.Adam(enc.parameters(), lr=1e-3)
v1 = torch.rand(8, 1, 64, 128, 128, device=dev); v2 = torch.rand(8, 1, 64, 128, 128, device=dev)
nce = InfoNCELoss(set_size=8, tau=0.1, device=dev)
def step4():
    opt.zero_grad(); l = nce(F.normalize(enc(v1), dim=1), F.normalize(enc(v2), dim=1)); l.backward(); opt.step(); return l
ms = timeit(step4)
print(f'cfg-4 global contrastive (3D encoder d4 tf32, 2 views of 8x1x64x128x128): {ms:.2f} ms/step, {2*8*64*128*128/ms/1e3:.1f} M voxels/s', flush=True)
del enc, opt, v1, v2; torch.cuda.empty_cache()

# cfg-4 local
pu = Partial_UNet(depth=5, n_decoder=3, use_3D=False, top_filter=32, midchannels_factor=1, head_channel=[128, 32], p_dropout=0.0).to(dev).train()
opt = torch.optim.Adam(pu.parameters(), lr=1e-3)
a = torch.rand(32, 1, 256, 256, device=dev); b = torch.rand(32, 1, 256, 256, device=dev)
loc = LocalInfoNCELoss(tau=0.1, K=3, n_region=20, device=dev)
def step4l():
    np.random.seed(0); opt.zero_grad(); l = loc(pu(a), pu(b)); l.backward(); opt.step(); return l
ms = timeit(step4l)
print(f'cfg-4 local contrastive (2D partial U-Net d5, 2 views of 32x1x256x256): {ms:.2f} ms/step', flush=True)
del pu, opt, a, b; torch.cuda.empty_cache()

# cfg-5
net = UNet(depth=4, use_3D=True, top_filter=32, midchannels_factor=2, p_dropout=0.0).to(dev).eval()
vol = torch.rand(1, 1, 32, 512, 512, device=dev)
ms = timeit(lambda: infer.sliding_window_predict(net, vol, (32, 128, 128), (32, 128, 128), batch=8, distributed=False), n=3, warm=1)
print(f'cfg-5 sliding-window inference 1x32x512x512, 16 windows of 32x128x128: {ms:.2f} ms/volume, {32*512*512/ms/1e3:.1f} M voxels/s', flush=True)
ms = timeit(lambda: infer.sliding_window_predict(net, vol, (32, 128, 128), (32, 64, 64), batch=8, distributed=False), n=2, warm=1)
print(f'cfg-5 with 50% overlap (49 windows): {ms:.2f} ms/volume', flush=True)
