import sys, os
sys.path.insert(0, '/root/repo/tests'); import conftest
import numpy as np, torch
from ich_b200 import ops
from oracle import losses_oracle as LO
fx = torch.load('tests/golden/losses.pt')
for c in fx['local']:
    f1, f2 = c['f1'].cuda(), c['f2'].cuda()
    np.random.seed(c['np_seed'])
    corners = LO.sample_regions(tuple(f1.shape), c['K'], c['n_region']).astype(np.int32)
    ct = torch.from_numpy(corners).cuda()
    P = ops.RegionGather.apply(f1, f2, ct, c['K'])
    bs = f1.shape[0]
    def gather(f):
        return torch.stack([torch.stack([f[b, h0:h0+c['K'], w0:w0+c['K'], :].reshape(-1) for (h0, w0) in corners[b]]) for b in range(bs)])
    Pr = torch.cat((gather(c['f1']), gather(c['f2'])), dim=1)
    print('P err', (P.cpu() - Pr).abs().max().item(), P.shape)
    v = ops.InfoNCE.apply(P, c['tau'])
    s = LO._cos_sim_matrix(Pr) / c['tau']
    a2 = s.shape[1]
    eye = torch.eye(a2, dtype=torch.bool)
    lse = torch.logsumexp(s.masked_fill(eye, float('-inf')), dim=2)
    ar = torch.arange(a2)
    pos = s[:, ar, (ar + a2 // 2) % a2]
    print('loss', v.item(), (lse - pos).mean().item(), c['value'].item())
    for b in range(bs):
        vb = ops.InfoNCE.apply(P[b:b+1].contiguous(), c['tau'])
        print('  b', b, vb.item(), (lse - pos)[b].mean().item())
