"""CPU restatement of the reference's hot-path losses. TEST INFRASTRUCTURE ONLY.

Closed forms of models/optim/LossFunctions.py (paths relative to /root/reference/code/src); each
is pinned against the reference module by oracle/make_golden.py + tests/test_oracle.py.
"""
import numpy as np
import torch


def _reduce(v, reduction):
    return v.mean() if reduction == 'mean' else v.sum() if reduction == 'sum' else v


def binary_dice_loss(pred, mask, reduction='mean', p=2, alpha=1.0, eps=1):
    """BinaryDiceLoss.forward, LossFunctions.py:39-63."""
    dims = tuple(range(1, pred.ndim))
    inter = (pred * mask).sum(dims)                                            # :52
    union = pred.pow(p).sum(dims) + mask.pow(p).sum(dims)                      # :53
    dl = 1 - (2 * inter + eps) / (union + eps)                                 # :54
    dl = torch.where(mask.sum(dims) > 0, dl, alpha * dl)                       # :56
    return _reduce(dl, reduction)


def tversky_loss(pred, mask, alpha=1.0, beta=0.5, gamma=0.5, reduction='mean', eps=1):
    """TverskyLoss.forward, LossFunctions.py:88-114."""
    dims = tuple(range(1, pred.ndim))
    tp = (pred * mask).sum(dims)                                               # :101
    fp = (pred * (1 - mask)).sum(dims)                                         # :102
    fn = ((1 - pred) * mask).sum(dims)                                         # :103
    tl = 1 - (tp + eps) / (tp + beta * fn + gamma * fp + eps)                  # :105
    tl = torch.where(mask.sum(dims) > 0, tl, alpha * tl)                       # :107
    return _reduce(tl, reduction)


def combo_loss(pred, mask, alpha=0.5, beta=0.5, reduction='mean', p=1):
    """ComboLoss.forward, LossFunctions.py:143-166 (BCE summed over voxels, :157)."""
    dims = tuple(range(1, pred.ndim))
    dice = binary_dice_loss(pred, mask, reduction='none', p=p)
    bce = -(beta * mask * torch.log(pred + 1e-14) + (1 - beta) * (1 - mask) * torch.log(1 - pred + 1e-14)).sum(dims)
    return _reduce(alpha * bce + (1 - alpha) * dice, reduction)


def _cos_sim_matrix(p, eps=1e-8):
    """nn.CosineSimilarity broadcast form used at LossFunctions.py:221,330."""
    n = p.norm(dim=-1, keepdim=True).clamp_min(eps)
    pn = p / n
    return pn @ pn.transpose(-1, -2)


def info_nce_loss(z1, z2, tau=0.5):
    """InfoNCELoss.forward, LossFunctions.py:208-230 in closed form:
    mean_i( logsumexp_{j != i} S_ij - S_{i,(i+N) mod 2N} ), S = cos/tau."""
    n = z1.shape[0]
    p = torch.cat((z1, z2), dim=0)
    s = _cos_sim_matrix(p) / tau
    eye = torch.eye(2 * n, dtype=torch.bool, device=s.device)
    lse = torch.logsumexp(s.masked_fill(eye, float('-inf')), dim=1)
    ar = torch.arange(2 * n, device=s.device)
    pos = s[ar, (ar + n) % (2 * n)]
    return (lse - pos).mean()


def sample_regions(feature_shape, K, n_region):
    """Region sampling of LocalInfoNCELoss.get_sample_region_mask, LossFunctions.py:279-306.
    Draws from the GLOBAL numpy RNG with the same two calls (:292-293). Returns int64 array
    [bs, n_region, 2] holding the (h, w) top-left corner of each K x K region (region id = index+1)."""
    bs, H, W, _ = feature_shape
    gh, gw = H // K, W // K
    idx_col = np.random.choice(gh * gw, n_region, replace=False)               # :292
    idx = np.random.rand(bs, gh * gw).argsort(axis=1)[:, idx_col]              # :293
    return np.stack([(idx // gw) * K, (idx % gw) * K], axis=-1).astype(np.int64)


def local_info_nce_loss(f1, f2, tau=0.5, K=3, n_region=13):
    """LocalInfoNCELoss.forward, LossFunctions.py:308-341 in closed form. f is read as (bs,H,W,C)."""
    bs = f1.shape[0]
    corners = sample_regions(tuple(f1.shape), K, n_region)

    def gather(f):
        out = []
        for b in range(bs):
            rows = []
            for a in range(n_region):
                h0, w0 = corners[b, a]
                rows.append(f[b, h0:h0 + K, w0:w0 + K, :].reshape(-1))         # (h, w, c) row-major :323-326
            out.append(torch.stack(rows))
        return torch.stack(out)

    p = torch.cat((gather(f1), gather(f2)), dim=1)                            # B x 2A x K*K*C
    s = _cos_sim_matrix(p) / tau
    a2 = 2 * n_region
    eye = torch.eye(a2, dtype=torch.bool, device=s.device)
    lse = torch.logsumexp(s.masked_fill(eye, float('-inf')), dim=2)
    ar = torch.arange(a2, device=s.device)
    pos = s[:, ar, (ar + n_region) % a2]
    return (lse - pos).mean()


def batch_binary_confusion(pred, target):
    """utils/tensor_utils.py:12-36."""
    t = target.reshape(target.shape[0], -1)
    p = pred.reshape(pred.shape[0], -1)
    return ((1 - p) * (1 - t)).sum(1), (p * (1 - t)).sum(1), ((1 - p) * t).sum(1), (p * t).sum(1)
