"""B200-native drop-in for the reference's `src.models.optim.LossFunctions` module.

Hot-path losses (BinaryDiceLoss, ComboLoss, InfoNCELoss, LocalInfoNCELoss) and TverskyLoss keep the
reference's constructor signatures, assertions and call protocol (/root/reference/code/src/models/optim/LossFunctions.py:14-63,
:65-114, :116-166, :168-230, :232-341) but run as fused one-pass CUDA kernels (ich_b200.ops).  The side-track losses the
module must still export (DiscountedL1 :343-409, GDL :411-448, HSCLoss :450-470) are plain torch
restatements -- they are not on the hot path.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from ich_b200 import ops  # noqa: E402


def _apply_reduction(v, reduction):
    if reduction == 'mean':
        return v.mean()
    if reduction == 'sum':
        return v.sum()
    if reduction == 'none':
        return v


class BinaryDiceLoss(nn.Module):
    """1 - (2*sum(p*m) + eps) / (sum(p^P) + sum(m^P) + eps) per sample, scaled by alpha when the mask is empty."""

    def __init__(self, reduction='mean', p=2, alpha=1.0, eps=1):
        super(BinaryDiceLoss, self).__init__()
        assert reduction in ['mean', 'none', 'sum'], f"Reduction mode: '{reduction}' is not supported. Use either 'mean', 'sum' or 'none'."
        self.reduction = reduction
        self.p = p
        self.alpha = alpha
        self.eps = eps

    def forward(self, pred, mask):
        assert pred.shape == mask.shape, f'Prediction and Mask should have the same dimensions! Given: Prediction {pred.shape} / Mask {mask.shape}'
        return ops.SegLoss.apply(pred, mask, self.p, self.eps, self.alpha, 0.0, 1.0, 0.5, self.reduction)


class ComboLoss(nn.Module):
    """alpha * BCE(beta-weighted, SUMMED over voxels) + (1 - alpha) * Dice, per sample."""

    def __init__(self, alpha=0.5, beta=0.5, reduction='mean', p=1):
        super(ComboLoss, self).__init__()
        assert alpha >= 0 and alpha <= 1, f'ValueError. alpha must in the range [0,1]. {alpha} given'
        assert beta >= 0 and beta <= 1, f'ValueError. beta must in the range [0,1]. {beta} given'
        self.alpha = alpha
        self.beta = beta
        self.reduction = reduction
        self.bin_dice_loss_fn = BinaryDiceLoss(reduction='none', p=p)

    def forward(self, pred, mask):
        assert pred.shape == mask.shape, f'Prediction and Mask should have the same dimensions! Given: Prediction {pred.shape} / Mask {mask.shape}'
        d = self.bin_dice_loss_fn
        red = self.reduction if self.reduction in ('mean', 'sum', 'none') else 'none'
        out = ops.SegLoss.apply(pred, mask, d.p, d.eps, d.alpha, self.alpha, 1.0 - self.alpha, self.beta, red)
        return out if self.reduction in ('mean', 'sum', 'none') else None   # the reference returns None for unknown reductions


class InfoNCELoss(nn.Module):
    """Global contrastive loss over the 2N x 2N cosine-similarity matrix (positives on the +-N diagonals)."""

    def __init__(self, set_size=None, tau=0.5, device='cuda'):
        assert set_size is not None, 'The set size is a mandatory parameter'
        super(InfoNCELoss, self).__init__()
        self.tau = tau
        self.device = device
        self.set_size = set_size
        self.neg_mask = self.get_neg_mask(set_size)

    def get_neg_mask(self, set_size):
        """Boolean mask of the negatives (everything but the main and the +-set_size diagonals); kept for API parity."""
        idx = torch.arange(2 * set_size, device=self.device)
        diff = (idx.unsqueeze(0) - idx.unsqueeze(1)).abs()
        return ~((diff == 0) | (diff == set_size))

    def forward(self, z1, z2):
        if z1.shape[0] != self.set_size or z2.shape[0] != self.set_size:
            raise RuntimeError(f'InfoNCELoss: batch ({z1.shape[0]}) must equal set_size ({self.set_size})')
        # multi-GPU (not in the reference): with ICH_B200_GLOBAL_NCE=1 the comparison set spans the batches of all ranks
        z1, z2 = ops.gather_rows(z1), ops.gather_rows(z2)
        p = torch.cat((z1, z2), dim=0).unsqueeze(0)          # [1, 2N, E]
        return ops.InfoNCE.apply(p, self.tau)


class LocalInfoNCELoss(nn.Module):
    """Local contrastive loss: n_region random K x K regions per sample, compared across the two views."""

    def __init__(self, tau=0.5, K=3, n_region=13, device='cuda'):
        super(LocalInfoNCELoss, self).__init__()
        self.tau = tau
        self.K = K
        self.n_region = n_region
        self.device = device
        self.pos_mask, self.neg_mask = self.get_masks(n_region)

    def get_masks(self, set_size):
        idx = torch.arange(2 * set_size, device=self.device)
        diff = (idx.unsqueeze(0) - idx.unsqueeze(1)).abs()
        pos_mask = diff == set_size
        return pos_mask, ~(pos_mask | (diff == 0))

    def sample_region_corners(self, feature_shape):
        """Same two draws from the GLOBAL numpy RNG as the reference (:292-293) -> [bs, n_region, 2] (h, w) corners."""
        bs, H, W, C = feature_shape
        gh, gw = H // self.K, W // self.K
        idx_col = np.random.choice(gh * gw, self.n_region, replace=False)
        idx = np.random.rand(bs, gh * gw).argsort(axis=1)[:, idx_col]
        return np.ascontiguousarray(np.stack([(idx // gw) * self.K, (idx % gw) * self.K], axis=-1), dtype=np.int32)

    def get_sample_region_mask(self, feature_shape):
        """Region-label mask (B x H x W, labels 1..n_region) as the reference returns it; not used by forward."""
        corners = self.sample_region_corners(feature_shape)
        out = torch.zeros(feature_shape[:-1], device=self.device)
        for b in range(corners.shape[0]):
            for a in range(self.n_region):
                h0, w0 = corners[b, a]
                out[b, h0:h0 + self.K, w0:w0 + self.K] = a + 1
        return out

    def _corners_to_device(self, corners, device):
        """Region corners (drawn on the host from numpy's global RNG, as the reference does) -> device WITHOUT a host synchronisation:
        a copy from pageable memory blocks the host until the stream reaches it (i.e. until both forward passes have finished), after
        which the whole backward pass is launched into an idle GPU.  Ring of pinned staging buffers, each guarded by an event."""
        t = torch.from_numpy(corners)
        if torch.device(device).type != 'cuda':
            return t.to(device)
        ring = getattr(self, '_corner_ring', None)
        if ring is None or ring[0][0].shape != t.shape:
            ring = [[torch.empty(t.shape, dtype=t.dtype).pin_memory(), None] for _ in range(4)]
            self._corner_ring, self._corner_turn = ring, 0
        slot = ring[self._corner_turn % len(ring)]
        self._corner_turn += 1
        if slot[1] is not None:
            slot[1].synchronize()          # the copy that last used this pinned buffer (4 calls ago) has completed
        slot[0].copy_(t)
        out = slot[0].to(device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return out

    def forward(self, f1, f2):
        bs, H, W, C = f1.shape       # dims are read as (bs, H, W, C) whatever the caller meant
        corners = self._corners_to_device(self.sample_region_corners(tuple(f1.shape)), f1.device)
        p = ops.RegionGather.apply(f1, f2, corners, self.K)   # [bs, 2A, K*K*C]
        return ops.InfoNCE.apply(p, self.tau)


class TverskyLoss(nn.Module):
    """1 - (TP + 1) / (TP + beta*FN + gamma*FP + 1) per sample, scaled by alpha when the mask is empty (reference :65-114).
    Runs on the engine: the Dice reduction pass + a finalize, and a backward pass that reads only the mask."""

    def __init__(self, alpha=1.0, beta=0.5, gamma=0.5, reduction='mean'):
        super(TverskyLoss, self).__init__()
        self.alpha = alpha
        self.beta = beta
        self.gamma = gamma
        self.reduction = reduction
        self.eps = 1

    def forward(self, pred, mask):
        assert pred.shape == mask.shape, f'Prediction and Mask should have the same dimensions! Given: Prediction {pred.shape} / Mask {mask.shape}'
        red = self.reduction if self.reduction in ('mean', 'sum', 'none') else 'none'
        out = ops.TverskyLossFn.apply(pred, mask, self.eps, self.alpha, self.beta, self.gamma, red)
        return out if self.reduction in ('mean', 'sum', 'none') else None   # the reference returns None for unknown reductions


# ---------------------------------------------------------------------------------------------------------------------
# Out-of-scope losses (plain torch; exported because other reference trainers import them by name)
# ---------------------------------------------------------------------------------------------------------------------
class DiscountedL1(nn.Module):
    """L1 on the mask, discounted by gamma ** (distance to the nearest non-mask border pixel)."""

    def __init__(self, gamma=0.99, reduction='mean', device='cuda'):
        super(DiscountedL1, self).__init__()
        assert reduction in ['mean', 'none', 'sum'], f"Reduction mode: '{reduction}' is not supported. Use either 'mean', 'sum' or 'none'."
        self.gamma = torch.tensor(gamma, device=device)
        self.L1 = nn.L1Loss(reduction='none')
        self.reduction = reduction
        self.device = device

    def get_dist_mask(self, mask):
        border = nn.functional.max_pool2d(mask, 3, stride=1, padding=1) - mask
        maps = []
        for m, b in zip(mask[:, 0], border[:, 0]):
            inside, edge = torch.nonzero(m), torch.nonzero(b)
            dist = torch.zeros(m.shape, device=self.device)
            dist[inside[:, 0], inside[:, 1]] = torch.cdist(inside[None].float(), edge[None].float(), p=2).min(dim=2)[0][0]
            maps.append(dist[None, None])
        return torch.cat(maps, dim=0)

    def forward(self, rec, im, mask):
        weight = (self.gamma.view(1, 1, 1, 1) ** self.get_dist_mask(mask)) * mask
        return _apply_reduction(self.L1(rec, im) * weight, self.reduction)


class GDL(nn.Module):
    """Gradient-difference loss on horizontal / vertical finite differences."""

    def __init__(self, reduction='mean', channels=1, device='cuda'):
        super(GDL, self).__init__()
        assert reduction in ['none', 'mean', 'sum'], f"Reduction startegy not supported. Must be one of ['none', 'mean', 'sum']. Given {reduction}."
        self.reduction = reduction
        k_h = torch.zeros(1, 1, 3, 3, device=device)
        k_h[0, 0, 1, 0], k_h[0, 0, 1, 1] = -1.0, 1.0
        k_v = torch.zeros(1, 1, 3, 3, device=device)
        k_v[0, 0, 0, 1], k_v[0, 0, 1, 1] = -1.0, 1.0
        self.w_h = k_h.repeat(1, channels, 1, 1)
        self.w_v = k_v.repeat(1, channels, 1, 1)

    def forward(self, im, rec):
        grad = lambda t, k: torch.abs(nn.functional.conv2d(t, k, padding=1))
        loss = torch.abs(grad(im, self.w_h) - grad(rec, self.w_h)) + torch.abs(grad(im, self.w_v) - grad(rec, self.w_v))
        return _apply_reduction(loss.sum(dim=[1, 2, 3]), self.reduction)


class HSCLoss(nn.Module):
    """Hypersphere-classifier loss with the pseudo-Huber map (FCDD)."""

    def __init__(self, reduction='mean'):
        super().__init__()
        assert reduction in ['none', 'mean'], f"Reduction mode not supported. Must be either 'none' or 'mean'. Given '{reduction}'."
        self.reduction = reduction

    def forward(self, x, y):
        a = (torch.sqrt(x ** 2 + 1) - 1).reshape(x.shape[0], -1).mean(-1)
        loss = torch.where(y == 1, -torch.log(1 - torch.exp(-a) + 1e-31), a)
        return _apply_reduction(loss, self.reduction)
