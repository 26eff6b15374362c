"""Sliding-window full-volume inference (BASELINE cfg-5), sharded per window across ranks.

The reference only has slice-batched 2-D inference (`UNet2D.segement_volume`, models/optim/UNet2D.py:272-314); the 3-D
windowing rule here is the one SURVEY section 8d fixes as the oracle: eval-mode network on each window, mean blending where
windows overlap, mask = pred >= 0.5 (UNet2D.py:220)."""
import torch
import torch.distributed as dist

from .dp import shard_indices


def window_starts(L, w, s):
    return sorted(set(list(range(0, max(L - w, 0) + 1, s)) + [max(L - w, 0)]))


def sliding_window_predict(net, vol, window, stride=None, batch=4, threshold=0.5, distributed=None):
    """vol [1, C, D, H, W] (cuda). Returns (pred fp32 [1, out, D, H, W], mask bool). With torch.distributed initialised
    the windows are sharded over ranks and the accumulators summed with one all-reduce."""
    stride = stride or window
    _, _, D, H, W = vol.shape
    wins = [(d0, h0, w0) for d0 in window_starts(D, window[0], stride[0]) for h0 in window_starts(H, window[1], stride[1])
            for w0 in window_starts(W, window[2], stride[2])]
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    mine = [wins[i] for i in shard_indices(len(wins))] if distributed else wins
    was_training = net.training
    net.eval()
    acc = cnt = None
    with torch.no_grad():
        for i in range(0, len(mine), batch):
            chunk = mine[i:i + batch]
            x = torch.cat([vol[:, :, d0:d0 + window[0], h0:h0 + window[1], w0:w0 + window[2]] for d0, h0, w0 in chunk], dim=0)
            p = net(x)
            if acc is None:
                acc = torch.zeros((1, p.shape[1], D, H, W), dtype=torch.float32, device=vol.device)
                cnt = torch.zeros((1, 1, D, H, W), dtype=torch.float32, device=vol.device)
            for j, (d0, h0, w0) in enumerate(chunk):
                acc[:, :, d0:d0 + window[0], h0:h0 + window[1], w0:w0 + window[2]] += p[j:j + 1]
                cnt[:, :, d0:d0 + window[0], h0:h0 + window[1], w0:w0 + window[2]] += 1
    if acc is None:   # a rank with no window (more ranks than windows)
        out_ch = net.final_conv.out_channels if hasattr(net.final_conv, 'out_channels') else 1
        acc = torch.zeros((1, out_ch, D, H, W), dtype=torch.float32, device=vol.device)
        cnt = torch.zeros((1, 1, D, H, W), dtype=torch.float32, device=vol.device)
    if distributed:
        dist.all_reduce(acc)
        dist.all_reduce(cnt)
    net.train(was_training)
    pred = acc / cnt
    return pred, pred >= threshold
