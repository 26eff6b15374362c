"""CPU stand-ins for the window / staging kernels, for the HOST-LOGIC tests of ich_b200.infer (window enumeration, sharding, which
buffers are exchanged).  Test infrastructure only: the product path calls the CUDA kernels and has no CPU fallback."""
import torch


def install_window_mocks(ops, infer):
    def gather(v, starts, window):
        wd, wh, ww = window
        out = torch.zeros((starts.shape[0], wd, wh, ww, 1), dtype=v.dtype)
        for n, (d0, h0, w0) in enumerate(starts.tolist()):
            blk = v[d0:d0 + wd, h0:h0 + wh, w0:w0 + ww]
            out[n, :blk.shape[0], :blk.shape[1], :blk.shape[2], 0] = blk
        return out

    def scatter(pred, starts, window, shape, overlap, thr, acc, cnt, mask):
        wd, wh, ww = window
        D, H, W = shape
        for n, (d0, h0, w0) in enumerate(starts.tolist()):
            p = pred[n, 0, :D - d0, :H - h0, :W - w0].float()
            sl = (slice(d0, d0 + wd), slice(h0, h0 + wh), slice(w0, w0 + ww))
            if overlap:
                acc[sl] += p
                cnt[sl] += 1
            else:
                if acc is not None:
                    acc[sl] = p
                if mask is not None:
                    mask[sl] = (p >= thr).to(torch.uint8)

    def blend(acc, cnt, thr, mask):
        acc.copy_(torch.where(cnt > 0, acc / cnt.clamp_min(1), torch.zeros_like(acc)))
        if mask is not None:
            mask.copy_((acc >= thr).to(torch.uint8))

    ops.window_gather, ops.window_scatter, ops.blend_threshold = gather, scatter, blend
    infer._cast = lambda v: v.float()


class ChannelLastAdapter(torch.nn.Module):
    """Lets a plain NCDHW torch module stand in for a drop-in network fed with staged [N, D, H, W, 1] windows."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        return self.inner(x.permute(0, 4, 1, 2, 3).float())
