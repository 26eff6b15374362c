"""Full-volume inference (BASELINE cfg-5).

`sliding_window_predict` is the 3-D window driver: the windowing rule is the one SURVEY section 8d fixes as the oracle (eval-mode
network on each window, mean blending where windows overlap, mask = pred >= 0.5 as in models/optim/UNet2D.py:220).  Everything
between the raw volume and the mask runs on the device in this library's kernels: staging (window / clip / cast,
`ich_stage_ct`), window extraction (`ich_window_gather`), the network, stitching + threshold (`ich_window_scatter`,
`ich_blend_threshold`).  Under torch.distributed the windows are sharded over the ranks; non-overlapping windows exchange only
the uint8 mask (one all-reduce(MAX) of D*H*W bytes = the gather of SURVEY section 8e), overlapping ones sum the fp32
accumulators.

`segement_volume` (sic) keeps the name, arguments and result conventions of the reference method
(models/optim/UNet2D.py:272-314): rot90, optional CT window, every slice (2-D nets) or the whole volume in 3-D windows
(3-D nets) through the network, `pred >= 0.5`, uint8 0 / 255 volume rotated back, optional NIfTI save (needs nibabel)."""
import numpy as np
import torch
import torch.distributed as dist

from . import config, ops
from .dp import shard_indices


_STARTS = {}      # (window corners of this rank, device) -> int32 device tensor


def window_starts(L, w, s):
    return sorted(set(list(range(0, max(L - w, 0) + 1, s)) + [max(L - w, 0)]))


def _one_channel_engine_volume(vol):
    """[1, 1, D, H, W] (fp32 NCDHW, or any staged engine tensor) -> [D, H, W] in the engine dtype (a 1-channel volume is its own
    channel-last form)."""
    if getattr(vol, '_ich_staged', False):
        return vol.reshape(vol.shape[1], vol.shape[2], vol.shape[3])
    _, _, D, H, W = vol.shape
    v = vol.reshape(D, H, W)
    if v.dtype != config.act_dtype():
        v = _cast(v)
    return v.contiguous()


def _cast(v):
    """Plain cast of an in-range fp32 volume to the engine dtype through the staging kernel (identity map, no clipping in effect)."""
    big = 1.0e30                       # window [-big, big] -> [-big, big]: scale exactly 1, offset exactly 0, clip never active
    out = torch.empty(v.shape, dtype=config.act_dtype(), device=v.device)
    v = v.contiguous().float()
    ops.call('ich_stage_ct', v.data_ptr(), 0, out.data_ptr(), config.dtype_code(out.dtype), v.numel(), -big, big, -big, big, ops._stream())
    return out


def sliding_window_predict(net, vol, window, stride=None, batch=4, threshold=0.5, distributed=None, return_pred=True):
    """vol [1, 1, D, H, W] on the device (fp32, values as the network expects them) or a staged engine tensor [1, D, H, W, 1]
    (ops.stage_ct + ops.staged).  Returns (pred fp32 [1, out_ch, D, H, W] or None, mask bool [1, 1, D, H, W]).

    Single-class heads (the reference's segmentation nets) take the device-side path described in the module docstring; multi-class
    heads fall back to stitching with torch indexing (no reference caller uses them for volumes)."""
    stride = tuple(stride or window)
    window = tuple(window)
    if getattr(vol, '_ich_staged', False):
        D, H, W = vol.shape[1:4]
        in_ch = vol.shape[4]
    else:
        in_ch, (D, H, W) = vol.shape[1], vol.shape[2:]
    if in_ch != 1:
        raise RuntimeError('ich_b200.infer: the window driver takes one-channel CT volumes (the reference nets have in_channels=1)')
    wins = [(d0, h0, w0) for d0 in window_starts(D, window[0], stride[0]) for h0 in window_starts(H, window[1], stride[1])
            for w0 in window_starts(W, window[2], stride[2])]
    # windows overlap when the stride is smaller than the window OR when the last window of an axis is clamped to the volume edge
    # (L not a multiple of the stride): either way the stitched value is the mean over the covering windows
    overlap = any(b - a < w for L, w, s in zip((D, H, W), window, stride) for a, b in zip(window_starts(L, w, s), window_starts(L, w, s)[1:]))
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    mine = [wins[i] for i in shard_indices(len(wins))] if distributed else wins
    dev = vol.device
    was_training = net.training
    net.eval()
    need_pred = return_pred or overlap
    acc = torch.zeros((D, H, W), dtype=torch.float32, device=dev) if need_pred else None
    cnt = torch.zeros((D, H, W), dtype=torch.float32, device=dev) if overlap else None
    mask = torch.zeros((D, H, W), dtype=torch.uint8, device=dev)
    v = _one_channel_engine_volume(vol)
    with torch.no_grad():
        if mine:
            # window corners on the device, uploaded once per (geometry, shard, device): the call itself then makes no host -> device copy
            # and no synchronisation, so a fixed-shape inference loop can be replayed as one CUDA graph (ich_b200.graph.GraphedStep)
            key = (tuple(mine), str(dev))
            starts_all = _STARTS.get(key)
            if starts_all is None:
                if len(_STARTS) > 64:
                    _STARTS.clear()
                starts_all = _STARTS[key] = torch.tensor(mine, dtype=torch.int32, device=dev)
        for i in range(0, len(mine), batch):
            starts = starts_all[i:i + batch].contiguous()
            x = ops.staged(ops.window_gather(v, starts, window))
            p = net(x)
            if isinstance(p, tuple):
                p = p[0]
            if p.shape[1] != 1:
                raise RuntimeError('ich_b200.infer: multi-class volume stitching is not built (no reference caller); use out_channels=1')
            ops.window_scatter(p, starts, window, (D, H, W), overlap, threshold, acc, cnt, None if overlap else mask)
    if distributed:
        if overlap:
            dist.all_reduce(acc)
            dist.all_reduce(cnt)
        else:
            dist.all_reduce(mask, op=dist.ReduceOp.MAX)        # windows are disjoint: MAX of the 0 / 1 bytes = gather of the masks
            if return_pred:
                dist.all_reduce(acc)
    if overlap:
        ops.blend_threshold(acc, cnt, threshold, mask)
    net.train(was_training)
    pred = acc.view(1, 1, D, H, W) if return_pred else None
    return pred, mask.view(1, 1, D, H, W).bool()


def _volume_array(vol):
    """nibabel image (duck-typed: get_fdata()), numpy array or torch tensor -> (numpy / torch array [H, W, S], affine or None)."""
    if hasattr(vol, 'get_fdata'):
        return vol.get_fdata(), getattr(vol, 'affine', None)
    return vol, None


def segement_volume(net, vol, save_fn=None, window=None, input_size=(256, 256), return_pred=False, batch_size=16, device=None,
                    window_3d=(32, 128, 128)):
    """Drop-in for `UNet2D.segement_volume(vol, save_fn, window, input_size, return_pred)` (models/optim/UNet2D.py:272-314), with the
    trainer's `self.unet / self.batch_size / self.device` passed explicitly.

    vol: a nibabel image (anything with get_fdata() / affine) or an array [H, W, S] of Hounsfield units (int16 / uint16 / fp32 numpy or
    torch).  Steps as in the reference: 90 degree counter-clockwise rotation in the (H, W) plane, CT window (center, width) -> [0, 1] when
    `window` is given, the network on every slice (2-D nets, batches of `batch_size` slices) or on 3-D windows of `window_3d`
    (3-D nets: `sliding_window_predict` over the [S, H, W] volume), pred >= 0.5, uint8 0 / 255, rotated back.  The window / cast runs in
    the staging kernel on the device; the raw volume crosses PCIe in its own dtype (int16: half the bytes of the reference's fp32).
    Slices are resized to `input_size` only if they differ from it (bilinear + nearest back, torch; the reference uses
    skimage.transform.resize, which is not in this image -- sizes that need no resize are exact).
    Returns the uint8 prediction volume [H, W, S] (a nibabel Nifti1Pair when nibabel is importable and the input was one) if
    `return_pred`; saves to `save_fn` when given (needs nibabel)."""
    data, affine = _volume_array(vol)
    p0 = next(net.parameters())
    dev = torch.device(device) if device is not None else p0.device
    t = torch.as_tensor(np.ascontiguousarray(data)) if not torch.is_tensor(data) else data
    if t.dtype == torch.float64:
        t = t.float()
    t = t.to(dev, non_blocking=True)
    t = torch.rot90(t, 1, (0, 1))                                    # np.rot90(vol, axes=(0, 1)), UNet2D.py:285
    Hr, Wr, S = t.shape
    vol_shw = t.permute(2, 0, 1).contiguous()                        # [S, H, W]: slices are the batch / depth axis
    if window:
        x = ops.stage_ct(vol_shw, win_center=window[0], win_width=window[1], out_range=(0, 1))       # UNet2D.py:286-287
    else:
        x = _cast(vol_shw.float())
    is_3d = any(isinstance(m, torch.nn.Conv3d) for m in net.modules())
    resize = (not is_3d) and input_size is not None and tuple(input_size) != (Hr, Wr)
    was_training = net.training
    net.eval()
    with torch.no_grad():
        if is_3d:
            _, m = sliding_window_predict(net, ops.staged(x.view(1, S, Hr, Wr, 1)), window_3d, batch=max(1, batch_size // 4), return_pred=False,
                                          distributed=False)
            mask = m.view(S, Hr, Wr)
        else:
            out = []
            for s in range(0, S, batch_size):                                                       # UNet2D.py:292-301
                xs = x[s:s + batch_size]
                if resize:
                    xs = torch.nn.functional.interpolate(xs.float().unsqueeze(1), size=tuple(input_size), mode='bilinear', antialias=True,
                                                         align_corners=False).squeeze(1).to(x.dtype)
                b, h, w = xs.shape
                pred = net(ops.staged(xs.contiguous().view(b, 1, h, w, 1), was_4d=True))
                if isinstance(pred, tuple):
                    pred = pred[0]
                m = pred >= 0.5
                if resize:
                    m = torch.nn.functional.interpolate(m.float(), size=(Hr, Wr), mode='nearest') >= 0.5
                out.append(m[:, 0])
            mask = torch.cat(out, 0)
    net.train(was_training)
    vol_pred = torch.rot90((mask.to(torch.uint8) * 255).permute(1, 2, 0), 1, (1, 0)).contiguous()   # UNet2D.py:303,309-311
    result = vol_pred.cpu().numpy()
    if save_fn is not None or (return_pred and affine is not None):
        try:
            import nibabel as nib
        except ImportError:
            if save_fn is not None:
                raise RuntimeError('ich_b200.infer.segement_volume: saving a NIfTI volume needs nibabel, which is not installed')
            nib = None
        if nib is not None:
            nii = nib.Nifti1Pair(result, affine if affine is not None else np.eye(4))
            if save_fn is not None:
                nib.save(nii, save_fn)
            result = nii
    if return_pred:
        return result
