"""CPU restatement of the reference's CT windowing and of the slice-wise volume segmentation rule. TEST INFRASTRUCTURE ONLY.

Paths relative to /root/reference/code/src.  Pinned against the reference function itself by oracle/make_golden.py
(tests/golden/window_ct.pt) and tests/test_oracle.py.  Nothing here is imported by the product path."""
import numpy as np
import torch


def window_ct(ct_scan, win_center=40, win_width=120, out_range=(0, 1)):
    """utils/ct_utils.py:13-36: rescale [center - width/2, center + width/2] linearly to out_range, clip to out_range.
    The reference evaluates it in float64 numpy (nibabel's get_fdata) -- so does this."""
    ct = np.asarray(ct_scan, dtype=np.float64)
    win_min = win_center - win_width / 2
    win_max = win_center + win_width / 2
    out = (out_range[1] - out_range[0]) * (ct - win_min) / (win_max - win_min) + out_range[0]
    return np.clip(out, out_range[0], out_range[1])


def segment_volume_slices(vol_hws, forward, window=None, batch_size=4):
    """UNet2D.segement_volume, models/optim/UNet2D.py:272-314, without the resize (input_size == slice size) and without NIfTI I/O:
    rot90 counter-clockwise (:285), optional window (:286-287), every slice through `forward` (a callable NCHW fp32 -> probabilities,
    the eval-mode 2-D net) in batches (:292-301), pred >= 0.5 -> uint8 0 / 255 (:299,303), concatenate, rot90 clockwise (:309)."""
    data = np.rot90(np.asarray(vol_hws, dtype=np.float64), axes=(0, 1))
    if window:
        data = window_ct(data, win_center=window[0], win_width=window[1], out_range=(0, 1))
    preds = []
    for s in range(0, data.shape[2], batch_size):
        x = torch.from_numpy(np.ascontiguousarray(data[:, :, s:s + batch_size])).float().permute(2, 0, 1).unsqueeze(1)   # B x 1 x H x W
        with torch.no_grad():
            p = forward(x)
        p = (p >= 0.5)
        preds.append(p.squeeze(1).permute(1, 2, 0).cpu().numpy().astype(np.uint8) * 255)
    vol_pred = np.concatenate(preds, axis=2)
    return np.rot90(vol_pred, axes=(1, 0)).astype(np.uint8)
