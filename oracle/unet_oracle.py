"""CPU (torch fp32/fp64) restatement of the reference U-Net hot path. TEST INFRASTRUCTURE ONLY.

The reference's arithmetic lives in PyTorch (un-pinned third-party dependency, torch 2.11 here), so
the oracle restates the reference's *graph* with torch CPU functional ops, driven by a plain
state_dict with the reference's key names.  Each function cites the reference lines it follows
(paths relative to /root/reference/code/src).  Pinned against the reference modules themselves by
oracle/make_golden.py (fixtures in tests/golden/) and tests/test_oracle.py.

Nothing here is imported by the product path.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _dims(use_3D):
    return (F.conv3d, F.conv_transpose3d, F.max_pool3d) if use_3D else (F.conv2d, F.conv_transpose2d, F.max_pool2d)


def batch_norm(x, sd, prefix, training, new_stats=None):
    """nn.BatchNorm{2,3}d as used at models/networks/UNet.py:154,156,159,161 (eps 1e-5, momentum 0.1)."""
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    red = [0] + list(range(2, x.ndim))
    shape = [1, -1] + [1] * (x.ndim - 2)
    if training:
        mean = x.mean(dim=red)
        var = x.var(dim=red, unbiased=False)
        if new_stats is not None:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                new_stats[prefix + '.running_mean'] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                new_stats[prefix + '.running_var'] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * n / max(n - 1, 1)
                new_stats[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
    else:
        mean, var = rm, rv
    xhat = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    return xhat * w.view(shape) + b.view(shape)


def conv_block(x, sd, prefix, training, use_3D, new_stats=None):
    """ConvBlock.forward, models/networks/UNet.py:163-177 with p_dropout == 0 (Dropout skipped at :175)."""
    conv, _, _ = _dims(use_3D)
    x = conv(x, sd[prefix + '.conv1.weight'], sd[prefix + '.conv1.bias'], padding=1)
    x = F.relu(batch_norm(x, sd, prefix + '.bn1', training, new_stats))
    x = conv(x, sd[prefix + '.conv2.weight'], sd[prefix + '.conv2.bias'], padding=1)
    x = F.relu(batch_norm(x, sd, prefix + '.bn2', training, new_stats))
    return x


def _n_blocks(sd, name):
    return len({k.split('.')[1] for k in sd if k.startswith(name + '.')})


def unet_forward(x, sd, use_3D=True, training=True, use_final_activation=True, return_bottleneck=False, new_stats=None):
    """UNet.forward, models/networks/UNet.py:93-127.  bilinear=True nets (nn.Upsample(scale_factor=2, tri/bilinear,
    align_corners=True) instead of ConvTranspose, :69-72) are recognised by the absence of up_samp weights in the state dict."""
    conv, convT, pool = _dims(use_3D)
    bilinear = 'up_samp.0.weight' not in sd
    res = []
    for i in range(_n_blocks(sd, 'down_block')):                       # :106-109
        x = conv_block(x, sd, f'down_block.{i}', training, use_3D, new_stats)
        res.append(x)
        x = pool(x, kernel_size=2, stride=2)
    x = conv_block(x, sd, 'bottleneck_block', training, use_3D, new_stats)   # :112
    xb = x
    for i, r in zip(range(_n_blocks(sd, 'up_block')), res[::-1]):     # :117-119
        if bilinear:
            x = F.interpolate(x, scale_factor=2, mode='trilinear' if use_3D else 'bilinear', align_corners=True)
        else:
            x = convT(x, sd[f'up_samp.{i}.weight'], sd[f'up_samp.{i}.bias'], stride=2)
        x = conv_block(torch.cat([r, x], dim=1), sd, f'up_block.{i}', training, use_3D, new_stats)
    x = conv(x, sd['final_conv.weight'], sd['final_conv.bias'])        # :122
    if use_final_activation:                                            # :85-91
        x = torch.sigmoid(x) if x.shape[1] == 1 else torch.softmax(x, dim=1)
    return (x, xb) if return_bottleneck else x


def unet_encoder_forward(x, sd, use_3D=True, training=True, return_bottleneck=False, new_stats=None):
    """UNet_Encoder.forward, models/networks/UNet.py:298-326; MLPHead.forward :196-209."""
    _, _, pool = _dims(use_3D)
    for i in range(_n_blocks(sd, 'down_block')):
        x = conv_block(x, sd, f'down_block.{i}', training, use_3D, new_stats)
        x = pool(x, kernel_size=2, stride=2)
    x = conv_block(x, sd, 'bottleneck_block', training, use_3D, new_stats)
    x = x.mean(dim=tuple(range(2, x.ndim)), keepdim=True)              # AdaptiveAvgPool(1) :318
    xb = x
    h = torch.flatten(x, 1)
    n_fc = _n_blocks({k[len('mlp_head.'):]: v for k, v in sd.items() if k.startswith('mlp_head.')}, 'fc_layers')
    for i in range(n_fc):
        h = F.linear(h, sd[f'mlp_head.fc_layers.{i}.weight'], sd[f'mlp_head.fc_layers.{i}.bias'])
        if i < n_fc - 1:
            h = F.relu(h)
    return (h, xb) if return_bottleneck else h


def partial_unet_forward(x, sd, use_3D=False, training=True, return_bottleneck=False, new_stats=None):
    """Partial_UNet.forward, models/networks/UNet.py:401-435; ConvHead.forward :230-243."""
    conv, convT, pool = _dims(use_3D)
    res = []
    for i in range(_n_blocks(sd, 'down_block')):
        x = conv_block(x, sd, f'down_block.{i}', training, use_3D, new_stats)
        res.append(x)
        x = pool(x, kernel_size=2, stride=2)
    x = conv_block(x, sd, 'bottleneck_block', training, use_3D, new_stats)
    xb = x
    n_dec = _n_blocks(sd, 'up_samp')
    for i, r in zip(range(n_dec), res[::-1][:n_dec]):
        x = convT(x, sd[f'up_samp.{i}.weight'], sd[f'up_samp.{i}.bias'], stride=2)
        x = conv_block(torch.cat([r, x], dim=1), sd, f'up_block.{i}', training, use_3D, new_stats)
    n_head = _n_blocks({k[len('final_conv.'):]: v for k, v in sd.items() if k.startswith('final_conv.')}, 'conv_layers')
    for i in range(n_head):
        x = conv(x, sd[f'final_conv.conv_layers.{i}.weight'], sd[f'final_conv.conv_layers.{i}.bias'])
        if i < n_head - 1:
            x = F.relu(x)
    return (x, xb) if return_bottleneck else x


def sliding_window_predict(vol, sd, window, stride, use_3D=True):
    """Oracle for cfg-5: eval-mode UNet on each window, mean blending of overlaps,
    mask = pred >= 0.5 (models/optim/UNet2D.py:220)."""
    _, _, D, H, W = vol.shape
    acc = torch.zeros_like(vol)
    cnt = torch.zeros_like(vol)
    starts = lambda L, w, s: sorted(set(list(range(0, max(L - w, 0) + 1, s)) + [max(L - w, 0)]))
    for d0 in starts(D, window[0], stride[0]):
        for h0 in starts(H, window[1], stride[1]):
            for w0 in starts(W, window[2], stride[2]):
                sl = (slice(None), slice(None), slice(d0, d0 + window[0]), slice(h0, h0 + window[1]), slice(w0, w0 + window[2]))
                with torch.no_grad():
                    acc[sl] += unet_forward(vol[sl], sd, use_3D=use_3D, training=False)
                cnt[sl] += 1
    pred = acc / cnt
    return pred, pred >= 0.5
