"""B200-native drop-in for the reference's `src.models.networks.GatedUNet` module (the U-Net of the ad-attention side-track,
scripts/ad_attention_unet-2D/adUNet2D_scripts.py:33): same classes (`UNet`, `ConvBlock`, `ConvLayer`, `GatedConv`), constructor
signatures, attribute tree and state-dict keys as /root/reference/code/src/models/networks/GatedUNet.py (UNet :5-112, ConvBlock
:114-165, ConvLayer :167-240, GatedConv :242-322).

The parameters live in ordinary nn.Conv / nn.BatchNorm sub-modules; `forward` runs the engine's kernels: a ConvLayer is the
Conv -> BatchNorm -> ReLU unit of the plain U-Net (`ops.ConvBnRelu`), a GatedConv is that unit times the sigmoid of a second conv of the
same input (`ops.ConvBias` + `ops.GateMul`, GatedUNet.py:303-322).  The engine covers what the reference U-Net constructor can build:
stride 1, dilation 1, zero "same" padding, activation 'relu' or 'none', no spectral norm; anything else raises NotImplementedError
instead of falling back silently.  CUDA only."""
import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from ich_b200 import ops  # noqa: E402
from src.models.networks.UNet import _begin_forward, _flush_nbt, _was_4d, _PENDING_NBT, _dropout_list, _filters, _not_built  # noqa: E402

_ACTIVATIONS = {'relu': nn.ReLU, 'lrelu': lambda: nn.LeakyReLU(0.2), 'prelu': nn.PReLU, 'selu': nn.SELU, 'tanh': nn.Tanh, 'sigmoid': nn.Sigmoid}


def _activation(name):
    if name == 'none':
        return None
    assert name in _ACTIVATIONS, f"Unsupported activation: {name}"
    return _ACTIVATIONS[name]()


def _check_engine_conv(conv, what):
    k = conv.kernel_size[0]
    if any(s != 1 for s in conv.stride) or any(d != 1 for d in conv.dilation) or conv.padding_mode != 'zeros' or \
            any(kk != k for kk in conv.kernel_size) or k not in (1, 3) or any(p != k // 2 for p in conv.padding):
        _not_built(f'{what} with kernel {conv.kernel_size}, stride {conv.stride}, padding {conv.padding}, dilation {conv.dilation}, '
                   f'padding_mode {conv.padding_mode}')


def _conv_norm_act(x, conv, norm, activation, training):
    """conv -> (BatchNorm) -> ReLU / identity on channel-last engine tensors."""
    if activation is not None and not isinstance(activation, nn.ReLU):
        _not_built(f'activation {type(activation).__name__} (the reference U-Net constructor only builds relu / none)')
    relu = activation is not None
    if norm is None:
        return ops.ConvBias.apply(x, conv.weight, conv.bias, relu)
    train = training or not norm.track_running_stats
    z = ops.ConvBnRelu.apply(x, conv.weight, conv.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var, train, relu)
    if train and norm.track_running_stats:
        _PENDING_NBT.append(norm.num_batches_tracked)
    return z


class ConvLayer(nn.Module):
    """Conv -> (BatchNorm) -> activation.  Reference: GatedUNet.py:167-240."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True, padding_mode='zeros',
                 activation='relu', batch_norm=True, use_3D=False, sn=False, power_iter=1):
        super(ConvLayer, self).__init__()
        self.activation = _activation(activation)
        conv_fn = nn.Conv3d if use_3D else nn.Conv2d
        batchnorm_fn = nn.BatchNorm3d if use_3D else nn.BatchNorm2d
        if sn:
            _not_built('spectral normalisation (sn=True)')
        self.conv = conv_fn(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, bias=bias,
                            padding_mode=padding_mode)
        self.norm = batchnorm_fn(out_channels) if batch_norm else None

    def forward_cl(self, x):
        _check_engine_conv(self.conv, 'ConvLayer')
        return _conv_norm_act(x, self.conv, self.norm, self.activation, self.training)

    def forward(self, x):
        _begin_forward()
        was_4d = _was_4d(x)
        out = ops.from_channels_last(self.forward_cl(ops.to_channels_last(x)), was_4d)
        _flush_nbt()
        return out


class GatedConv(nn.Module):
    """Gated convolution (Yu et al. 2018): act(BN(conv_feat(x))) * sigmoid(conv_gate(x)).  Reference: GatedUNet.py:242-322."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True, padding_mode='zeros',
                 activation='relu', batch_norm=True, use_3D=False):
        super(GatedConv, self).__init__()
        self.activation = _activation(activation)
        conv_fn = nn.Conv3d if use_3D else nn.Conv2d
        batchnorm_fn = nn.BatchNorm3d if use_3D else nn.BatchNorm2d
        self.conv_feat = conv_fn(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, bias=bias,
                                 padding_mode=padding_mode)
        self.conv_gate = conv_fn(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, bias=bias,
                                 padding_mode=padding_mode)
        self.sigmoid = nn.Sigmoid()
        self.norm = batchnorm_fn(out_channels) if batch_norm else None

    def forward_cl(self, x):
        _check_engine_conv(self.conv_feat, 'GatedConv')
        feat = _conv_norm_act(x, self.conv_feat, self.norm, self.activation, self.training)         # :312-317
        gate = ops.ConvBias.apply(x, self.conv_gate.weight, self.conv_gate.bias, False)             # :319 (sigmoid inside GateMul)
        return ops.GateMul.apply(feat, gate)                                                         # :321

    def forward(self, x):
        _begin_forward()
        was_4d = _was_4d(x)
        out = ops.from_channels_last(self.forward_cl(ops.to_channels_last(x)), was_4d)
        _flush_nbt()
        return out


class ConvBlock(nn.Module):
    """Two ConvLayer / GatedConv layers (+ Dropout if p > 0).  Reference: GatedUNet.py:114-165."""

    def __init__(self, in_channels, out_channels, mid_channels=None, kernel_size=3, use_3D=False, p_dropout=0.0, use_gatedConv=False):
        super(ConvBlock, self).__init__()
        assert 0.0 <= p_dropout <= 1.0, f'Dropout probaility must be in [0.0, 1.0]. Given {p_dropout}.'
        self.dropout = nn.Dropout(p=p_dropout)
        mid_channels = mid_channels if mid_channels else out_channels
        layer = GatedConv if use_gatedConv else ConvLayer
        kw = dict(kernel_size=kernel_size, padding=1, bias=True, padding_mode='zeros', activation='relu', batch_norm=True, use_3D=use_3D)
        self.conv1 = layer(in_channels=in_channels, out_channels=mid_channels, **kw)
        self.conv2 = layer(in_channels=mid_channels, out_channels=out_channels, **kw)

    def forward_cl(self, x):
        if self.dropout.p > 0.0 and self.training:
            _not_built('Dropout inside the GatedUNet blocks (the shipped ad-attention config trains with p_dropout = 0.0)')
        return self.conv2.forward_cl(self.conv1.forward_cl(x))

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        out = ops.from_channels_last(self.forward_cl(ops.to_channels_last(input)), was_4d)
        _flush_nbt()
        return out


class UNet(nn.Module):
    """2-D / 3-D U-Net with optional gated convolutions.  Reference: GatedUNet.py:5-112."""

    def __init__(self, depth=5, use_3D=False, bilinear=False, in_channels=1, out_channels=1, top_filter=64, midchannels_factor=2,
                 p_dropout=0.5, use_final_activation=True, use_gatedConv=False):
        super(UNet, self).__init__()
        p_dropout_list = _dropout_list(p_dropout, depth)
        self.return_bottleneck = False
        self.down_block = nn.ModuleList()
        self.up_samp = nn.ModuleList()
        self.up_block = nn.ModuleList()
        down, bottleneck = _filters(in_channels, top_filter, depth)
        up_filters = [(top_filter * 2 ** d, top_filter * 2 ** (d - 1)) for d in range(depth - 1, 0, -1)]
        for down_ch, up_ch, p in zip(down, up_filters, p_dropout_list[:-1]):                     # construction order of :57-68
            self.down_block.append(ConvBlock(down_ch[0], down_ch[1], mid_channels=down_ch[1] // midchannels_factor, use_3D=use_3D, p_dropout=p,
                                             use_gatedConv=use_gatedConv))
            if bilinear or use_gatedConv:
                self.up_block.append(ConvBlock(int(1.5 * up_ch[0]), up_ch[1], mid_channels=up_ch[1], use_3D=use_3D, use_gatedConv=use_gatedConv))
                self.up_samp.append(nn.Upsample(scale_factor=2, mode='trilinear' if use_3D else 'bilinear', align_corners=True))
            else:
                self.up_block.append(ConvBlock(up_ch[0], up_ch[1], mid_channels=up_ch[1], use_3D=use_3D, use_gatedConv=False))
                convT = nn.ConvTranspose3d if use_3D else nn.ConvTranspose2d
                self.up_samp.append(convT(up_ch[0], up_ch[1], kernel_size=2, stride=2))
        self.bottleneck_block = ConvBlock(bottleneck[0], bottleneck[1], mid_channels=bottleneck[1] // midchannels_factor, use_3D=use_3D,
                                          p_dropout=p_dropout_list[-1], use_gatedConv=use_gatedConv)
        self.downpool = nn.MaxPool3d(kernel_size=2, stride=2) if use_3D else nn.MaxPool2d(kernel_size=2, stride=2)
        if use_gatedConv:
            self.final_conv = GatedConv(top_filter, out_channels, kernel_size=1, use_3D=use_3D, activation='none', batch_norm=False)
        else:
            self.final_conv = nn.Conv3d(top_filter, out_channels, kernel_size=1) if use_3D else nn.Conv2d(top_filter, out_channels, kernel_size=1)
        if use_final_activation:
            self.final_activation = nn.Softmax(dim=1) if out_channels > 1 else nn.Sigmoid()
        else:
            self.final_activation = nn.Identity()

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        fd = 2 if isinstance(self.downpool, nn.MaxPool3d) else 1
        x = ops.to_channels_last(input)
        f = 2 ** len(self.down_block)
        d, h, w = x.shape[1:4]
        if (fd == 2 and d % f) or h % f or w % f:
            raise RuntimeError(f'ich_b200: spatial size {(d, h, w) if fd == 2 else (h, w)} must be divisible by {f} '
                               f'(the reference has no pad/crop logic either, GatedUNet.py:101-103)')
        res = []
        for block in self.down_block:                                                       # :92-95
            x = block.forward_cl(x)
            skip, x = ops.PoolSkip.apply(x, fd, 0)
            res.append(skip)
        x = self.bottleneck_block.forward_cl(x)                                             # :97
        x_bottleneck = x
        for up, block, r in zip(self.up_samp, self.up_block, res[::-1]):                    # :101-103
            if isinstance(up, nn.Upsample):
                x = ops.UpsampleCat.apply(x, r, fd)
            else:
                x = ops.UpConvCat.apply(x, r, up.weight, up.bias, fd, None)
            x = block.forward_cl(x)
        act = 0 if isinstance(self.final_activation, nn.Identity) else (2 if isinstance(self.final_activation, nn.Softmax) else 1)
        if isinstance(self.final_conv, GatedConv):                                          # :105 with the gated 1x1 head of :76
            g = self.final_conv.forward_cl(x)
            c = g.shape[-1]
            eye = torch.eye(c, dtype=torch.float32, device=g.device).view(c, c, *([1] * (2 if was_4d else 3)))
            out = ops.Head.apply(g, eye, None, act)          # identity 1x1 head: only its activation + fp32 NC(D)HW layout are used
        else:
            out = ops.Head.apply(x, self.final_conv.weight, self.final_conv.bias, act)
        _flush_nbt()
        if was_4d:
            out = out.squeeze(2)
        if self.return_bottleneck:
            return out, ops.from_channels_last(x_bottleneck, was_4d)
        return out
