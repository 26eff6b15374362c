"""B200-native drop-in for the reference's `src.models.networks.UNet` module.

Same classes, constructor signatures, attribute tree and state-dict keys as the reference
(/root/reference/code/src/models/networks/UNet.py: UNet :18-127, ConvBlock :129-177, MLPHead :179-209,
ConvHead :211-243, UNet_Encoder :245-326, Partial_UNet :328-435), so the reference trainers
(`UNet2D`, `Contrastive`, ...) and scripts run on top of it unchanged.  The parameters live in ordinary
nn.Conv/BatchNorm/ConvTranspose/Linear sub-modules (state-dict compatible in both directions), but `forward` never
calls them: every arithmetic step runs in the hand-written sm_100a kernels of libich_b200.so through
`ich_b200.ops` (channel-last bf16/fp32 activations, tcgen05 implicit-GEMM convs).  CUDA only -- no CPU fallback.
"""
import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from ich_b200 import config as _cfg, ops  # noqa: E402


def _dropout_list(p_dropout, depth):
    """Reference validation of the p_dropout argument (UNet.py:47-53)."""
    if isinstance(p_dropout, float):
        return [p_dropout] * depth
    if isinstance(p_dropout, list):
        assert len(p_dropout) == depth, (f'p_dropout provided as list should have the same length as depth. '
                                         f'p_dropout {len(p_dropout)} vs depth {depth}.')
        return p_dropout
    raise TypeError(f'p_dropout list not supported. Should be float or list of float. Given {type(p_dropout)}.')


def _filters(in_channels, top_filter, depth):
    """Filter plan of UNet.py:61-63."""
    down = [(in_channels, top_filter)] + [(top_filter * 2 ** d, top_filter * 2 ** (d + 1)) for d in range(depth - 2)]
    bottleneck = (top_filter * 2 ** (depth - 2), top_filter * 2 ** (depth - 1))
    return down, bottleneck


# BatchNorm's num_batches_tracked counters of one forward pass are bumped together (one multi-tensor launch instead of one per layer)
_PENDING_NBT = []


def _flush_nbt():
    if _PENDING_NBT:
        torch._foreach_add_(_PENDING_NBT, 1)
        _PENDING_NBT.clear()


def _begin_forward():
    """Start of a top-level forward: counters queued by a forward pass that raised half-way are dropped, not applied to this pass."""
    _PENDING_NBT.clear()


def _was_4d(input):
    """2-D nets take NCHW (4-D) inputs; an already staged engine tensor [N, 1, H, W, C] carries the flag (ops.staged)."""
    return input.dim() == 4 or bool(getattr(input, '_ich_was_4d', False))


def _not_built(what):
    raise NotImplementedError(f'ich_b200: {what} is not part of the B200 hot path yet (SURVEY section 8f); '
                              f'refusing to fall back silently')


class ConvBlock(nn.Module):
    """[Conv k3 p1 -> BatchNorm -> ReLU] x 2 (+ Dropout if p > 0). Reference: UNet.py:129-177."""

    def __init__(self, in_channels, out_channels, mid_channels=None, kernel_size=3, use_3D=False, p_dropout=0.0):
        super(ConvBlock, self).__init__()
        assert 0.0 <= p_dropout <= 1.0, f'Dropout probaility must be in [0.0, 1.0]. Given {p_dropout}.'
        self.activation = nn.ReLU()
        self.dropout = nn.Dropout(p=p_dropout)
        mid_channels = mid_channels if mid_channels else out_channels
        conv = nn.Conv3d if use_3D else nn.Conv2d
        bn = nn.BatchNorm3d if use_3D else nn.BatchNorm2d
        self.conv1 = conv(in_channels=in_channels, out_channels=mid_channels, kernel_size=kernel_size, padding=1)
        self.bn1 = bn(mid_channels)
        self.conv2 = conv(in_channels=mid_channels, out_channels=out_channels, kernel_size=kernel_size, padding=1)
        self.bn2 = bn(out_channels)

    def _unit(self, x, conv, bn, concat_c=0, drop_p=0.0):
        if conv.padding[0] * 2 + 1 != conv.kernel_size[0]:
            _not_built(f'kernel_size={conv.kernel_size} with padding={conv.padding}')
        training = self.training or not bn.track_running_stats
        if not training and not concat_c and _cfg.get('fold_eval_bn') and not torch.is_grad_enabled():
            # inference: BatchNorm folded into the conv (one launch, bias + ReLU epilogue)
            w_eff, b_eff = ops.folded_eval_unit(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
            return ops.conv_forward(x, w_eff, b_eff, relu=True)
        z = ops.ConvBnRelu.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, True, concat_c,
                                 drop_p)
        if concat_c:
            z, buf = z
            z._ich_concat_buf = buf          # the decoder's UpConvCat completes this buffer in place (zero-copy skip concat)
        if training and bn.track_running_stats:
            _PENDING_NBT.append(bn.num_batches_tracked)      # flushed at the end of the enclosing forward (_flush_nbt)
        return z

    def forward_cl(self, x, concat_c=0):
        """Channel-last engine path: x [N, D, H, W, C] in the engine dtype. concat_c > 0: the block output is laid out as the
        first channel slab of a [.., C + concat_c] buffer (it will be the skip half of a decoder concat).
        nn.Dropout (reference UNet.py:175-176) is fused into the second unit's BN-apply / BN-backward kernels (counter-based
        Philox mask, no mask tensor); being RNG-dependent it matches the reference statistically, not element-wise."""
        drop_p = float(self.dropout.p) if (self.dropout.p > 0.0 and self.training) else 0.0
        if self._pad_mid(x):
            return self._forward_cl_padded_mid(x, concat_c, drop_p)
        x = self._unit(x, self.conv1, self.bn1)
        return self._unit(x, self.conv2, self.bn2, concat_c, drop_p)

    # ---- 8-channel mid tensors (top_filter = 16 nets with midchannels_factor = 2: BASELINE.json configs[0]) ------------------------------
    # The tcgen05 kernels work on 16-channel K chunks / cout blocks; an 8-channel tensor between the two units of a block would send both
    # full-resolution convs of that block to the CUDA-core kernels (measured: 9.1 of 10.2 ms per cfg-1 step).  Instead the block runs
    # with its mid tensor zero-padded to 16 channels: unit 1 gets 8 extra all-zero filters (their BatchNorm output is exactly 0), unit 2
    # gets 8 extra all-zero input channels.  The padding is built with differentiable torch ops on the (tiny) parameters, so autograd
    # hands every parameter exactly its own slice of the gradient; the statistics of the real channels are untouched.
    def _pad_mid(self, x):
        mid = self.conv1.out_channels
        return (_cfg.get('tensor_cores') and x.dtype == torch.bfloat16 and mid % 16 == 8 and self.conv2.out_channels % 16 == 0
                and self.conv1.kernel_size[0] == 3 and (x.shape[-1] == 1 or x.shape[-1] % 16 == 0))

    def _forward_cl_padded_mid(self, x, concat_c, drop_p):
        import torch.nn.functional as F
        c1, b1, c2, b2 = self.conv1, self.bn1, self.conv2, self.bn2
        nd = c1.weight.dim()
        if not (self.training or not b1.track_running_stats) and not concat_c and _cfg.get('fold_eval_bn') and not torch.is_grad_enabled():
            # inference with BatchNorm folded into the convs (ops.folded_eval_unit): pad the FOLDED weights, cached on the folded tensors
            def padded(conv, bn, pad_out):
                w, b = ops.folded_eval_unit(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
                cached = getattr(w, '_ich_padded', None)
                if cached is None:
                    cached = (F.pad(w, (0, 0) * (nd - 1) + (0, 8)), F.pad(b, (0, 8))) if pad_out else (F.pad(w, (0, 0) * (nd - 2) + (0, 8)), b)
                    w._ich_padded = cached
                return cached
            w1, bias1 = padded(c1, b1, True)
            w2, bias2 = padded(c2, b2, False)
            return ops.conv_forward(ops.conv_forward(x, w1, bias1, relu=True), w2, bias2, relu=True)
        w1 = F.pad(c1.weight, (0, 0) * (nd - 1) + (0, 8))                       # [mid + 8, Cin, k..]: 8 all-zero filters
        w2 = F.pad(c2.weight, (0, 0) * (nd - 2) + (0, 8))                       # [Cout, mid + 8, k..]: 8 all-zero input channels
        bias1 = F.pad(c1.bias, (0, 8)) if c1.bias is not None else None
        gamma = F.pad(b1.weight, (0, 8), value=1.0) if b1.weight is not None else None
        beta = F.pad(b1.bias, (0, 8)) if b1.bias is not None else None
        training = self.training or not b1.track_running_stats
        rm = F.pad(b1.running_mean, (0, 8)) if b1.running_mean is not None else None
        rv = F.pad(b1.running_var, (0, 8), value=1.0) if b1.running_var is not None else None
        z = ops.ConvBnRelu.apply(x, w1, bias1, gamma, beta, rm, rv, training, True, 0, 0.0)
        if training and b1.track_running_stats:
            with torch.no_grad():                                               # the kernel updated the padded copies in place
                b1.running_mean.copy_(rm[:-8])
                b1.running_var.copy_(rv[:-8])
            _PENDING_NBT.append(b1.num_batches_tracked)
        training2 = self.training or not b2.track_running_stats
        out = ops.ConvBnRelu.apply(z, w2, c2.bias, b2.weight, b2.bias, b2.running_mean, b2.running_var, training2, True, concat_c, drop_p)
        if concat_c:
            out, buf = out
            out._ich_concat_buf = buf
        if training2 and b2.track_running_stats:
            _PENDING_NBT.append(b2.num_batches_tracked)
        return out

    def forward_cl_head(self, x, final_conv, act):
        """This block followed by the single-class head `final_conv` (+ Sigmoid if act == 1), with the second unit and the head fused
        (ops.ConvBnReluHead; the caller checked ops.head_fusable).  Returns the network output, fp32 [N, 1, D, H, W]."""
        x = self._unit(x, self.conv1, self.bn1)
        conv, bn = self.conv2, self.bn2
        training = self.training or not bn.track_running_stats
        out = ops.ConvBnReluHead.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                       final_conv.weight, final_conv.bias, act)
        if training and bn.track_running_stats:
            _PENDING_NBT.append(bn.num_batches_tracked)
        return out

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        out = ops.from_channels_last(self.forward_cl(ops.to_channels_last(input)), was_4d)
        _flush_nbt()
        return out


class MLPHead(nn.Module):
    """Linear/ReLU projection head (no ReLU after the last layer). Reference: UNet.py:179-209.  The parameters live in nn.Linear
    sub-modules (state-dict compatible); the arithmetic runs in ich_linear_fwd / ich_linear_bwd (B x 256 matrices: latency-bound)."""

    def __init__(self, Neurons_layer=[512, 256, 128]):
        nn.Module.__init__(self)
        self.fc_layers = nn.ModuleList(nn.Linear(in_features=n_in, out_features=n_out)
                                       for n_in, n_out in zip(Neurons_layer[:-1], Neurons_layer[1:]))
        self.relu = nn.ReLU()

    def forward(self, x):
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1])
        n = len(self.fc_layers)
        for i, linear in enumerate(self.fc_layers):
            x = ops.Linear.apply(x, linear.weight, linear.bias, i < n - 1)
        return x.reshape(*lead, x.shape[-1])


class ConvHead(nn.Module):
    """1x1-conv/ReLU projection head. Reference: UNet.py:211-243."""

    def __init__(self, channel_layer=[128, 256, 32], use_3D=False):
        super(ConvHead, self).__init__()
        conv = nn.Conv3d if use_3D else nn.Conv2d
        self.conv_layers = nn.ModuleList(conv(in_channels=n_in, out_channels=n_out, kernel_size=1)
                                         for n_in, n_out in zip(channel_layer[:-1], channel_layer[1:]))
        self.relu = nn.ReLU()

    def forward_cl(self, x):
        n = len(self.conv_layers)
        for i, conv in enumerate(self.conv_layers):
            x = ops.ConvBias.apply(x, conv.weight, conv.bias, i < n - 1)
        return x

    def forward(self, x):
        was_4d = x.dim() == 4
        return ops.from_channels_last(self.forward_cl(ops.to_channels_last(x)), was_4d)


class _UNetBase(nn.Module):
    """Shared encoder / decoder plumbing of UNet, UNet_Encoder and Partial_UNet."""

    def _build_encoder(self, depth, use_3D, in_channels, top_filter, midchannels_factor, p_dropout_list):
        down, bottleneck = _filters(in_channels, top_filter, depth)
        self.down_block = nn.ModuleList()
        for ch, p in zip(down, p_dropout_list[:-1]):
            self.down_block.append(ConvBlock(ch[0], ch[1], mid_channels=ch[1] // midchannels_factor, use_3D=use_3D, p_dropout=p))
        return down, bottleneck

    def _build_bottleneck(self, bottleneck, use_3D, midchannels_factor, p):
        self.bottleneck_block = ConvBlock(bottleneck[0], bottleneck[1], mid_channels=bottleneck[1] // midchannels_factor, use_3D=use_3D,
                                          p_dropout=p)
        self.downpool = nn.MaxPool3d(kernel_size=2, stride=2) if use_3D else nn.MaxPool2d(kernel_size=2, stride=2)

    def _build_decoder(self, up_filters, use_3D, bilinear):
        self.up_samp = nn.ModuleList()
        self.up_block = nn.ModuleList()
        for ch in up_filters:
            if bilinear:
                self.up_block.append(ConvBlock(int(1.5 * ch[0]), ch[1], mid_channels=ch[1], use_3D=use_3D))
                self.up_samp.append(nn.Upsample(scale_factor=2, mode='trilinear' if use_3D else 'bilinear', align_corners=True))
            else:
                self.up_block.append(ConvBlock(ch[0], ch[1], mid_channels=ch[1], use_3D=use_3D))
                convT = nn.ConvTranspose3d if use_3D else nn.ConvTranspose2d
                self.up_samp.append(convT(ch[0], ch[1], kernel_size=2, stride=2))

    @property
    def _fd(self):
        return 2 if isinstance(self.downpool, nn.MaxPool3d) else 1

    def _check_grid(self, x, n_pool):
        f = 2 ** n_pool
        d, h, w = x.shape[1:4]
        if (self._fd == 2 and d % f) or h % f or w % f:
            raise RuntimeError(f'ich_b200: spatial size {(d, h, w) if self._fd == 2 else (h, w)} must be divisible by {f} '
                               f'(the reference has no pad/crop logic either, UNet.py:117-119)')

    def _encode(self, x):
        if len(getattr(self, 'up_samp', [])):           # a decoder concatenates with the skip tensors: sizes must halve exactly
            self._check_grid(x, len(self.down_block))   # (encoder-only nets pool with floor, as nn.MaxPool does, UNet.py:82,313)
        res = []
        n_down = len(self.down_block)
        ups = list(getattr(self, 'up_samp', []))
        for i, block in enumerate(self.down_block):         # UNet.py:106-109
            # skip tensor i is consumed by decoder stage (n_down - 1 - i): the pool kernel lays it out inside that stage's concat
            # buffer on the way (no copy pass for torch.cat([res, up], 1), UNet.py:119)
            j = n_down - 1 - i
            concat_c = ups[j].out_channels if (_cfg.get('zero_copy_concat') and j < len(ups)
                                               and isinstance(ups[j], (nn.ConvTranspose3d, nn.ConvTranspose2d))) else 0
            x = block.forward_cl(x)
            out = ops.PoolSkip.apply(x, self._fd, concat_c)  # skip tensor + pooled tensor; their gradients meet in one kernel
            skip, x = out[0], out[1]
            if concat_c:
                skip._ich_concat_buf = out[2]
            res.append(skip)
        return self.bottleneck_block.forward_cl(x), res     # UNet.py:112

    def _decode(self, x, res, head=None):
        """head = (final_conv, act): the last block runs fused with the single-class head and the network output is returned."""
        last = len(self.up_block) - 1
        for i, (up, block, r) in enumerate(zip(self.up_samp, self.up_block, res[::-1])):     # UNet.py:117-119
            if isinstance(up, nn.Upsample):                                     # bilinear=True: UNet.py:69-72
                x = ops.UpsampleCat.apply(x, r, self._fd)
            else:
                x = ops.UpConvCat.apply(x, r, up.weight, up.bias, self._fd, getattr(r, '_ich_concat_buf', None))
            x = block.forward_cl_head(x, *head) if (head is not None and i == last) else block.forward_cl(x)
        return x


class UNet(_UNetBase):
    """2-D / 3-D U-Net. Reference: UNet.py:18-127."""

    def __init__(self, depth=5, use_3D=False, bilinear=False, in_channels=1, out_channels=1, top_filter=64, midchannels_factor=2,
                 p_dropout=0.5, use_final_activation=True):
        super(UNet, self).__init__()
        p_dropout_list = _dropout_list(p_dropout, depth)
        self.return_bottleneck = False
        # module lists registered as down_block, up_samp, up_block (state-dict key order) and FILLED in the reference's interleaved order
        # down_block[i], up_block[i], up_samp[i] (UNet.py:66-76): the parameter initialisation draws from the global RNG in construction
        # order, so the same torch.manual_seed gives the same initial weights as the reference
        down, bottleneck = _filters(in_channels, top_filter, depth)
        up_filters = [(top_filter * 2 ** d, top_filter * 2 ** (d - 1)) for d in range(depth - 1, 0, -1)]
        self.down_block = nn.ModuleList()
        self.up_samp = nn.ModuleList()
        self.up_block = nn.ModuleList()
        for down_ch, up_ch, p in zip(down, up_filters, p_dropout_list[:-1]):
            self.down_block.append(ConvBlock(down_ch[0], down_ch[1], mid_channels=down_ch[1] // midchannels_factor, use_3D=use_3D, p_dropout=p))
            if bilinear:
                self.up_block.append(ConvBlock(int(1.5 * up_ch[0]), up_ch[1], mid_channels=up_ch[1], use_3D=use_3D))
                self.up_samp.append(nn.Upsample(scale_factor=2, mode='trilinear' if use_3D else 'bilinear', align_corners=True))
            else:
                self.up_block.append(ConvBlock(up_ch[0], up_ch[1], mid_channels=up_ch[1], use_3D=use_3D))
                convT = nn.ConvTranspose3d if use_3D else nn.ConvTranspose2d
                self.up_samp.append(convT(up_ch[0], up_ch[1], kernel_size=2, stride=2))
        self._build_bottleneck(bottleneck, use_3D, midchannels_factor, p_dropout_list[-1])
        self.final_conv = nn.Conv3d(top_filter, out_channels, kernel_size=1) if use_3D else nn.Conv2d(top_filter, out_channels, kernel_size=1)
        if use_final_activation:
            self.final_activation = nn.Softmax(dim=1) if out_channels > 1 else nn.Sigmoid()
        else:
            self.final_activation = nn.Identity()

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        x = ops.to_channels_last(input)
        x, res = self._encode(x)
        x_bottleneck = x
        act = 0 if isinstance(self.final_activation, nn.Identity) else (2 if isinstance(self.final_activation, nn.Softmax) else 1)
        tail = self.up_block[-1] if len(self.up_block) else None
        if tail is not None and not (tail.training and tail.dropout.p > 0.0) and tail.conv2.kernel_size[0] == 3 and \
                ops.head_fusable(x.dtype, tail.conv2.out_channels, self.final_conv.weight, act, tail.training):
            out = self._decode(x, res, head=(self.final_conv, act))                          # UNet.py:119 + :122, fused
        else:
            x = self._decode(x, res)
            fc = self.final_conv
            if fc.in_channels > 64 and fc.out_channels <= 8:
                # the head kernel keeps its weights in shared memory (<= 64 input channels): wider nets (top_filter = 128) run the 1x1 conv
                # on the conv kernels and use the head kernel for the activation + fp32 NC(D)HW layout only (identity weights)
                logits = ops.ConvBias.apply(x, fc.weight, fc.bias, False)
                c = fc.out_channels
                eye = torch.eye(c, dtype=torch.float32, device=logits.device).view(c, c, *([1] * (fc.weight.dim() - 2)))
                out = ops.Head.apply(logits, eye, None, act)
            else:
                out = ops.Head.apply(x, fc.weight, fc.bias, act)                                 # UNet.py:122
        _flush_nbt()
        if was_4d:
            out = out.squeeze(2)
        if self.return_bottleneck:
            return out, ops.from_channels_last(x_bottleneck, was_4d)
        return out


class UNet_Encoder(_UNetBase):
    """U-Net encoder + global average pool + MLP head. Reference: UNet.py:245-326."""

    def __init__(self, depth=5, use_3D=False, in_channels=1, MLP_head=[256, 128], top_filter=64, midchannels_factor=2, p_dropout=0.5):
        super(UNet_Encoder, self).__init__()
        p_dropout_list = _dropout_list(p_dropout, depth)
        self.return_bottleneck = False
        down, bottleneck = self._build_encoder(depth, use_3D, in_channels, top_filter, midchannels_factor, p_dropout_list)
        self._build_bottleneck(bottleneck, use_3D, midchannels_factor, p_dropout_list[-1])
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1)) if use_3D else nn.AdaptiveAvgPool2d((1, 1))
        self.mlp_head = MLPHead(Neurons_layer=[bottleneck[1]] + MLP_head)

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        x, _ = self._encode(ops.to_channels_last(input))
        _flush_nbt()
        pooled = ops.GlobalAvgPool.apply(x)                                     # UNet.py:318, fp32 [N, C]
        out = self.mlp_head(pooled)                                             # UNet.py:321
        if self.return_bottleneck:
            return out, pooled.view(*pooled.shape, *([1, 1] if was_4d else [1, 1, 1]))
        return out


class Partial_UNet(_UNetBase):
    """Full encoder, first n_decoder up-stages, ConvHead, no activation. Reference: UNet.py:328-435."""

    def __init__(self, depth=5, n_decoder=3, use_3D=False, bilinear=False, in_channels=1, head_channel=[64, 32], top_filter=64,
                 midchannels_factor=2, p_dropout=0.5):
        super(Partial_UNet, self).__init__()
        p_dropout_list = _dropout_list(p_dropout, depth)
        self.return_bottleneck = False
        down, bottleneck = self._build_encoder(depth, use_3D, in_channels, top_filter, midchannels_factor, p_dropout_list)
        up_filters = [(top_filter * 2 ** d, top_filter * 2 ** (d - 1)) for d in range(depth - 1, depth - 1 - n_decoder, -1)]
        self.n_decoder = n_decoder
        self._build_decoder(up_filters, use_3D, bilinear)
        self._build_bottleneck(bottleneck, use_3D, midchannels_factor, p_dropout_list[-1])
        self.final_conv = ConvHead(channel_layer=[up_filters[-1][1]] + head_channel, use_3D=use_3D)

    def forward(self, input):
        _begin_forward()
        was_4d = _was_4d(input)
        x, res = self._encode(ops.to_channels_last(input))
        x_bottleneck = x
        x = self._decode(x, res[::-1][:self.n_decoder][::-1])                  # UNet.py:425-427
        out = ops.from_channels_last(self.final_conv.forward_cl(x), was_4d)     # UNet.py:430
        _flush_nbt()
        if self.return_bottleneck:
            return out, ops.from_channels_last(x_bottleneck, was_4d)
        return out
