"""CPU: dry run of the host-side plumbing (ich_b200.ops + the drop-in modules) with the kernel launches replaced by a recorder.

No arithmetic is checked here (that is what the `-m gpu` parity tests do through the real library): every `call(name, *args)` the
autograd Functions make is validated against the prototype include/ich_b200.h declares for `name` (argument count, pointer vs
integer vs floating-point kind), the host-only query entry points (`ich_*_supported`, `ich_conv_tc_variant`) run for real, and
the whole forward + loss + backward graph of each drop-in network must flow (shapes, save-for-backward, gradient plumbing) in
both engine precisions.  Catches host-side slips -- a missing argument, a pointer in an integer slot, a wrong gradient arity --
without a GPU."""
import ctypes

import numpy as np
import pytest
import torch

from ich_b200 import _lib, config, ops

PROTOS = _lib.parse_header()
_INT = (ctypes.c_int, ctypes.c_longlong, ctypes.c_uint)
_FLT = (ctypes.c_float, ctypes.c_double)


class Recorder:
    def __init__(self):
        self.trace = []

    def __call__(self, name, *args):
        assert name in PROTOS, f'{name} is not declared in include/ich_b200.h'
        argtypes = PROTOS[name][1]
        assert len(args) == len(argtypes), f'{name}: {len(args)} arguments, the header declares {len(argtypes)}'
        for i, (a, t) in enumerate(zip(args, argtypes)):
            if t is ctypes.c_void_p:
                ok = a is None or (isinstance(a, int) and not isinstance(a, bool)) or isinstance(a, ctypes.c_void_p)
            elif t in _INT:
                ok = isinstance(a, (int, np.integer)) and not isinstance(a, float)
            else:
                assert t in _FLT
                ok = isinstance(a, (int, float)) and not isinstance(a, bool)
            assert ok, f'{name}: argument {i} = {a!r} does not fit {t.__name__}'
        self.trace.append(name)


@pytest.fixture
def dry(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(ops, 'call', rec)
    monkeypatch.setattr(ops, '_stream', lambda: 0)
    monkeypatch.setattr(ops, '_require_cuda', lambda t, what: None)
    return rec


def _grads_ok(net):
    for name, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_unet3d_training_step_plumbing(dry, prec):
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss
    with config.override(precision=prec):
        net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
        x = torch.rand(2, 1, 8, 16, 16).requires_grad_(True)
        m = (torch.rand(2, 1, 8, 16, 16) > 0.9).float().requires_grad_(True)
        out = net(x)
        assert out.shape == x.shape and out.dtype == torch.float32 and out.is_contiguous()
        ComboLoss(alpha=0.5, beta=0.5, reduction='mean', p=1)(out, m).backward()
    _grads_ok(net)
    t = dry.trace
    for name in ('ich_layout_nc_to_nl', 'ich_bn_finalize', 'ich_maxpool2_bwd', 'ich_seg_loss_fwd', 'ich_seg_loss_bwd'):
        assert name in t, name
    assert t.count('ich_bn_finalize') == 10        # 5 ConvBlocks x 2 units
    assert net.down_block[0].bn1.num_batches_tracked == 1 and net.up_block[-1].bn2.num_batches_tracked == 1


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_fused_head_plumbing(dry, prec):
    """ICH_B200_FUSE_HEAD: the last ConvBlock unit and the single-class head run as one op in train and eval mode; multi-class
    heads, dropout on the last block and SyncBN keep the unfused path."""
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss
    x = torch.rand(2, 1, 8, 16, 16)
    with config.override(precision=prec, fuse_head=True):
        net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
        ComboLoss()(net(x), torch.zeros_like(x)).backward()
        _grads_ok(net)
        t = dry.trace
        assert t.count('ich_bn_head_fwd') == 1 and t.count('ich_bn_head_bwd') == 1
        assert 'ich_head_fwd' not in t and 'ich_head1_bwd' not in t
        assert t.count('ich_affine_act') == 9 and t.count('ich_bn_act_bwd') == 9 and t.count('ich_bn_finalize') == 10
        assert net.up_block[-1].bn2.num_batches_tracked == 1
        net.eval()
        dry.trace.clear()
        with torch.no_grad():
            assert net(x).shape == x.shape
        assert dry.trace.count('ich_bn_head_fwd') == 1
        dry.trace.clear()
        net3 = UNet(depth=3, use_3D=True, top_filter=16, out_channels=3, p_dropout=0.0).train()
        net3(x).sum().backward()
        assert 'ich_bn_head_fwd' not in dry.trace and 'ich_head_fwd' in dry.trace
        dry.trace.clear()
        with config.override(sync_bn=True):
            old = ops.SYNC_BN_COMM
            ops.SYNC_BN_COMM = (2, lambda t: None)
            try:
                net.train()
                ComboLoss()(net(x), torch.zeros_like(x)).backward()
            finally:
                ops.SYNC_BN_COMM = old
        assert 'ich_bn_head_fwd' not in dry.trace and 'ich_bn_act_bwd_sync' in dry.trace


@pytest.mark.parametrize('kw', [dict(bilinear=True), dict(out_channels=3), dict(use_final_activation=False), dict(p_dropout=0.3)])
def test_unet2d_variants_plumbing(dry, kw):
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import BinaryDiceLoss, TverskyLoss
    args = dict(depth=3, use_3D=False, top_filter=16, midchannels_factor=1, p_dropout=0.0)
    args.update(kw)
    net = UNet(**args).train()
    net.return_bottleneck = True
    x = torch.rand(2, 1, 16, 32)
    out, bott = net(x)
    assert out.shape == (2, args.get('out_channels', 1), 16, 32) and bott.shape == (2, 64, 4, 8)
    loss = BinaryDiceLoss(reduction='mean', p=2, alpha=0.2)(out, torch.rand_like(out).round()) + \
        TverskyLoss(alpha=0.2, beta=0.7, gamma=0.3)(out, torch.rand_like(out).round()) + bott.sum()
    loss.backward()
    _grads_ok(net)
    net.eval()
    with torch.no_grad():
        assert net(x)[0].shape == out.shape


def test_folded_eval_bn_plumbing(dry):
    """ICH_B200_FOLD_EVAL_BN: under eval + no_grad every unit but the fused last one is ONE conv launch (BatchNorm folded into the
    weights, bias + ReLU epilogue); training mode, eval with autograd and the default configuration keep the BatchNorm kernels.
    The folded weights are a cache keyed on the versions of the tensors they derive from."""
    from src.models.networks.UNet import UNet
    x = torch.rand(1, 1, 8, 16, 16)
    net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).eval()
    with config.override(fold_eval_bn=True):
        with torch.no_grad():
            assert net(x).shape == x.shape
        t = dry.trace
        assert t.count('ich_bn_finalize') == 1 and t.count('ich_bn_head_fwd') == 1 and 'ich_affine_act' not in t
        conv = net.down_block[1].conv1
        w1, b1 = ops.folded_eval_unit(conv.weight, conv.bias, net.down_block[1].bn1.weight, net.down_block[1].bn1.bias,
                                      net.down_block[1].bn1.running_mean, net.down_block[1].bn1.running_var)
        bn = net.down_block[1].bn1
        want = conv.weight.detach() * (bn.weight / torch.sqrt(bn.running_var + 1e-5)).view(-1, 1, 1, 1, 1).detach()
        assert torch.allclose(w1, want, rtol=1e-6, atol=1e-8)
        assert torch.allclose(b1, (bn.bias + (conv.bias - bn.running_mean) * bn.weight / torch.sqrt(bn.running_var + 1e-5)).detach(), atol=1e-7)
        w2, _ = ops.folded_eval_unit(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        assert w2 is w1                                             # cached
        with torch.no_grad():
            bn.running_var.mul_(2.0)                                # e.g. load_state_dict copying in place
        w3, _ = ops.folded_eval_unit(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        assert w3 is not w1 and not torch.allclose(w3, w1)
        dry.trace.clear()
        net(x).sum().backward()                                     # eval WITH autograd: unfolded path
        assert dry.trace.count('ich_bn_finalize') == 10
        net.train()
        dry.trace.clear()
        with torch.no_grad():
            net(x)
        assert dry.trace.count('ich_bn_finalize') == 10
    net.eval()
    dry.trace.clear()
    with torch.no_grad():
        net(x)
    assert dry.trace.count('ich_bn_finalize') == 10                 # default: off


def test_weight_packs_refreshed_from_the_optimizer_hook(dry, monkeypatch):
    """The packs the optimizer step invalidated are re-derived in ONE batched launch from the global optimizer post-step hook; the
    refresh at the start of the next forward pass then has nothing to do.  ICH_B200_REFRESH_AFTER_STEP=0 restores refresh-at-forward."""
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss
    monkeypatch.setattr(ops, '_REFRESH_ANY_DEVICE', True)
    x = torch.rand(1, 1, 8, 16, 16)
    net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    ComboLoss()(net(x), torch.zeros_like(x)).backward()
    n_single = dry.trace.count('ich_permute5')                  # first use: one derivation per pack
    assert n_single > 10 and 'ich_permute5_batch' not in dry.trace
    dry.trace.clear()
    opt.step()
    assert dry.trace == ['ich_permute5_batch']                  # from the hook
    assert ops.refresh_packs() == 0                             # nothing left for the forward pass
    dry.trace.clear()
    ComboLoss()(net(x), torch.zeros_like(x)).backward()
    assert 'ich_permute5' not in dry.trace and 'ich_permute5_batch' not in dry.trace
    with config.override(refresh_after_step=False):
        dry.trace.clear()
        opt.step()
        assert dry.trace == []
        net(x)
        assert dry.trace[0] == 'ich_layout_nc_to_nl' or dry.trace[0] == 'ich_permute5_batch'
        assert dry.trace.count('ich_permute5_batch') == 1 and 'ich_permute5' not in dry.trace
    # an unrelated optimizer in the same process does not disturb anything
    lin = torch.nn.Linear(2, 2)
    o2 = torch.optim.SGD(lin.parameters(), lr=0.1)
    lin(torch.rand(1, 2)).sum().backward()
    dry.trace.clear()
    o2.step()
    assert dry.trace == []


def test_contrastive_nets_plumbing(dry):
    from src.models.networks.UNet import UNet_Encoder, Partial_UNet
    from src.models.optim.LossFunctions import InfoNCELoss, LocalInfoNCELoss
    enc = UNet_Encoder(depth=3, use_3D=True, top_filter=16, MLP_head=[32, 8], p_dropout=0.0).train()
    x1, x2 = torch.rand(2, 1, 8, 16, 16), torch.rand(2, 1, 8, 16, 16)
    z1, z2 = enc(x1), enc(x2)
    assert z1.shape == (2, 8)
    InfoNCELoss(set_size=2, tau=0.1, device='cpu')(torch.nn.functional.normalize(z1, dim=1), torch.nn.functional.normalize(z2, dim=1)).backward()
    _grads_ok(enc)
    pu = Partial_UNet(depth=4, n_decoder=2, use_3D=False, top_filter=16, midchannels_factor=1, head_channel=[32, 8], p_dropout=0.0).train()
    f1, f2 = pu(torch.rand(2, 1, 32, 32)), pu(torch.rand(2, 1, 32, 32))
    assert f1.shape == (2, 8, 16, 16)
    np.random.seed(0)
    LocalInfoNCELoss(tau=0.1, K=2, n_region=3, device='cpu')(f1, f2).backward()
    _grads_ok(pu)
    for name in ('ich_avgpool_fwd', 'ich_infonce_fwd', 'ich_infonce_bwd', 'ich_region_gather', 'ich_region_scatter'):
        assert name in dry.trace, name


def test_frozen_parameters_skip_their_weight_gradients(dry):
    """Frozen params => no wgrad launch for them (transfer_weights + rgetattr(net, key).requires_grad = False)."""
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss
    net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
    x = torch.rand(1, 1, 8, 16, 16)
    ComboLoss()(net(x), torch.zeros_like(x)).backward()
    wgrads = [n for n in dry.trace if 'wgrad' in n]
    for p in net.down_block.parameters():
        p.requires_grad = False
    net.zero_grad()
    dry.trace.clear()
    ComboLoss()(net(x), torch.zeros_like(x)).backward()
    assert len([n for n in dry.trace if 'wgrad' in n]) < len(wgrads)
    assert all(p.grad is None for p in net.down_block.parameters())
    assert all(p.grad is not None for p in net.up_block.parameters())
