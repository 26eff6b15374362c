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
    weights, bias + ReLU epilogue; the default since round 2); training mode, eval with autograd and ICH_B200_FOLD_EVAL_BN=0 keep the BatchNorm kernels.
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
    assert dry.trace.count('ich_bn_finalize') == 1                  # default: on (round 2, after the full-size cfg-5 parity run)
    with config.override(fold_eval_bn=False):
        dry.trace.clear()
        with torch.no_grad():
            net(x)
        assert dry.trace.count('ich_bn_finalize') == 10             # ICH_B200_FOLD_EVAL_BN=0: the BatchNorm kernels


def test_weight_packs_refreshed_from_the_optimizer_hook(dry, monkeypatch):
    """The packs the optimizer step invalidated are re-derived in ONE batched launch from the global optimizer post-step hook; the
    refresh at the start of the next forward pass then has nothing to do.  ICH_B200_REFRESH_AFTER_STEP=0 restores refresh-at-forward."""
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss
    monkeypatch.setattr(ops, '_REFRESH_ANY_DEVICE', True)
    x = torch.rand(1, 1, 8, 16, 16)
    # (32 base filters: a 16-filter net with midchannels_factor = 2 pads its 8-channel mid tensor with per-step derived weights, see
    # ConvBlock._forward_cl_padded_mid -- those are packed on use, outside the batched refresh)
    net = UNet(depth=3, use_3D=True, top_filter=32, midchannels_factor=2, p_dropout=0.0).train()
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


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_gated_unet_plumbing(dry, prec, golden):
    """Drop-in GatedUNet (reference models/networks/GatedUNet.py): state-dict keys of the golden runs load, and forward + backward flow
    through ConvBnRelu / ConvBias / GateMul / UpsampleCat / Head for the gated and the plain variant."""
    from src.models.networks.GatedUNet import UNet, GatedConv, ConvLayer
    from src.models.optim.LossFunctions import BinaryDiceLoss
    fxs = golden('gated_unet.pt')
    with config.override(precision=prec):
        for name, fx in fxs.items():
            net = UNet(**fx['kwargs'])
            assert list(net.state_dict().keys()) == list(fx['state_dict'].keys()), name
            net.load_state_dict(fx['state_dict'])
            net.train()
            dry.trace.clear()
            out = net(fx['x'])
            assert out.shape == fx['out_train'].shape and out.dtype == torch.float32
            BinaryDiceLoss(**fx['loss_kwargs'])(out, fx['mask']).backward()
            _grads_ok(net)
            gated = fx['kwargs']['use_gatedConv']
            assert ('ich_gate_mul_fwd' in dry.trace) == gated and ('ich_gate_mul_bwd' in dry.trace) == gated
            assert ('ich_upsample2_fwd' in dry.trace) == gated
            assert isinstance(net.down_block[0].conv1, GatedConv if gated else ConvLayer)
    with pytest.raises(NotImplementedError):
        UNet(depth=3, top_filter=8, p_dropout=0.5, use_gatedConv=True).train()(torch.rand(1, 1, 16, 16))
    with pytest.raises(NotImplementedError):
        ConvLayer(4, 4, 3, padding=1, activation='lrelu')(torch.rand(1, 4, 8, 8))


def test_mlp_head_runs_on_the_engine(dry):
    from src.models.networks.UNet import MLPHead, UNet_Encoder
    head = MLPHead([32, 16, 8])
    out = head(torch.rand(4, 32))
    assert out.shape == (4, 8)
    out.sum().backward()
    _grads_ok(head)
    assert dry.trace.count('ich_linear_fwd') == 2 and dry.trace.count('ich_linear_bwd') == 2
    dry.trace.clear()
    enc = UNet_Encoder(depth=3, use_3D=True, top_filter=16, MLP_head=[32, 8], p_dropout=0.0).train()
    enc(torch.rand(2, 1, 8, 16, 16)).sum().backward()
    _grads_ok(enc)
    assert 'ich_linear_fwd' in dry.trace and 'ich_avgpool_fwd' in dry.trace


def test_staging_and_volume_inference_plumbing(dry):
    """ops.stage_ct / staged inputs / window gather-scatter / segement_volume: argument plumbing against the header prototypes."""
    from ich_b200 import infer
    from src.models.networks.UNet import UNet
    with config.override(precision='bf16'):
        raw = torch.randint(-1000, 2000, (6, 16, 16), dtype=torch.int16)
        x = ops.stage_ct(raw, 40, 120, (0, 1))
        assert x.dtype == torch.bfloat16 and x.shape == raw.shape and dry.trace[-1] == 'ich_stage_ct'
        with pytest.raises(RuntimeError):
            ops.staged(torch.zeros(1, 1, 4, 4))                      # not an engine-layout tensor
        net3 = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).eval()
        vol = torch.rand(1, 1, 8, 32, 16)
        for stride, ret in (((8, 16, 16), True), ((8, 16, 16), False), ((4, 8, 8), True)):
            dry.trace.clear()
            pred, mask = infer.sliding_window_predict(net3, vol, (8, 16, 16), stride, batch=3, distributed=False, return_pred=ret)
            assert mask.shape == vol.shape and mask.dtype == torch.bool and (pred is None) == (not ret)
            assert 'ich_window_gather' in dry.trace and 'ich_window_scatter' in dry.trace
            assert ('ich_blend_threshold' in dry.trace) == (stride != (8, 16, 16))
            assert 'ich_layout_nc_to_nl' not in dry.trace          # windows enter the network already in the engine layout
        # segement_volume: 2-D net on every slice, 3-D net on windows; [H, W, S] int16 Hounsfield units in, uint8 0 / 255 out
        hu = np.random.RandomState(0).randint(-1000, 2000, size=(16, 32, 5)).astype(np.int16)
        net2 = UNet(depth=3, use_3D=False, top_filter=16, midchannels_factor=1, p_dropout=0.0).eval()
        dry.trace.clear()
        seg = infer.segement_volume(net2, hu, window=(40, 120), input_size=(32, 16), return_pred=True, batch_size=2, device='cpu')
        assert seg.shape == hu.shape and seg.dtype == np.uint8 and dry.trace.count('ich_stage_ct') == 1
        assert dry.trace.count('ich_bn_head_fwd') == 3              # 5 slices in batches of 2
        seg3 = infer.segement_volume(net3, np.zeros((16, 32, 8), np.int16), window=(40, 120), return_pred=True, device='cpu', window_3d=(8, 16, 16))
        assert seg3.shape == (16, 32, 8)
        assert infer.segement_volume(net2, hu, window=None, input_size=None, device='cpu') is None


def test_graphed_step_needs_fresh_optimizer():
    from ich_b200.graph import GraphedStep
    lin = torch.nn.Linear(4, 4)
    opt = torch.optim.Adam(lin.parameters())
    step = GraphedStep(lambda x: x, opt)
    assert all(g['capturable'] for g in opt.param_groups) and step.graph is None
    lin(torch.rand(2, 4)).sum().backward()
    opt2 = torch.optim.Adam(lin.parameters())
    opt2.step()
    with pytest.raises(RuntimeError):
        GraphedStep(lambda x: x, opt2)


def test_graphed_step_refuses_other_shapes():
    """A replay only fits the shapes it was captured for; `copy_` into the static inputs would broadcast a smaller batch silently."""
    from ich_b200.graph import GraphedStep

    class FakeGraph:
        replays = 0

        def replay(self):
            FakeGraph.replays += 1
    step = GraphedStep(lambda x: x)
    step.graph, step.static_in, step.static_out = FakeGraph(), [torch.zeros(4, 3)], 'out'
    assert step(torch.ones(4, 3)) == 'out' and FakeGraph.replays == 1 and float(step.static_in[0].sum()) == 12.0
    with pytest.raises(RuntimeError, match='captured for inputs'):
        step(torch.ones(1, 3))              # broadcastable: must not be copied in
    with pytest.raises(RuntimeError, match='captured for inputs'):
        step(torch.ones(4, 3, dtype=torch.float64))
    with pytest.raises(RuntimeError, match='captured with 1 inputs'):
        step(torch.ones(4, 3), torch.ones(4, 3))
    assert FakeGraph.replays == 1


def test_padded_mid_channels_plumbing(dry):
    """top_filter = 16 nets with midchannels_factor = 2 (BASELINE.json configs[0]): the 8-channel mid tensor of the first block is carried
    zero-padded to 16 channels (tcgen05 granularity); every parameter still receives a gradient of its own shape, running statistics
    keep their 8 entries, eval + folding takes the padded folded weights."""
    from src.models.networks.UNet import UNet
    with config.override(precision='bf16'):
        net = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
        blk = net.down_block[0]
        x = torch.rand(1, 1, 8, 16, 16)
        assert blk._pad_mid(torch.empty(1, 8, 16, 16, 1, dtype=torch.bfloat16))
        net(x).sum().backward()
        _grads_ok(net)
        assert blk.bn1.running_mean.shape == (8,) and blk.bn1.num_batches_tracked == 1 and blk.bn2.num_batches_tracked == 1
        assert blk.conv1.weight.grad.shape == (8, 1, 3, 3, 3) and blk.conv2.weight.grad.shape == (16, 8, 3, 3, 3)
        net.eval()
        dry.trace.clear()
        with torch.no_grad():
            assert net(x).shape == x.shape
        assert dry.trace.count('ich_bn_finalize') == 1              # folded everywhere but the fused last unit
    with config.override(precision='fp32'):
        assert not blk._pad_mid(torch.empty(1, 8, 16, 16, 1, dtype=torch.float32))     # fp32 verification mode: untouched


def test_wide_head_plumbing(dry):
    """top_filter = 128 with a multi-class head: the 1x1 conv runs on the conv kernels, the head kernel does activation + layout."""
    from src.models.networks.UNet import UNet
    net = UNet(depth=2, use_3D=False, top_filter=128, out_channels=3, p_dropout=0.0).train()
    out = net(torch.rand(1, 1, 8, 16))
    assert out.shape == (1, 3, 8, 16)
    out.sum().backward()
    _grads_ok(net)
    assert dry.trace.count('ich_head_fwd') == 1 and 'ich_head_dlogit' in dry.trace


def test_encoder_floor_pooling_plumbing(dry):
    from src.models.networks.UNet import UNet_Encoder, UNet
    enc = UNet_Encoder(depth=3, use_3D=True, top_filter=16, midchannels_factor=1, MLP_head=[32, 8], p_dropout=0.0).train()
    out = enc(torch.rand(1, 1, 10, 22, 18))
    assert out.shape == (1, 8)
    out.sum().backward()
    _grads_ok(enc)
    with pytest.raises(RuntimeError):                     # a decoder still needs sizes that halve exactly
        UNet(depth=3, use_3D=True, top_filter=16, p_dropout=0.0)(torch.rand(1, 1, 10, 22, 18))
