"""CPU: the C-ABI library builds / loads and exports every symbol include/ich_b200.h declares; the drop-in modules keep
the reference's API surface (constructor validation, attribute tree, state-dict keys) and refuse to run without CUDA."""
import ctypes
import os
import sys

import pytest
import torch

from ich_b200 import _lib

PKG = _lib.PKG_ROOT


def test_library_exports_every_declared_symbol():
    path = _lib.build()
    assert os.path.exists(path)
    protos = _lib.parse_header()
    assert len(protos) >= 30
    l = ctypes.CDLL(path)
    for name in protos:
        assert hasattr(l, name), f'{name} declared in include/ich_b200.h but not exported'
    assert _lib.lib().ich_abi_version() == 1
    assert _lib.lib().ich_last_error() is not None


def test_state_dict_keys_match_reference(golden):
    from src.models.networks.UNet import UNet, UNet_Encoder, Partial_UNet
    for name, cls in [('unet3d_combo.pt', UNet), ('unet2d_dice.pt', UNet), ('unet3d_softmax.pt', UNet),
                      ('encoder_infonce.pt', UNet_Encoder), ('partial_local_infonce.pt', Partial_UNet)]:
        fx = golden(name)
        net = cls(**fx['kwargs'])
        assert list(net.state_dict().keys()) == list(fx['state_dict'].keys()), name
        for k, v in net.state_dict().items():
            assert v.shape == fx['state_dict'][k].shape and v.dtype == fx['state_dict'][k].dtype, (name, k)
        net.load_state_dict(fx['state_dict'])
    net = UNet(depth=4, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0)
    assert len(net.state_dict()) == 106 and sum(p.numel() for p in net.parameters()) == 962481      # SURVEY 8c anchors
    net = UNet(depth=4, use_3D=True, top_filter=32, midchannels_factor=2, p_dropout=0.0)
    assert sum(p.numel() for p in net.parameters()) == 3845729


def test_constructor_validation():
    from src.models.networks.UNet import UNet, ConvBlock, UNet_Encoder, Partial_UNet
    from src.models.optim import LossFunctions as LF
    with pytest.raises(TypeError):
        UNet(p_dropout='0.5')
    with pytest.raises(AssertionError):
        UNet(depth=4, p_dropout=[0.1, 0.2])
    with pytest.raises(AssertionError):
        ConvBlock(1, 8, p_dropout=1.5)
    with pytest.raises(AssertionError):
        LF.BinaryDiceLoss(reduction='avg')
    with pytest.raises(AssertionError):
        LF.ComboLoss(alpha=1.5)
    with pytest.raises(AssertionError):
        LF.InfoNCELoss()
    for n in ['BinaryDiceLoss', 'TverskyLoss', 'ComboLoss', 'InfoNCELoss', 'LocalInfoNCELoss', 'DiscountedL1', 'GDL', 'HSCLoss']:
        assert hasattr(LF, n)
    net = UNet(depth=3, top_filter=8, p_dropout=[0.0, 0.1, 0.2])
    assert net.return_bottleneck is False and net.down_block[1].dropout.p == 0.1 and net.bottleneck_block.dropout.p == 0.2
    assert isinstance(UNet(out_channels=3, depth=2, top_filter=4).final_activation, torch.nn.Softmax)
    assert isinstance(UNet(use_final_activation=False, depth=2, top_filter=4).final_activation, torch.nn.Identity)
    enc = UNet_Encoder(depth=3, top_filter=8, MLP_head=[16, 4])
    assert enc.mlp_head.fc_layers[0].in_features == 32
    pu = Partial_UNet(depth=4, n_decoder=2, top_filter=8, head_channel=[16, 4])
    assert len(pu.up_samp) == 2 and pu.final_conv.conv_layers[0].in_channels == 16


def test_losses_masks_and_sampling_match_reference_semantics():
    import numpy as np
    from src.models.optim import LossFunctions as LF
    from oracle import losses_oracle as LO
    l = LF.InfoNCELoss(set_size=3, tau=0.1, device='cpu')
    eye = torch.diag(torch.ones(6)) + torch.diag(torch.ones(3), 3) + torch.diag(torch.ones(3), -3)
    assert torch.equal(l.neg_mask, ~eye.bool())
    loc = LF.LocalInfoNCELoss(tau=0.1, K=3, n_region=4, device='cpu')
    np.random.seed(5)
    a = loc.sample_region_corners((2, 12, 15, 8))
    np.random.seed(5)
    b = LO.sample_regions((2, 12, 15, 8), 3, 4)
    assert (a == b).all() and a.shape == (2, 4, 2)
    np.random.seed(5)
    m = loc.get_sample_region_mask((2, 12, 15, 8))
    assert m.shape == (2, 12, 15) and sorted(m.unique().tolist()) == [0, 1, 2, 3, 4] and (m > 0).sum().item() == 2 * 4 * 9


def test_no_cpu_fallback(golden):
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss, InfoNCELoss
    fx = golden('unet3d_combo.pt')
    net = UNet(**fx['kwargs'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net(fx['x'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ComboLoss()(torch.rand(1, 1, 4, 4), torch.rand(1, 1, 4, 4))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        InfoNCELoss(set_size=2, device='cpu')(torch.rand(2, 4), torch.rand(2, 4))


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it (a product path that routes through the oracle or any
    CPU restatement would void every parity claim)."""
    import re
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(_lib.__file__)))
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f), errors='ignore').read()
                if re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M) or 'oracle/' in text:
                    offenders.append(os.path.join(root, f))
    assert not offenders, offenders


@pytest.mark.skipif(not os.path.isdir('/root/reference/code'), reason='needs the reference tree (present in the build container only)')
def test_same_seed_same_initial_weights_as_reference():
    """The drop-in builds its sub-modules in the reference's construction order (UNet.py:66-76 interleaves down_block[i], up_block[i],
    up_samp[i]), so `torch.manual_seed(s); UNet(...)` draws the same initial parameters as the reference (seeded-run comparability,
    UNet2D_scripts.py:53-60).  Checked in a child process: both package trees are called `src`."""
    import subprocess
    code = '''
import sys, torch
sys.dont_write_bytecode = True
sys.path.insert(0, sys.argv[1])
from src.models.networks.UNet import UNet, UNet_Encoder, Partial_UNet
cases = [(UNet, dict(depth=4, use_3D=True, top_filter=8, p_dropout=0.0)), (UNet, dict(depth=3, bilinear=True, top_filter=8)),
         (UNet_Encoder, dict(depth=3, top_filter=8, MLP_head=[16, 8])), (Partial_UNet, dict(depth=4, n_decoder=2, top_filter=8, head_channel=[8, 4]))]
for cls, kw in cases:
    torch.manual_seed(3)
    sd = cls(**kw).state_dict()
    print(cls.__name__, len(sd), ' '.join(f'{v.double().sum().item():.10e}' for v in sd.values()))
'''
    outs = [subprocess.run([sys.executable, '-c', code, path], capture_output=True, text=True, check=True).stdout
            for path in ('/root/reference/code', os.path.join(PKG, 'code'))]
    assert outs[0] == outs[1] and outs[0].count('\n') == 4


def test_slab_kernel_planner_limits():
    """Host-only planner of the tcgen05 slab kernel (ich_conv_tc_plan_info, no driver needed): for every conv shape of the BASELINE
    configs the tiling must respect the hardware limits -- 227 KB of shared memory, 512 TMEM columns, >= 2 pipeline stages -- and the
    128-channel-multiple layers must get 128-wide cout blocks (3-D ones in kd-split mode)."""
    l = _lib.lib()
    shapes = []
    for c, (d, h, w) in ((32, (64, 128, 128)), (64, (32, 64, 64)), (128, (16, 32, 32)), (256, (8, 16, 16))):      # cfg-3 levels
        shapes += [(8, d, h, w, c, c, 3), (8, d, h, w, 2 * c, c, 3), (8, d, h, w, c // 2 if c > 32 else 16, c, 3), (8, d, h, w, c, 2 * c, 3)]
    for c, s in ((32, 512), (64, 256), (128, 128), (256, 64), (512, 32)):                                       # cfg-2 levels (2-D)
        shapes += [(32, 1, s, s, c, c, 1), (32, 1, s, s, 2 * c, c, 1), (32, 1, s, s, max(c // 2, 16), c, 1)]
    for n, d, h, w, cin, cout, kd in shapes:
        out = (ctypes.c_longlong * 10)()
        rc = l.ich_conv_tc_plan_info(n, d, h, w, cin, cout, kd, 3, 3, ctypes.cast(out, ctypes.c_void_p))
        assert rc == 0, (n, d, h, w, cin, cout, kd)
        nb, r, t, nacc, stages, kds, smem, tmem, cw, items = list(out)
        assert smem <= 227 * 1024 and tmem <= 512 and nacc * t * nb <= 512 and stages >= 2 and cout % nb == 0 and items > 0
        if cout % 128 == 0 and h * w >= 1024:
            assert nb == 128 and kds == (1 if kd == 3 else 0), (cin, cout, kd, nb, kds)
        else:
            assert nb <= 64 and kds == 0


def test_stream_kernel_planner_limits():
    """Host-only planner of the plane-streaming kernel (ich_conv_tc_stream_plan_info): on every 3x3x3 shape it can be given -- the
    full-resolution layers of cfg-3 / cfg-5 in both directions, the mid-resolution layers it takes under ICH_TC_STREAM=2, flat tiling
    (W < 128), short volumes -- the accumulator ring must fit the 512 TMEM columns, an item must not have more tiles than the two issuer
    warps' tile tables hold, shared memory <= 227 KB, >= 2 stages; the ring shared by the tiles is planned exactly for cout blocks of 32
    (where it measured faster), and Cin / Cout that are not multiples of 16 are refused."""
    l = _lib.lib()
    shapes = [(8, 64, 128, 128, 16, 32), (8, 64, 128, 128, 32, 16), (8, 64, 128, 128, 64, 32), (8, 64, 128, 128, 32, 64), (8, 64, 128, 128, 32, 32),
              (4, 32, 128, 128, 64, 32), (1, 4, 8, 128, 16, 16), (8, 32, 64, 64, 32, 64), (8, 32, 64, 64, 64, 32), (8, 32, 64, 64, 64, 64),
              (8, 32, 64, 64, 128, 64), (8, 16, 32, 32, 128, 128), (2, 20, 12, 256, 32, 32), (1, 5, 7, 40, 48, 48)]
    for n, d, h, w, cin, cout in shapes:
        out = (ctypes.c_longlong * 10)()
        rc = l.ich_conv_tc_stream_plan_info(n, d, h, w, cin, cout, ctypes.cast(out, ctypes.c_void_p))
        assert rc == 0, (n, d, h, w, cin, cout)
        nb, r, t, slots, stages, resident, smem, tmem, cring, items = list(out)
        assert cout % nb == 0 and nb in (16, 32, 48, 64) and slots in (4, 8)
        assert t * slots * nb <= 512 and tmem <= 512 and t <= 4 and 1 <= r <= h
        assert smem <= 227 * 1024 and stages >= 2 and items > 0
        assert cring == (1 if (nb <= 32 and slots == 4) else 0), (nb, slots, cring)
        if cin * 27 * nb * 2 <= 96 * 1024:
            assert resident == 1                      # small weight blocks stay in shared memory for the CTA's lifetime
    out = (ctypes.c_longlong * 10)()
    for bad in ((8, 64, 128, 128, 8, 32), (8, 64, 128, 128, 32, 24), (8, 2, 128, 128, 32, 32), (8, 64, 128, 192, 32, 32)):
        assert l.ich_conv_tc_stream_plan_info(*bad, ctypes.cast(out, ctypes.c_void_p)) != 0, bad


def test_shared_accumulator_ring_invariants():
    """Model of the streaming kernel's shared accumulator ring (conv_tc_stream.cu, SParams::cring): tile t keeps output plane g in cell
    (slots * t + g) mod (slots * T).  While plane g is the newest acquired one, the planes g - slots + 1 .. g of every tile are alive
    (three accumulating, the rest draining): no two alive (tile, plane) pairs may share a cell, the cell a tile takes for a new plane must be
    the one plane g - slots of its neighbour tile has left (that is what the per-plane `tempty` barrier guarantees), and the three live
    planes of a tile straddle the end of the ring -- the MMA splits in two -- for exactly 2 of every slots * T planes (per-tile rings: 2 of
    every `slots`)."""
    for slots, T in ((4, 4), (4, 3), (4, 2), (4, 1), (8, 4), (8, 2)):
        ring = slots * T
        cell = lambda t, g: (slots * t + g) % ring                                    # noqa: E731
        splits = [0] * T
        for g in range(slots, slots + 6 * ring):                                      # six whole turns of the ring
            alive = {}
            for t in range(T):
                for q in range(g - slots + 1, g + 1):
                    c = cell(t, q)
                    assert c not in alive, (slots, T, g, t, q, alive[c])
                    alive[c] = (t, q)
            for t in range(T):
                assert cell(t, g) == cell((t + 1) % T, g - slots)                         # freed by the neighbour's plane g - slots
                lo = g - 2                                                               # live window lo .. g (three planes)
                if cell(t, lo) + 3 > ring:
                    splits[t] += 1
        assert splits == [12] * T, (slots, T, splits)                                 # 2 per turn and tile
