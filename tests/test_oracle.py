"""CPU: the oracle restatement reproduces every golden vector generated from the unmodified reference modules
(oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import unet_oracle as UO, losses_oracle as LO


def test_unet3d_combo(golden):
    fx = golden('unet3d_combo.pt')
    x = fx['x']
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in fx['state_dict'].items()}
    ns = {}
    out = UO.unet_forward(x, sd, use_3D=True, training=True, new_stats=ns)
    assert torch.allclose(out, fx['out_train'], atol=2e-6)
    loss = LO.combo_loss(out, fx['mask'], **fx['loss_kwargs'])
    assert abs(loss.item() - fx['loss'].item()) <= 1e-5 * abs(fx['loss'].item())
    loss.backward()
    for k, g in fx['grads'].items():
        ref = g
        err = (sd[k].grad - ref).norm() / (ref.norm() + 1e-12)
        if k.endswith('conv1.bias') or k.endswith('conv2.bias'):      # pre-BN biases: true gradient 0, reference = fp noise
            assert sd[k].grad.abs().max() < 1e-2 * fx['grads']['final_conv.bias'].abs().max()
        else:
            assert err < 1e-3, (k, err.item())
    for k, v in ns.items():
        assert torch.allclose(v.float(), fx['state_dict_after'][k].float(), atol=1e-6), k
    ev = UO.unet_forward(x, fx['state_dict_after'], use_3D=True, training=False)
    assert torch.allclose(ev, fx['out_eval'], atol=2e-6)


def test_unet2d_dice(golden):
    fx = golden('unet2d_dice.pt')
    out = UO.unet_forward(fx['x'], fx['state_dict'], use_3D=False, training=True)
    assert torch.allclose(out, fx['out_train'], atol=2e-6)
    loss = LO.binary_dice_loss(out, fx['mask'], **fx['loss_kwargs'])
    assert abs(loss.item() - fx['loss'].item()) < 1e-6


def test_softmax_and_bottleneck(golden):
    fx = golden('unet3d_softmax.pt')
    out, xb = UO.unet_forward(fx['x'], fx['state_dict'], use_3D=True, training=True, return_bottleneck=True)
    assert torch.allclose(out, fx['out_train'], atol=5e-6)
    assert torch.allclose(xb, fx['bottleneck'], atol=5e-6)


def test_encoder_infonce(golden):
    fx = golden('encoder_infonce.pt')
    z1 = torch.nn.functional.normalize(UO.unet_encoder_forward(fx['x1'], fx['state_dict']), dim=1)
    z2 = torch.nn.functional.normalize(UO.unet_encoder_forward(fx['x2'], fx['state_dict']), dim=1)
    assert torch.allclose(z1, fx['z1'], atol=2e-6) and torch.allclose(z2, fx['z2'], atol=2e-6)
    assert abs(LO.info_nce_loss(z1, z2, fx['tau']).item() - fx['loss'].item()) < 5e-6


def test_partial_local_infonce(golden):
    fx = golden('partial_local_infonce.pt')
    f1 = UO.partial_unet_forward(fx['x1'], fx['state_dict'], use_3D=False)
    f2 = UO.partial_unet_forward(fx['x2'], fx['state_dict'], use_3D=False)
    assert torch.allclose(f1, fx['f1'], atol=5e-6)
    np.random.seed(fx['np_seed'])
    assert abs(LO.local_info_nce_loss(f1, f2, **fx['loss_kwargs']).item() - fx['loss'].item()) < 5e-6


def test_loss_vectors(golden):
    fx = golden('losses.pt')
    for c in fx['cases']:
        fn = LO.binary_dice_loss if c['kind'] == 'dice' else LO.combo_loss
        p = fx['pred'].clone().requires_grad_(True)
        v = fn(p, fx['mask'], **c['kwargs'])
        assert torch.allclose(v, c['value'], rtol=1e-6, atol=1e-6)
        v.sum().backward()
        assert torch.allclose(p.grad, c['grad'], rtol=1e-4, atol=1e-6)
    for c in fx['infonce']:
        assert abs(LO.info_nce_loss(c['z1'], c['z2'], c['tau']).item() - c['value'].item()) < 5e-6
    for c in fx['local']:
        np.random.seed(c['np_seed'])
        assert abs(LO.local_info_nce_loss(c['f1'], c['f2'], c['tau'], c['K'], c['n_region']).item() - c['value'].item()) < 5e-6


def test_tversky_vectors(golden):
    """TverskyLoss (LossFunctions.py:65-114): oracle vs the reference module's values and input gradients."""
    fx = golden('tversky.pt')
    for c in fx['cases']:
        p = fx['pred'].clone().requires_grad_(True)
        v = LO.tversky_loss(p, fx['mask'], **c['kwargs'])
        assert v.shape == c['value'].shape and torch.allclose(v, c['value'], rtol=1e-6, atol=1e-7)
        v.sum().backward()
        assert torch.allclose(p.grad, c['grad'], rtol=1e-5, atol=1e-8)


def test_sliding_window_identity():
    """Non-overlapping windows == whole-volume eval forward when the window tiles the volume exactly is NOT true in
    general (zero padding at window borders), but a single window covering the volume must be the identity."""
    torch.manual_seed(0)
    fx_sd = torch.load(__import__('os').path.join(__import__('os').path.dirname(__file__), 'golden', 'unet3d_combo.pt'))['state_dict_after']
    vol = torch.rand(1, 1, 8, 16, 16)
    whole = UO.unet_forward(vol, fx_sd, training=False)
    pred, mask = UO.sliding_window_predict(vol, fx_sd, (8, 16, 16), (8, 16, 16))
    assert torch.equal(pred, whole) and torch.equal(mask, whole >= 0.5)
    pred2, _ = UO.sliding_window_predict(vol, fx_sd, (8, 8, 8), (8, 4, 4))
    assert pred2.shape == vol.shape and torch.isfinite(pred2).all()


def test_bilinear_decoders(golden):
    """bilinear=True nets (nn.Upsample align_corners=True, reference UNet.py:69-72): 3-D trilinear and 2-D bilinear."""
    fxs = golden('unet_bilinear.pt')
    for name, fx in fxs.items():
        use_3D = fx['kwargs']['use_3D']
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in fx['state_dict'].items()}
        out = UO.unet_forward(fx['x'], sd, use_3D=use_3D, training=True)
        assert torch.allclose(out, fx['out_train'], atol=2e-6), name
        loss = LO.combo_loss(out, fx['mask'], **fx['loss_kwargs'])
        assert abs(loss.item() - fx['loss'].item()) <= 1e-5 * abs(fx['loss'].item())
        loss.backward()
        for k, g in fx['grads'].items():
            if k.endswith('conv1.bias') or k.endswith('conv2.bias'):
                continue
            assert (sd[k].grad - g).norm() / (g.norm() + 1e-12) < 1e-3, (name, k)



def test_window_ct_vectors(golden):
    """oracle/ct_oracle.window_ct against utils/ct_utils.py:13-36 run on int16 / uint16 / uint8 / fp32 Hounsfield-unit arrays."""
    import numpy as np
    from oracle import ct_oracle as CO
    cases = golden('window_ct.pt')
    assert len(cases) == 12
    for c in cases:
        got = CO.window_ct(c['hu'].numpy(), c['center'], c['width'], c['out_range'])
        assert np.array_equal(got, c['out'].numpy()), (c['dtype'], c['center'])


def test_segment_volume_slices_rule():
    """Slice-wise volume segmentation rule (UNet2D.py:272-314 without resize / NIfTI): rotation there and back, batching, 0 / 255."""
    import numpy as np
    from oracle import ct_oracle as CO
    hu = np.random.RandomState(1).randint(-100, 200, size=(6, 4, 5)).astype(np.int16)
    fwd = lambda x: (x > 0.5).float()                                  # a "network" that thresholds the windowed input
    seg = CO.segment_volume_slices(hu, fwd, window=(40, 120), batch_size=2)
    want = ((CO.window_ct(hu, 40, 120) > 0.5) * 255).astype(np.uint8)  # rot90 there and back is the identity on a pointwise rule
    assert seg.shape == hu.shape and np.array_equal(seg, want)
