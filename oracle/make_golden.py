"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from
/root/reference/code, CPU fp32) on seeded inputs, and check the oracle restatement against them.

Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import importlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('ICH_REFERENCE_CODE', '/root/reference/code')
OUT = os.path.join(ROOT, 'tests', 'golden')


def load_reference():
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k == 'src' or k.startswith('src.')]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    unet = importlib.import_module('src.models.networks.UNet')
    losses = importlib.import_module('src.models.optim.LossFunctions')
    sys.path.remove(REF)
    return unet, losses


def grads_of(net):
    return {k: p.grad.clone() for k, p in net.named_parameters()}


def bilinear_fixtures(ref_unet, ref_loss, UO, LO, report):
    """bilinear=True decoders (reference UNet.py:69-72): 3-D trilinear and 2-D bilinear nets, ComboLoss, all gradients."""
    cases = {}
    for name, use_3D, shape, seed in [('3d', True, (2, 1, 8, 16, 16), 21), ('2d', False, (2, 1, 32, 32), 22)]:
        torch.manual_seed(seed)
        kw = dict(depth=3, use_3D=use_3D, bilinear=True, in_channels=1, out_channels=1, top_filter=8, midchannels_factor=2, p_dropout=0.0)
        net = ref_unet.UNet(**kw).train()
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(*shape, generator=g)
        mask = (torch.rand(*shape, generator=g) > 0.9).float()
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        lk = dict(alpha=0.5, beta=0.5, reduction='mean', p=1)
        out = net(x)
        loss = ref_loss.ComboLoss(**lk)(out, mask)
        loss.backward()
        cases[name] = dict(kwargs=kw, x=x, mask=mask, state_dict=sd0, out_train=out.detach(), loss=loss.detach(), grads=grads_of(net),
                           loss_kwargs=lk)
        o = UO.unet_forward(x, sd0, use_3D=use_3D, training=True)
        report[f'bilinear {name} out'] = (o - out).abs().max().item()
        report[f'bilinear {name} loss'] = abs(LO.combo_loss(o, mask, **lk).item() - loss.item()) / abs(loss.item())
    torch.save(cases, os.path.join(OUT, 'unet_bilinear.pt'))


def tversky_fixtures(ref_loss, LO, report):
    """TverskyLoss known-answer vectors (values + input gradients) from the reference module -> tests/golden/tversky.pt."""
    g = torch.Generator().manual_seed(9)
    pred = torch.rand(4, 1, 4, 8, 8, generator=g)
    mask = (torch.rand(4, 1, 4, 8, 8, generator=g) > 0.8).float()
    mask[1] = 0                      # empty mask -> alpha scaling
    mask[3] = 1                      # full mask
    pred[0, 0, 0, 0, :2] = torch.tensor([0.0, 1.0])
    cases = []
    for lk in [dict(alpha=1.0, beta=0.5, gamma=0.5, reduction='mean'), dict(alpha=0.2, beta=0.7, gamma=0.3, reduction='none'),
               dict(alpha=0.5, beta=0.3, gamma=0.9, reduction='sum'), dict(alpha=0.0, beta=1.0, gamma=0.0, reduction='mean')]:
        p = pred.clone().requires_grad_(True)
        v = ref_loss.TverskyLoss(**lk)(p, mask)
        v.sum().backward()
        cases.append(dict(kwargs=lk, value=v.detach(), grad=p.grad.clone()))
        report[f'TverskyLoss {lk}'] = ((LO.tversky_loss(pred, mask, **lk) - v).abs().max() / v.abs().max().clamp_min(1e-12)).item()
    torch.save(dict(pred=pred, mask=mask, cases=cases), os.path.join(OUT, 'tversky.pt'))


def window_ct_fixture(report):
    """utils/ct_utils.py:13-36 run on Hounsfield-unit arrays of the dtypes a CT pipeline holds (int16 / uint16 / uint8 / fp32)."""
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == 'src' or k.startswith('src.')]:
        del sys.modules[k]
    ct_utils = importlib.import_module('src.utils.ct_utils')
    sys.path.remove(REF)
    sys.path.insert(0, ROOT)
    from oracle import ct_oracle as CO
    rng = np.random.RandomState(0)
    cases = []
    for dtype, lo, hi in ((np.int16, -1200, 3000), (np.uint16, 0, 4095), (np.uint8, 0, 255), (np.float32, -1200, 3000)):
        hu = rng.uniform(lo, hi, size=(3, 17, 19)).astype(dtype)
        for center, width, rng_out in ((40, 120, (0, 1)), (50, 100, (0, 255)), (600, 2800, (0, 1))):
            want = ct_utils.window_ct(hu.astype(np.float64), win_center=center, win_width=width, out_range=rng_out)
            got = CO.window_ct(hu, center, width, rng_out)
            report[f'window_ct/{np.dtype(dtype).name}/{center}-{width}'] = float(np.abs(got - want).max())
            cases.append(dict(hu=torch.from_numpy(hu.astype(np.int32 if dtype == np.uint16 else dtype)), dtype=np.dtype(dtype).name, center=center, width=width,
                              out_range=rng_out, out=torch.from_numpy(want)))
    torch.save(cases, os.path.join(OUT, 'window_ct.pt'))


def gated_unet_fixtures(report):
    """models/networks/GatedUNet.py: UNet(use_gatedConv=True / False) -- the ad-attention side-track's U-Net (ConvLayer / GatedConv
    blocks, bilinear decoder when gated).  Golden forward / loss / gradients for the drop-in's GatedUNet."""
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == 'src' or k.startswith('src.')]:
        del sys.modules[k]
    gated = importlib.import_module('src.models.networks.GatedUNet')
    losses = importlib.import_module('src.models.optim.LossFunctions')
    sys.path.remove(REF)
    out = {}
    for name, kw, shape in (('gated2d', dict(depth=3, use_3D=False, in_channels=2, out_channels=1, top_filter=8, midchannels_factor=2, p_dropout=0.0,
                                              use_gatedConv=True), (2, 2, 16, 16)),
                            ('plain3d', dict(depth=3, use_3D=True, in_channels=1, out_channels=1, top_filter=8, midchannels_factor=2, p_dropout=0.0,
                                             use_gatedConv=False), (2, 1, 8, 16, 16)),
                            ('gated3d', dict(depth=3, use_3D=True, in_channels=1, out_channels=2, top_filter=8, midchannels_factor=1, p_dropout=0.0,
                                             use_gatedConv=True), (1, 1, 8, 16, 16))):
        torch.manual_seed(5)
        net = gated.UNet(**kw).train()
        g = torch.Generator().manual_seed(5)
        x = torch.rand(*shape, generator=g)
        oshape = (shape[0], kw['out_channels']) + tuple(shape[2:])
        mask = (torch.rand(*oshape, generator=g) > 0.8).float()
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        lossf = losses.BinaryDiceLoss(reduction='mean', p=2, alpha=1.0)
        o = net(x)
        loss = lossf(o, mask)
        loss.backward()
        sd1 = {k: v.clone() for k, v in net.state_dict().items()}
        net.eval()
        with torch.no_grad():
            oe = net(x)
        out[name] = dict(kwargs=kw, x=x, mask=mask, state_dict=sd0, out_train=o.detach(), loss=loss.detach(), grads=grads_of(net),
                         state_dict_after=sd1, out_eval=oe, loss_kwargs=dict(reduction='mean', p=2, alpha=1.0))
        report[f'gated_unet/{name}/n_keys'] = 0.0
    torch.save(out, os.path.join(OUT, 'gated_unet.pt'))


def main():
    sys.path.insert(0, ROOT)
    from oracle import unet_oracle as UO, losses_oracle as LO
    ref_unet, ref_loss = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    report = {}
    only = [a for a in sys.argv if a in ('--bilinear-only', '--tversky-only', '--window-ct-only', '--gated-only')]
    if only:      # add fixtures without rewriting the others
        if '--bilinear-only' in only:
            bilinear_fixtures(ref_unet, ref_loss, UO, LO, report)
        if '--tversky-only' in only:
            tversky_fixtures(ref_loss, LO, report)
        if '--window-ct-only' in only:
            window_ct_fixture(report)
        if '--gated-only' in only:
            gated_unet_fixtures(report)
        for k, v in report.items():
            print(f'{k:60s} oracle-vs-reference {v:.3e}')
        assert max(report.values()) < 5e-5
        return

    # ---- 1. supervised 3-D U-Net + ComboLoss (cfg-1 graph, shrunk) ------------------------------------
    torch.manual_seed(0)
    kw = dict(depth=3, use_3D=True, in_channels=1, out_channels=1, top_filter=8, midchannels_factor=2, p_dropout=0.0)
    net = ref_unet.UNet(**kw).train()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 8, 16, 16, generator=g)
    mask = (torch.rand(2, 1, 8, 16, 16, generator=g) > 0.9).float()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    lossf = ref_loss.ComboLoss(alpha=0.5, beta=0.5, reduction='mean', p=1)
    out = net(x)
    loss = lossf(out, mask)
    loss.backward()
    sd1 = {k: v.clone() for k, v in net.state_dict().items()}
    net.eval()
    with torch.no_grad():
        out_eval = net(x)
    fx = dict(kwargs=kw, x=x, mask=mask, state_dict=sd0, out_train=out.detach(), loss=loss.detach(), grads=grads_of(net),
              state_dict_after=sd1, out_eval=out_eval, loss_kwargs=dict(alpha=0.5, beta=0.5, reduction='mean', p=1))
    torch.save(fx, os.path.join(OUT, 'unet3d_combo.pt'))
    ns = {}
    o = UO.unet_forward(x, sd0, use_3D=True, training=True, new_stats=ns)
    report['unet3d out'] = (o - out).abs().max().item()
    report['unet3d loss'] = abs(LO.combo_loss(o, mask).item() - loss.item()) / abs(loss.item())
    report['unet3d running stats'] = max((ns[k].float() - sd1[k].float()).abs().max().item() for k in ns)
    sd1e = dict(sd1)
    report['unet3d eval'] = (UO.unet_forward(x, sd1e, use_3D=True, training=False) - out_eval).abs().max().item()

    # ---- 2. supervised 2-D U-Net + BinaryDiceLoss (cfg-2 graph, shrunk) -------------------------------
    torch.manual_seed(1)
    kw = dict(depth=3, use_3D=False, in_channels=1, out_channels=1, top_filter=8, midchannels_factor=1, p_dropout=0.0)
    net = ref_unet.UNet(**kw).train()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(3, 1, 32, 32, generator=g)
    mask = (torch.rand(3, 1, 32, 32, generator=g) > 0.9).float()
    mask[2] = 0          # one sample without positives -> the alpha branch (LossFunctions.py:56)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    lk = dict(reduction='mean', p=2, alpha=0.2)
    out = net(x)
    loss = ref_loss.BinaryDiceLoss(**lk)(out, mask)
    loss.backward()
    torch.save(dict(kwargs=kw, x=x, mask=mask, state_dict=sd0, out_train=out.detach(), loss=loss.detach(),
                    grads=grads_of(net), loss_kwargs=lk), os.path.join(OUT, 'unet2d_dice.pt'))
    o = UO.unet_forward(x, sd0, use_3D=False, training=True)
    report['unet2d out'] = (o - out).abs().max().item()
    report['unet2d loss'] = abs(LO.binary_dice_loss(o, mask, **lk).item() - loss.item())

    # ---- 3. multi-class head (softmax) ---------------------------------------------------------------
    torch.manual_seed(2)
    kw = dict(depth=2, use_3D=True, in_channels=2, out_channels=3, top_filter=8, midchannels_factor=2, p_dropout=0.0)
    net = ref_unet.UNet(**kw).train()
    x = torch.rand(1, 2, 4, 8, 8, generator=torch.Generator().manual_seed(2))
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net.return_bottleneck = True
    out, xb = net(x)
    torch.save(dict(kwargs=kw, x=x, state_dict=sd0, out_train=out.detach(), bottleneck=xb.detach()),
               os.path.join(OUT, 'unet3d_softmax.pt'))
    o, ob = UO.unet_forward(x, sd0, use_3D=True, training=True, return_bottleneck=True)
    report['softmax out'] = max((o - out).abs().max().item(), (ob - xb).abs().max().item())

    # ---- 4. encoder + global InfoNCE (cfg-4 global, shrunk) --------------------------------------------
    torch.manual_seed(3)
    kw = dict(depth=3, use_3D=True, in_channels=1, MLP_head=[24, 16], top_filter=8, midchannels_factor=2, p_dropout=0.0)
    net = ref_unet.UNet_Encoder(**kw).train()
    g = torch.Generator().manual_seed(3)
    x1 = torch.rand(4, 1, 8, 16, 16, generator=g)
    x2 = torch.rand(4, 1, 8, 16, 16, generator=g)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    z1 = torch.nn.functional.normalize(net(x1), dim=1)
    z2 = torch.nn.functional.normalize(net(x2), dim=1)
    loss = ref_loss.InfoNCELoss(set_size=4, tau=0.1, device='cpu')(z1, z2)
    loss.backward()
    torch.save(dict(kwargs=kw, x1=x1, x2=x2, state_dict=sd0, z1=z1.detach(), z2=z2.detach(), loss=loss.detach(),
                    grads=grads_of(net), tau=0.1), os.path.join(OUT, 'encoder_infonce.pt'))
    o1 = torch.nn.functional.normalize(UO.unet_encoder_forward(x1, sd0, use_3D=True), dim=1)
    o2 = torch.nn.functional.normalize(UO.unet_encoder_forward(x2, sd0, use_3D=True), dim=1)
    report['encoder z'] = max((o1 - z1).abs().max().item(), (o2 - z2).abs().max().item())
    report['infonce'] = abs(LO.info_nce_loss(o1, o2, tau=0.1).item() - loss.item())

    # ---- 5. partial U-Net + local InfoNCE (cfg-4 local, shrunk) -----------------------------------------
    torch.manual_seed(4)
    kw = dict(depth=3, n_decoder=1, use_3D=False, in_channels=1, head_channel=[16, 8], top_filter=8,
              midchannels_factor=1, p_dropout=0.0)
    net = ref_unet.Partial_UNet(**kw).train()
    g = torch.Generator().manual_seed(4)
    x1 = torch.rand(2, 1, 32, 32, generator=g)
    x2 = torch.rand(2, 1, 32, 32, generator=g)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    f1, f2 = net(x1), net(x2)
    lk = dict(tau=0.1, K=3, n_region=4)
    np.random.seed(7)
    loss = ref_loss.LocalInfoNCELoss(device='cpu', **lk)(f1, f2)
    loss.backward()
    torch.save(dict(kwargs=kw, x1=x1, x2=x2, state_dict=sd0, f1=f1.detach(), f2=f2.detach(), loss=loss.detach(),
                    grads=grads_of(net), loss_kwargs=lk, np_seed=7), os.path.join(OUT, 'partial_local_infonce.pt'))
    o1 = UO.partial_unet_forward(x1, sd0, use_3D=False)
    o2 = UO.partial_unet_forward(x2, sd0, use_3D=False)
    report['partial f'] = max((o1 - f1).abs().max().item(), (o2 - f2).abs().max().item())
    np.random.seed(7)
    report['local infonce'] = abs(LO.local_info_nce_loss(o1, o2, **lk).item() - loss.item())

    # ---- 6. loss known-answer vectors (values + input grads) -----------------------------------------
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(3, 1, 4, 8, 8, generator=g).clamp(1e-4, 1 - 1e-4)
    mask = (torch.rand(3, 1, 4, 8, 8, generator=g) > 0.8).float()
    mask[1] = 0
    pred[0, 0, 0, 0, 0] = 0.0       # exercises log(0 + 1e-14)
    pred[0, 0, 0, 0, 1] = 1.0
    cases = []
    for name, cls, lk in [('dice', 'BinaryDiceLoss', dict(reduction='mean', p=2, alpha=0.2, eps=1)),
                          ('dice', 'BinaryDiceLoss', dict(reduction='none', p=1, alpha=1.0, eps=1)),
                          ('dice', 'BinaryDiceLoss', dict(reduction='sum', p=3, alpha=0.5, eps=2)),
                          ('combo', 'ComboLoss', dict(alpha=0.5, beta=0.5, reduction='mean', p=1)),
                          ('combo', 'ComboLoss', dict(alpha=0.3, beta=0.7, reduction='sum', p=2)),
                          ('combo', 'ComboLoss', dict(alpha=0.5, beta=0.5, reduction='none', p=1))]:
        p = pred.clone().requires_grad_(True)
        v = getattr(ref_loss, cls)(**lk)(p, mask)
        v.sum().backward()
        cases.append(dict(kind=name, kwargs=lk, value=v.detach(), grad=p.grad.clone()))
        fn = LO.binary_dice_loss if name == 'dice' else LO.combo_loss
        report[f'{cls} {lk}'] = ((fn(pred, mask, **lk) - v).abs().max() / v.abs().max()).item()
    nce = []
    for n, e, tau in [(4, 16, 0.1), (8, 128, 0.5), (40, 128, 0.1)]:
        g = torch.Generator().manual_seed(n)
        z1 = torch.randn(n, e, generator=g).requires_grad_(True)
        z2 = torch.randn(n, e, generator=g).requires_grad_(True)
        v = ref_loss.InfoNCELoss(set_size=n, tau=tau, device='cpu')(z1, z2)
        v.backward()
        nce.append(dict(z1=z1.detach(), z2=z2.detach(), tau=tau, value=v.detach(), g1=z1.grad.clone(), g2=z2.grad.clone()))
        report[f'InfoNCE n={n}'] = abs(LO.info_nce_loss(z1.detach(), z2.detach(), tau).item() - v.item())
    loc = []
    for (bs, H, W, C), K, A, tau, seed in [((2, 12, 12, 4), 3, 5, 0.5, 11), ((3, 9, 16, 8), 2, 6, 0.1, 12)]:
        g = torch.Generator().manual_seed(seed)
        f1 = torch.randn(bs, H, W, C, generator=g).requires_grad_(True)
        f2 = torch.randn(bs, H, W, C, generator=g).requires_grad_(True)
        np.random.seed(seed)
        v = ref_loss.LocalInfoNCELoss(tau=tau, K=K, n_region=A, device='cpu')(f1, f2)
        v.backward()
        loc.append(dict(f1=f1.detach(), f2=f2.detach(), tau=tau, K=K, n_region=A, np_seed=seed, value=v.detach(),
                        g1=f1.grad.clone(), g2=f2.grad.clone()))
        np.random.seed(seed)
        report[f'LocalInfoNCE {bs,H,W,C}'] = abs(LO.local_info_nce_loss(f1.detach(), f2.detach(), tau, K, A).item() - v.item())
    torch.save(dict(pred=pred, mask=mask, cases=cases, infonce=nce, local=loc), os.path.join(OUT, 'losses.pt'))

    bilinear_fixtures(ref_unet, ref_loss, UO, LO, report)
    tversky_fixtures(ref_loss, LO, report)

    worst = 0.0
    for k, v in report.items():
        print(f'{k:60s} oracle-vs-reference {v:.3e}')
        worst = max(worst, v)
    assert worst < 5e-5, f'oracle deviates from the reference: {worst}'
    print('oracle pinned; fixtures written to', OUT)


if __name__ == '__main__':
    main()
