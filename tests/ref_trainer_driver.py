"""Helper of tests/test_parity_r2.py::test_reference_trainer_on_the_drop_in (run as a child process; both package trees are called `src`).

    python tests/ref_trainer_driver.py dropin|reference <device> <init_state.pt> <out.json>

Imports the UNMODIFIED reference trainer `src.models.optim.UNet2D.UNet2D` from baseline/_ref/code (the I/O-only dependencies that are not
installed -- skimage, nibabel -- are stubbed with empty modules) and runs its own train() / evaluate()
loops on a small synthetic Dataset.  `dropin`: `src.models.networks.UNet`, `src.models.optim.LossFunctions` and
`src.utils.tensor_utils` resolve to THIS repo's drop-in modules (the package shadowing of INTEGRATION.md, option B) and everything
else -- the trainer, the transforms, the utilities -- to the reference tree.  `reference`: the reference tree only."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200')
REF = os.path.join(ROOT, 'baseline', '_ref', 'code')


def main():
    mode, device, init_path, out_path = sys.argv[1:5]
    sys.dont_write_bytecode = True
    for name in ('skimage', 'skimage.io', 'skimage.transform', 'skimage.filters', 'skimage.util', 'skimage.morphology', 'skimage.measure',
                 'skimage.exposure', 'nibabel'):
        sys.modules[name] = types.ModuleType(name)
    for sub in ('io', 'transform', 'filters', 'util', 'morphology', 'measure', 'exposure'):
        setattr(sys.modules['skimage'], sub, sys.modules['skimage.' + sub])
    import torch
    if mode == 'dropin':
        sys.path.insert(0, PKG)
        sys.path.insert(0, os.path.join(PKG, 'code'))
        import importlib
        for pkg, sub in (('src', ''), ('src.models', 'models'), ('src.models.networks', 'models/networks'), ('src.models.optim', 'models/optim'),
                         ('src.utils', 'utils')):
            importlib.import_module(pkg).__path__.append(os.path.join(REF, 'src', sub))      # the rest of `src` comes from the reference
        from ich_b200 import config
        config.set(precision=os.environ.get('ICH_B200_PRECISION', 'fp32'))
    else:
        sys.path.insert(0, REF)
    from src.models.optim.UNet2D import UNet2D                 # the reference's trainer in both modes
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import BinaryDiceLoss
    import src.models.optim.UNet2D as trainer_mod
    import src.models.networks.UNet as net_mod
    assert 'baseline' in trainer_mod.__file__
    assert ('baseline' in net_mod.__file__) == (mode == 'reference'), net_mod.__file__

    class Slices(torch.utils.data.Dataset):
        """(input [1, H, W] float, target [1, H, W] bool, volID, slice number) -- the contract of UNet2D.py:135,214."""

        def __init__(self, n=24, size=64, seed=0):
            g = torch.Generator().manual_seed(seed)
            self.x = torch.rand(n, 1, size, size, generator=g)
            yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
            cy, cx = torch.randint(16, size - 16, (n,), generator=g), torch.randint(16, size - 16, (n,), generator=g)
            self.m = ((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2 < 81)[:, None]
            self.m[::4] = False                                  # some empty slices (alpha-weighted Dice branch)
            self.x = self.x + 0.8 * self.m.float()               # lesions are brighter: learnable in a few steps

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], self.m[i], torch.tensor(i // 8), torch.tensor(i % 8)

    torch.manual_seed(0)
    net = UNet(depth=3, use_3D=False, in_channels=1, out_channels=1, top_filter=16, midchannels_factor=1, p_dropout=0.0)
    net.load_state_dict(torch.load(init_path))
    trainer = UNet2D(net, n_epoch=10, batch_size=8, lr=1e-3, loss_fn=BinaryDiceLoss, loss_fn_kwargs=dict(reduction='mean', p=2, alpha=0.2),
                     weight_decay=1e-6, num_workers=0, device=device, print_progress=False)
    ckpt = out_path + '.ckpt'
    if os.path.exists(ckpt):       # (the reference's own resume path cannot load its checkpoints on torch >= 2.6: numpy scalars vs weights_only)
        os.remove(ckpt)
    torch.manual_seed(1)                                       # DataLoader shuffling order
    trainer.train(Slices(), valid_dataset=Slices(seed=1), checkpoint_path=ckpt)
    trainer.evaluate(Slices(seed=2), print_to_logger=False)
    ck = torch.load(ckpt, map_location='cpu', weights_only=False)      # the reference stores numpy scalars in it
    json.dump({'evolution': trainer.outputs['train']['evolution'], 'dice': trainer.outputs['eval']['dice'],
               'ckpt_epochs': ck['n_epoch_finished'], 'ckpt_keys': sorted(ck['net_state'].keys())[:3], 'n_keys': len(ck['net_state'])}, open(out_path, 'w'))


if __name__ == '__main__':
    main()
