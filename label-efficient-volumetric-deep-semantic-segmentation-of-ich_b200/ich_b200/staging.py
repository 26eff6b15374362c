"""Input staging for the training loop.

The reference trainers do `input.to(device).float()` from un-pinned memory inside the hot loop
(models/optim/UNet2D.py:137-138, Contrastive.py:134-135): a synchronous H2D copy per step.  `DevicePrefetcher` wraps any
iterable of batches (a DataLoader): batch k+1 is pinned and copied on a side stream while step k computes, and tensors are
yielded already on the device, so the trainer's `.to(device)` becomes a no-op.  `window_ct` is the on-device version of
utils/ct_utils.py:13-36 (window / rescale / clip), for pipelines that ship raw Hounsfield units to the GPU.
"""
import torch


def _map(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map(o, fn) for o in obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    return obj


class DevicePrefetcher:
    """for batch in DevicePrefetcher(loader, device): ...   (one batch of look-ahead, pinned staging, side stream)."""

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        if not self.cuda:
            return batch
        with torch.cuda.stream(self.stream):
            def put(t):
                if t.is_cuda:
                    return t
                if not t.is_pinned():
                    t = t.pin_memory()
                return t.to(self.device, non_blocking=True)
            return _map(batch, put)

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while True:
            cur = nxt
            if self.cuda:
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                _map(cur, lambda t: t.record_stream(torch.cuda.current_stream(self.device)) if t.is_cuda else None)
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                yield cur
                return
            yield cur


def window_ct(ct_scan, win_center=40, win_width=120, out_range=(0, 1)):
    """utils/ct_utils.py:13-36 on a torch tensor (any device): rescale [center - width/2, center + width/2] to out_range, clip."""
    win_min = win_center - win_width / 2
    win_max = win_center + win_width / 2
    out = (out_range[1] - out_range[0]) * (ct_scan.float() - win_min) / (win_max - win_min) + out_range[0]
    return out.clamp_(out_range[0], out_range[1])
