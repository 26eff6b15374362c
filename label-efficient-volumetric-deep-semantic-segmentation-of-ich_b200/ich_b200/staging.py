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
    """for batch in DevicePrefetcher(loader, device): ...

    One batch of look-ahead: batch k+1 is copied host->device on a side stream while step k runs.  Device buffers are
    three static sets per (shape, dtype) slot used round-robin (no allocator traffic in the loop), so a yielded batch stays
    valid until two further batches have been requested."""

    SETS = 3

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._bufs = {}
        self._turn = 0

    def __len__(self):
        return len(self.loader)

    def _buffer(self, slot, t):
        key = (slot, self._turn % self.SETS, tuple(t.shape), t.dtype)
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            self._bufs[key] = b
        return b

    def _stage(self, batch):
        if not self.cuda:
            return batch
        # the buffer set being overwritten was handed out SETS batches ago; order the copy after the compute stream's work
        # enqueued so far (which includes every kernel that read it), then overlap with the step launched next
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        # Device buffers are created under the COMPUTE stream (the caching allocator keeps one pool per stream: a buffer first
        # requested under a fresh side stream is a synchronous cudaMalloc -- measured +2..8 ms per step until all sets exist).
        slot = [0]
        pairs = []

        def reserve(t):
            if t.is_cuda:
                return t
            i = slot[0]
            slot[0] += 1
            if not t.is_pinned():
                t = t.pin_memory()
            buf = self._buffer(i, t)
            pairs.append((buf, t))
            return buf
        out = _map(batch, reserve)
        with torch.cuda.stream(self.stream):
            for buf, t in pairs:
                buf.copy_(t, non_blocking=True)
        self._turn += 1
        return out

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
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                yield cur
                return
            yield cur


def window_ct(ct_scan, win_center=40, win_width=120, out_range=(0, 1)):
    """utils/ct_utils.py:13-36 on a torch tensor: rescale [center - width/2, center + width/2] to out_range, clip.  CUDA tensors
    (int16 / uint16 / uint8 / fp32 Hounsfield units) go through the staging kernel (`ops.stage_ct`, one pass, fp32 result; pass
    dtype there to get the engine dtype directly); CPU tensors take the torch expression below (host-side pipelines)."""
    if ct_scan.is_cuda:
        from . import ops
        return ops.stage_ct(ct_scan, win_center, win_width, out_range, dtype=torch.float32)
    win_min = win_center - win_width / 2
    win_max = win_center + win_width / 2
    out = (out_range[1] - out_range[0]) * (ct_scan.float() - win_min) / (win_max - win_min) + out_range[0]
    return out.clamp_(out_range[0], out_range[1])
