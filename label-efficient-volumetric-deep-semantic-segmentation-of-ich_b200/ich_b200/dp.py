"""Data-parallel training for the drop-in U-Net: one process per GPU (torchrun), gradient all-reduce over NCCL
(NVLink 5 / NVSwitch) bucketed in backward order and overlapped with the rest of backward.  The reference has no
multi-GPU U-Net path; semantics chosen = DistributedDataParallel's: per-rank BatchNorm statistics,
gradients averaged over ranks, parameters / buffers broadcast from rank 0 once.

Installed on the module itself so the reference trainer loop (zero_grad / forward / backward / step) stays unchanged:

    net = UNet(...).to(device); ich_b200.dp.install(net)        # no-op when torch.distributed is not initialised
"""
import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params, device, dtype):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.flat = torch.zeros(n, dtype=dtype, device=device)
        self.pending = 0
        self.work = None


class GradAllReducer:
    def __init__(self, module, bucket_bytes=8 << 20, group=None, average=True):
        self.group = group
        self.world = dist.get_world_size(group)
        self.average = average
        params = [p for p in module.parameters() if p.requires_grad]
        # autograd produces gradients roughly in reverse registration order: final_conv first, down_block.0 last
        params = params[::-1]
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in params:
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self.buckets.append(_Bucket(cur, p.device, p.dtype))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, cur[0].device, cur[0].dtype))
        self._slot = {}
        for b in self.buckets:
            for p, off in zip(b.params, b.offsets):
                self._slot[p] = (b, off)
        self._callback_queued = False
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in params]
        for b in self.buckets:
            b.pending = len(b.params)

    def _hook(self, p):
        b, _ = self._slot[p]
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b):
        """Gather the bucket's gradients into its flat buffer with ONE multi-tensor copy (not one copy kernel per parameter), make
        the gradients alias the buffer, and start the all-reduce.  Parameters without a gradient this step (frozen / unused)
        contribute zeros and keep `grad = None`."""
        src, dst = [], []
        for p, off in zip(b.params, b.offsets):
            view = b.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                src.append(p.grad)
                dst.append(view)
                p.grad = view                   # the gradient now aliases the bucket: the all-reduce updates it in place
        if src:
            torch._foreach_copy_(dst, src)
        op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.group) == 'nccl') else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        b.avg_done = op == dist.ReduceOp.AVG

    def _finish(self):
        for b in self.buckets:
            if b.pending != 0:                  # some parameters received no gradient this step: reduce what is there
                self._launch(b)
            b.work.wait()
            if self.average and not b.avg_done:
                b.flat.div_(self.world)
            b.work = None
            b.pending = len(b.params)
        self._callback_queued = False

    def remove(self):
        for h in self._handles:
            h.remove()


def broadcast_state(module, src=0, group=None):
    """Parameters and BatchNorm buffers from rank `src` to every rank (once, at start)."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def install(module, bucket_bytes=8 << 20, group=None):
    """Attach the overlapped gradient all-reduce to `module`; returns the reducer (or None when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    broadcast_state(module, 0, group)
    red = GradAllReducer(module, bucket_bytes, group)
    module._ich_grad_reducer = red
    return red


def shard_indices(n_items, rank=None, world=None):
    """Contiguous, balanced shard of range(n_items) for this rank (volumes or sliding windows)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))
