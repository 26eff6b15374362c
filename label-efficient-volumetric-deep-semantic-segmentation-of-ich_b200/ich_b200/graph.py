"""Opt-in CUDA-graph replay of a whole training step (zero_grad + forward + loss + backward [+ gradient all-reduce] + optimizer step).

A cfg-3 step is ~123 launches of this library plus ~40 torch launches (fills, multi-tensor Adam); the reference trainers read
`loss.item()` every step (models/optim/UNet2D.py:146-147), which drains the launch queue, so the host time of building the next
step (0.7-0.9 ms at 1 GPU, more with eight processes sharing the host) is GPU idle time.  One `cudaGraphLaunch` per step removes it.

    step = GraphedStep(lambda x, m: train_step(x, m), optimizer)      # train_step does zero_grad / forward / loss / backward / step
    for x, m in loader:
        loss = step(x, m)          # first `warmup` calls run eagerly, the next one captures, the rest replay
        loss.item()

What makes the step capturable: the library never allocates (workspaces are torch tensors from the graph's private pool), every launch
goes to torch's current stream, TMA descriptors and the batched weight-pack refresh travel as kernel parameters, BatchNorm counters
are device tensors.  Not capturable (raises from torch during capture): fused dropout (the Philox seed is drawn on the host every
step) and LocalInfoNCELoss (region corners are drawn from numpy's RNG on the host every step, as the reference does)."""
import torch

from . import _lib


class GraphedStep:
    def __init__(self, step_fn, optimizer=None, warmup=3, strict=True):
        """step_fn(*device_tensors) -> tensor: a training step (pass its optimizer) or any fixed-shape inference call, e.g.
        `lambda vol: infer.sliding_window_predict(net, vol, window, return_pred=False)[1]` (its window corners are cached on the device
        after the first call, so the rest is pure kernel launches).
        strict=False: a step that cannot be captured keeps running eagerly (`failed` holds the reason) instead of raising."""
        self.step_fn = step_fn
        self.strict = strict
        self.failed = None
        self.warmup = max(1, int(warmup))
        self.calls = 0
        self.graph = None
        self.static_in = None
        self.static_out = None
        self.kernels_per_replay = 0
        if optimizer is not None:
            for g in optimizer.param_groups:
                if 'capturable' in g:
                    if any(len(optimizer.state.get(p, {})) for p in g['params']) and not g['capturable']:
                        raise RuntimeError('ich_b200.GraphedStep: create it before the first optimizer step (Adam keeps its step counter on the '
                                           'host unless capturable=True is set before the state is initialised)')
                    g['capturable'] = True

    def _capture(self, inputs):
        self.static_in = [torch.empty_like(t) for t in inputs]
        for s, t in zip(self.static_in, inputs):
            s.copy_(t)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launches()
        try:
            # (capture runs on a side stream: the AccumulateGrad nodes of the warm-up steps belong to the default stream, which autograd reports)
            warn = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
            if warn is not None:
                warn(False)
            with torch.cuda.graph(graph):
                self.static_out = self.step_fn(*self.static_in)
        except Exception as e:      # noqa: BLE001 -- e.g. a host synchronisation inside the step (dropout seed, numpy-drawn region corners)
            if self.strict:
                raise
            self.failed = repr(e)
            torch.cuda.synchronize()
            from . import ops
            ops.invalidate_packs()          # pack refreshes recorded during the aborted capture never ran: re-derive on next use
            return False
        self.graph = graph
        self.kernels_per_replay = _lib.launches() - before
        torch.cuda.synchronize()
        return True

    def __call__(self, *inputs):
        if self.graph is None:
            if self.calls < self.warmup or self.failed:
                self.calls += 1
                return self.step_fn(*inputs)
            if not self._capture(inputs):  # capturing does not execute: fall through to the first replay on these inputs
                return self.step_fn(*inputs)
        if len(inputs) != len(self.static_in):
            raise RuntimeError(f'ich_b200.GraphedStep: captured with {len(self.static_in)} inputs, called with {len(inputs)}')
        for s, t in zip(self.static_in, inputs):
            if s.shape != t.shape or s.dtype != t.dtype:       # copy_ would broadcast / cast silently: a replay only fits the captured shapes
                raise RuntimeError(f'ich_b200.GraphedStep: captured for inputs of {tuple(s.shape)} {s.dtype}, called with {tuple(t.shape)} {t.dtype} '
                                   '(a partial last batch? use drop_last=True as the reference loaders do, or run that step eagerly)')
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out
