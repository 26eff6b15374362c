"""torch.autograd.Function wrappers over the C-ABI kernels. They own save-for-backward and workspaces (torch tensors);
the library never allocates.  Internal activations are channel-last [N, D, H, W, C] tensors (2-D nets: D = 1) in the
engine dtype (bf16 or fp32, ich_b200.config).  Call sites cite the reference lines they replace."""
import contextlib
import weakref

import torch
from torch.autograd import Function

from . import config
from . import _lib as _libmod
from ._lib import call

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
BN_SUM_COPIES = 16      # ICH_BN_SUM_COPIES of include/ich_b200.h: replicated partial sums of the BN-backward reduction


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'ich_b200.{what}: the hot path runs only on CUDA (sm_100a); there is no CPU fallback')


def _rows(t):
    """(ptr, channel pitch) of channel-last rows [N, D, H, W, C] (or [M, C]); the tensor may be a channel slab
    (last-dim slice) of a wider contiguous buffer."""
    if t.dim() == 2:
        return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else t.shape[1])
    n, d, h, w, c = t.shape
    st = t.stride()
    if c > 1 and st[4] != 1:
        raise RuntimeError('ich_b200: activation is not channel-last')
    ld, mult = None, 1
    for size, stride in ((w, st[3]), (h, st[2]), (d, st[1]), (n, st[0])):
        if size > 1:
            if ld is None:
                if stride % mult:
                    raise RuntimeError(f'ich_b200: irregular activation strides {st} for shape {tuple(t.shape)}')
                ld = stride // mult
            elif stride != ld * mult:
                raise RuntimeError(f'ich_b200: irregular activation strides {st} for shape {tuple(t.shape)}')
        mult *= size
    return t.data_ptr(), (c if ld is None else ld)


def _dt(t):
    return config.dtype_code(t.dtype)


def _alias(buf, shape, strides, offset=0):
    """Fresh tensor over `buf`'s storage (own autograd version counter): lets one Function write a channel slab of a buffer
    that another Function's saved tensors alias, without tripping autograd's in-place checks. The kernels, not torch, do
    the writes, and the slabs never overlap."""
    return torch.empty(0, dtype=buf.dtype, device=buf.device).set_(buf.untyped_storage(), buf.storage_offset() + offset, shape, strides)


# ---------------------------------------------------------------------------------------------------------------
# weight packs (derived caches, never serialised; invalidated when the optimizer mutates the parameter in place)
# ---------------------------------------------------------------------------------------------------------------
def _as5d(w):
    return w if w.dim() == 5 else w.unsqueeze(2)


# source dims: conv [co, ci, kd, kh, kw]; transposed conv [ci, co, i, j, l].  kind -> (perm, flipped source dims, bf16?, final shape(a, b, taps))
_PACK_SPEC = {
    'conv_fwd':        ((2, 3, 4, 1, 0), (),        False, lambda a, b, t: (t * b, a)),     # [taps*Cin][Cout] fp32
    'conv_dgrad':      ((2, 3, 4, 0, 1), (2, 3, 4), False, lambda a, b, t: (t * a, b)),     # [taps'*Cout][Cin] fp32, taps flipped
    'conv_fwd_tc':     ((2, 3, 4, 0, 1), (),        True,  lambda a, b, t: (t, a, b)),      # [taps][Cout][Cin] bf16
    'conv_dgrad_tc':   ((2, 3, 4, 1, 0), (2, 3, 4), True,  lambda a, b, t: (t, b, a)),      # [taps'][Cin][Cout] bf16
    'conv_fwd_tc_s':   ((3, 4, 2, 0, 1), (2,),      True,  lambda a, b, t: (t, a, b)),      # plane-streaming kernel: [kh][kw][2-kd][Cout][Cin] bf16
    'conv_dgrad_tc_s': ((3, 4, 2, 1, 0), (3, 4),    True,  lambda a, b, t: (t, b, a)),      # same for the data-gradient conv
    'convT_fwd_tc':    ((2, 3, 4, 1, 0), (),        True,  lambda a, b, t: (t * b, a)),     # [taps*Cout][Cin] bf16
    'convT_dgrad_tc':  ((0, 2, 3, 4, 1), (),        True,  lambda a, b, t: (a, t * b)),     # [Cin][taps*Cout] bf16
    'convT_fwd':       ((0, 2, 3, 4, 1), (),        False, lambda a, b, t: (a, t * b)),     # [Cin][taps*Cout] fp32
    'convT_dgrad':     ((2, 3, 4, 1, 0), (),        False, lambda a, b, t: (t * b, a)),     # [taps*Cout][Cin] fp32
}

# every parameter that has weight packs (weak): refresh_packs() re-derives the packs that went stale with an optimizer step in ONE launch
_PACKED_PARAMS = weakref.WeakSet()


def _pack_key(param):
    return (param._version, param.data_ptr(), param.device)


def _pack(param, kind):
    """Kernel-layout copy of a conv / transposed-conv weight (derived cache, never serialised).  The buffers are persistent per
    (parameter, kind); `fresh` is the set of kinds that match the parameter's current version."""
    cache = getattr(param, '_ich_packs', None)
    key = _pack_key(param)
    if cache is None or cache['key'][1:] != key[1:]:
        cache = {'key': key, 'bufs': {}, 'fresh': set()}
        param._ich_packs = cache
        _PACKED_PARAMS.add(param)
    elif cache['key'] != key:                      # the optimizer mutated the parameter in place
        cache['key'] = key
        cache['fresh'] = set()
    if kind not in cache['fresh']:
        if kind not in _PACK_SPEC:
            raise KeyError(kind)
        perm, flips, bf16, shape_of = _PACK_SPEC[kind]
        w = _as5d(param.detach())
        a, b = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] * w.shape[4]
        src = w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()
        dst = cache['bufs'].get(kind)
        if dst is None:
            dst = torch.empty(shape_of(a, b, taps), dtype=torch.bfloat16 if bf16 else torch.float32, device=w.device)
            cache['bufs'][kind] = dst
        call('ich_permute5', src.data_ptr(), dst.data_ptr(), 1 if bf16 else 0, *w.shape, *perm, sum(1 << f for f in flips), _stream())
        cache['fresh'].add(kind)
    return cache['bufs'][kind]


def invalidate_packs(module_or_params=None):
    """Drop the derived weight packs (and folded inference weights) of a module / an iterable of parameters / everything (None).
    The caches are keyed on the parameter's autograd version counter, which `load_state_dict` and torch optimizers bump; in-place
    writes through `param.data` (`p.data.copy_()`, EMA updates, manual weight surgery) do NOT bump it -- call this after such writes."""
    if module_or_params is None:
        params = list(_PACKED_PARAMS)
    elif hasattr(module_or_params, 'parameters'):
        params = list(module_or_params.parameters())
    else:
        params = list(module_or_params)
    for p in params:
        for attr in ('_ich_packs', '_ich_folded'):
            if hasattr(p, attr):
                delattr(p, attr)
        _PACKED_PARAMS.discard(p)


_REFRESH_ANY_DEVICE = False     # tests: let refresh_packs() batch CPU tensors too (the launch itself is mocked there)


def refresh_packs():
    """Every pack that was in use before the last in-place parameter update is re-derived now, all of them in one kernel launch
    (ich_permute5_batch) instead of one launch per layer and kind spread over forward and backward.  Called right after every
    optimizer step (global post-step hook below) and again at the start of a forward pass (normally a no-op then).  Returns the
    number of packs re-derived."""
    import ctypes
    jobs = []
    dev = None
    for param in list(_PACKED_PARAMS):
        cache = getattr(param, '_ich_packs', None)
        if cache is None or not (param.is_cuda or _REFRESH_ANY_DEVICE) or param.dtype != torch.float32 or not param.is_contiguous():
            continue
        key = _pack_key(param)
        if cache['key'][1:] != key[1:]:
            continue                                # storage moved (.to(), load_state_dict on a new tensor): _pack rebuilds lazily
        if cache['key'] != key:
            cache['key'] = key
            cache['fresh'] = set()
        stale = [k for k in cache['bufs'] if k not in cache['fresh']]
        if not stale:
            continue
        if dev is None:
            dev = param.device
        if param.device != dev:
            continue
        w = _as5d(param.detach())
        for kind in stale:
            perm, flips, bf16, _ = _PACK_SPEC[kind]
            jobs.append((w.data_ptr(), cache['bufs'][kind].data_ptr(), 1 if bf16 else 0, tuple(w.shape), perm, sum(1 << f for f in flips)))
            cache['fresh'].add(kind)
    if not jobs:
        return 0
    n = len(jobs)
    src = (ctypes.c_void_p * n)(*[j[0] for j in jobs])
    dst = (ctypes.c_void_p * n)(*[j[1] for j in jobs])
    dt = (ctypes.c_int * n)(*[j[2] for j in jobs])
    dims = (ctypes.c_int * (5 * n))(*[d for j in jobs for d in j[3]])
    perm = (ctypes.c_int * (5 * n))(*[q for j in jobs for q in j[4]])
    flip = (ctypes.c_int * n)(*[j[5] for j in jobs])
    with (torch.cuda.device(dev) if dev.type == 'cuda' else contextlib.nullcontext()):
        call('ich_permute5_batch', n, ctypes.cast(src, ctypes.c_void_p), ctypes.cast(dst, ctypes.c_void_p), ctypes.cast(dt, ctypes.c_void_p),
             ctypes.cast(dims, ctypes.c_void_p), ctypes.cast(perm, ctypes.c_void_p), ctypes.cast(flip, ctypes.c_void_p), _stream())
    return n


def _refresh_after_optimizer_step(optimizer, args, kwargs):
    """Global optimizer post-step hook: re-derive the weight packs right after the parameters changed, while the GPU is still busy with
    the step that was just launched.  Deriving them at the start of the next forward pass instead costs 0.3-0.6 ms of host work (version
    checks + job arrays for ~33 packs) that sits on the critical path whenever the trainer synchronises every step (`loss.item()`,
    models/optim/UNet2D.py:146).  The refresh at the start of the forward pass stays and finds nothing to do (~50 us)."""
    if config.get('refresh_after_step') and len(_PACKED_PARAMS):
        refresh_packs()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _POST_STEP_HOOK = _reg_post_hook(_refresh_after_optimizer_step)
except ImportError:          # older torch: the forward-pass refresh does all the work
    _POST_STEP_HOOK = None


def _ksize(weight):
    w = _as5d(weight)
    return w.shape[2], w.shape[3], w.shape[4]


def _tc_variant(x, cin, cout, k):
    """0: CUDA-core path; 1: slab tcgen05 kernel; 2: plane-streaming tcgen05 kernel (different weight pack)."""
    if not (config.get('tensor_cores') and x.dtype == torch.bfloat16):
        return 0
    from ._lib import lib
    n, d, h, w, _ = x.shape
    return int(lib().ich_conv_tc_variant(n, d, h, w, cin, cout, *k))


def _use_tc(x, cin, cout, k):
    return _tc_variant(x, cin, cout, k) != 0


def _cin1_tc(x, cin, cout, k):
    """First layer (Cin = 1) on the tcgen05 im2col kernels (bf16 engine only)."""
    if cin != 1 or k[1] != 3 or k[2] != 3 or not (config.get('tensor_cores') and x.dtype == torch.bfloat16):
        return False
    from ._lib import lib
    n, d, h, w, _ = x.shape
    return bool(lib().ich_conv_cin1_tc_supported(n, d, h, w, cout, k[0]))


# SyncBN plumbing: (world_size, all_reduce_sum(tensor) -> None in place).  None = resolve from torch.distributed at call time.
# Tests install a fake pair to emulate several ranks on one GPU.
SYNC_BN_COMM = None


def _sync_bn_comm():
    """(world, allreduce) when SyncBN is on and there is more than one rank, else None."""
    if not config.get('sync_bn'):
        return None
    if SYNC_BN_COMM is not None:
        return SYNC_BN_COMM if SYNC_BN_COMM[0] > 1 else None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.get_world_size(), (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))


class _Timed:
    """Tags the conv launches made inside the block with their direction ('fwd' / 'dgrad' / 'wgrad') for the per-launch profile that
    bench.py collects through `_lib.PROFILE` (no effect otherwise)."""

    def __init__(self, kind, flops=0.0):
        self.kind = kind

    def __enter__(self):
        _libmod.TAG = self.kind
        return self

    def __exit__(self, *a):
        _libmod.TAG = None
        return False


def conv_forward(x, weight, bias, relu=False):
    """y = conv(x) (+ bias) on channel-last rows; picks the tcgen05 kernel when eligible, else the FFMA kernel."""
    n, d, h, w, cin = x.shape
    cout = weight.shape[0]
    k = _ksize(weight)
    with _Timed('fwd', 2.0 * n * d * h * w * cin * cout * k[0] * k[1] * k[2]):
        return _conv_forward(x, weight, bias, relu)


def _conv_forward(x, weight, bias, relu):
    n, d, h, w, cin = x.shape
    cout = weight.shape[0]
    k = _ksize(weight)
    y = torch.empty((n, d, h, w, cout), dtype=x.dtype, device=x.device)
    xp, xld = _rows(x)
    var = _tc_variant(x, cin, cout, k)
    if var:
        call('ich_conv_tc_fwd', xp, xld, _p(_pack(weight, 'conv_fwd_tc_s' if var == 2 else 'conv_fwd_tc')), _p(bias), y.data_ptr(), cout,
             n, d, h, w, cin, cout, *k, int(relu), _stream())
    elif _cin1_tc(x, cin, cout, k):
        call('ich_conv_cin1_tc_fwd', xp, xld, _p(_pack(weight, 'conv_fwd')), _p(bias), y.data_ptr(), cout, None, None, n, d, h, w, cout, k[0],
             int(relu), _stream())
    else:
        call('ich_conv_fwd', xp, xld, _p(_pack(weight, 'conv_fwd')), _p(bias), y.data_ptr(), cout, _dt(x), n, d, h, w, cin, cout, *k,
             int(relu), _stream())
    return y


def conv_dgrad(dy, weight, want_colsum=False):
    """want_colsum: also return the fp64 per-channel sums of the stored gradient when the tcgen05 kernel can take them in its epilogue
    (the transposed conv that produced the tensor needs them as its bias gradient); attached to the result as `_ich_colsum`."""
    n, d, h, w, cout = dy.shape
    cin = weight.shape[1]
    k = _ksize(weight)
    with _Timed('dgrad', 2.0 * n * d * h * w * cin * cout * k[0] * k[1] * k[2]):
        return _conv_dgrad(dy, weight, want_colsum)


def _conv_dgrad(dy, weight, want_colsum=False):
    n, d, h, w, cout = dy.shape
    cin = weight.shape[1]
    k = _ksize(weight)
    dx = torch.empty((n, d, h, w, cin), dtype=dy.dtype, device=dy.device)
    yp, yld = _rows(dy)
    var = _tc_variant(dy, cout, cin, k)
    if var and want_colsum and k[1] == 3:
        sums = torch.empty((2, cin), dtype=torch.float64, device=dy.device)
        call('ich_conv_tc_fwd_stats', yp, yld, _p(_pack(weight, 'conv_dgrad_tc_s' if var == 2 else 'conv_dgrad_tc')), dx.data_ptr(), cin,
             sums[0].data_ptr(), sums[1].data_ptr(), n, d, h, w, cout, cin, *k, _stream())
        dx._ich_colsum = sums[0]
    elif var:
        call('ich_conv_tc_fwd', yp, yld, _p(_pack(weight, 'conv_dgrad_tc_s' if var == 2 else 'conv_dgrad_tc')), None, dx.data_ptr(), cin,
             n, d, h, w, cout, cin, *k, 0, _stream())
    else:
        call('ich_conv_fwd', yp, yld, _p(_pack(weight, 'conv_dgrad')), None, dx.data_ptr(), cin, _dt(dy), n, d, h, w, cout, cin, *k, 0, _stream())
    return dx


def conv_wgrad(x, dy, weight):
    n, d, h, w, cin = x.shape
    cout = weight.shape[0]
    k = _ksize(weight)
    with _Timed('wgrad', 2.0 * n * d * h * w * cin * cout * k[0] * k[1] * k[2]):
        return _conv_wgrad(x, dy, weight)


def _conv_wgrad(x, dy, weight):
    n, d, h, w, cin = x.shape
    cout = weight.shape[0]
    k = _ksize(weight)
    dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
    xp, xld = _rows(x)
    yp, yld = _rows(dy)
    from ._lib import lib
    if config.get('tensor_cores') and x.dtype == torch.bfloat16 and lib().ich_conv_tc_wgrad_supported(n, d, h, w, cin, cout, *k):
        call('ich_conv_tc_wgrad', xp, xld, yp, yld, dw.data_ptr(), n, d, h, w, cin, cout, *k, _stream())
    elif _cin1_tc(x, cin, cout, k) and yld % 8 == 0:
        call('ich_conv_cin1_tc_wgrad', xp, xld, yp, yld, dw.data_ptr(), n, d, h, w, cout, k[0], _stream())
    else:
        call('ich_conv_wgrad', xp, xld, yp, yld, _dt(x), dw.data_ptr(), n, d, h, w, cin, cout, *k, _stream())
    return dw


def col_sum(x):
    """fp64 per-channel sum over all rows of a channel-last tensor."""
    c = x.shape[-1]
    s = torch.empty(c, dtype=torch.float64, device=x.device)
    xp, xld = _rows(x)
    call('ich_colstats', xp, xld, _dt(x), x.numel() // c, c, s.data_ptr(), None, _stream())
    return s


# ---------------------------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------------------------
class ToChannelsLast(Function):
    """NC(D)HW fp32 (what the trainers feed, models/optim/UNet2D.py:137) -> engine N(D)HWC."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, 'ToChannelsLast')
        x = x.contiguous().float()
        ctx.was_4d = x.dim() == 4
        if ctx.was_4d:
            x = x.unsqueeze(2)
        n, c, d, h, w = x.shape
        out = torch.empty((n, d, h, w, c), dtype=config.act_dtype(), device=x.device)
        call('ich_layout_nc_to_nl', x.data_ptr(), out.data_ptr(), _dt(out), n, c, d * h * w, c, _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        if g is None or not config.get('input_grad'):
            return None
        gx = FromChannelsLast.forward(None, g.contiguous())
        return gx.squeeze(2) if ctx.was_4d else gx


class FromChannelsLast(Function):
    """engine N(D)HWC -> NC(D)HW fp32 contiguous (callers .view() the result)."""

    @staticmethod
    def forward(ctx, x):
        n, d, h, w, c = x.shape
        out = torch.empty((n, c, d, h, w), dtype=torch.float32, device=x.device)
        xp, xld = _rows(x)
        call('ich_layout_nl_to_nc', xp, _dt(x), xld, out.data_ptr(), n, c, d * h * w, _stream())
        if ctx is not None:
            ctx.dt = x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        n, c, d, h, w = g.shape
        out = torch.empty((n, d, h, w, c), dtype=ctx.dt, device=g.device)
        call('ich_layout_nc_to_nl', g.data_ptr(), out.data_ptr(), config.dtype_code(ctx.dt), n, c, d * h * w, c, _stream())
        return out


def to_channels_last(x):
    refresh_packs()                         # start of a forward pass: re-derive the weight packs the optimizer step invalidated
    if getattr(x, '_ich_staged', False):    # already [N, D, H, W, C] in the engine dtype (ops.stage_ct / ops.staged)
        return x
    return ToChannelsLast.apply(x)


def from_channels_last(x, was_4d=False):
    out = FromChannelsLast.apply(x)
    return out.squeeze(2) if was_4d else out


def _conv_bn_stats(x, weight, bias, gamma, beta, running_mean, running_var, training):
    """Conv (bias skipped in training mode: BatchNorm cancels it) with the batch statistics taken in the conv epilogue where the
    tcgen05 kernels apply, then bn_finalize.  Returns (y [pre-BN conv output], stats [scale, shift, mean, invstd], synced)."""
    n, d, h, w, cin = x.shape
    cout = weight.shape[0]
    m = n * d * h * w
    dev = x.device
    # training: BatchNorm cancels the conv bias, so the kernel skips it and bn_finalize folds it into running_mean
    stats = torch.empty((4, cout), dtype=torch.float32, device=dev)      # scale, shift, mean, invstd
    sums = torch.empty((2, cout), dtype=torch.float64, device=dev)
    k = _ksize(weight)
    var = _tc_variant(x, cin, cout, k) if (training and k[1] == 3) else 0
    if var:
        # tcgen05 conv with the batch statistics accumulated in its epilogue (no separate pass over y)
        y = torch.empty((n, d, h, w, cout), dtype=x.dtype, device=dev)
        xp, xld = _rows(x)
        with _Timed('fwd', 2.0 * m * cin * cout * k[0] * k[1] * k[2]):
            call('ich_conv_tc_fwd_stats', xp, xld, _p(_pack(weight, 'conv_fwd_tc_s' if var == 2 else 'conv_fwd_tc')), y.data_ptr(), cout, sums[0].data_ptr(),
                 sums[1].data_ptr(), n, d, h, w, cin, cout, *k, _stream())
    elif training and _cin1_tc(x, cin, cout, k):
        # first layer: tcgen05 im2col kernel, batch statistics fused the same way
        y = torch.empty((n, d, h, w, cout), dtype=x.dtype, device=dev)
        xp, xld = _rows(x)
        with _Timed('fwd', 2.0 * m * cin * cout * k[0] * k[1] * k[2]):
            call('ich_conv_cin1_tc_fwd', xp, xld, _p(_pack(weight, 'conv_fwd')), None, y.data_ptr(), cout, sums[0].data_ptr(), sums[1].data_ptr(),
                 n, d, h, w, cout, k[0], 0, _stream())
    else:
        y = conv_forward(x, weight, None)
        if training:
            call('ich_colstats', y.data_ptr(), cout, _dt(y), m, cout, sums[0].data_ptr(), sums[1].data_ptr(), _stream())
    comm = _sync_bn_comm() if training else None
    m_stat = m
    if comm is not None:
        # SyncBN: per-channel sum / sum of squares over ALL ranks (equal local batch sizes, as under the DP sampler)
        comm[1](sums)
        m_stat = m * comm[0]
    call('ich_bn_finalize', sums[0].data_ptr(), sums[1].data_ptr(), m_stat, cout, _p(gamma), _p(beta), _p(bias), _p(running_mean),
         _p(running_var), BN_MOMENTUM, BN_EPS, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
         int(training), _stream())
    return y, stats, comm is not None


def folded_eval_unit(weight, bias, gamma, beta, running_mean, running_var):
    """Inference only (eval mode, no autograd): BatchNorm folded into the conv, so that Conv -> BN -> ReLU (reference UNet.py:173-174)
    is ONE conv launch with a bias + ReLU epilogue -- no bn_finalize, no BN-apply pass, no pre-BN tensor.
    w' = w * scale[co], b' = beta + (conv_bias - running_mean) * scale with scale = gamma / sqrt(running_var + eps).
    Derived cache on the weight tensor, keyed on the versions / storage of everything it depends on."""
    deps = (weight, bias, gamma, beta, running_mean, running_var)
    key = tuple((t._version, t.data_ptr()) for t in deps if t is not None)
    cache = getattr(weight, '_ich_folded', None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            scale = torch.rsqrt(running_var.float() + BN_EPS)
            if gamma is not None:
                scale = scale * gamma.float()
            shift = -running_mean.float() * scale
            if bias is not None:
                shift = shift + bias.float() * scale
            if beta is not None:
                shift = shift + beta.float()
            w_eff = (weight.detach().float() * scale.view(-1, *([1] * (weight.dim() - 1)))).contiguous()
        cache = (key, w_eff, shift.contiguous())
        weight._ich_folded = cache
    return cache[1], cache[2]


# ---------------------------------------------------------------------------------------------------------------
# Conv -> BatchNorm -> ReLU  (one unit of ConvBlock.forward, models/networks/UNet.py:173-174)
# ---------------------------------------------------------------------------------------------------------------
class ConvBnRelu(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, relu, concat_c=0, drop_p=0.0):
        """drop_p > 0: nn.Dropout(p) applied to the unit's output (reference UNet.py:175-176) fused into the BN-apply kernel; the
        mask is a function of (seed, voxel, channel) and is regenerated in backward.  The seed is drawn from torch's default CPU
        generator, so torch.manual_seed makes runs reproducible (UNet2D_scripts.py:53-60)."""
        n, d, h, w, _ = x.shape
        cout = weight.shape[0]
        m = n * d * h * w
        dev = x.device
        y, stats, synced = _conv_bn_stats(x, weight, bias, gamma, beta, running_mean, running_var, training)
        if concat_c:
            # zero-copy skip connection: z is written as the first channel slab of the [.., cout + concat_c] buffer that the
            # decoder's ConvTranspose later completes (torch.cat([res, up], 1) of reference UNet.py:119 without the copy)
            ctot = cout + concat_c
            buf = torch.empty((n, d, h, w, ctot), dtype=y.dtype, device=dev)
            z = _alias(buf, (n, d, h, w, cout), (d * h * w * ctot, h * w * ctot, w * ctot, ctot, 1))
            zld = ctot
        else:
            buf = None
            z = torch.empty_like(y)
            zld = cout
        drop_p = float(drop_p)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0
        if drop_p > 0.0:
            call('ich_affine_act_drop', y.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), z.data_ptr(), zld, _dt(y), m, cout,
                 int(relu), drop_p, seed, _stream())
        else:
            call('ich_affine_act', y.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), z.data_ptr(), zld, _dt(y), m, cout, int(relu),
                 _stream())
        ctx.save_for_backward(x, weight, y, stats)
        ctx.training, ctx.relu, ctx.drop = training, relu, (drop_p, seed)
        ctx.sync = synced
        ctx.colsum = bool(getattr(x, '_ich_upcat', False))      # the input is UpConvCat's buffer: its backward wants the column sums of dx
        if concat_c:
            ctx.mark_non_differentiable(buf)
            return z, buf
        return z

    @staticmethod
    def backward(ctx, dz, *unused):
        x, weight, y, stats = ctx.saved_tensors
        n, d, h, w, cout = y.shape
        m = n * d * h * w
        try:
            dzp, dzld = _rows(dz)
        except RuntimeError:
            dz = dz.contiguous()
            dzp, dzld = dz.data_ptr(), cout
        dy = torch.empty_like(y)
        sums = torch.empty((BN_SUM_COPIES * 2, cout), dtype=torch.float64, device=y.device)
        dgamma = torch.empty(cout, dtype=torch.float32, device=y.device)
        dbeta = torch.empty(cout, dtype=torch.float32, device=y.device)
        comm = _sync_bn_comm() if (ctx.training and ctx.sync) else None
        if comm is not None:
            # SyncBN backward: reduction pass -> per-rank d(gamma), d(beta) -> all-reduce of the partial sums -> apply pass
            args = (dzp, dzld, y.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
                    sums.data_ptr(), dy.data_ptr(), cout, None, None, _dt(y), m, cout, int(ctx.relu), 1, ctx.drop[0], ctx.drop[1])
            call('ich_bn_act_bwd_sync', *args, 1, m, _stream())
            local = sums.view(BN_SUM_COPIES, 2, cout).sum(0)
            dbeta.copy_(local[0])
            dgamma.copy_(local[1])
            comm[1](sums)
            call('ich_bn_act_bwd_sync', *args, 2, m * comm[0], _stream())
        elif ctx.drop[0] > 0.0:
            call('ich_bn_act_bwd_drop', dzp, dzld, y.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
                 stats[3].data_ptr(), sums.data_ptr(), dy.data_ptr(), cout, dgamma.data_ptr(), dbeta.data_ptr(), _dt(y), m, cout, int(ctx.relu),
                 int(ctx.training), ctx.drop[0], ctx.drop[1], _stream())
        else:
            call('ich_bn_act_bwd', dzp, dzld, y.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
                 stats[3].data_ptr(), sums.data_ptr(), dy.data_ptr(), cout, dgamma.data_ptr(), dbeta.data_ptr(), _dt(y), m, cout, int(ctx.relu),
                 int(ctx.training), _stream())
        need = ctx.needs_input_grad
        dx = conv_dgrad(dy, weight, want_colsum=ctx.colsum) if need[0] else None
        dw = conv_wgrad(x, dy, weight) if need[1] else None
        db = None
        if need[2]:
            # training: d(loss)/d(bias) is exactly 0 (BatchNorm removes the mean); eval: sum of dy
            db = torch.zeros(cout, dtype=torch.float32, device=y.device) if ctx.training else col_sum(dy).float()
        return dx, dw, db, (dgamma if need[3] else None), (dbeta if need[4] else None), None, None, None, None, None, None


class ConvBnReluHead(Function):
    """Last unit of the last decoder ConvBlock fused with the single-class head: Conv -> BN -> ReLU -> final 1x1 conv -> Sigmoid /
    Identity (reference UNet.py:173-174 then :122).  z = ReLU(BN(y)) feeds final_conv only, so it is never written: the head kernel
    reads the conv output y, and the backward rebuilds dz = d(logit) * w_head inside the two BatchNorm-backward passes.  Returns the
    network output, fp32 [N, 1, D, H, W]."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, head_weight, head_bias, act):
        n, d, h, w, _ = x.shape
        cout = weight.shape[0]
        y, stats, synced = _conv_bn_stats(x, weight, bias, gamma, beta, running_mean, running_var, training)
        if synced:
            raise RuntimeError('ich_b200.ConvBnReluHead: not available with SyncBN (the caller must take the unfused path)')
        out = torch.empty((n, 1, d, h, w), dtype=torch.float32, device=x.device)
        hw = head_weight.detach().reshape(cout).float().contiguous()
        call('ich_bn_head_fwd', y.data_ptr(), cout, _dt(y), stats[0].data_ptr(), stats[1].data_ptr(), hw.data_ptr(), _p(head_bias), out.data_ptr(),
             n * d * h * w, cout, 1, act, _stream())
        ctx.save_for_backward(x, weight, y, stats, hw, out)
        ctx.training, ctx.act, ctx.head_shape = training, act, head_weight.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, y, stats, hw, out = ctx.saved_tensors
        n, d, h, w, cout = y.shape
        m = n * d * h * w
        dev = y.device
        dout = dout.contiguous().float()
        dy = torch.empty_like(y)
        sums = torch.empty(BN_SUM_COPIES * (3 * cout + 1), dtype=torch.float64, device=dev)
        dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
        dhw = torch.empty(cout, dtype=torch.float32, device=dev)
        dhb = torch.empty(1, dtype=torch.float32, device=dev)
        call('ich_bn_head_bwd', y.data_ptr(), cout, _dt(y), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
             hw.data_ptr(), out.data_ptr(), dout.data_ptr(), sums.data_ptr(), dy.data_ptr(), cout, dgamma.data_ptr(), dbeta.data_ptr(),
             dhw.data_ptr(), dhb.data_ptr(), m, cout, 1, int(ctx.training), ctx.act, _stream())
        need = ctx.needs_input_grad
        dx = conv_dgrad(dy, weight) if need[0] else None
        dw = conv_wgrad(x, dy, weight) if need[1] else None
        db = None
        if need[2]:
            db = torch.zeros(cout, dtype=torch.float32, device=dev) if ctx.training else col_sum(dy).float()
        return (dx, dw, db, (dgamma if need[3] else None), (dbeta if need[4] else None), None, None, None,
                (dhw.view(ctx.head_shape) if need[8] else None), (dhb if need[9] else None), None)


def head_fusable(x_dtype, cout, head_weight, act, training):
    """Can the last ConvBlock unit (cout channels) run fused with this head (ConvBnReluHead)?"""
    if not config.get('fuse_head') or head_weight.shape[0] != 1 or head_weight[0, 0].numel() != 1 or act not in (0, 1):
        return False
    if training and _sync_bn_comm() is not None:
        return False
    from ._lib import lib
    return bool(lib().ich_bn_head_supported(config.dtype_code(x_dtype), cout))


class ConvBias(Function):
    """Plain conv + bias (+ReLU): the 1x1 ConvHead layers (models/networks/UNet.py:228,238-242)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        z = conv_forward(x, weight, bias, relu)
        ctx.save_for_backward(x, weight, z)
        ctx.relu = relu
        return z

    @staticmethod
    def backward(ctx, dz):
        x, weight, z = ctx.saved_tensors
        cout = z.shape[-1]
        m = z.numel() // cout
        dev = z.device
        dz = dz.contiguous()
        need = ctx.needs_input_grad
        if ctx.relu:
            # dy = dz * [z > 0]; reuse the BN-backward kernel in eval form with scale = 1, shift = 0 (it also yields db)
            stats = torch.zeros((4, cout), dtype=torch.float32, device=dev)
            stats[0].fill_(1.0)
            stats[3].fill_(1.0)
            sums = torch.empty((BN_SUM_COPIES * 2, cout), dtype=torch.float64, device=dev)
            dy = torch.empty_like(z)
            db = torch.empty(cout, dtype=torch.float32, device=dev)
            call('ich_bn_act_bwd', dz.data_ptr(), cout, z.data_ptr(), cout, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
                 stats[3].data_ptr(), sums.data_ptr(), dy.data_ptr(), cout, None, db.data_ptr(), _dt(z), m, cout, 1, 0, _stream())
        else:
            dy = dz
            db = col_sum(dy).float() if need[2] else None
        dx = conv_dgrad(dy, weight) if need[0] else None
        dw = conv_wgrad(x, dy, weight) if need[1] else None
        return dx, dw, (db if need[2] else None), None


# ---------------------------------------------------------------------------------------------------------------
# MaxPool 2x (models/networks/UNet.py:82,109)
# ---------------------------------------------------------------------------------------------------------------
class MaxPool2(Function):
    @staticmethod
    def forward(ctx, x, fd):
        n, d, h, w, c = x.shape
        y = torch.empty((n, d // fd, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        xp, xld = _rows(x)
        call('ich_maxpool2_fwd', xp, xld, y.data_ptr(), c, _dt(x), n, d, h, w, c, fd, _stream())
        ctx.save_for_backward(x)
        ctx.fd = fd
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, d, h, w, c = x.shape
        dy = dy.contiguous()
        odd = d % ctx.fd or h % 2 or w % 2          # floor pooling: the un-pooled tail receives no gradient
        dx = (torch.zeros if odd else torch.empty)((n, d, h, w, c), dtype=x.dtype, device=x.device)
        xp, xld = _rows(x)
        call('ich_maxpool2_bwd', xp, xld, dy.data_ptr(), c, dx.data_ptr(), c, _dt(x), n, d, h, w, c, ctx.fd, None, 0, _stream())
        return dx, None


class PoolSkip(Function):
    """Encoder step of UNet.forward (UNet.py:107-109): the block output is kept as the skip tensor AND max-pooled. One Function
    with two outputs, so that the two gradients of the same tensor (through the decoder concat and through the pooled path)
    are combined inside the max-pool backward kernel instead of a separate autograd add.
    concat_c > 0: the pool kernel also writes the tensor into the first channel slab of a [.., C + concat_c] buffer -- the decoder's
    torch.cat([res, up], 1) buffer (UNet.py:119), which UpConvCat then completes in place (no separate copy pass)."""

    @staticmethod
    def forward(ctx, x, fd, concat_c=0):
        n, d, h, w, c = x.shape
        y = torch.empty((n, d // fd, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        xp, xld = _rows(x)
        ctx.save_for_backward(x)
        ctx.fd = fd
        ctx.set_materialize_grads(False)      # an unused skip output (encoder-only nets) arrives as None in backward, not as a zero tensor
        if concat_c:
            ctot = c + concat_c
            buf = torch.empty((n, d, h, w, ctot), dtype=x.dtype, device=x.device)
            call('ich_maxpool2_fwd_skip', xp, xld, y.data_ptr(), c, buf.data_ptr(), ctot, _dt(x), n, d, h, w, c, fd, _stream())
            skip = _alias(buf, (n, d, h, w, c), (d * h * w * ctot, h * w * ctot, w * ctot, ctot, 1))
            ctx.mark_non_differentiable(buf)
            return skip, y, buf
        call('ich_maxpool2_fwd', xp, xld, y.data_ptr(), c, _dt(x), n, d, h, w, c, fd, _stream())
        skip = _alias(x, tuple(x.shape), x.stride())     # same storage, fresh tensor object
        return skip, y

    @staticmethod
    def backward(ctx, dskip, dy, *unused):
        (x,) = ctx.saved_tensors
        n, d, h, w, c = x.shape
        if dy is None:
            return (dskip.contiguous() if dskip is not None else None), None, None
        dy = dy.contiguous()
        odd = d % ctx.fd or h % 2 or w % 2          # floor pooling (encoder-only nets): the un-pooled tail receives no pooled gradient
        if odd and dskip is not None:
            raise RuntimeError('ich_b200.PoolSkip: a skip connection needs spatial sizes divisible by the pooling factor')
        dx = (torch.zeros if odd else torch.empty)((n, d, h, w, c), dtype=x.dtype, device=x.device)
        xp, xld = _rows(x)
        sp, sld = (None, 0)
        if dskip is not None:
            try:
                sp, sld = _rows(dskip)
            except RuntimeError:
                dskip = dskip.contiguous()
                sp, sld = dskip.data_ptr(), c
        call('ich_maxpool2_bwd', xp, xld, dy.data_ptr(), c, dx.data_ptr(), c, _dt(x), n, d, h, w, c, ctx.fd, sp, sld, _stream())
        return dx, None, None


# ---------------------------------------------------------------------------------------------------------------
# ConvTranspose k2 s2 + torch.cat([res, up], 1) (models/networks/UNet.py:117-119): the up-sampled tensor is written
# straight into its channel slab of the concat buffer.
# ---------------------------------------------------------------------------------------------------------------
class UpConvCat(Function):
    @staticmethod
    def forward(ctx, x, res, weight, bias, fd, buf=None):
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        cres = res.shape[-1]
        ctot = cres + cout
        oshape = (n, d * fd, h * 2, w * 2, ctot)
        if tuple(res.shape[:4]) != oshape[:4]:
            raise RuntimeError(f'ich_b200: skip tensor {tuple(res.shape)} does not match the up-sampled grid {oshape}')
        m_out = oshape[0] * oshape[1] * oshape[2] * oshape[3]
        if buf is not None and tuple(buf.shape) == oshape and buf.dtype == x.dtype and buf.is_contiguous() and res.data_ptr() == buf.data_ptr():
            out = _alias(buf, oshape, buf.stride())      # the skip half is already in place (ConvBnRelu wrote it there)
        else:
            out = torch.empty(oshape, dtype=x.dtype, device=x.device)
            rp, rld = _rows(res)
            call('ich_slab_copy', rp, rld, out.data_ptr(), ctot, _dt(x), m_out, cres, _stream())
        up = out[..., cres:]
        xp, xld = _rows(x)
        from ._lib import lib
        use_tc = bool(config.get('tensor_cores') and x.dtype == torch.bfloat16 and cres % 8 == 0 and
                      lib().ich_convT2_tc_supported(n, d, h, w, cin, cout, fd) and
                      lib().ich_conv_tc_supported(n, d, h, w, 4 * fd * cout, cin, 1, 1, 1) and
                      lib().ich_convT2_tc_wgrad_supported(n, d, h, w, cin, cout, fd))
        if use_tc:
            call('ich_convT2_tc_fwd', xp, xld, _p(_pack(weight, 'convT_fwd_tc')), _p(bias), up.data_ptr(), ctot, n, d, h, w, cin, cout, fd,
                 _stream())
        else:
            call('ich_convT2_fwd', xp, xld, _p(_pack(weight, 'convT_fwd')), _p(bias), up.data_ptr(), ctot, _dt(x), n, d, h, w, cin, cout, fd,
                 _stream())
        ctx.save_for_backward(x, weight)
        ctx.fd, ctx.cres, ctx.use_tc = fd, cres, use_tc
        # the consuming conv's data-gradient epilogue then sums the bias gradient on the way (ICH_B200_DGRAD_COLSUM=0: a separate pass)
        out._ich_upcat = bool(use_tc and bias is not None and config.get('dgrad_colsum'))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        cres, fd = ctx.cres, ctx.fd
        dout = dout.contiguous()
        ctot = dout.shape[-1]
        m_out = dout.numel() // ctot
        need = ctx.needs_input_grad
        dres = None
        if need[1]:
            dres = dout[..., :cres]                      # zero-copy: consumers take the channel pitch explicitly
        dup = dout[..., cres:]
        dx = dw = db = None
        colsum = getattr(dout, '_ich_colsum', None)           # per-channel sums of dout from the epilogue of the kernel that produced it
        if need[3] and colsum is not None and colsum.shape[0] == ctot:
            db = colsum[cres:].float()
            need = (need[0], need[1], need[2], False, need[4])
        from ._lib import lib
        if ctx.use_tc and (need[0] or need[2]) and config.get('convt_direct') and \
                lib().ich_convT2_tc_dgrad_supported(n, d, h, w, cin, cout, fd) and lib().ich_convT2_tc_wgrad_direct_supported(n, d, h, w, cin, cout, fd):
            # both gradients are 1x1 GEMMs over [voxel][tap*Cout + co]; their operand chunks are read IN PLACE from the fine-grid gradient
            # through one strided tensor map per tap (no re-packed copy); the bias gradient is the column-sum pass at the end
            if need[0]:
                dx = torch.empty_like(x)
                call('ich_convT2_tc_dgrad', dup.data_ptr(), ctot, _p(_pack(weight, 'convT_dgrad_tc')), dx.data_ptr(), cin, n, d, h, w, cin, cout, fd,
                     _stream())
            if need[2]:
                dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
                xp, xld = _rows(x)
                call('ich_convT2_tc_wgrad_direct', xp, xld, dup.data_ptr(), ctot, dw.data_ptr(), n, d, h, w, cin, cout, fd, _stream())
            need = (False, need[1], False, need[3], need[4])
        elif ctx.use_tc and (need[0] or need[2]):
            # re-pack the up-sampled gradient to the coarse grid once ([voxel][tap*Cout + co]); both gradients are then 1x1 GEMMs
            taps = 4 * fd
            g = torch.empty((n, d, h, w, taps * cout), dtype=dout.dtype, device=dout.device)
            if need[3] and cout <= 256 and 256 % max(1, cout // 8) == 0 and cout % 8 == 0:
                # the bias gradient (sum of the up-sampled gradient over voxels) rides along with the re-pack
                bsum = torch.empty(cout, dtype=torch.float64, device=x.device)
                call('ich_space_to_depth2_sum', dup.data_ptr(), ctot, g.data_ptr(), _dt(dout), n, d, h, w, cout, fd, bsum.data_ptr(), _stream())
                db = bsum.float()
            else:
                call('ich_space_to_depth2', dup.data_ptr(), ctot, g.data_ptr(), _dt(dout), n, d, h, w, cout, fd, _stream())
            if need[0]:
                dx = torch.empty_like(x)
                call('ich_conv_tc_fwd', g.data_ptr(), taps * cout, _p(_pack(weight, 'convT_dgrad_tc')), None, dx.data_ptr(), cin, n, d, h, w,
                     taps * cout, cin, 1, 1, 1, 0, _stream())
            if need[2]:
                dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
                xp, xld = _rows(x)
                call('ich_convT2_tc_wgrad', xp, xld, g.data_ptr(), taps * cout, dw.data_ptr(), n, d, h, w, cin, cout, fd, _stream())
            need = (False, need[1], False, need[3] and db is None, need[4])
        if need[0]:
            dx = torch.empty_like(x)
            call('ich_convT2_dgrad', dup.data_ptr(), ctot, _p(_pack(weight, 'convT_dgrad')), dx.data_ptr(), cin, _dt(x), n, d, h, w, cin, cout,
                 fd, _stream())
        if need[2]:
            dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
            xp, xld = _rows(x)
            call('ich_convT2_wgrad', xp, xld, dup.data_ptr(), ctot, _dt(x), dw.data_ptr(), n, d, h, w, cin, cout, fd, _stream())
        if need[3]:
            s = torch.empty(cout, dtype=torch.float64, device=x.device)
            call('ich_colstats', dup.data_ptr(), ctot, _dt(dout), m_out, cout, s.data_ptr(), None, _stream())
            db = s.float()
        return dx, dres, dw, db, None, None


class UpsampleCat(Function):
    """bilinear=True decoder stage (models/networks/UNet.py:69-72,117-119): nn.Upsample(scale_factor=2, tri/bilinear,
    align_corners=True) written straight into its channel slab of torch.cat([res, up], 1)."""

    @staticmethod
    def forward(ctx, x, res, fd):
        n, d, h, w, c = x.shape
        cres = res.shape[-1]
        ctot = cres + c
        oshape = (n, d * fd, h * 2, w * 2, ctot)
        if tuple(res.shape[:4]) != oshape[:4]:
            raise RuntimeError(f'ich_b200: skip tensor {tuple(res.shape)} does not match the up-sampled grid {oshape}')
        out = torch.empty(oshape, dtype=x.dtype, device=x.device)
        m_out = oshape[0] * oshape[1] * oshape[2] * oshape[3]
        rp, rld = _rows(res)
        call('ich_slab_copy', rp, rld, out.data_ptr(), ctot, _dt(x), m_out, cres, _stream())
        up = out[..., cres:]
        xp, xld = _rows(x)
        call('ich_upsample2_fwd', xp, xld, up.data_ptr(), ctot, _dt(x), n, d, h, w, c, fd, _stream())
        ctx.geom = (n, d, h, w, c, cres, fd)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, d, h, w, c, cres, fd = ctx.geom
        dout = dout.contiguous()
        ctot = dout.shape[-1]
        need = ctx.needs_input_grad
        dx = None
        if need[0]:
            dx = torch.empty((n, d, h, w, c), dtype=dout.dtype, device=dout.device)
            call('ich_upsample2_bwd', dout[..., cres:].data_ptr(), ctot, dx.data_ptr(), c, _dt(dout), n, d, h, w, c, fd, _stream())
        return dx, (dout[..., :cres] if need[1] else None), None


# ---------------------------------------------------------------------------------------------------------------
# final 1x1 conv + Sigmoid / Softmax / Identity (models/networks/UNet.py:84-91,122) -> fp32 NC(D)HW
# ---------------------------------------------------------------------------------------------------------------
class Head(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        n, d, h, w, cin = x.shape
        cout = weight.shape[0]
        s = d * h * w
        out = torch.empty((n, cout, d, h, w), dtype=torch.float32, device=x.device)
        wf = weight.detach().reshape(cout, cin).float().contiguous()
        xp, xld = _rows(x)
        call('ich_head_fwd', xp, xld, _dt(x), wf.data_ptr(), _p(bias), out.data_ptr(), n, s, cin, cout, act, _stream())
        ctx.save_for_backward(x, weight, out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, out = ctx.saved_tensors
        n, d, h, w, cin = x.shape
        cout = weight.shape[0]
        s = d * h * w
        dout = dout.contiguous().float()
        need = ctx.needs_input_grad
        xp, xld = _rows(x)
        if cout == 1:
            wf = weight.detach().reshape(cin).float().contiguous()
            dx = torch.empty_like(x) if need[0] else None
            dw = torch.empty(cin, dtype=torch.float32, device=x.device)
            db = torch.empty(1, dtype=torch.float32, device=x.device)
            call('ich_head1_bwd', xp, xld, _dt(x), wf.data_ptr(), out.data_ptr(), dout.data_ptr(), _p(dx), cin, dw.data_ptr(), db.data_ptr(),
                 n * s, cin, ctx.act, _stream())
            return dx, (dw.view(weight.shape) if need[1] else None), (db if need[2] else None), None
        dl = torch.empty((n, d, h, w, cout), dtype=x.dtype, device=x.device)
        call('ich_head_dlogit', out.data_ptr(), dout.data_ptr(), dl.data_ptr(), _dt(x), n, s, cout, ctx.act, _stream())
        dx = conv_dgrad(dl, weight) if need[0] else None
        dw = conv_wgrad(x, dl, weight) if need[1] else None
        db = col_sum(dl).float() if need[2] else None
        return dx, dw, db, None


# ---------------------------------------------------------------------------------------------------------------
# AdaptiveAvgPool(1) (models/networks/UNet.py:295,318) -> fp32 [N, C]
# ---------------------------------------------------------------------------------------------------------------
class GlobalAvgPool(Function):
    @staticmethod
    def forward(ctx, x):
        n, d, h, w, c = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        xp, xld = _rows(x)
        call('ich_avgpool_fwd', xp, xld, _dt(x), out.data_ptr(), n, d * h * w, c, _stream())
        ctx.shape, ctx.dt = tuple(x.shape), x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        n, d, h, w, c = ctx.shape
        g = g.contiguous().float()
        dx = torch.empty(ctx.shape, dtype=ctx.dt, device=g.device)
        call('ich_avgpool_bwd', g.data_ptr(), dx.data_ptr(), c, config.dtype_code(ctx.dt), n, d * h * w, c, _stream())
        return dx


# ---------------------------------------------------------------------------------------------------------------
# MLPHead layers (models/networks/UNet.py:179-209): Linear (+ReLU) on the pooled [B, C] fp32 matrix
# ---------------------------------------------------------------------------------------------------------------
class Linear(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        _require_cuda(x, 'Linear')
        x = x.contiguous().float()
        w = weight.detach().contiguous().float()
        b, k = x.shape
        n = w.shape[0]
        out = torch.empty((b, n), dtype=torch.float32, device=x.device)
        call('ich_linear_fwd', x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), b, k, n, int(relu), _stream())
        ctx.save_for_backward(x, w, out)
        ctx.relu, ctx.has_bias = bool(relu), bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, out = ctx.saved_tensors
        b, k = x.shape
        n = w.shape[0]
        dout = dout.contiguous().float()
        need = ctx.needs_input_grad
        dx = torch.empty_like(x) if need[0] else None
        want_w = need[1] or (need[2] and ctx.has_bias)
        dw = torch.empty_like(w) if want_w else None
        db = torch.empty(n, dtype=torch.float32, device=x.device) if (want_w and ctx.has_bias) else None
        call('ich_linear_bwd', dout.data_ptr(), out.data_ptr(), x.data_ptr(), w.data_ptr(), _p(dx), _p(dw), _p(db), b, k, n, int(ctx.relu), _stream())
        return dx, (dw if need[1] else None), (db if need[2] else None), None


# ---------------------------------------------------------------------------------------------------------------
# GatedConv output (models/networks/GatedUNet.py:303-322): out = feat * sigmoid(gate)
# ---------------------------------------------------------------------------------------------------------------
class GateMul(Function):
    @staticmethod
    def forward(ctx, feat, gate):
        feat, gate = feat.contiguous(), gate.contiguous()
        out = torch.empty_like(feat)
        call('ich_gate_mul_fwd', feat.data_ptr(), gate.data_ptr(), out.data_ptr(), _dt(feat), feat.numel(), _stream())
        ctx.save_for_backward(feat, gate)
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, gate = ctx.saved_tensors
        dout = dout.contiguous()
        dfeat, dgate = torch.empty_like(feat), torch.empty_like(gate)
        call('ich_gate_mul_bwd', feat.data_ptr(), gate.data_ptr(), dout.data_ptr(), dfeat.data_ptr(), dgate.data_ptr(), _dt(feat), feat.numel(), _stream())
        return dfeat, dgate


# ---------------------------------------------------------------------------------------------------------------
# input staging + sliding-window plumbing
# ---------------------------------------------------------------------------------------------------------------
_SRC_DTYPE = {torch.float32: 0, torch.int16: 1, torch.uint16: 2, torch.uint8: 3}


def stage_ct(raw, win_center=40, win_width=120, out_range=(0, 1), dtype=None):
    """utils/ct_utils.py:13-36 (window / rescale / clip) + `.float()` + the cast to the engine dtype in one pass over a raw CT tensor on
    the device (int16 / uint16 / uint8 / fp32, any shape).  Returns a tensor of the same shape in `dtype` (default: the engine dtype).
    For a one-channel input [N, 1, D, H, W] the result viewed as [N, D, H, W, 1] IS the engine's channel-last layout (see `staged`)."""
    _require_cuda(raw, 'stage_ct')
    if raw.dtype not in _SRC_DTYPE:
        raw = raw.float()
    raw = raw.contiguous()
    dtype = dtype or config.act_dtype()
    out = torch.empty(raw.shape, dtype=dtype, device=raw.device)
    lo, hi = win_center - win_width / 2, win_center + win_width / 2
    call('ich_stage_ct', raw.data_ptr(), _SRC_DTYPE[raw.dtype], out.data_ptr(), config.dtype_code(dtype), raw.numel(), float(lo), float(hi),
         float(out_range[0]), float(out_range[1]), _stream())
    return out


def staged(x_cl, was_4d=False):
    """Mark an engine-layout tensor [N, D, H, W, C] (engine dtype) so that the drop-in networks take it as is instead of converting from
    NC(D)HW fp32 (ops.to_channels_last passes it through).  was_4d: the logical input is NCHW (2-D nets, D = 1), so the network
    output is squeezed back to 4-D."""
    if x_cl.dim() != 5 or x_cl.dtype != config.act_dtype():
        raise RuntimeError(f'ich_b200.staged: expected [N, D, H, W, C] in {config.act_dtype()}, got {tuple(x_cl.shape)} {x_cl.dtype}')
    x_cl._ich_staged = True
    x_cl._ich_was_4d = bool(was_4d)
    return x_cl


def window_gather(vol, starts, window):
    """vol [D, H, W] (one channel, any engine dtype) + starts int32 [n, 3] on the device -> windows [n, wd, wh, ww, 1]."""
    d, h, w = vol.shape
    n = starts.shape[0]
    out = torch.empty((n,) + tuple(window) + (1,), dtype=vol.dtype, device=vol.device)
    call('ich_window_gather', vol.data_ptr(), _dt(vol), d, h, w, starts.data_ptr(), n, *window, out.data_ptr(), _stream())
    return out


def window_scatter(pred, starts, window, shape, overlap, threshold, acc, cnt, mask):
    """Stitch window predictions fp32 [n, 1, wd, wh, ww] into the volume buffers (see ich_window_scatter)."""
    d, h, w = shape
    pred = pred.contiguous().float()
    call('ich_window_scatter', pred.data_ptr(), d, h, w, starts.data_ptr(), starts.shape[0], *window, int(overlap), float(threshold),
         _p(acc), _p(cnt), _p(mask), _stream())


def blend_threshold(acc, cnt, threshold, mask):
    call('ich_blend_threshold', acc.data_ptr(), cnt.data_ptr(), acc.numel(), float(threshold), _p(mask), _stream())


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
_RED = {'none': 0, 'mean': 1, 'sum': 2}


class SegLoss(Function):
    """BinaryDiceLoss / ComboLoss (models/optim/LossFunctions.py:39-63,143-166) in one reduction pass + one backward pass."""

    @staticmethod
    def forward(ctx, pred, mask, p, eps, alpha_empty, w_bce, w_dice, beta, reduction):
        _require_cuda(pred, 'SegLoss')
        pred_c = pred.contiguous().float()
        mask_c = mask.detach().contiguous().float()
        b = pred_c.shape[0]
        s = pred_c.numel() // b
        acc = torch.empty((b, 5), dtype=torch.float64, device=pred.device)
        per = torch.empty(b, dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        call('ich_seg_loss_fwd', pred_c.data_ptr(), mask_c.data_ptr(), b, s, float(p), float(eps), float(alpha_empty), float(w_bce),
             float(w_dice), float(beta), _RED[reduction], acc.data_ptr(), per.data_ptr(), loss.data_ptr(), _stream())
        ctx.save_for_backward(pred_c, mask_c, acc)
        ctx.args = (float(p), float(eps), float(alpha_empty), float(w_bce), float(w_dice), float(beta), reduction)
        return per if reduction == 'none' else loss

    @staticmethod
    def backward(ctx, g):
        pred, mask, acc = ctx.saved_tensors
        p, eps, alpha_empty, w_bce, w_dice, beta, reduction = ctx.args
        b = pred.shape[0]
        s = pred.numel() // b
        g = g.float()
        if reduction == 'none':
            gscale = g.contiguous()
        else:
            gscale = (g / b if reduction == 'mean' else g).expand(b).contiguous()
        dpred = torch.empty_like(pred)
        call('ich_seg_loss_bwd', pred.data_ptr(), mask.data_ptr(), acc.data_ptr(), gscale.data_ptr(), b, s, p, eps, alpha_empty, w_bce, w_dice,
             beta, dpred.data_ptr(), _stream())
        # the trainers also set mask.requires_grad_ (models/optim/UNet2D.py:138) but never read the result -> None
        return dpred, None, None, None, None, None, None, None, None


class TverskyLossFn(Function):
    """TverskyLoss (models/optim/LossFunctions.py:65-114): the Dice reduction pass (TP, sum p, sum m per sample in fp64) + a
    finalize; the gradient is affine in the mask, so the backward pass reads only the mask."""

    @staticmethod
    def forward(ctx, pred, mask, eps, alpha_empty, beta, gamma, reduction):
        _require_cuda(pred, 'TverskyLoss')
        pred_c = pred.contiguous().float()
        mask_c = mask.detach().contiguous().float()
        b = pred_c.shape[0]
        s = pred_c.numel() // b
        acc = torch.empty((b, 5), dtype=torch.float64, device=pred.device)
        per = torch.empty(b, dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        call('ich_tversky_loss_fwd', pred_c.data_ptr(), mask_c.data_ptr(), b, s, float(eps), float(alpha_empty), float(beta), float(gamma),
             _RED[reduction], acc.data_ptr(), per.data_ptr(), loss.data_ptr(), _stream())
        ctx.save_for_backward(mask_c, acc)
        ctx.args = (float(eps), float(alpha_empty), float(beta), float(gamma), reduction, pred.shape)
        return per if reduction == 'none' else loss

    @staticmethod
    def backward(ctx, g):
        mask, acc = ctx.saved_tensors
        eps, alpha_empty, beta, gamma, reduction, shape = ctx.args
        b = mask.shape[0]
        s = mask.numel() // b
        g = g.float()
        if reduction == 'none':
            gscale = g.contiguous()
        else:
            gscale = (g / b if reduction == 'mean' else g).expand(b).contiguous()
        dpred = torch.empty(shape, dtype=torch.float32, device=mask.device)
        call('ich_tversky_loss_bwd', mask.data_ptr(), acc.data_ptr(), gscale.data_ptr(), b, s, eps, alpha_empty, beta, gamma,
             dpred.data_ptr(), _stream())
        return dpred, None, None, None, None, None, None


class InfoNCE(Function):
    """mean_rows( logsumexp_{j != i} S_ij - S_{i, pos(i)} ), S = cos / tau, on P [B][2A][E] (closed form of
    models/optim/LossFunctions.py:208-230 and :328-339)."""

    @staticmethod
    def forward(ctx, P, tau):
        _require_cuda(P, 'InfoNCE')
        P = P.contiguous().float()
        b, r, e = P.shape
        dev = P.device
        Pn = torch.empty_like(P)
        aux = torch.empty((3, b * r), dtype=torch.float32, device=dev)     # invn, lse, rowloss
        loss = torch.empty((), dtype=torch.float32, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        call('ich_infonce_fwd', P.data_ptr(), b, r, e, float(tau), Pn.data_ptr(), aux[0].data_ptr(), aux[1].data_ptr(), aux[2].data_ptr(),
             loss.data_ptr(), counter.data_ptr(), _stream())
        ctx.save_for_backward(Pn, aux)
        ctx.tau = float(tau)
        return loss

    @staticmethod
    def backward(ctx, g):
        Pn, aux = ctx.saved_tensors
        b, r, e = Pn.shape
        g = g.contiguous().float()
        dP = torch.empty_like(Pn)
        call('ich_infonce_bwd', Pn.data_ptr(), aux[0].data_ptr(), aux[1].data_ptr(), b, r, e, ctx.tau, g.data_ptr(), dP.data_ptr(), _stream())
        return dP, None


# Cross-rank contrastive set: (world, rank, all_gather(t) -> [world * n, E]).  None = resolve from torch.distributed at call time.
GLOBAL_NCE_COMM = None


def _global_nce_comm():
    if not config.get('global_nce'):
        return None
    if GLOBAL_NCE_COMM is not None:
        return GLOBAL_NCE_COMM if GLOBAL_NCE_COMM[0] > 1 else None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None

    def gather(t):
        out = torch.empty((dist.get_world_size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out
    return dist.get_world_size(), dist.get_rank(), gather


class GatherRows(Function):
    """z [n, E] of this rank -> rows of all ranks [world * n, E] (rank-major), for a loss that every rank evaluates identically on
    the gathered set.  Backward keeps this rank's rows and multiplies by `world`: the data-parallel gradient all-reduce AVERAGES the
    parameter gradients over ranks, and only this rank back-propagates through its own rows."""

    @staticmethod
    def forward(ctx, z, world, rank, gather):
        ctx.world, ctx.rank, ctx.n = world, rank, z.shape[0]
        return gather(z.detach())

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n] * float(ctx.world), None, None, None


def gather_rows(z):
    """All ranks' rows when the global contrastive set is enabled (ICH_B200_GLOBAL_NCE=1 under torch.distributed), else z itself."""
    comm = _global_nce_comm()
    return z if comm is None else GatherRows.apply(z, *comm)


class RegionGather(Function):
    """f1, f2 [bs][H][W][C] + corners [bs][A][2] -> P [bs][2A][K*K*C] (models/optim/LossFunctions.py:321-328)."""

    @staticmethod
    def forward(ctx, f1, f2, corners, K):
        _require_cuda(f1, 'RegionGather')
        f1 = f1.contiguous().float()
        f2 = f2.contiguous().float()
        bs, H, W, C = f1.shape
        corners = corners.to(torch.int32).contiguous()
        A = corners.shape[1]
        P = torch.empty((bs, 2 * A, K * K * C), dtype=torch.float32, device=f1.device)
        for view, f in enumerate((f1, f2)):
            call('ich_region_gather', f.data_ptr(), corners.data_ptr(), P.data_ptr(), bs, H, W, C, A, K, view, _stream())
        ctx.save_for_backward(corners)
        ctx.geom = (bs, H, W, C, A, K)
        return P

    @staticmethod
    def backward(ctx, dP):
        (corners,) = ctx.saved_tensors
        bs, H, W, C, A, K = ctx.geom
        dP = dP.contiguous()
        outs = []
        for view in range(2):
            df = torch.empty((bs, H, W, C), dtype=torch.float32, device=dP.device)
            call('ich_region_scatter', dP.data_ptr(), corners.data_ptr(), df.data_ptr(), bs, H, W, C, A, K, view, _stream())
            outs.append(df)
        return outs[0], outs[1], None, None


def confusion_matrix(pred, target, threshold=None):
    """batch_binary_confusion_matrix (utils/tensor_utils.py:12-36), optionally fused with the >= threshold of UNet2D.py:220.
    Returns (tn, fp, fn, tp), each [B] fp32."""
    _require_cuda(pred, 'confusion_matrix')
    assert pred.shape == target.shape, f'Shapes do not match! {pred.shape} =/= {target.shape}'
    assert pred.ndim > 1, f'The tensor must have more that a single dimension. {pred.ndim} dimension passed.'
    p = pred.contiguous().float()
    t = target.contiguous().float()
    b = p.shape[0]
    out = torch.empty((b, 4), dtype=torch.float64, device=p.device)
    call('ich_confusion', p.data_ptr(), t.data_ptr(), b, p.numel() // b, -1.0 if threshold is None else float(threshold), out.data_ptr(),
         _stream())
    o = out.float()
    return o[:, 0], o[:, 1], o[:, 2], o[:, 3]
