"""Algorithmic work of one C-ABI launch: (family, FLOP, bytes) from the entry point's name and its arguments as declared in
include/ich_b200.h.  Used by bench.py to turn the per-launch CUDA-event times (`_lib.PROFILE`) into achieved TFLOP/s for the
tensor-bound kernels (2*M*N*K) and achieved GB/s for the HBM-bound ones (every operand
tensor touched once).  Measurement helper only -- nothing on the product path imports it."""


def _es(a):
    return 2 if a.get('dtype', 1) == 1 else 4


def _vox(a):
    return a['N'] * a['D'] * a['H'] * a['W']


def _conv(a, tag):
    cin = a.get('Cin', 1)
    taps = a.get('KD', 1) * a.get('KH', 3) * a.get('KW', 3)
    f = 2.0 * _vox(a) * cin * a['Cout'] * taps
    # DRAM bytes if every operand crosses once: input + output (data-gradient: the same two tensors the other way round; weight gradient:
    # both read) + the weights (bf16 pack read, or the fp32 gradient written)
    b = float(_vox(a) * (cin + a['Cout']) * _es(a) + taps * cin * a['Cout'] * (4 if tag == 'wgrad' else 2))
    return ('conv_' + (tag or 'fwd'), f, b)


def _convT(a, tag):
    taps = 4 * a['FD']          # `_vox` = coarse voxels; the fine tensor has `taps` voxels per coarse one
    return ('convT', 2.0 * _vox(a) * a['Cin'] * a['Cout'] * taps, float(_vox(a) * (a['Cin'] + taps * a['Cout']) * _es(a)))


def _bw(family, fn):
    return lambda a, tag: (family, None, float(fn(a)))


def _maxpool_bwd(a):
    e = _vox(a) * a['C'] * _es(a)
    return e * (2 + 1.0 / (4 * a['FD'])) + (e if a.get('dskip') else 0)


TABLE = {
    'ich_conv_tc_fwd': _conv, 'ich_conv_tc_fwd_stats': _conv, 'ich_conv_fwd': _conv, 'ich_conv_cin1_tc_fwd': _conv,
    'ich_conv_tc_wgrad': lambda a, t: _conv(a, 'wgrad'), 'ich_conv_wgrad': lambda a, t: _conv(a, 'wgrad'),
    'ich_conv_cin1_tc_wgrad': lambda a, t: _conv(a, 'wgrad'),
    'ich_convT2_tc_fwd': _convT, 'ich_convT2_fwd': _convT, 'ich_convT2_tc_dgrad': _convT, 'ich_convT2_tc_wgrad_direct': _convT, 'ich_convT2_tc_wgrad': _convT, 'ich_convT2_wgrad': _convT, 'ich_convT2_dgrad': _convT,
    'ich_affine_act': _bw('bn_apply', lambda a: 2 * a['M'] * a['C'] * _es(a)),
    'ich_affine_act_drop': _bw('bn_apply', lambda a: 2 * a['M'] * a['C'] * _es(a)),
    'ich_bn_act_bwd': _bw('bn_bwd', lambda a: 5 * a['M'] * a['C'] * _es(a)),
    'ich_bn_act_bwd_drop': _bw('bn_bwd', lambda a: 5 * a['M'] * a['C'] * _es(a)),
    'ich_bn_act_bwd_sync': _bw('bn_bwd', lambda a: (2 if a['phase'] == 1 else 3) * a['M'] * a['C'] * _es(a)),
    'ich_colstats': _bw('bn_stats', lambda a: a['M'] * a['C'] * _es(a)),
    'ich_maxpool2_fwd': _bw('pool', lambda a: _vox(a) * a['C'] * _es(a) * (1 + 1.0 / (4 * a['FD']))),
    'ich_maxpool2_fwd_skip': _bw('pool', lambda a: _vox(a) * a['C'] * _es(a) * (2 + 1.0 / (4 * a['FD']))),
    'ich_maxpool2_bwd': _bw('pool', _maxpool_bwd),
    'ich_slab_copy': _bw('concat_copy', lambda a: 2 * a['M'] * a['C'] * _es(a)),
    'ich_space_to_depth2': _bw('space_to_depth', lambda a: 2 * _vox(a) * 4 * a['FD'] * a['C'] * _es(a)),
    'ich_space_to_depth2_sum': _bw('space_to_depth', lambda a: 2 * _vox(a) * 4 * a['FD'] * a['C'] * _es(a)),
    'ich_layout_nc_to_nl': _bw('layout', lambda a: a['N'] * a['C'] * a['S'] * (4 + _es(a))),
    'ich_layout_nl_to_nc': _bw('layout', lambda a: a['N'] * a['C'] * a['S'] * (4 + _es(a))),
    'ich_stage_ct': _bw('layout', lambda a: a['M'] * ((4, 2, 2, 1)[a['src_dtype']] + _es(a))),
    'ich_head_fwd': _bw('head', lambda a: a['N'] * a['S'] * (a['Cin'] * _es(a) + 4 * a['Cout'])),
    'ich_head1_bwd': _bw('head', lambda a: a['M'] * (a['Cin'] * _es(a) * (2 if a.get('dx') else 1) + 8)),
    'ich_bn_head_fwd': _bw('head', lambda a: a['M'] * (a['C'] * _es(a) + 4)),
    'ich_bn_head_bwd': _bw('head', lambda a: a['M'] * (3 * a['C'] * _es(a) + 16)),
    'ich_seg_loss_fwd': _bw('loss', lambda a: a['B'] * a['S'] * 8), 'ich_seg_loss_bwd': _bw('loss', lambda a: a['B'] * a['S'] * 12),
    'ich_tversky_loss_fwd': _bw('loss', lambda a: a['B'] * a['S'] * 8), 'ich_tversky_loss_bwd': _bw('loss', lambda a: a['B'] * a['S'] * 8),
    'ich_confusion': _bw('loss', lambda a: a['B'] * a['S'] * 8),
    'ich_threshold_confusion': _bw('loss', lambda a: a['B'] * a['S'] * 9),
    'ich_avgpool_fwd': _bw('pool', lambda a: a['N'] * a['S'] * a['C'] * _es(a)),
    'ich_avgpool_bwd': _bw('pool', lambda a: a['N'] * a['S'] * a['C'] * _es(a)),
    'ich_upsample2_fwd': _bw('upsample', lambda a: _vox(a) * a['C'] * _es(a) * (1 + 4 * a['FD'])),
    'ich_upsample2_bwd': _bw('upsample', lambda a: _vox(a) * a['C'] * _es(a) * (1 + 4 * a['FD'])),
}

TENSOR_FAMILIES = ('conv_fwd', 'conv_dgrad', 'conv_wgrad', 'convT')


def algo_work(name, args, tag=None):
    """(family, flop or None, bytes or None) of one launch; unknown / latency-bound entry points -> ('small', None, None)."""
    fn = TABLE.get(name)
    if fn is None:
        return ('small', None, None)
    try:
        fam, f, b = fn(args, tag)
    except KeyError:
        return ('small', None, None)
    if name == 'ich_conv_tc_fwd' and args.get('KH') == 1 and tag is None:
        fam = 'convT'          # the transposed conv's data-gradient GEMM (ops.UpConvCat.backward) and the 1x1 heads
    return (fam, f, b)


def summarise(records, steps):
    """records = _lib.PROFILE entries (name, args, tag, e0, e1) -> {family: {ms_per_step, launches_per_step, tflops | gbs}}."""
    fam = {}
    for name, args, tag, e0, e1 in records:
        f, fl, by = algo_work(name, args, tag)
        a = fam.setdefault(f, {'ms': 0.0, 'n': 0, 'flop': 0.0, 'bytes': 0.0})
        a['ms'] += e0.elapsed_time(e1)
        a['n'] += 1
        a['flop'] += fl or 0.0
        a['bytes'] += by or 0.0
    out = {}
    for f, a in fam.items():
        o = {'ms_per_step': a['ms'] / steps, 'launches_per_step': a['n'] / steps, 'algo_bytes_per_step': a['bytes'] / steps}
        if a['flop'] and a['ms']:
            o['tflops'] = a['flop'] / (a['ms'] * 1e-3) / 1e12
        if a['bytes'] and a['ms']:
            o['gbs'] = a['bytes'] / (a['ms'] * 1e-3) / 1e9
        out[f] = o
    return out
