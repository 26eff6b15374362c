"""Engine knobs. The reference's JSON configs and scripts stay unchanged, so knobs come from the environment
 and can be overridden programmatically (tests, bench)."""
import os
import contextlib
import torch

_state = {
    # 'bf16': bf16 activations + tcgen05 tensor-core convs (fp32 accumulate); 'fp32': fp32 activations, FFMA convs
    'precision': os.environ.get('ICH_B200_PRECISION', 'bf16').lower(),
    # compute d(loss)/d(input) when the caller set input.requires_grad_ (models/optim/UNet2D.py:137). The reference
    # trainers never read it; default off saves the first layer's data-gradient.
    'input_grad': os.environ.get('ICH_B200_INPUT_GRAD', '0') == '1',
    # 1 = the max-pool kernel also lays the encoder skip tensor out inside the decoder's concat buffer (no copy pass for
    # torch.cat([res, up], 1)); measured 0.15-0.2 ms/step SLOWER at cfg-3 than the separate slab copy in the decoder -> off
    'zero_copy_concat': os.environ.get('ICH_B200_ZERO_COPY_CONCAT', '0') == '1',
    # SyncBN (optional): BatchNorm batch statistics and the BN-backward sums are all-reduced over the ranks, so
    # N GPUs x local batch behave exactly like one GPU with the global batch.  Default off = DistributedDataParallel semantics.
    'sync_bn': os.environ.get('ICH_B200_SYNC_BN', '0') == '1',
    # global contrastive set: InfoNCELoss compares against the embeddings of ALL ranks (all-gather) instead of the
    # local batch only.  Default off = the reference's per-process semantics (each rank an independent replica of the loss).
    'global_nce': os.environ.get('ICH_B200_GLOBAL_NCE', '0') == '1',
    # single-class heads (out_channels = 1): run the last ConvBlock unit fused with final_conv + Sigmoid (ops.ConvBnReluHead): the
    # BN+ReLU output of that unit is never materialised and its gradient is rebuilt inside the BatchNorm-backward passes
    # (measured at cfg-3: 15.57 -> 15.19 ms/step, end to end 16.25 -> 15.99; 0 = separate BN-apply / head kernels)
    'fuse_head': os.environ.get('ICH_B200_FUSE_HEAD', '1') == '1',
    # inference (eval mode under torch.no_grad): fold BatchNorm into the conv weights, one conv launch with a bias + ReLU epilogue per
    # unit instead of conv + bn_finalize + BN-apply (ops.folded_eval_unit).  Parity at cfg-5's real size with and without folding:
    # tests/test_parity_r2.py::test_cfg5_full_volume_sliding_window; measured 5.38 -> 4.56 ms per 32x512x512 volume.  0 = unfolded.
    'fold_eval_bn': os.environ.get('ICH_B200_FOLD_EVAL_BN', '1') == '1',
    # re-derive the kernel-layout weight packs from a global optimizer post-step hook (hidden behind the GPU's backlog) instead of at
    # the start of the next forward pass (0 = only there)
    'refresh_after_step': os.environ.get('ICH_B200_REFRESH_AFTER_STEP', '1') == '1',
    # transposed-conv backward reads the up-sampled gradient in place (one strided tensor map per tap) instead of re-packing it to the
    # coarse grid first (ich_space_to_depth2: a 4 B/element pass); 0 = the re-pack path
    'convt_direct': os.environ.get('ICH_B200_CONVT_DIRECT', '1') == '1',
    # the transposed conv's bias gradient (column sums of the up-sampled gradient) is taken in the statistics epilogue of the data-gradient
    # conv that produces that tensor; 0 = a separate column-sum pass over it
    'dgrad_colsum': os.environ.get('ICH_B200_DGRAD_COLSUM', '1') == '1',
    # use the tcgen05 kernels when the shape is eligible (bf16 mode only)
    'tensor_cores': os.environ.get('ICH_B200_TENSOR_CORES', '1') == '1',
}


def precision():
    return _state['precision']


def act_dtype():
    return torch.bfloat16 if _state['precision'] == 'bf16' else torch.float32


def dtype_code(dt):
    if dt == torch.float32:
        return 0
    if dt == torch.bfloat16:
        return 1
    raise TypeError(f'ich_b200: unsupported activation dtype {dt}')


def get(key):
    return _state[key]


def set(**kw):
    for k, v in kw.items():
        if k not in _state:
            raise KeyError(k)
        if k == 'precision' and v not in ('bf16', 'fp32'):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        _state[k] = v


@contextlib.contextmanager
def override(**kw):
    old = dict(_state)
    set(**kw)
    try:
        yield
    finally:
        _state.update(old)
