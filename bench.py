#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: "3D U-Net fwd+bwd voxels/s ...; conv tensor-pipe util vs BF16 peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's arm, default workload cfg3 (BASELINE.json configs[2])
    python bench.py --config cfg1|cfg2|cfg3|cfg4g|cfg4l|cfg5  # the other BASELINE.json configs, same JSON line
    python bench.py --impl reference ...                     # the reference's own modules on the HOST cores (baseline/_ref), bounded sample
    python bench.py --impl cudnn [--cudnn-mode bf16|tf32]    # side leg: the UNMODIFIED reference modules on the same GPU through
                                                             # PyTorch / cuDNN (the on-box incumbent; none of this repo's kernels)

A training step = zero_grad + forward + loss + backward (+ gradient all-reduce) + Adam step; a cfg5 step = one full volume through the
sliding-window driver.  One process per GPU under torchrun for N > 1.  Prints ONE JSON line (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200')
REF_CODE = os.path.join(ROOT, 'baseline', '_ref', 'code')       # git-ignored copy of the unmodified reference (made by __graft_entry__.build)
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = '3D U-Net fwd+bwd voxels/s'      # BASELINE.json `metric` (the tensor-pipe part is the `roofline` object)

_SEG3 = dict(depth=4, use_3D=True, in_channels=1, out_channels=1, midchannels_factor=2, p_dropout=0.0)
_COMBO = dict(alpha=0.5, beta=0.5, reduction='mean', p=1)
# SURVEY section 8d.  conv_flop = algorithmic 2*M*N*K of every conv / transposed conv, x3 (fwd + dgrad + wgrad), per GPU-step.
# cpu_sample = the bounded sample one step of the host-core reference arm runs (batch, spatial shape).
WORKLOADS = {
    'cfg1': dict(kind='seg', desc='cfg-1: 3D U-Net depth4 tf16 mcf2, batch 2 of 1x64x128x128, ComboLoss(Dice+BCE), Adam',
                 net='UNet', net_kw=dict(_SEG3, top_filter=16), loss='ComboLoss', loss_kw=_COMBO, batch=2, shape=(64, 128, 128),
                 conv_flop=571.4e9, cpu_sample=(2, (64, 128, 128))),
    'cfg2': dict(kind='seg', desc='cfg-2: 2D U-Net depth5 tf32 mcf1, batch 32 of 1x512x512 slices, BinaryDiceLoss(p=2, alpha=0.2), Adam',
                 net='UNet', net_kw=dict(depth=5, use_3D=False, in_channels=1, out_channels=1, top_filter=32, midchannels_factor=1, p_dropout=0.0),
                 loss='BinaryDiceLoss', loss_kw=dict(reduction='mean', p=2, alpha=0.2), batch=32, shape=(512, 512),
                 conv_flop=9241.7e9, cpu_sample=(1, (512, 512))),
    'cfg3': dict(kind='seg', desc='cfg-3: 3D U-Net depth4 tf32 mcf2, batch 8/GPU of 1x64x128x128, ComboLoss(Dice+BCE), Adam',
                 net='UNet', net_kw=dict(_SEG3, top_filter=32), loss='ComboLoss', loss_kw=_COMBO, batch=8, shape=(64, 128, 128),
                 conv_flop=9118.5e9, cpu_sample=(1, (64, 128, 128))),
    'cfg4g': dict(kind='nce_global', desc='cfg-4 global: UNet_Encoder 3D depth4 tf32 mcf2 MLP[512,128], 2 views of 8/GPU x 1x64x128x128, InfoNCE(tau=0.1), Adam',
                  net='UNet_Encoder', net_kw=dict(depth=4, use_3D=True, in_channels=1, top_filter=32, midchannels_factor=2, MLP_head=[512, 128], p_dropout=0.0),
                  batch=8, shape=(64, 128, 128), conv_flop=2 * 3 * 543.6e9, cpu_sample=(2, (32, 128, 128))),
    'cfg4l': dict(kind='nce_local', desc='cfg-4 local: Partial_UNet 2D depth5 n_decoder3 tf32 mcf1 head[128,32], 2 views of 32/GPU x 1x256x256, '
                                         'LocalInfoNCE(tau=0.1, K=3, n_region=20), Adam',
                  net='Partial_UNet', net_kw=dict(depth=5, n_decoder=3, use_3D=False, in_channels=1, top_filter=32, midchannels_factor=1,
                                                  head_channel=[128, 32], p_dropout=0.0),
                  batch=32, shape=(256, 256), conv_flop=2 * 3 * 658.0e9, cpu_sample=(2, (256, 256))),
    'cfg5': dict(kind='infer', desc='cfg-5: sliding-window inference, eval-mode cfg-3 net, one 1x32x512x512 volume per GPU-step, 16 windows of 32x128x128 '
                                    '(stride = window), mask = pred >= 0.5',
                 net='UNet', net_kw=dict(_SEG3, top_filter=32), batch=1, shape=(32, 512, 512), window=(32, 128, 128),
                 conv_flop=3039.5e9, cpu_sample=(1, (32, 128, 128))),
}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Sampled through NVML in-process (the
    library nvidia-smi itself queries): forking `nvidia-smi` every 100 ms from a Python thread stalled the launch loop of the
    end-to-end pass (GIL + driver locks); `nvidia-smi` stays as the fallback when pynvml is unavailable."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    # nvmlClocksEventReason* bit masks
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        return [str(sm), str(self.max_sm)] + ['Active' if mask & bit else 'Not Active' for _, bit in self.BITS]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([f.strip() for f in out.split(',')])
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.5)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = [n for n, _ in self.BITS]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith('active') for s in self.samples)]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ---------------------------------------------------------------------------------------------------------------------
# workloads: the same job description drives this repo's modules, the reference modules on cuDNN, and the reference on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def import_modules(code_dir):
    """The `src` package tree of either arm (this repo's drop-in under PKG/code, or the unmodified reference under baseline/_ref/code)."""
    sys.path.insert(0, code_dir)
    import importlib
    nets = importlib.import_module('src.models.networks.UNet')
    losses = importlib.import_module('src.models.optim.LossFunctions')
    return nets, losses


def plain_sliding_window(net, vol, window, batch, post=lambda t: t):
    """cfg-5 rule in plain torch, for the reference arms: eval-mode net on non-overlapping windows, pred >= 0.5."""
    _, _, D, H, W = vol.shape
    wd, wh, ww = window
    wins = [(d, h, w) for d in range(0, D, wd) for h in range(0, H, wh) for w in range(0, W, ww)]
    out = torch.empty((1, 1, D, H, W), dtype=torch.float32, device=vol.device)
    with torch.no_grad():
        for i in range(0, len(wins), batch):
            ch = wins[i:i + batch]
            x = torch.cat([vol[:, :, d:d + wd, h:h + wh, w:w + ww] for d, h, w in ch], 0)
            p = post(net(x)).float()
            for j, (d, h, w) in enumerate(ch):
                out[:, :, d:d + wd, h:h + wh, w:w + ww] = p[j:j + 1]
    return out >= 0.5


class Job:
    """One workload on one device for one arm.  `host` = this rank's step inputs in (pinned) host memory; `step(*device_inputs)` runs one
    step and returns a tensor whose value the end-to-end pass reads back (the loss, or the predicted mask)."""

    def __init__(self, wl, nets, losses, dev, rank, batch, spatial, arm, cudnn_mode='bf16', world=1, shard='volume'):
        import numpy as np
        self.wl, self.arm, self.dev, self.kind = wl, arm, dev, wl['kind']
        self.shard_windows = shard == 'window' and world > 1 and wl['kind'] == 'infer' and arm == 'b200' 
        self.rank, self.np = rank, np
        torch.manual_seed(0)
        net = getattr(nets, wl['net'])(**wl['net_kw']).to(dev)
        self.autocast = arm == 'cudnn' and cudnn_mode == 'bf16'
        self.fmt = None
        if arm == 'cudnn':
            torch.backends.cudnn.benchmark = True
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            if self.autocast:
                self.fmt = torch.channels_last_3d if len(spatial) == 3 else torch.channels_last
                net = net.to(memory_format=self.fmt)
        self.net = net.eval() if self.kind == 'infer' else net.train()
        self.opt = None if self.kind == 'infer' else torch.optim.Adam(net.parameters(), lr=1e-3)
        g = torch.Generator().manual_seed(0 if self.shard_windows else rank)      # window sharding: every rank holds the SAME volume
        shape = (batch, 1) + tuple(spatial)
        pin = (lambda t: t.pin_memory()) if dev.type == 'cuda' else (lambda t: t)
        self.units = batch
        for s in spatial:
            self.units *= s
        if self.kind == 'seg':
            self.lossf = getattr(losses, wl['loss'])(**wl['loss_kw'])
            self.host = [pin(torch.rand(*shape, generator=g)), pin((torch.rand(*shape, generator=g) > 0.98).float())]
        elif self.kind == 'nce_global':
            # set_size = the LOCAL batch (what the module checks against); with ICH_B200_GLOBAL_NCE=1 the drop-in's loss gathers the
            # embeddings of all ranks itself (ops.gather_rows)
            self.lossf = losses.InfoNCELoss(set_size=batch, tau=0.1, device=str(dev))
            self.host = [pin(torch.rand(*shape, generator=g)), pin(torch.rand(*shape, generator=g))]
            self.units *= 2
        elif self.kind == 'nce_local':
            self.lossf = losses.LocalInfoNCELoss(tau=0.1, K=3, n_region=20, device=str(dev))
            self.host = [pin(torch.rand(*shape, generator=g)), pin(torch.rand(*shape, generator=g))]
            self.units *= 2
        else:
            with torch.no_grad():          # BN running statistics randomised once, seeded
                gg = torch.Generator().manual_seed(1)
                for name, b in net.named_buffers():
                    if name.endswith('running_mean'):
                        b.copy_((torch.randn(b.shape, generator=gg) * 0.1).to(dev))
                    elif name.endswith('running_var'):
                        b.copy_((torch.rand(b.shape, generator=gg) + 0.5).to(dev))
            vol = torch.rand(*shape, generator=g)
            if arm == 'b200':
                # this repo's arm takes the volume as RAW int16 Hounsfield units (what a CT pipeline holds) and windows it on the device
                # (ich_stage_ct, window 40 / 120 as utils/ct_utils.py:13); the reference arms get the host-windowed fp32 volume in [0, 1]
                # that the reference's own segement_volume would upload (UNet2D.py:286-297)
                self.host = [pin((vol * 120.0 - 20.0).round().to(torch.int16))]
            else:
                self.host = [pin(vol)]

    def _fwd(self, x):
        if self.fmt is not None:
            x = x.contiguous(memory_format=self.fmt)
        if self.autocast:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                return self.net(x)
        return self.net(x)

    def step(self, *inp):
        import torch.nn.functional as F
        if self.kind == 'infer':
            if self.arm == 'b200':
                from ich_b200 import infer, ops
                _, _, D, H, W = inp[0].shape
                x = ops.staged(ops.stage_ct(inp[0], 40, 120, (0, 1)).view(1, D, H, W, 1))
                return infer.sliding_window_predict(self.net, x, self.wl['window'], batch=8, distributed=self.shard_windows, return_pred=False)[1]
            return plain_sliding_window(self._fwd, inp[0], self.wl['window'], 8)
        self.opt.zero_grad()
        if self.kind == 'seg':
            loss = self.lossf(self._fwd(inp[0]).float(), inp[1])
        elif self.kind == 'nce_global':
            loss = self.lossf(F.normalize(self._fwd(inp[0]).float(), dim=1), F.normalize(self._fwd(inp[1]).float(), dim=1))
        else:
            self.np.random.seed(self.rank)
            loss = self.lossf(self._fwd(inp[0]).float(), self._fwd(inp[1]).float())
        loss.backward()
        self.opt.step()
        return loss


def readback(t):
    """Device -> host read of a step's result: loss.item() for training steps, the predicted mask for inference.  Returns bytes."""
    if t.numel() == 1:
        t.item()
        return 4
    h = t.to('cpu')
    return h.numel() * h.element_size()


# ---------------------------------------------------------------------------------------------------------------------
# host-core reference arm
# ---------------------------------------------------------------------------------------------------------------------
def oracle_port_step(wl, batch, spatial, threads):
    """Fallback when baseline/_ref is absent: the oracle port (same graph on torch CPU ops).  seg workloads only."""
    from oracle import unet_oracle as UO, losses_oracle as LO
    sys.path.insert(0, os.path.join(PKG, 'code'))
    from src.models.networks.UNet import UNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in UNet(**wl['net_kw']).state_dict().items()}
    g = torch.Generator().manual_seed(0)
    shape = (batch, 1) + tuple(spatial)
    x = torch.rand(*shape, generator=g)
    m = (torch.rand(*shape, generator=g) > 0.98).float()
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3)
    lossf = LO.combo_loss if wl.get('loss') == 'ComboLoss' else LO.binary_dice_loss

    def step():
        opt.zero_grad()
        out = UO.unet_forward(x, sd, use_3D=wl['net_kw']['use_3D'], training=wl['kind'] != 'infer')
        if wl['kind'] == 'infer':
            return out
        loss = lossf(out, m, **wl['loss_kw'])
        loss.backward()
        opt.step()
        return loss
    return step


def run_reference(args, emit, wl):
    """The reference's own CPU implementation of the path on this box's host cores: the UNMODIFIED reference modules from baseline/_ref
    (kind 'reference'), else the oracle port (kind 'port').  Each step is a bounded sample of the workload (WORKLOADS[..]['cpu_sample'])."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    batch, spatial = wl['cpu_sample']
    if os.path.isdir(REF_CODE):
        nets, losses = import_modules(REF_CODE)
        job = Job(wl, nets, losses, torch.device('cpu'), 0, batch, spatial, 'reference')
        step, kind, units = (lambda: job.step(*job.host)), 'reference', job.units
    else:
        if wl['kind'] not in ('seg', 'infer'):
            emit({'impl': 'reference', 'unavailable': 'baseline/_ref is absent and the oracle port covers the segmentation workloads only'})
            return
        step, kind = oracle_port_step(wl, batch, spatial, threads), 'port'
        units = batch
        for s in spatial:
            units *= s
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step()
        if r.numel() == 1:
            r.item()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    sample = (f'{"unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port"}, torch {torch.__version__} CPU fp32, '
              f'{threads} threads; one step = batch {batch} of 1x{"x".join(map(str, spatial))} of the same net / loss / optimizer '
              f'({"the full cfg-1 step" if args.config == "cfg1" else "a bounded sample of the per-GPU step"})')
    emit({'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voxels/s', 'n_gpus': args.gpus, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
          'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': wl['desc']},
          'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': threads, 'kind': kind, 'sample': sample},
          'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0})


def cpu_baseline_subprocess(config):
    """cpu_baseline of this repo's arm: the host-core reference arm run in a child process (the two `src` package trees cannot share
    one interpreter), 1 warm-up + 2 timed steps of its bounded sample."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', config, '--steps', '2', '--warmup', '1'],
                           capture_output=True, text=True, timeout=600, env={k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')})
        line = json.loads(r.stdout.strip().splitlines()[-1])
        return line.get('cpu_baseline')
    except Exception as e:      # noqa: BLE001
        return {'value': None, 'unit': 'voxels/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {e!r}'}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    # stdout carries exactly ONE JSON line: route everything else written to fd 1 (NCCL banners, library chatter) to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + '\n').encode())
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'cudnn'])
    ap.add_argument('--config', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--cudnn-mode', default='bf16', choices=['bf16', 'tf32'],
                    help='--impl cudnn: bf16 autocast + channels_last(_3d), or fp32 modules with TF32 convs (torch defaults)')
    ap.add_argument('--precision', default=os.environ.get('ICH_B200_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--graph', type=int, default=int(os.environ.get('ICH_B200_CUDA_GRAPH', '-1')),
                    help='1 = run the step through ich_b200.graph.GraphedStep (one CUDA-graph replay per step; under torchrun the bucketed NCCL '
                         'all-reduce is captured with it -- measured at 2 and 8 GPUs), 0 = eager launches; default 1')
    ap.add_argument('--shard', default='volume', choices=['volume', 'window'],
                    help='cfg5 at N > 1: one volume per GPU-step (weak scaling, no data-path collective) or the windows of ONE volume sharded '
                         'over the ranks with a uint8 mask exchange (strong scaling)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    if args.impl == 'reference':
        return run_reference(args, emit, wl)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    batch = args.batch or wl['batch']

    if args.impl == 'cudnn':
        if not os.path.isdir(REF_CODE):
            if rank == 0:
                emit({'impl': 'cudnn', 'unavailable': 'baseline/_ref (copy of the unmodified reference) is absent: run __graft_entry__.build() where /root/reference exists'})
            return
        nets, losses = import_modules(REF_CODE)
        job = Job(wl, nets, losses, dev, rank, batch, wl['shape'], 'cudnn', args.cudnn_mode, world)
        if world > 1 and job.opt is not None:
            job.net = torch.nn.parallel.DistributedDataParallel(job.net, device_ids=[local])
        _lib = None
    else:
        from ich_b200 import config, dp, _lib, profile
        config.set(precision=args.precision)
        nets, losses = import_modules(os.path.join(PKG, 'code'))
        job = Job(wl, nets, losses, dev, rank, batch, wl['shape'], 'b200', world=world, shard=args.shard)
        if job.opt is not None:
            dp.install(job.net)
    dev_in = [t.to(dev) for t in job.host]
    strong = bool(getattr(job, 'shard_windows', False))
    units_step = job.units if strong else world * job.units

    step_fn = job.step
    if args.graph < 0:
        args.graph = 1
    capturable = (job.opt is not None and job.kind != 'nce_local') or (job.kind == 'infer' and not getattr(job, 'shard_windows', False))
    if args.graph and args.impl == 'b200' and capturable:     # (LocalInfoNCE draws its regions on the host every step: never captured)
        from ich_b200.graph import GraphedStep
        # calls 1-2 eager, call 3 (still warm-up, >= 3 enforced above) captures; a step that cannot be captured keeps running eagerly
        step_fn = GraphedStep(job.step, job.opt, warmup=2, strict=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / 1e3

    for _ in range(args.warmup):
        step_fn(*dev_in)
    sampler = ClockSampler(local) if rank == 0 else None      # one NVML poller per job, not one per rank
    if sampler is not None:
        sampler.start()
    l0 = _lib.launches() if _lib else 0
    t_dev = timed(lambda: step_fn(*dev_in), args.steps)
    launches = (_lib.launches() - l0) if _lib else 0
    if _lib and getattr(step_fn, 'graph', None) is not None:       # replayed: the launches recorded at capture time run once per replay
        launches = step_fn.kernels_per_replay * args.steps

    # end to end through the public API: every step's inputs come from pinned HOST memory (copied inside the timed region,
    # one batch of look-ahead on a side stream = ich_b200.staging.DevicePrefetcher) and the result is read back to the host
    sys.path.insert(0, PKG)
    from ich_b200.staging import DevicePrefetcher
    d2h = [0]

    def e2e_run():
        for inp in DevicePrefetcher([tuple(job.host)] * args.steps, dev):      # one prefetcher per pass, like one per DataLoader epoch
            d2h[0] = readback(step_fn(*inp))
    e2e_run()
    t_e2e = timed(e2e_run, 1)
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    clocks = sampler.summary() if sampler is not None else None

    # per-launch profile (second pass, CUDA events around EVERY C-ABI launch on the launching stream): tensor-bound conv kernels as
    # achieved TFLOP/s, HBM-bound families as achieved GB/s
    roof = None
    if _lib:
        n_prof = min(args.steps, 3)
        _lib.PROFILE = []
        t_prof = timed(lambda: job.step(*dev_in), n_prof)
        prof, _lib.PROFILE = _lib.PROFILE, None
        fam = profile.summarise(prof, n_prof)
        pk, pk_src = peaks()
        burst = bool(clocks and clocks['sm_mhz'] and clocks['sm_max_mhz'] and clocks['sm_mhz'] >= 0.9 * clocks['sm_max_mhz'])
        peak_tf = pk['bf16_tflops'] if burst else pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        t_ms = sum(fam[f]['ms_per_step'] for f in profile.TENSOR_FAMILIES if f in fam)
        t_flop = sum(fam[f].get('tflops', 0.0) * fam[f]['ms_per_step'] for f in profile.TENSOR_FAMILIES if f in fam)   # TFLOP/s * ms = GFLOP
        achieved = t_flop / t_ms if t_ms else 0.0
        hbm = pk['hbm_gbs']
        roof = {'bound': 'tensor', 'kernel': 'implicit-GEMM conv fwd + dgrad + wgrad and transposed-conv GEMMs (every launch of the step)',
                'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf if peak_tf else None,
                'peak_source': f'{pk_src}: {"burst" if burst else "sustained"} bf16 peak (median SM clock under load '
                               f'{clocks["sm_mhz"] if clocks else None} MHz vs max {clocks["sm_max_mhz"] if clocks else None})',
                'traffic': None, 'traffic_note': 'per-launch DRAM bytes from ncu are in profiles/ (README there), not re-measured by this run',
                'algorithmic_bytes_per_step': sum(fam[f].get('algo_bytes_per_step', 0.0) for f in profile.TENSOR_FAMILIES if f in fam),
                'conv_ms_per_step': t_ms, 'conv_share_of_step': t_ms / (1e3 * t_prof / n_prof) if t_prof else None,
                'algorithmic_conv_tflop_per_step': t_flop / 1e3,
                'tensor_families': {f: {'tflops': fam[f].get('tflops'), 'frac': (fam[f].get('tflops') or 0.0) / peak_tf, 'ms_per_step': fam[f]['ms_per_step']}
                                    for f in profile.TENSOR_FAMILIES if f in fam},
                'hbm_families': {f: {'gbs': v.get('gbs'), 'frac': (v.get('gbs') or 0.0) / hbm, 'ms_per_step': v['ms_per_step'],
                                     'launches_per_step': v['launches_per_step']}
                                 for f, v in fam.items() if f not in profile.TENSOR_FAMILIES},
                'hbm_peak_gbs': hbm}

        # DRAM bytes the same launches moved, from the committed ncu launch list of this workload (a number taken under ncu is never
        # re-measured here; it is comparable because `achieved` sums over exactly these launches)
        try:
            with open(os.path.join(ROOT, 'profiles', 'conv_dram_traffic.json')) as f:
                tr = json.load(f).get(args.config)
        except (OSError, ValueError):
            tr = None
        if tr and tr.get('batch_per_gpu') == batch and args.precision == 'bf16':
            roof['traffic'] = tr['dram_read_bytes_per_step'] + tr['dram_write_bytes_per_step']
            roof['traffic_note'] = ('bytes per step = dram__bytes_read.sum + dram__bytes_write.sum over the conv launches of one step (the launches '
                                    '`achieved` sums over); ' + tr['source'])

    cpu = None
    if rank == 0 and world == 1 and args.impl == 'b200' and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args.config)

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in job.host)
        line = {
            'metric': METRIC, 'value': units_step * args.steps / t_dev, 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
            'vs_baseline': None, 'dtype': ('bf16' if args.precision == 'bf16' else 'f32') if args.impl == 'b200' else ('bf16' if args.cudnn_mode == 'bf16' else 'tf32'),
            'data': 'synthetic', 'config': {'workload': wl['desc']},
            'run': {'global_batch': world * batch, 'parallelism': f'dp{world}' + (' (windows of one volume sharded, uint8 mask all-reduce)' if strong else ''),
                    'cuda_graph': bool(getattr(step_fn, 'graph', None) is not None),
                    'l2': 'working set per step (GBs of activations) >> 126 MB L2, no flush needed'},
            'e2e': {'value': units_step * args.steps / t_e2e, 'unit': 'voxels/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h[0], 'ms_per_step': 1e3 * t_e2e / args.steps},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        if args.impl == 'cudnn':
            line['impl'] = 'cudnn'
            line['run']['cudnn'] = (f'unmodified reference modules (baseline/_ref) on torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, '
                                    + ('bf16 autocast + channels_last' if args.cudnn_mode == 'bf16' else 'fp32 modules, TF32 convs')
                                    + ', cudnn.benchmark=True')
        if roof:
            line['roofline'] = roof
        if cpu:
            line['cpu_baseline'] = cpu
        emit(line)
    if world > 1:
        if getattr(step_fn, 'graph', None) is not None:
            # live CUDA graphs hold the NCCL communicator and destroy_process_group() hung with them at 8 GPUs: the job is done, so leave
            # without the teardown -- and never later than 20 s from now, whatever the barrier does
            threading.Timer(20.0, lambda: os._exit(0)).start()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    except BaseException:       # noqa: BLE001 -- print, then leave without NCCL / CUDA teardown (a rank that dies in a collective job must not hang the others)
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
