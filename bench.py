#!/usr/bin/env python
"""Benchmark of the hot path: 3-D U-Net (BASELINE.json configs[2]: depth 4, 32 base filters, batch 8 of 1x64x128x128 per
GPU, Dice+BCE) training step = zero_grad + forward + loss + backward (+ gradient all-reduce) + Adam step.

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200')
for p in (ROOT, PKG, os.path.join(PKG, 'code')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

NET_KW = dict(depth=4, use_3D=True, in_channels=1, out_channels=1, top_filter=32, midchannels_factor=2, p_dropout=0.0)
LOSS_KW = dict(alpha=0.5, beta=0.5, reduction='mean', p=1)
BATCH, PATCH = 8, (64, 128, 128)
WORKLOAD = 'cfg-3: 3D U-Net depth4 tf32 mcf2, batch 8/GPU of 1x64x128x128, ComboLoss(Dice+BCE), Adam'
METRIC = '3D U-Net fwd+bwd voxels/s'      # BASELINE.json `metric` (the tensor-pipe part is the `roofline` object)
CONV_FLOP_PER_STEP = 9118.5e9            # SURVEY section 8d, algorithmic 2*M*N*K x3 (fwd + dgrad + wgrad), per GPU-step


def ncu_traffic():
    """DRAM bytes of the dominant conv launch (u2.c1 forward, 64->32 @ 8x64x128x128, plane-streaming kernel) from the committed
    `ncu --set full` extract: dram__bytes_read.sum + dram__bytes_write.sum of the first row of profiles/r01_ncu_u2c1_v6_raw.csv
    (rows: forward, data-gradient, weight-gradient of that layer).  Algorithmic bytes of that launch: input 1.074 GB + output
    0.537 GB (bf16, each touched once)."""
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', 'r01_ncu_u2c1_v6_raw.csv'))))
        hdr, units, first = rows[0], rows[1], rows[2]
        tot = 0.0
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(name)
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[units[i]]
            tot += float(first[i]) * scale
        return tot
    except Exception:
        return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


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


def cpu_reference_step(sample_shape, threads, max_seconds=25.0, steps=3, warmup=1):
    """The reference's CPU path (oracle port: same graph on torch CPU ops) on a bounded sample; returns voxels/s."""
    from oracle import unet_oracle as UO, losses_oracle as LO
    from src.models.networks.UNet import UNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = UNet(**NET_KW)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    x = torch.rand(*sample_shape, generator=g)
    m = (torch.rand(*sample_shape, generator=g) > 0.98).float()
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = UO.unet_forward(x, sd, use_3D=True, training=True)
        loss = LO.combo_loss(out, m, **LOSS_KW)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > max_seconds and times:
            break
    best = min(times)
    return x.numel() / best, best, len(times)


def run_reference(args, emit):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    shape = (1, 1, 32, 128, 128)
    vox = shape[0] * shape[2] * shape[3] * shape[4]
    from oracle import unet_oracle as UO, losses_oracle as LO
    from src.models.networks.UNet import UNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = UNet(**NET_KW)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    x = torch.rand(*shape, generator=g)
    m = (torch.rand(*shape, generator=g) > 0.98).float()
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)

    def step():
        opt.zero_grad()
        loss = LO.combo_loss(UO.unet_forward(x, sd, use_3D=True, training=True), m, **LOSS_KW)
        loss.backward()
        opt.step()
        return loss.item()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = vox * args.steps / dt
    sample = f'batch 1 of 1x32x128x128 (1/16 of the per-GPU batch) per step, {threads} threads, torch CPU fp32'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voxels/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'reference_sample': sample},
            'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    emit(line)


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
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('ICH_B200_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args, emit)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from ich_b200 import config, dp, ops, _lib
    from src.models.networks.UNet import UNet
    from src.models.optim.LossFunctions import ComboLoss

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    config.set(precision=args.precision)

    torch.manual_seed(0)
    net = UNet(**NET_KW).to(dev).train()
    dp.install(net)
    lossf = ComboLoss(**LOSS_KW)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    shape = (args.batch, 1) + PATCH
    g = torch.Generator().manual_seed(rank)
    xh = torch.rand(*shape, generator=g).pin_memory()
    mh = (torch.rand(*shape, generator=g) > 0.98).float().pin_memory()
    xd, md = xh.to(dev), mh.to(dev)
    vox_step = world * shape[0] * PATCH[0] * PATCH[1] * PATCH[2]

    def step(x, m):
        opt.zero_grad()
        out = net(x)
        loss = lossf(out, m)
        loss.backward()
        opt.step()
        return loss

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
        step(xd, md)
    sampler = ClockSampler(local) if rank == 0 else None      # one NVML poller per job, not one per rank
    if sampler is not None:
        sampler.start()
    l0 = _lib.launches()
    t_dev = timed(lambda: step(xd, md), args.steps)
    launches = _lib.launches() - l0

    # dominant kernel (implicit-GEMM conv fwd / dgrad / wgrad): CUDA events around every launch over a second timed pass
    ops.PROFILE = []
    t_prof = timed(lambda: step(xd, md), min(args.steps, 3))
    prof, ops.PROFILE = ops.PROFILE, None
    conv_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in prof)
    conv_flop = sum(f for _, f, _, _ in prof)
    by_kind = {}
    for k, f, e0, e1 in prof:
        a = by_kind.setdefault(k, [0.0, 0.0])
        a[0] += f
        a[1] += e0.elapsed_time(e1)
    pk, pk_src = peaks()
    peak_tf = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
    achieved = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    # end to end through the public API: every step's inputs come from pinned HOST memory (copied inside the timed region,
    # one batch of look-ahead on a side stream = ich_b200.staging.DevicePrefetcher) and the loss is read back to the host
    from ich_b200.staging import DevicePrefetcher

    prefetch = DevicePrefetcher([(xh, mh)] * args.steps, dev)      # one prefetcher for the run, like one per DataLoader in a trainer

    def e2e_run():
        for x, m in prefetch:
            step(x, m).item()
    e2e_run()
    t_e2e = timed(e2e_run, 1)
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, best, n = cpu_reference_step((1, 1, 32, 128, 128), threads)
        cpu = {'value': v, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port',
               'sample': f'oracle port, batch 1 of 1x32x128x128 (same net), fwd+loss+bwd, best of {n}, {best:.2f} s/step'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': vox_step * args.steps / t_dev, 'unit': 'voxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': world * args.batch, 'parallelism': f'dp{world}',
                       'l2': 'working set (>10 GB of activations per step) >> 126 MB L2, no flush needed',
                       'tensor_cores': bool(config.get('tensor_cores'))},
            'e2e': {'value': vox_step * args.steps / t_e2e, 'unit': 'voxels/s', 'h2d_bytes_per_step': xh.numel() * 4 + mh.numel() * 4,
                    'd2h_bytes_per_step': 4, 'ms_per_step': 1e3 * t_e2e / args.steps},
            'gpu_launches': launches,
            'roofline': {'bound': 'tensor', 'kernel': 'implicit-GEMM conv fwd+dgrad+wgrad (all launches of the step)', 'achieved': achieved,
                         'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf if peak_tf else None, 'traffic': ncu_traffic(),
                         'traffic_note': 'DRAM bytes of the largest conv launch (u2.c1 fwd) from ncu --set full; algorithmic 1.61e9',
                         'peak_source': pk_src + ' (sustained)', 'conv_ms_per_step': conv_ms / max(1, min(args.steps, 3)),
                         'conv_share_of_step': conv_ms / (t_prof * 1e3) if t_prof else None,
                         'by_kind_tflops': {k: (a[0] / (a[1] * 1e-3) / 1e12 if a[1] else 0.0) for k, a in by_kind.items()}},
            'clocks': sampler.summary(),
        }
        if cpu:
            line['cpu_baseline'] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
