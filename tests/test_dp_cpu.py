"""CPU / gloo, world_size 2: the host-side data-parallel logic (gradient bucketing + overlapped all-reduce hooks, state
broadcast, window sharding for sliding-window inference).  The kernels are not involved: plain torch modules stand in."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from ich_b200 import dp, infer
        torch.manual_seed(100 + rank)                       # different init per rank: install() must broadcast rank 0's
        net = torch.nn.Sequential(torch.nn.Conv3d(1, 4, 3, padding=1), torch.nn.BatchNorm3d(4), torch.nn.ReLU(),
                                  torch.nn.Conv3d(4, 1, 1), torch.nn.Sigmoid())
        red = dp.install(net, bucket_bytes=64)              # tiny buckets -> several all-reduces in flight
        assert red is not None and len(red.buckets) >= 2
        ref = [p.detach().clone() for p in net.parameters()]
        gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
        dist.all_gather(gathered, ref[0])
        assert all(torch.equal(g, gathered[0]) for g in gathered)

        # two steps (the second exercises grads that already alias the buckets, and zero_grad(set_to_none=True))
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        for step in range(2):
            torch.manual_seed(rank * 10 + step)
            x = torch.rand(2, 1, 4, 8, 8)
            opt.zero_grad()
            net(x).mean().backward()
            local = [p.grad.detach().clone() for p in net.parameters()]
            # reference: recompute the un-reduced local gradient without hooks and average over ranks by hand
            import copy
            twin = copy.deepcopy(net)
            for p in twin.parameters():
                p.grad = None
            twin.train()
            # BatchNorm buffers were already updated by the first forward; parameters are identical -> same gradient
            twin[1].running_mean.copy_(torch.zeros(4)); twin[1].running_var.copy_(torch.ones(4))
            twin(x).mean().backward()
            for p_l, p_t in zip(local, twin.parameters()):
                want = p_t.grad.detach().clone()
                dist.all_reduce(want)
                want /= world
                assert torch.allclose(p_l, want, atol=1e-6), (step, (p_l - want).abs().max())
            opt.step()
        # a frozen parameter (transfer_weights + requires_grad = False, models/optim/Contrastive.py:251-253): its bucket never
        # completes through the hooks, is reduced at the end of backward with zeros in its slot, and its grad stays None
        net[0].bias.requires_grad = False
        opt.zero_grad()
        torch.manual_seed(rank * 10 + 7)
        x = torch.rand(2, 1, 4, 8, 8)
        net(x).mean().backward()
        assert net[0].bias.grad is None
        want = net[0].weight.grad.detach().clone()          # already averaged: identical on both ranks
        g = [torch.zeros_like(want) for _ in range(world)]
        dist.all_gather(g, want)
        assert torch.equal(g[0], g[1]) and want.abs().sum() > 0
        opt.step()
        net[0].bias.requires_grad = True
        # parameters stay identical across ranks after the steps
        for p in net.parameters():
            g = [torch.zeros_like(p) for _ in range(world)]
            dist.all_gather(g, p.detach())
            assert torch.equal(g[0], g[1])

        # sliding-window inference: windows sharded over ranks == single-process result
        torch.manual_seed(0)
        vol = torch.rand(1, 1, 8, 16, 16)
        net.eval()
        dp.broadcast_state(net)                              # per-rank BatchNorm buffers diverged during training (DDP semantics)
        from ich_b200 import config as _config, ops as _ops
        import host_mocks
        host_mocks.install_window_mocks(_ops, infer)         # the window kernels replaced by CPU stand-ins: host logic only
        wnet = host_mocks.ChannelLastAdapter(net)
        with _config.override(precision='fp32'):
            pred_d, mask_d = infer.sliding_window_predict(wnet, vol, (4, 8, 8), (4, 4, 8), batch=2, distributed=True)
            pred_s, mask_s = infer.sliding_window_predict(wnet, vol, (4, 8, 8), (4, 4, 8), batch=3, distributed=False)
            assert torch.allclose(pred_d, pred_s, atol=1e-6) and torch.equal(mask_d, mask_s)
            # disjoint windows: only the uint8 mask is exchanged (and the prediction when asked for)
            pred_d, mask_d = infer.sliding_window_predict(wnet, vol, (4, 8, 8), batch=2, distributed=True)
            none_d, mask_m = infer.sliding_window_predict(wnet, vol, (4, 8, 8), batch=2, distributed=True, return_pred=False)
            pred_s, mask_s = infer.sliding_window_predict(wnet, vol, (4, 8, 8), batch=3, distributed=False)
            assert none_d is None and torch.equal(mask_m, mask_s) and torch.equal(mask_d, mask_s) and torch.allclose(pred_d, pred_s, atol=1e-6)
        assert sorted(sum((dp.shard_indices(7, r, world) for r in range(world)), [])) == list(range(7))

        # cross-rank contrastive set (ICH_B200_GLOBAL_NCE): gather_rows over gloo; the loss is evaluated identically on every rank,
        # this rank's gradient = world * its rows of d(loss)/d(gathered), so that the data-parallel AVERAGE is the true gradient
        from ich_b200 import config, ops
        with config.override(global_nce=True):
            torch.manual_seed(7)
            z_all = torch.randn(world * 3, 5)                # same on every rank; rank r owns rows 3r .. 3r+2
            z = z_all[3 * rank:3 * rank + 3].clone().requires_grad_(True)
            gathered = ops.gather_rows(z)
            assert gathered.shape == (world * 3, 5) and torch.equal(gathered.detach(), z_all)
            wts = torch.arange(world * 3 * 5, dtype=torch.float32).view(world * 3, 5)
            (gathered * wts).sum().backward()
            assert torch.allclose(z.grad, world * wts[3 * rank:3 * rank + 3])
        with config.override(global_nce=False):
            z = torch.randn(3, 5, requires_grad=True)
            assert ops.gather_rows(z) is z
        # the drop-in U-Net under the reducer, kernels replaced by a no-op recorder (host plumbing only): every parameter's gradient
        # must end up aliasing its bucket after backward -- including the fused last-unit + head op, whose Function hands back the
        # gradients of final_conv and of the last block at once -- and a second step must reuse the same buckets
        from src.models.networks.UNet import UNet
        trace = []
        ops.call = lambda name, *a: trace.append(name)
        ops._stream = lambda: 0
        ops._require_cuda = lambda t, what: None
        unet = UNet(depth=3, use_3D=True, top_filter=16, midchannels_factor=2, p_dropout=0.0).train()
        red = dp.install(unet, bucket_bytes=32 << 10)
        assert len(red.buckets) >= 3
        for _ in range(2):
            unet.zero_grad()
            unet(torch.rand(1, 1, 8, 16, 16)).sum().backward()
            for b in red.buckets:
                assert b.pending == len(b.params) and b.work is None
                lo, hi = b.flat.data_ptr(), b.flat.data_ptr() + b.flat.numel() * 4
                assert all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in b.params)
        assert 'ich_bn_head_bwd' in trace
        out.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_window_sharding_gloo_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == 'ok', f'rank {rank}: {msg}'


def test_shard_indices_balanced():
    from ich_b200.dp import shard_indices
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_sliding_window_matches_oracle_rule_single_process():
    """Host logic of the window driver vs the oracle's stitching rule, with the oracle network as the stand-in module."""
    import os as _os
    from ich_b200 import infer
    from oracle import unet_oracle as UO
    fx = torch.load(_os.path.join(_os.path.dirname(__file__), 'golden', 'unet3d_combo.pt'))
    sd = fx['state_dict_after']

    class OracleNet(torch.nn.Module):
        training = False

        class final_conv:  # noqa: N801
            out_channels = 1

        def forward(self, x):
            return UO.unet_forward(x, sd, use_3D=True, training=False)

    from ich_b200 import config as _config, ops as _ops
    import host_mocks
    saved = (_ops.window_gather, _ops.window_scatter, _ops.blend_threshold, infer._cast)
    host_mocks.install_window_mocks(_ops, infer)             # the window kernels replaced by CPU stand-ins: host logic only
    try:
        vol = torch.rand(1, 1, 8, 32, 16, generator=torch.Generator().manual_seed(3))
        with _config.override(precision='fp32'):
            for window, stride in (((8, 16, 16), (8, 16, 16)), ((8, 16, 16), (8, 8, 8)), ((4, 16, 8), (4, 12, 8))):
                got, gm = infer.sliding_window_predict(host_mocks.ChannelLastAdapter(OracleNet()), vol, window, stride, batch=3, distributed=False)
                want, wm = UO.sliding_window_predict(vol, sd, window, stride)
                assert torch.allclose(got, want, atol=1e-6) and torch.equal(gm, wm)
    finally:
        _ops.window_gather, _ops.window_scatter, _ops.blend_threshold, infer._cast = saved


def test_prefetcher_and_window_ct_host_logic():
    """DevicePrefetcher yields every batch once, in order, with nested structures intact (CPU device = passthrough);
    window_ct matches the reference formula (utils/ct_utils.py:13-36)."""
    import numpy as np
    from ich_b200.staging import DevicePrefetcher, window_ct
    batches = [(torch.full((2, 1, 4, 4), float(i)), torch.zeros(2, 1, 4, 4, dtype=torch.bool), torch.tensor([i, i]), {'k': torch.tensor(i)})
               for i in range(5)]
    got = list(DevicePrefetcher(batches, 'cpu'))
    assert len(got) == 5 and all(torch.equal(g[0], b[0]) and g[3]['k'].item() == i for i, (g, b) in enumerate(zip(got, batches)))
    assert list(DevicePrefetcher([], 'cpu')) == []
    hu = torch.tensor([-1000.0, -20.0, 40.0, 100.0, 3000.0])
    ref = (hu.numpy() - (40 - 60)) / 120.0
    ref = np.clip(ref, 0, 1)
    assert np.allclose(window_ct(hu).numpy(), ref)
    assert np.allclose(window_ct(hu, 50, 100, (0, 255)).numpy(), np.clip(255 * (hu.numpy() - 0) / 100.0, 0, 255))
