"""CPU: the parts of bench.py that do not need a GPU -- the host-core reference arm prints ONE JSON line with the contract's keys on a
bounded sample, the workload registry covers every BASELINE.json config, and the per-launch algorithmic-work table (ich_b200/profile.py)
knows the entry points the training step launches."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'cfg5', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'voxels/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('cfg-5')


def test_workloads_cover_every_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert len(base['configs']) == 5
    assert set(bench.WORKLOADS) == {'cfg1', 'cfg2', 'cfg3', 'cfg4g', 'cfg4l', 'cfg5'}      # configs[3] has a global and a local part
    w = bench.WORKLOADS
    assert w['cfg3']['batch'] == 8 and w['cfg3']['shape'] == (64, 128, 128) and w['cfg3']['net_kw']['top_filter'] == 32
    assert w['cfg1']['batch'] == 2 and w['cfg1']['net_kw']['top_filter'] == 16 and w['cfg1']['cpu_sample'] == (2, (64, 128, 128))
    assert w['cfg2']['net_kw']['use_3D'] is False and w['cfg2']['batch'] == 32 and w['cfg2']['shape'] == (512, 512)
    assert w['cfg5']['shape'] == (32, 512, 512) and w['cfg5']['window'] == (32, 128, 128)
    for v in w.values():
        assert v['conv_flop'] > 0 and v['desc'].startswith('cfg-')


def test_algorithmic_work_table():
    from ich_b200 import _lib, profile
    protos = _lib.parse_header()
    for name in profile.TABLE:
        assert name in protos or name == 'ich_threshold_confusion', name
    args = dict(N=8, D=64, H=128, W=128, Cin=64, Cout=32, KD=3, KH=3, KW=3)
    fam, flop, by = profile.algo_work('ich_conv_tc_fwd_stats', args, 'fwd')
    assert fam == 'conv_fwd' and abs(flop - 927.7e9) < 1e9                                 # u2.c1 = 927.7 GFLOP
    assert by == 8 * 64 * 128 * 128 * (64 + 32) * 2 + 27 * 64 * 32 * 2                      # input + output once (bf16) + the weight pack
    assert profile.algo_work('ich_conv_tc_wgrad', args)[0] == 'conv_wgrad'
    fam, flop, by = profile.algo_work('ich_bn_act_bwd', dict(M=1 << 20, C=32, dtype=1))
    assert fam == 'bn_bwd' and flop is None and by == 5 * (1 << 20) * 32 * 2
    assert profile.algo_work('ich_bn_finalize', {})[0] == 'small'
    fam, flop, by = profile.algo_work('ich_convT2_tc_dgrad', dict(N=8, D=32, H=64, W=64, Cin=64, Cout=32, FD=2))
    assert fam == 'convT' and abs(flop - 34.4e9) < 0.1e9                                   # upT2 = 34.4 GFLOP
    assert by == 8 * 32 * 64 * 64 * (64 + 8 * 32) * 2                                       # coarse tensor + the 8x larger fine one


def test_conv_dram_traffic_file_matches_the_default_workload():
    """`roofline.traffic` comes from a committed ncu launch list: the file must name the default workload at its per-GPU batch and sit within
    a small factor of the algorithmic bytes (SURVEY section 8d layer table), else the number in the bench line is stale."""
    import json
    import bench
    tr = json.load(open(os.path.join(ROOT, 'profiles', 'conv_dram_traffic.json')))['cfg3']
    assert tr['batch_per_gpu'] == bench.WORKLOADS['cfg3']['batch']
    assert os.path.exists(os.path.join(ROOT, tr['source'].split(' ')[0]))
    total = tr['dram_read_bytes_per_step'] + tr['dram_write_bytes_per_step']
    assert 17.5e9 < total < 1.25 * 17.5e9              # algorithmic: 17.55 GB per step (every conv operand once, fwd + dgrad + wgrad)
