#!/bin/bash
# one-shot GPU validation: parity suite, smoke, then the bench with and without the fused head
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/f2_tests.txt 2>&1; echo "pytest rc $?" >> gpurun_out/f2_tests.txt
tail -5 gpurun_out/f2_tests.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/f2_smoke.txt
tail -3 gpurun_out/f2_smoke.txt
timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f2_bench_unfused.json 2> gpurun_out/f2_bench_unfused.err
ICH_B200_FUSE_HEAD=1 timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f2_bench_fused.json 2> gpurun_out/f2_bench_fused.err
python - <<'PY'
import json
for n in ('unfused', 'fused'):
    try:
        d = json.load(open(f'gpurun_out/f2_bench_{n}.json'))
        print(n, round(d['ms_per_step'], 3), 'ms/step; e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'], 3))
    except Exception as e:
        print(n, 'failed', e)
PY
