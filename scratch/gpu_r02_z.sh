#!/bin/bash
# round 2, GPU call Z: the other BASELINE.json workloads on the final tree (one box)
mkdir -p gpurun_out; O=gpurun_out
for c in cfg5 cfg4g cfg1 cfg2; do
  timeout 150 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/r02z_bench_$c.json 2> $O/r02z_bench_$c.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02z_bench_*.json')):
    try:
        j = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, round(j['ms_per_step'], 3), round(j['e2e']['ms_per_step'], 3), round(j['roofline']['achieved'], 1))
    except Exception as e:
        print(f, 'ERR', e)
PY
