#!/bin/bash
# partitioned alive-key path: parity tests, kernel time, per-kernel breakdown (ncu launch list; cold-cache, serialised)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned" 2>&1 | tail -12 | tee gpurun_out/part_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra "$@" 2>gpurun_out/part_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-52s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms  %.3e msg/s' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']))" "$@" || tail -5 gpurun_out/part_err.log; }
{
for a in "$@"; do run $a; done
} 2>&1 | tee gpurun_out/part_sweep.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/part_launches.csv python bench.py --mode alive --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > gpurun_out/part_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/part_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[-9:]:
    print(r[ki][:60], r[vi])
PY
