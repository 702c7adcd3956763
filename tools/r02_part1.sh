#!/bin/bash
# first GPU session of the partitioned alive-key path: its parity tests, then kernel time against the direct path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partitioned" 2>&1 | tail -25 | tee gpurun_out/part_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra "$@" 2>gpurun_out/part_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-52s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms  %.3e msg/s' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']))" "$@" || tail -5 gpurun_out/part_err.log; }
{
run --mode alive
run --mode alive --tombstones 500
KTA_ALIVE_PART_MIN=0 run --mode alive --tombstones 500
run --mode alive --distinct-keys 1000000
run --mode alive --zipf-keys --geometric-values
run --mode alive --key-mode 1 --tombstones 500
run --mode hll
} 2>&1 | tee gpurun_out/part_sweep.log
