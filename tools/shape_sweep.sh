#!/bin/bash
# kernel time / roofline fraction of the scan kernels across workload shapes (one line each)
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-52s kernel %.4f ms  %.0f GB/s  frac %.3f  %.3e msg/s' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['value']))" "$@"; }
if [ $# -gt 0 ]; then for c in "$@"; do run $c; done; exit 0; fi
run --mode fused
run --mode fused --run-len 500
run --mode fused --partitions 256
run --mode fused --partitions 256 --run-len 500
run --mode fused --key-mode 1
run --mode fused --key-mode 2
run --mode counters
run --mode counters --run-len 500
run --mode counters --partitions 256
run --mode alive
