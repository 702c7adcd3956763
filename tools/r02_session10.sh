#!/bin/bash
# round-2 session 10: the session-3 build (simple in-line table pass) against table size — does slot probing matter?
set -u
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify "$@" 2>$out/r02s10_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-60s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step']))" "$@" || tail -3 $out/r02s10_last.err; }
{
export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_s3.so
for kib in 0 262144 524288 1048576 4194304; do run --mode alive --alive-table-kib $kib; done
run --mode alive --alive-table-kib 1048576 --distinct-keys 1000000
run --config C2 --steps 3 --warmup 1 --alive-table-kib 1048576
} 2>&1 | tee $out/r02s10_sweep.log
