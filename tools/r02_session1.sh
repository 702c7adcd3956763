#!/bin/bash
# round-2 session 1: parity of the new alive-key table, then what it costs at different table sizes / with and without L2 hints
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/r02s1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r02s1_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra "$@" 2>$out/r02s1_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-64s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step']))" "$@" || tail -3 $out/r02s1_last.err; }
{
echo "== hints on (default build) =="
for kib in 0 98304 114688 163840 262144 1048576; do run --mode alive --alive-table-kib $kib; done
run --mode alive --tombstones 0
run --mode alive --distinct-keys 1000000
run --mode alive --distinct-keys 100000000
run --mode alive --key-mode 1
run --mode alive --partitions 256
echo "== hints off =="
export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_nohints.so
for kib in 0 98304 114688 262144; do run --mode alive --alive-table-kib $kib; done
unset KTA_LIB
echo "== shapes (r2-prep changes) =="
run --mode fused
run --mode fused --run-len 500
run --mode fused --partitions 256
run --mode fused --key-mode 1
run --mode fused --key-mode 2
run --mode counters
run --mode counters --run-len 500
run --mode counters --partitions 256
} 2>&1 | tee $out/r02s1_sweep.log
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o $out/r02s1_prof_alive \
    python bench.py --mode alive --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > $out/r02s1_prof_alive.log 2>&1
ls -la $out/r02s1_prof_alive.ncu-rep
