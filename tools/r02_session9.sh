#!/bin/bash
# round-2 session 9: in-line exact mode with a queued, prefetched, dense table pass
set -u
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
timeout 180 python tools/sanitize_driver.py exact ragged ring log 2>&1 | tail -30 | tee $out/r02s9_quick.log
rc=${PIPESTATUS[0]}; if [ $rc -ne 0 ]; then echo "quick check failed rc=$rc: stopping" | tee -a $out/r02s9_quick.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r02s9_tests.log
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify "$@" 2>$out/r02s9_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-60s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step']))" "$@" || tail -3 $out/r02s9_last.err; }
{
for v in base stage0 t768; do
  echo "== variant $v =="
  export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_$v.so
  run --mode alive
  run --mode alive --distinct-keys 1000000
done
export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_base.so
for kib in 524288; do echo "== table $kib KiB =="; KTA_ALIVE_TABLE_KIB=$kib run --mode alive; done
KTA_ALIVE_TABLE_KIB=524288 run --config C2 --steps 3 --warmup 1
run --config C2 --steps 3 --warmup 1
run --mode fused
run --mode alive --distinct-keys 100000000
run --mode alive --key-mode 1
run --mode alive --run-len 500
run --mode alive --partitions 256
unset KTA_LIB
} 2>&1 | tee $out/r02s9_sweep.log
export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_base.so
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o $out/r02s9_prof_alive \
    python bench.py --mode alive --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > $out/r02s9_prof_alive.log 2>&1
