#!/bin/bash
# round-2 session 4: pipelined table pass
set -u
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r02s4_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify "$@" 2>$out/r02s4_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-60s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step']))" "$@" || tail -3 $out/r02s4_last.err; }
{
for v in base stage1 t768; do
  echo "== variant $v =="
  export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_$v.so
  run --mode alive
  run --mode alive --distinct-keys 1000000
  if [ $v = base ]; then
    run --mode fused
    run --mode alive --tombstones 0
    run --mode alive --distinct-keys 100000000
    run --mode alive --key-mode 1
    run --mode alive --run-len 500
    run --config C2 --steps 3 --warmup 1
    run --config C3 --steps 5 --warmup 2
    run --mode hll
    run --mode counters
  fi
done
unset KTA_LIB
} 2>&1 | tee $out/r02s4_sweep.log
export KTA_LIB=$PWD/kafka_topic_analyzer_b200/libkta_gpu_exp_base.so
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o $out/r02s4_prof_alive \
    python bench.py --mode alive --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > $out/r02s4_prof_alive.log 2>&1
unset KTA_LIB
cli=kafka_topic_analyzer_b200/csrc/cli/kafka-topic-analyzer
for f in "" "-c"; do $cli -t bench -b none --synthetic n=40000000,partitions=64,distinct_keys=2000000 --feed push $f 2>&1 >/dev/null | grep feed= | tee -a $out/r02s4_push.log; done
for f in "" "-c"; do $cli -t bench -b none --synthetic n=40000000,partitions=64,distinct_keys=2000000 --feed batch $f 2>&1 >/dev/null | grep feed= | tee -a $out/r02s4_push.log; done
