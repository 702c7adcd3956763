#!/bin/bash
# round-2 measurement set of the shipped build (one GPU).  usage: tools/r02_final.sh [tag]
set -u
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
git rev-parse HEAD > $out/${tag}_head.txt 2>/dev/null || true
timeout 180 python tools/sanitize_driver.py 2>&1 | tail -8 | tee $out/${tag}_quick.log
rc=${PIPESTATUS[0]}; if [ $rc -ne 0 ]; then echo "quick check failed rc=$rc: stopping" | tee -a $out/${tag}_quick.log; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $out/${tag}_tests.log
# the driver's command, then the other named configurations
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_C1.json 2> $out/${tag}_bench_C1.err; tail -c 300 $out/${tag}_bench_C1.json; echo
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_C1_reference.json 2> $out/${tag}_bench_C1_reference.err; tail -c 200 $out/${tag}_bench_C1_reference.json; echo
timeout 900 python bench.py --config C2 --steps 5 --warmup 2 --no-e2e --no-cpu --no-extra > $out/${tag}_bench_C2.json 2> $out/${tag}_bench_C2.err; tail -c 300 $out/${tag}_bench_C2.json; echo
timeout 900 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra > $out/${tag}_bench_C3.json 2> $out/${tag}_bench_C3.err; tail -c 300 $out/${tag}_bench_C3.json; echo
timeout 900 python bench.py --config C4 --steps 10 --warmup 3 > $out/${tag}_bench_C4.json 2> $out/${tag}_bench_C4.err; tail -c 300 $out/${tag}_bench_C4.json; echo
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify "$@" 2>$out/${tag}_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-52s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms  %.3e msg/s' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']))" "$@" || tail -3 $out/${tag}_last.err; }
{
run --mode hll
run --mode hll --tombstones 500
run --mode hll --run-len 500
run --mode hll --partitions 256
run --mode hll --partitions 256 --shard-world 8
run --mode hll --key-mode 1
run --mode hll --key-mode 2
run --mode hll --zipf-keys --geometric-values
run --mode counters
run --mode counters --run-len 500
run --mode counters --partitions 256
run --mode alive
run --mode alive --tombstones 500
run --mode fused --tombstones 500
run --mode alive --distinct-keys 1000000
run --mode alive --distinct-keys 100000000
run --mode alive --run-len 500 --tombstones 500
run --mode alive --key-mode 1 --tombstones 500
} 2>&1 | tee $out/${tag}_shape_sweep.log
# per-launch durations of the driver's command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > $out/${tag}_launches.log 2>&1
for mode in hll alive counters; do
  ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o $out/${tag}_prof_$mode \
      python bench.py --mode $mode --tombstones 500 --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > $out/${tag}_prof_$mode.log 2>&1
  ls -la $out/${tag}_prof_$mode.ncu-rep
done
cli=kafka_topic_analyzer_b200/csrc/cli/kafka-topic-analyzer
for feed in push batch device; do for f in "" "-c"; do
  $cli -t bench -b none --synthetic n=40000000,partitions=64,distinct_keys=4000000 --feed $feed $f 2>&1 >/dev/null | grep feed= | sed "s/^/[$f] /" | tee -a $out/${tag}_cli_feeds.log
done; done
{ python tools/logdecode_bench.py 256 56; python tools/logdecode_bench.py 1024 14; python tools/logdecode_bench.py 256 56 lz4; python tools/logdecode_bench.py 256 56 snappy; python tools/logdecode_bench.py 256 56 gzip; } 2>&1 | grep -E "encoded|decode" | tee $out/${tag}_logdecode_bench.log
tools/sanitize.sh $tag
