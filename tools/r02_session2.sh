#!/bin/bash
# round-2 session 2: parity with the seen cache, where L2 residency ends (table-size sweep without the cache), the cache's effect,
# compute-sanitizer, one ncu capture
set -u
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r02s2_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify "$@" 2>$out/r02s2_last.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-72s kernel %.4f ms  %.0f GB/s  frac %.3f  step %.4f ms' % (' '.join(sys.argv[1:]), r['kernel_ms'], r['achieved'], r['frac'], d['ms_per_step']))" "$@" || tail -3 $out/r02s2_last.err; }
{
echo "== seen cache on (default) =="
run --mode alive
run --mode fused
run --mode alive --tombstones 0
run --mode alive --distinct-keys 1000000
run --mode alive --distinct-keys 100000000
run --mode alive --key-mode 1
run --mode alive --partitions 256
run --config C2 --steps 3 --warmup 1
echo "== no cache: where does L2 residency end?  table sized for load 0.6, keys scaled =="
export KTA_ALIVE_NO_CACHE=1
run --mode alive --distinct-keys 300000 --alive-table-kib 4096
run --mode alive --distinct-keys 600000 --alive-table-kib 8192
run --mode alive --distinct-keys 1250000 --alive-table-kib 16384
run --mode alive --distinct-keys 2500000 --alive-table-kib 32768
run --mode alive --distinct-keys 3750000 --alive-table-kib 49152
run --mode alive --distinct-keys 5000000 --alive-table-kib 65536
run --mode alive --distinct-keys 7500000 --alive-table-kib 98304
run --mode alive
unset KTA_ALIVE_NO_CACHE
echo "== other modes =="
run --mode hll
run --mode counters
} 2>&1 | tee $out/r02s2_sweep.log
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o $out/r02s2_prof_alive \
    python bench.py --mode alive --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-verify > $out/r02s2_prof_alive.log 2>&1
ls -la $out/r02s2_prof_alive.ncu-rep
tools/sanitize.sh r02s2
