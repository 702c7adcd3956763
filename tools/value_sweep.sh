#!/bin/bash
# BASELINE.json configs[4]: value-size sweep 64 B – 64 KiB.  Value BYTES are never read by the reference
# (src/metric.rs:235 uses only v.len()), so kernel time is independent of the value size and the *logical* topic
# GB/s (records x (key + value bytes) / time) grows linearly with it — it is not a physical bandwidth.
for v in 64 256 1024 4096 16384 65536; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --value-mean $v "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value mean %6d B  n_gpus %d  %.3e msg/s  kernel %.4f ms  alg %.0f GB/s (%.3f of measured HBM peak)  logical topic %.0f GB/s' % ($v, d['n_gpus'], d['value'], r['kernel_ms'], r['achieved'], r['frac'], d['logical_topic_gb_s']))"
done
