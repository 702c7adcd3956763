#!/bin/bash
# round-2 multi-GPU session: N>1 parity on real hardware (NCCL merge, exact alive-key exchange), then scaling lines.
# usage: tools/r02_multigpu.sh <N> [tag]
set -u
N=${1:-2}; tag=${2:-r02mg$N}; quick=${3:-}
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
nvidia-smi -L | tee $out/${tag}_gpus.txt
[ -z "$quick" ] && timeout 300 python bench.py --mode alive --tombstones 500 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --no-verify 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('alive (1 GPU): kernel %.4f ms frac %.3f' % (r['kernel_ms'], r['frac']))" | tee $out/${tag}_alive_1gpu.log
[ -z "$quick" ] && timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $out/${tag}_tests.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
# every line below carries "verified": the merged state of all ranks equals rank 0's own single-engine scan of all shards
tr --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_C1.json 2> $out/${tag}_bench_C1.err; tail -c 400 $out/${tag}_bench_C1.json; echo
tr --steps 20 --warmup 3 --mode fused --tombstones 500 --no-e2e > $out/${tag}_bench_C1_fused.json 2> $out/${tag}_bench_C1_fused.err; tail -c 300 $out/${tag}_bench_C1_fused.json; echo
[ -z "$quick" ] && tr --steps 20 --warmup 3 --mode counters --no-e2e > $out/${tag}_bench_C1_counters.json 2> $out/${tag}_bench_C1_counters.err; tail -c 300 $out/${tag}_bench_C1_counters.json; echo
tr --config C3 --steps 10 --warmup 3 --no-e2e > $out/${tag}_bench_C3.json 2> $out/${tag}_bench_C3.err; tail -c 300 $out/${tag}_bench_C3.json; echo
tail -3 $out/${tag}_bench_*.err; exit 0
