#!/bin/bash
# round-2 multi-GPU session: N>1 parity on real hardware (NCCL merge, exact alive-key exchange), then scaling lines.
# usage: tools/r02_multigpu.sh <N> [tag]
set -u
N=${1:-2}; tag=${2:-r02mg$N}
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
nvidia-smi -L | tee $out/${tag}_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $out/${tag}_tests.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
# every line below carries "verified": the merged state of all ranks equals rank 0's own single-engine scan of all shards
tr --steps 20 --warmup 3 --no-cpu > $out/${tag}_bench_C1.json 2> $out/${tag}_bench_C1.err; tail -c 400 $out/${tag}_bench_C1.json; echo
tr --steps 20 --warmup 3 --mode hll --no-e2e > $out/${tag}_bench_C1_hll.json 2> $out/${tag}_bench_C1_hll.err; tail -c 300 $out/${tag}_bench_C1_hll.json; echo
tr --steps 20 --warmup 3 --mode counters --no-e2e > $out/${tag}_bench_C1_counters.json 2> $out/${tag}_bench_C1_counters.err; tail -c 300 $out/${tag}_bench_C1_counters.json; echo
tr --config C3 --steps 10 --warmup 3 --no-e2e > $out/${tag}_bench_C3.json 2> $out/${tag}_bench_C3.err; tail -c 300 $out/${tag}_bench_C3.json; echo
tail -3 $out/${tag}_bench_*.err
