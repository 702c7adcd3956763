#!/bin/bash
# minimal multi-GPU check: the driver's SCALE command on N GPUs (C1, sketch mode, e2e) and the C3 shape; both lines verified
set -u
N=${1:-8}; tag=r02mg$N; out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
nvidia-smi -L | tee $out/${tag}_gpus.txt
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
tr --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_C1.json 2> $out/${tag}_bench_C1.err; tail -c 600 $out/${tag}_bench_C1.json; echo
tr --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu > $out/${tag}_bench_C3.json 2> $out/${tag}_bench_C3.err; tail -c 400 $out/${tag}_bench_C3.json; echo
tail -2 $out/${tag}_bench_*.err; exit 0
