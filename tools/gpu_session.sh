#!/bin/bash
# One gpurun call's worth of measurement, in the order that matters if the call is cut short:
#   tools/gpu_session.sh <tag> [tests] [sweep] [bench] [launches] [ncu-fused] [ncu-counters] [ncu-alive] [budget]
# e.g.  gpurun --timeout 900 -- 'tools/gpu_session.sh r02 tests sweep bench launches ncu-fused budget'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/ afterwards
# (tools/ncu_summary.py turns a .ncu-rep into the text summary and refreshes profiles/traffic.json).
set -u
tag=${1:?tag}; shift
out=gpurun_out; mkdir -p $out
has() { local x; for x in "${ARGS[@]}"; do [ "$x" = "$1" ] && return 0; done; return 1; }
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(tests sweep bench)

if has tests; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/${tag}_tests.log
fi
if has sweep; then
  tools/shape_sweep.sh 2>&1 | tee $out/${tag}_shape_sweep.log
fi
if has bench; then
  python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json; echo
fi
if has launches; then   # per-launch durations of the bench command (cold-cache, serialised: shares only)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > $out/${tag}_launches.log 2>&1
fi
for mode in fused counters alive; do
  if has ncu-$mode; then
    k=scan_kernel
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $out/${tag}_prof_$mode \
        python bench.py --mode $mode --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > $out/${tag}_prof_$mode.log 2>&1
    ls -la $out/${tag}_prof_$mode.ncu-rep
  fi
done
if has budget; then     # per-instruction execution counts of the fused kernel (input of profiles/*_instruction_budget.md)
  ncu -i $out/${tag}_prof_fused.ncu-rep --page source --csv --print-source sass > $out/${tag}_fused_sass_counts.csv 2>/dev/null
  wc -l $out/${tag}_fused_sass_counts.csv
fi
