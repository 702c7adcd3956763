#!/bin/bash
# round-2 session 12: staged RecordBatch decoder
set -u
out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
timeout 600 python -m pytest tests/test_logdecode.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -8 | tee $out/r02s12_tests.log
timeout 600 python tools/logdecode_bench.py 2>&1 | tail -5 | tee $out/r02s12_logdecode.log
timeout 600 python tools/logdecode_bench.py 1024 14 2>&1 | tail -4 | tee -a $out/r02s12_logdecode.log
timeout 600 python tools/logdecode_bench.py 64 170 2>&1 | tail -4 | tee -a $out/r02s12_logdecode.log
timeout 300 python tools/sanitize_driver.py log 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --num-cuda-barriers 32768 --print-limit 10 python tools/sanitize_driver.py log 2>&1 | grep -E "ok|SUMMARY|Error" | head -8 | tee $out/r02s12_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --num-cuda-barriers 32768 --print-limit 10 python tools/sanitize_driver.py log 2>&1 | grep -E "ok|SUMMARY|Error|hazard" | head -8 | tee $out/r02s12_racecheck.log
