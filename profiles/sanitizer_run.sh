#!/bin/bash
# compute-sanitizer passes over the CUDA path (run on a B200 box):
#   memcheck over a subset of the GPU tests, racecheck over a tiny workload that touches every kernel
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
if [ "$1" != "race-only" ]; then
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_trace_classes.py -m gpu -x -q -k "golden or edge_cases or chunked or tally" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
fi
for part in ${PARTS:-assemble call}; do
timeout 1200 compute-sanitizer --tool racecheck --print-limit 0 python profiles/racecheck_workload.py $part 2>&1 | grep -A3 "Error:\|RACECHECK SUMMARY\|workload done" | grep -v "Host Frame" > gpurun_out/racecheck_$part.log
grep "RACECHECK SUMMARY" gpurun_out/racecheck_$part.log; grep -c "Error:" gpurun_out/racecheck_$part.log
done
