#!/bin/bash
# compute-sanitizer over a small-N subset of every kernel family: logs go to gpurun_out/ and,
# summarised, to profiles/.  The in-kernel NVLink exchanges of sharded runs are not covered: the tools serialise kernels
# of one process, and two ranks that wait for each other inside a kernel would not make progress under them.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  ( timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/sanitize_$tool.log
  echo "== $tool"; tail -n 12 gpurun_out/sanitize_$tool.log
done
