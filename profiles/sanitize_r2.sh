#!/bin/bash
# compute-sanitizer over the smoke run (one small scan through every stage of the hot path + the offset-0 scorer):
# memcheck, racecheck (shared-memory hazards: the prefilter's stream ring, ignore words, mbarriers), initcheck, synccheck.
cd "$(dirname "$0")/.."
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
