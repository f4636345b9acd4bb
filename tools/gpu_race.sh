#!/bin/bash
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python tools/sanitize_target.py jpeg > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
grep -E "RACECHECK SUMMARY|hazard|racecheck rc|exercised|ERROR" gpurun_out/racecheck.log | head -30
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python tools/sanitize_target.py jpeg > gpurun_out/synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/synccheck.log
grep -E "SUMMARY|rc=|exercised|Barrier|divergent" gpurun_out/synccheck.log | head -20
