#!/bin/bash
# round 2, call V: racecheck + synccheck of the staged line kernels (every converter, modes 1-3)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 10 python tools/sanitize_target.py staged > gpurun_out/racecheck_staged_i.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_staged_i.log
grep -E "RACECHECK SUMMARY|hazard|racecheck rc|exercised" gpurun_out/racecheck_staged_i.log | head -20
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 10 python tools/sanitize_target.py staged > gpurun_out/synccheck_staged_i.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/synccheck_staged_i.log
grep -E "ERROR SUMMARY|synccheck rc|exercised" gpurun_out/synccheck_staged_i.log | head
