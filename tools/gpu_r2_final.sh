#!/bin/bash
# round 2, last call: the library as finally built - all GPU parity tests + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
