#!/bin/bash
# experiment: occupancy variants of the fused DXT1 kernel
for mb in 0 1 7 8; do
  echo "== UGB200_DXT_MINB=$mb"
  UGB200_DXT_MINB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['value'])"
done
