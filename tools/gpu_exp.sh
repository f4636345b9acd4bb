#!/bin/bash
for v in "" "UGB200_DXT_BPT1=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['value'])"
done
