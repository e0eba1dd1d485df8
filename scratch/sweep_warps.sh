#!/bin/bash
for w in 4 8 12 16; do
  echo -n "warps_per_sm=$w "
  JMPC_WARPS_PER_SM=$w python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
