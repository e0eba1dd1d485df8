#!/bin/bash
for cfg in "4 16" "2 16" "2 14" "2 12" "1 15" "1 14" "1 13" "4 12"; do
  set -- $cfg
  echo -n "wpb=$1 warps_per_sm=$2 "
  JMPC_WPB=$1 JMPC_WARPS_PER_SM=$2 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
