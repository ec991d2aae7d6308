#!/bin/bash
# A/B of an environment switch: tools/ab_bench.sh VAR=value [repeats]
for r in $(seq 1 ${2:-2}); do
  for v in "" "$1"; do
    env $v python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print('${v:-default}'.ljust(28),'value',b['value'],'single',b['single_stream']['ms_per_step'],'e2e',b['e2e']['value'])"
  done
done
