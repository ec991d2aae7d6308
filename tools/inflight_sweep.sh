#!/bin/bash
# bench.py at several numbers of steps in flight (handles/streams per GPU); extra bench args via $BENCH_ARGS
for n in "$@"; do
  python bench.py --inflight "$n" --steps 96 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | tail -1 > /tmp/b.json
  python - "$n" <<'PY'
import json, sys
b = json.load(open("/tmp/b.json"))
print("inflight", sys.argv[1], "value", b["value"], "ms/step", b["ms_per_step"], "e2e", b["e2e"]["value"], flush=True)
PY
done
