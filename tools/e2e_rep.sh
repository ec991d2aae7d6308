for e in 4 5 6; do for r in 1 2 3; do python bench.py --no-cpu-baseline --e2e-inflight $e 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print('e2e_inflight',b['e2e']['host_threads'],'value',b['value'],'e2e',b['e2e']['value'])"; done; done
