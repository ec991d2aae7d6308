for cfg in "1 128" "2 128" "1 64" "2 64" "4 64"; do set -- $cfg; export HMDPOSE_TF32_CHUNK=$1 HMDPOSE_TF32_BN=$2;
python bench.py --precision parity --headline-only --no-cpu-baseline --rounds 3 --steps 20 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $1 bn $2:', j['value'], j['e2e']['value'], j['single_stream']['ms_per_step'], j['per_kernel']['gemm_tf32_kernel'])";
python tools/tf32_probe.py net 2>&1 | grep "net\[" | tr '\n' ' '; echo; done
