timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size or golden or fork_join" 2>&1 | tail -3
for v in 1 0 1 0; do GSCAN_SPLIT_ENC_TAIL=$v timeout 200 python bench.py --no-cpu-baseline --no-decode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('split $v', round(d['value']), round(d['ms_per_step'],4), s['dec_bwd_sweep'], s['dec_wgrad_gemms'], s['encoder_side_bwd'])"; done
GSCAN_CHAIN_TIMES=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep "chain. backward" | tail -1
