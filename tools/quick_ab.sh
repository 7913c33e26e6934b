# quick A/B of the shadow cut points (GSCAN_SHADOW_CUTS) + chain timeline + parity of the full-size case
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size_against_oracle or alternative" 2>&1 | tail -3
for cuts in 45 70,40,15 75,50,25 80,60,40,20 70,45,25,10; do GSCAN_SHADOW_CUTS=$cuts timeout 200 python bench.py --no-cpu-baseline --no-decode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('cuts $cuts', round(d['value']), round(d['ms_per_step'],4), s['dec_bwd_sweep'], s['dec_wgrad_gemms'], s['encoder_side_bwd'])"; done
GSCAN_CHAIN_TIMES=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep "chain. backward" | tail -1
