# quick A/B: parity of the full-size cases, bench with / without the shadow schedules, chain timeline
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size or golden or alternative or fork_join" 2>&1 | tail -3
for cuts in 35,65,90 30,60,85 25,50,75,92; do GSCAN_SHADOW_FWD_CUTS=$cuts timeout 200 python bench.py --no-cpu-baseline --no-decode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('fwd cuts $cuts', round(d['value']), round(d['ms_per_step'],4), s['dec_fwd_sweep'], s['out_proj'], s['out_proj_bwd'])"; done
GSCAN_SHADOW=0 timeout 200 python bench.py --no-cpu-baseline --no-decode 2>/dev/null | cut -c1-90
GSCAN_CHAIN_TIMES=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep "chain. forward" | tail -1
