# A/B of the SM budgets of the persistent GEMMs that run beside helper chains (GSCAN_CAP_PRELUDE / GSCAN_CAP_POST)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pre in 0 64 96 120; do for post in 0 96 120; do
GSCAN_CAP_PRELUDE=$pre GSCAN_CAP_POST=$post python bench.py --no-cpu-baseline --no-decode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('pre $pre post $post', round(d['value']), round(d['ms_per_step'],4), s['encoder_side'], s['dec_wgrad_gemms'], s['encoder_side_bwd'])"
done; done
