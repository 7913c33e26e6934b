# A/B of the SM budgets of the persistent GEMMs that run beside helper chains (GSCAN_CAP_PRELUDE / GSCAN_CAP_POST)
for pre in 96 112 120 132 148; do for post in 96 120 148; do
GSCAN_CAP_PRELUDE=$pre GSCAN_CAP_POST=$post python bench.py --no-cpu-baseline --no-decode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('pre $pre post $post', round(d['value']), round(d['ms_per_step'],4), s['encoder_side'], s['dec_wgrad_gemms'], s['encoder_side_bwd'])"
done; done
