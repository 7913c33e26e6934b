# quick A/B on one B200: parity of the main paths, then a short bench line (no CPU / eager side legs)
set -x
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size or greedy or medium or ragged or dropout or dense" 2>&1 | tail -4 > gpurun_out/quick_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
tail -4 gpurun_out/quick_tests.log; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print('RESULT', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['roofline']['stage_ms'], 'decode', round(d['decode']['seqs_per_sec']) if d.get('decode') else None)
PY
