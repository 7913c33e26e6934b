# GPU box check: parity tests, bench line, ncu launch list of one step (see tools/launch_table.py)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t.log
python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
for lim in 124 100; do GSCAN_GROUP_SM_LIMIT=$lim python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_lim$lim.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/t.log; for f in gpurun_out/bench_b.json gpurun_out/bench_lim*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'])"; done
