# GPU box check: parity tests, bench line, chain timeline, ncu launch list of one step (tools/after_sweep.py)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t.log
timeout 200 python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
GSCAN_CHAIN_TIMES=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep chain | tail -2 > gpurun_out/chain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/t.log; cat gpurun_out/chain.log; for f in gpurun_out/bench_b.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'])"; done
