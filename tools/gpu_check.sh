python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t.log
python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/t.log; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
