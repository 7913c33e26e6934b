# Round 2, call 1: baseline of the round-2 build (timeline code compiled out), phase timeline, ncu --set full with
# source of BOTH shipping sweeps, compute-sanitizer memcheck over the small parity cases.
set -x
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()"   # no-op unless a source is newer than the shipped .so
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
GSCAN_TIMELINE=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep timeline | tail -2 > gpurun_out/r02_c1_timeline.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'dec_.wd_v3' -s 6 -c 2 -o gpurun_out/r02_c1_sweeps python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/r02_c1_ncu.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny_aux or comp_small" > gpurun_out/r02_c1_memcheck.log 2>&1
tail -5 gpurun_out/r02_c1_memcheck.log; cat gpurun_out/r02_c1_timeline.log; cut -c1-300 gpurun_out/r02_c1_bench.json
