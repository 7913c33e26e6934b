# Round-end capture on one B200: parity tests, bench lines of the three workloads, chain timeline, ncu launch list,
# ncu --set full of the kernels added in v5.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/v5_tests.log
timeout 600 python bench.py > gpurun_out/v5_bench.json 2> gpurun_out/v5_bench.err
timeout 300 python bench.py --workload comp_aux --no-cpu-baseline --no-decode > gpurun_out/v5_bench_comp_aux.json 2>/dev/null
timeout 300 python bench.py --workload tlen --no-cpu-baseline --no-decode > gpurun_out/v5_bench_tlen.json 2>/dev/null
GSCAN_CHAIN_TIMES=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode 2>&1 | grep chain | tail -4 > gpurun_out/v5_chain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 110 --csv --log-file gpurun_out/v5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/v5_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_value_z|tc_group_tn|encoder_fwd_res|encoder_bwd_res|head_bwd_fused' -s 10 -c 6 -o gpurun_out/v5_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/v5_ncu_full.log 2>&1
tail -3 gpurun_out/v5_tests.log; cat gpurun_out/v5_chain.log; cut -c1-400 gpurun_out/v5_bench.json
