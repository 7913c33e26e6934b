# Round 2 FINAL evidence bundle on one B200 (everything lands in gpurun_out/r02f_*; summaries are copied to profiles/
# by tools/r02_collect.py).  ~12 minutes of box time.
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()" > /dev/null 2>&1
O=gpurun_out
# 1. tests, error margins, CNN density
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02f_tests.log
timeout 300 python tools/error_margins.py > $O/r02f_margins.json 2> $O/r02f_margins.err
timeout 300 python tools/cnn_density_bench.py > $O/r02f_cnn.json 2> $O/r02f_cnn.err
# 2. the bench lines (default workload with every leg; the two other single-GPU workloads BASELINE.json names)
timeout 900 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err
timeout 300 python bench.py --workload comp_aux --no-cpu-baseline --no-decode --no-gpu-eager > $O/r02f_bench_comp_aux.json 2> /dev/null
timeout 300 python bench.py --workload tlen --no-cpu-baseline --no-decode --no-gpu-eager > $O/r02f_bench_tlen.json 2> /dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_reference.json 2> /dev/null
# 3. micro-benchmarks
for u in ubench_exchange ubench_mvtile ubench_hmma_rates ubench_tc_matvec; do
  echo "## tools/$u.cu"; timeout 120 tools/bin/$u 2>&1
done > $O/r02f_ubench.txt
timeout 120 tools/bin/tc_gemm_dev 2>&1 | grep "M=" > $O/r02f_tc_gemm_dev.txt
# 4. timelines: per-warp stamps inside the sweeps, kernel-by-kernel trace of a step, pipelined chain times
GSCAN_TIMELINE=$O/r02f_tl timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > /dev/null 2> $O/r02f_timeline.err
timeout 200 python tools/step_trace.py $O/r02f_step_trace.md > /dev/null 2>&1
GSCAN_CHAIN_TIMES=2 timeout 300 python bench.py --steps 20 --warmup 8 --no-cpu-baseline --no-decode --no-gpu-eager > /dev/null 2> $O/r02f_chain.err
timeout 120 python tools/host_sync_probe.py > $O/r02f_host_probe.txt 2>&1
# 5. ncu: launch list of a step, full sets of the sweeps and of the other kernels this round wrote
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 140 --csv --log-file $O/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > $O/r02f_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'dec_.wd_v3' -s 6 -c 2 -o $O/r02f_sweeps -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > $O/r02f_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'attn_value_zm' -s 3 -c 1 -o $O/r02f_zm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > $O/r02f_ncu_zm.log 2>&1
# 6. sanitizer passes over the training step and the greedy decode of small cases
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(tiny_aux or comp_small or dense) and (forward_loss or greedy_decode_matches or encode_input or step_api or dense)" > $O/r02f_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comp_small and (forward_loss or greedy_decode_matches)" > $O/r02f_synccheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comp_small and (forward_loss or greedy_decode_matches)" > $O/r02f_racecheck.log 2>&1
for f in tests memcheck synccheck racecheck; do echo "== $f"; tail -3 $O/r02f_$f.log; done
cut -c1-300 $O/r02f_bench.json
