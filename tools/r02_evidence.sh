# Round 2 evidence bundle on one B200: GPU tests, sanitizer passes, ncu of the shipping sweeps, launch list, CNN density
# bench, error margins.  Everything lands in gpurun_out/r02e_* (summaries are copied to profiles/ by hand).
set -x
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02e_tests.log
timeout 300 python tools/error_margins.py > gpurun_out/r02e_margins.json 2> gpurun_out/r02e_margins.err
timeout 300 python tools/cnn_density_bench.py > gpurun_out/r02e_cnn.json 2> gpurun_out/r02e_cnn.err
SAN="tests/test_gpu_parity.py -x -q -m gpu -k (tiny_aux or comp_small) and (forward_loss or greedy_decode_matches or encode_input or step_api)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(tiny_aux or comp_small or dense) and (forward_loss or greedy_decode_matches or encode_input or step_api or dense)" > gpurun_out/r02e_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comp_small and (forward_loss or greedy_decode_matches)" > gpurun_out/r02e_synccheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comp_small and (forward_loss or greedy_decode_matches)" > gpurun_out/r02e_racecheck.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > gpurun_out/r02e_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'dec_.wd_v3' -s 6 -c 2 -o gpurun_out/r02e_sweeps python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-eager > gpurun_out/r02e_ncu.log 2>&1
for f in tests memcheck synccheck racecheck; do echo "== $f"; tail -4 gpurun_out/r02e_$f.log; done; cat gpurun_out/r02e_margins.json gpurun_out/r02e_cnn.json
