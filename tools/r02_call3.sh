# Round 2, call 3: the whole GPU test suite (incl. the reference drivers on pkg.Model and the 120-step golden),
# then the new bench line (decode e2e / cpu baseline / reference eager on the B200).
set -x
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c3_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
tail -15 gpurun_out/r02_c3_tests.log; tail -5 gpurun_out/r02_c3_bench.err; cut -c1-400 gpurun_out/r02_c3_bench.json
