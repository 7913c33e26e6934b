# what the driver runs at round end, in one call: GPU tests, smoke(), both bench arms
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-200 gpurun_out/final_bench.json
