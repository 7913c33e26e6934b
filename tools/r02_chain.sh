#!/bin/bash
# pipelined chain timelines (GSCAN_CHAIN_TIMES=2) of knob variants.  usage: bash tools/r02_chain.sh "A=1" "B=2 C=3" ...
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()" >/dev/null 2>&1
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg GSCAN_CHAIN_TIMES=2 timeout 300 python bench.py --steps 20 --warmup 8 --no-cpu-baseline --no-decode 2>gpurun_out/chain.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value %.0f  ms %.4f' % (d['value'], d['ms_per_step']))"
  grep "chain. forward" gpurun_out/chain.err | sed -n 40p
  grep "chain. backward" gpurun_out/chain.err | sed -n 40p
done
