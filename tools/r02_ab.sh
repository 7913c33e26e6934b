#!/bin/bash
# A/B of environment knobs through the bench itself (the host runs ahead of the GPU there; a CUPTI trace of a few steps
# is host-limited at the start of each step).  usage: bash tools/r02_ab.sh "A=1 B=2" "A=2" ...
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()" >/dev/null 2>&1
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-decode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value %.0f  ms %.4f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
