#!/bin/bash
# A/B of the forward-pass schedules (GSCAN_FWD_SCHED, GSCAN_CNN_AFTER, GSCAN_CAP_PRELUDE) on one B200
python -c "import multimodal_seq2seq_gscan_b200 as p; p.build()" >/dev/null 2>&1
mkdir -p gpurun_out
export STEP_TRACE_BRIEF=1
run() { echo "== $*"; env "$@" timeout 200 python tools/step_trace.py gpurun_out/trace_$(echo "$*" | tr ' =' '__').md 2>&1 | tail -1; }
run GSCAN_FWD_SCHED=0
run GSCAN_FWD_SCHED=1
run GSCAN_FWD_SCHED=2
run GSCAN_FWD_SCHED=1 GSCAN_CNN_AFTER=1
run GSCAN_FWD_SCHED=2 GSCAN_CNN_AFTER=1
echo "== bench"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
