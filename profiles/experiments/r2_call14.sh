#!/bin/bash
# round 2, call 14 (2 GPUs): image path with 1..4 trace streams (ring of 6), fold kernel shapes (incl. the 8-rank
# register footprint forced on 2 ranks), layer path; multi-rank parity on 2 GPUs
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29800
run() { n=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-e2e $EXTRA 2>$O/r2_v13_$n.err | grep '^{' > $O/r2_v13_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v13_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"] and round(d["ms_per_step_pipelined_order"],4), "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4))
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v13_$n.err").read()[-1500:])
PY
}
EXTRA="--one-block-per-rank"
run a_base
run a_s1 VR_TRACE_STREAMS=1
run a_s2 VR_TRACE_STREAMS=2
run a_s4 VR_TRACE_STREAMS=4
run a_nr8 VR_FOLD_NR8=1
run a_m VR_FOLD_LIGHT=2
run a_m_g2 VR_FOLD_LIGHT=2 VR_FOLD_GRID=2
run a_l_g4 VR_FOLD_LIGHT=1 VR_FOLD_GRID=4
run a_g2 VR_FOLD_GRID=2
EXTRA=""
run b_default
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q -k all_gpus > $O/r2_v13_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_v13_pytest_2gpu.log
tail -4 $O/r2_v13_pytest_2gpu.log
