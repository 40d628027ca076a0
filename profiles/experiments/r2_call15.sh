#!/bin/bash
# round 2, call 15 (2 GPUs): is the per-frame loop host-bound?  host issue time per frame of the Python loop vs
# ONE vr_comm_render_frames call for the whole batch
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29900
run() { n=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps $STEPS --warmup 5 --no-cpu --no-e2e --one-block-per-rank 2>$O/r2_v14_$n.err | grep '^{' > $O/r2_v14_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v14_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"] and round(d["ms_per_step_pipelined_order"],4), "batch", d["ms_per_step_batch_call"] and round(d["ms_per_step_batch_call"],4), "host", d["host_issue_ms_per_frame"], "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4))
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v14_$n.err").read()[-1500:])
PY
}
STEPS=30
run base
run s1 VR_TRACE_STREAMS=1
run s0 VR_TRACE_STREAMS=0
run g2 VR_FOLD_GRID=2
STEPS=200
run base200
run s4_200 VR_TRACE_STREAMS=4
