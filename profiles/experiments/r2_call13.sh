#!/bin/bash
# round 2, call 13 (2 GPUs): image path (one block per rank) and layer path with the light exchange kernels,
# side-stream traces and the relaxed ring dependency -- A/B against the previous behaviour
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29700
run() { n=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-e2e $EXTRA 2>$O/r2_v12_$n.err | grep '^{' > $O/r2_v12_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v12_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"] and round(d["ms_per_step_pipelined_order"],4), "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4))
    print("    per-rank", d["per_rank_ms"]["rows"])
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v12_$n.err").read()[-1500:])
PY
}
EXTRA="--one-block-per-rank"
run a_default
run a_nostreams VR_TRACE_STREAMS=0
run a_heavy VR_FOLD_LIGHT=0
run a_old VR_FOLD_LIGHT=0 VR_TRACE_STREAMS=0
run a_cta6 VR_CTAS_PER_SM=6
run a_pull VR_PUSH=0
EXTRA=""
run b_light
run b_heavy VR_FOLD_LIGHT=0
