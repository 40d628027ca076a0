#!/bin/bash
# round 2, call 16 (8 GPUs): c3 at N = 8 with the side-stream traces / ring of 6 / batch call, A/B of the fold
# kernel shape, then the scaling sweep N = 8, 4, 2 with parity
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() { n=$1; np=$2; shift; shift; port=$((port+1))
  env "$@" timeout 300 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np --steps 30 --warmup 5 $EXTRA 2>$O/r2_v15_$n.err | grep '^{' > $O/r2_v15_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v15_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"] and round(d["ms_per_step_pipelined_order"],4), "batch", d.get("ms_per_step_batch_call") and round(d["ms_per_step_batch_call"],4), "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4), "parity", (d.get("parity") or {}).get("bit_exact"), "t1", (d.get("t1_same_run") or {}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    print("    per-rank", [[r[1], r[3]] for r in d["per_rank_ms"]["rows"]])
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v15_$n.err").read()[-1500:])
PY
}
EXTRA="--no-cpu --no-e2e"
run n8_base 8
run n8_m 8 VR_FOLD_LIGHT=2
run n8_m_g2 8 VR_FOLD_LIGHT=2 VR_FOLD_GRID=2
run n8_g2 8 VR_FOLD_GRID=2
run n8_s0 8 VR_TRACE_STREAMS=0
run n8_rowmajor 8 VR_TILE_ORDER=1
run n8_pull 8 VR_PUSH=0
EXTRA="--no-e2e"
run n8_full 8
run n4_full 4
run n2_full 2
