#!/bin/bash
# round 2, call 17 (2 GPUs): do the library's streams alias onto the 8 default hardware queues (false
# dependencies between the side-stream traces and the exchange)?  CUDA_DEVICE_MAX_CONNECTIONS = 8 (default) vs 32
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=30100
run() { n=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-e2e $EXTRA 2>$O/r2_v16_$n.err | grep '^{' > $O/r2_v16_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v16_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"] and round(d["ms_per_step_pipelined_order"],4), "batch", d.get("ms_per_step_batch_call") and round(d["ms_per_step_batch_call"],4), "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4))
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v16_$n.err").read()[-1500:])
PY
}
EXTRA="--one-block-per-rank"
run a_c8
run a_c32 CUDA_DEVICE_MAX_CONNECTIONS=32
run a_c32_s4 CUDA_DEVICE_MAX_CONNECTIONS=32 VR_TRACE_STREAMS=4
run a_c32_s2 CUDA_DEVICE_MAX_CONNECTIONS=32 VR_TRACE_STREAMS=2
EXTRA=""
run b_c8
run b_c32 CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 200 env CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 N=1 c32 ms', round(d['ms_per_step'],4), round(d['render_ms_per_frame'],4))"
timeout 200 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 N=1 c8 ms', round(d['ms_per_step'],4), round(d['render_ms_per_frame'],4))"
