#!/bin/bash
# round 2, call 18 (2 GPUs): pushed ray layers (the sampler stores every entry into the receive pool of the rank
# that folds its tile; the fold reads local memory) vs pulled; multi-rank parity incl. ranks sharing one GPU
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q > $O/r2_v17_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_v17_pytest_2gpu.log
tail -5 $O/r2_v17_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=30200
run() { n=$1; shift; port=$((port+1))
  env "$@" timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e $EXTRA 2>$O/r2_v17_$n.err | grep '^{' > $O/r2_v17_$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v17_$n.json"))
    print("$n", "ms", round(d["ms_per_step"],4), "render_alone", round(d["render_ms_per_frame"],4), "comp_aligned", round(d["composite_ms_per_frame"],4), "parity", (d.get("parity") or {}).get("bit_exact"))
except Exception as e:
    print("$n FAILED", e); print(open("$O/r2_v17_$n.err").read()[-1500:])
PY
}
EXTRA=""
run b_push
EXTRA="--no-cpu"
run b_pull VR_LAYER_PUSH=0
