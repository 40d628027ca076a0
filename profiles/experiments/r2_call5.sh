#!/bin/bash
# round 2, call 5 (2 GPUs): refined sparse march (7 CTAs/SM) on c2 / c3-N1 / general, parity tests, then the
# exchange timeline (globaltimer stamps) of the 2-rank image path, pushed and pulled
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r2_v4_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_v4_pytest.log
tail -4 $O/r2_v4_pytest.log
run() { n=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-c3 2>/dev/null | grep '^{' > $O/r2_v4_c2_$n.json
  env "$@" timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > $O/r2_v4_c3_$n.json
  env "$@" timeout 300 python bench.py --samples 887 --steps 20 --warmup 5 --no-cpu --no-c3 2>/dev/null | grep '^{' > $O/r2_v4_p1_$n.json
  python - <<PY
import json
a=json.load(open("$O/r2_v4_c2_$n.json")); b=json.load(open("$O/r2_v4_c3_$n.json")); c=json.load(open("$O/r2_v4_p1_$n.json"))
print("$n", "c2 ms", round(a["ms_per_step"],4), "c3 ms", round(b["ms_per_step"],4), "c3 render", round(b["render_ms_per_frame"],4), "p1 ms", round(c["ms_per_step"],4))
PY
}
run general VR_NO_SPARSE=1
run sparse VR_NO_SPARSE=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for push in 1 0; do
VR_TIMELINE=1 VR_PUSH=$push timeout 600 $TR --master-port 2953$push bench.py --gpus 2 --steps 20 --warmup 5 --one-block-per-rank --no-cpu 2> $O/r2_v4_tl_push$push.err | grep '^{' > $O/r2_v4_tl_push$push.json
python - <<PY
import json
d=json.load(open("$O/r2_v4_tl_push$push.json"))
print("push=$push ms", round(d["ms_per_step"],4), "render", round(d["render_ms_per_frame"],4), "in-step composite", round(d["composite_in_step_ms"],4))
print(d["exchange_timeline"])
print(d["per_rank_ms"]["rows"])
PY
done
