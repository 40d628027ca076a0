#!/bin/bash
# round 2, call 11 (8 GPUs): multi-rank parity on 2/4/8 GPUs, then c3 at N = 8, 4, 2, 1 (the driver's scaling sweep),
# N = 8 also with the pulled exchange and with the exchange timeline
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -x -q --durations=5 > $O/r2_v10_pytest_8gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_v10_pytest_8gpu.log
tail -6 $O/r2_v10_pytest_8gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2 1; do
  timeout 600 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 2> $O/r2_v10_bench_c3_n$n.err | grep '^{' > $O/r2_v10_bench_c3_n$n.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_v10_bench_c3_n$n.json"))
    print("N=$n ms", round(d["ms_per_step"],4), "render", round(d["render_ms_per_frame"],4), "composite", round(d.get("composite_ms_per_frame") or 0,4), "parity", (d.get("parity") or {}).get("bit_exact"), (d.get("parity") or {}).get("ok"), "t1", (d.get("t1_same_run") or {}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("ms_per_frame"))
    print("   per-rank", d.get("per_rank_ms",{}).get("rows"))
except Exception as e:
    print("N=$n FAILED", e); print(open("$O/r2_v10_bench_c3_n$n.err").read()[-1500:])
PY
done
VR_PUSH=0 timeout 600 $TR --nproc-per-node 8 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu 2> $O/r2_v10_n8_pull.err | grep '^{' > $O/r2_v10_n8_pull.json
VR_TIMELINE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29619 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu 2> $O/r2_v10_n8_tl.err | grep '^{' > $O/r2_v10_n8_tl.json
python - <<PY
import json
for f in ("r2_v10_n8_pull","r2_v10_n8_tl"):
    try:
        d=json.load(open("$O/"+f+".json"))
        print(f, "ms", round(d["ms_per_step"],4), "serial", round(d["ms_per_step_serial_order"],4), "piped", d["ms_per_step_pipelined_order"], "render", round(d["render_ms_per_frame"],4))
        print("   per-rank", d["per_rank_ms"]["rows"])
        if d.get("exchange_timeline"): 
            for r in d["exchange_timeline"]["rows_us"]: print("   tl", r)
    except Exception as e: print(f, "FAILED", e)
PY
