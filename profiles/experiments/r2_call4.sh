#!/bin/bash
# round 2, call 4 (1 GPU): parity with the sparse march, then A/B of the sampler: general march (VR_NO_SPARSE=1)
# vs sparse march at 9 / 8 / 7 resident CTAs per SM (56 / 64 / 72 registers), c2 and c3 (N=1)
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_v3_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_v3_pytest.log
tail -5 $O/r2_v3_pytest.log
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-c3 2>/dev/null | grep '^{' > $O/r2_v3_c2_$n.json
  env "$@" timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > $O/r2_v3_c3_$n.json
  python - <<PY
import json
a=json.load(open("$O/r2_v3_c2_$n.json")); b=json.load(open("$O/r2_v3_c3_$n.json"))
print("$n", "c2 ms", round(a["ms_per_step"],4), "c3 ms", round(b["ms_per_step"],4), "c3 render", round(b["render_ms_per_frame"],4))
PY
}
run general VR_NO_SPARSE=1
run sparse_mb9 VR_NO_SPARSE=0
run sparse_mb8 VR_LIB_NAME=libvr_mb8.so
run sparse_mb7 VR_LIB_NAME=libvr_mb7.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 \
    -o $O/r2_v3_c2_sparse_mb9_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-c3 > $O/r2_v3_ncu1.log 2>&1
VR_LIB_NAME=libvr_mb8.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 \
    -o $O/r2_v3_c2_sparse_mb8_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-c3 > $O/r2_v3_ncu2.log 2>&1
ls -la $O | grep r2_v3
