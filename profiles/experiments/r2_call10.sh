#!/bin/bash
# round 2, call 10 (1 GPU): brick march (tensor TMA) debug + parity + A/B; layers fold rewrite; full gpu suite
O=gpurun_out; mkdir -p $O
touch tests/__init__.py
for a in "32 100" "48 100" "64 887"; do timeout 120 python profiles/experiments/brick_debug.py $a 2>&1 | tail -1; done
rm -f tests/__init__.py
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2_v9_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_v9_pytest.log
tail -16 $O/r2_v9_pytest.log
run() { n=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-c3 2>/dev/null | grep '^{' > $O/r2_v9_c2_$n.json
  env "$@" timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > $O/r2_v9_c3_$n.json
  env "$@" timeout 300 python bench.py --samples 887 --steps 20 --warmup 5 --no-cpu --no-c3 2>/dev/null | grep '^{' > $O/r2_v9_p1_$n.json
  env "$@" timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | grep '^{' > $O/r2_v9_c5_$n.json
  python - <<PY
import json
def L(f):
    try: return json.load(open(f))
    except Exception as e: return None
a=L("$O/r2_v9_c2_$n.json"); b=L("$O/r2_v9_c3_$n.json"); c=L("$O/r2_v9_p1_$n.json"); d=L("$O/r2_v9_c5_$n.json")
print("$n", "c2 ms", a and round(a["ms_per_step"],4), "c3 ms", b and round(b["ms_per_step"],4), "c3 render/comp", b and (round(b["render_ms_per_frame"],4), round(b["composite_ms_per_frame"],4)), "p1 ms", c and round(c["ms_per_step"],4), "c5 ms", d and (round(d["ms_per_step"],4), round(d["render_ms_per_frame"],4), round(d["composite_ms_per_frame"],4)))
PY
}
run general VR_NO_SPARSE=1 VR_NO_BRICK=1
run product VR_NO_SPARSE=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 \
    -o $O/r2_v9_p1_brick_full -f python bench.py --samples 887 --steps 2 --warmup 3 --no-cpu --no-c3 > $O/r2_v9_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:layers_fold -s 3 -c 1 \
    -o $O/r2_v9_c3n1_fold_full -f python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $O/r2_v9_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file $O/r2_v9_launches_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/r2_v9_ncu3.log 2>&1
ls -la $O | grep r2_v9 | head -30
