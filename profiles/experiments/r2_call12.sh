#!/bin/bash
# round 2, call 12 (1 GPU): exchange kernels with the footprint of one sampler CTA (light), exchange/trace on
# their own streams for N = 1 too; full gpu suite; c3/c5 at N = 1 light vs heavy
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2_v11_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_v11_pytest.log
tail -14 $O/r2_v11_pytest.log
run() { n=$1; shift
  env "$@" timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu 2>$O/r2_v11_c3_$n.err | grep '^{' > $O/r2_v11_c3_$n.json
  env "$@" timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>$O/r2_v11_c5_$n.err | grep '^{' > $O/r2_v11_c5_$n.json
  python - <<PY
import json
def L(f):
    try: return json.load(open(f))
    except Exception as e: return None
b=L("$O/r2_v11_c3_$n.json"); d=L("$O/r2_v11_c5_$n.json")
print("$n", "c3 ms", b and round(b["ms_per_step"],4), "c3 render/comp", b and (round(b["render_ms_per_frame"],4), round(b["composite_ms_per_frame"],4)), "c5 ms", d and (round(d["ms_per_step"],4), round(d["render_ms_per_frame"],4), round(d["composite_ms_per_frame"],4)))
PY
}
run light VR_FOLD_LIGHT=1
run heavy VR_FOLD_LIGHT=0
run light6 VR_FOLD_LIGHT=1 VR_CTAS_PER_SM=6
timeout 300 python bench.py --steps 30 --warmup 5 --no-c3 2>$O/r2_v11_c2.err | grep '^{' > $O/r2_v11_c2.json
python - <<PY
import json
d=json.load(open("$O/r2_v11_c2.json")); print("c2 ms", round(d["ms_per_step"],4), "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"] and d["parity"]["bit_exact"])
PY
