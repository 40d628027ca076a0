#!/bin/bash
# round 2, last call (1 GPU): full regression + ncu --set full captures of the kernels added this round
T=r2_v29; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${T}_pytest.log; tail -4 $O/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:trace_multi_kernel -s 1 -c 1 -o $O/${T}_c5_multi_full python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > $O/${T}_ncu_c5.log 2>&1
timeout 200 $NCU -k regex:png_ -s 3 -c 3 -o $O/${T}_png_full python profiles/experiments/png_once.py > $O/${T}_ncu_png.log 2>&1
timeout 200 $NCU -k regex:utrace_kernel -s 1 -c 1 -o $O/${T}_utrace_full python profiles/experiments/unstructured_time.py 65 > $O/${T}_ncu_utrace.log 2>&1
ls -la $O/${T}_*
