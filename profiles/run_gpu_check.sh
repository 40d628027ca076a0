#!/bin/bash
# GPU-box check for one round snapshot: parity tests, smoke, bench lines, ncu launch list + full capture
# usage (from the repo root, under gpurun): bash profiles/run_gpu_check.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; cat $O/${TAG}_bench_c2.json
timeout 300 python bench.py --workload c3 --steps 20 --no-cpu > $O/${TAG}_bench_c3_n1.json 2> $O/${TAG}_bench_c3.err; cat $O/${TAG}_bench_c3_n1.json
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; cat $O/${TAG}_bench_c4.json
timeout 300 python bench.py --workload c5 --steps 10 --no-cpu > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; cat $O/${TAG}_bench_c5.json
timeout 300 python bench.py --samples 887 --steps 20 --no-cpu > $O/${TAG}_bench_c2_p1.json 2> $O/${TAG}_bench_c2_p1.err; cat $O/${TAG}_bench_c2_p1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cat $O/${TAG}_bench_ref.json
# launch list of the same bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3_n1.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench_c3.log 2>&1
# one full capture of the dominant kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 2 \
    -o $O/${TAG}_trace_full -f python bench.py --steps 3 --warmup 3 --no-cpu > $O/${TAG}_ncu_full.log 2>&1
ls -la $O
