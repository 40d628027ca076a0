#!/bin/bash
# round 2, call 1: baselines on this round's box + full ncu captures of the sampler at the dense point (P1),
# of the c3 per-block launches (path B, N=1), and the c5 launch list
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/r2_v0_gpu.txt 2>&1
timeout 300 python bench.py --samples 887 --steps 20 --no-cpu > $O/r2_v0_bench_c2_p1.json 2> $O/r2_v0_bench_c2_p1.err
timeout 300 python bench.py --workload c3 --steps 20 --no-cpu > $O/r2_v0_bench_c3_n1.json 2> $O/r2_v0_bench_c3.err
timeout 300 python bench.py --workload c5 --steps 10 --no-cpu > $O/r2_v0_bench_c5.json 2> $O/r2_v0_bench_c5.err
timeout 300 python bench.py --workload c4 --steps 2 --no-cpu > $O/r2_v0_bench_c4.json 2> $O/r2_v0_bench_c4.err
VR_COUNT_SAMPLES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 \
    -o $O/r2_v0_p1_full -f python bench.py --samples 887 --steps 2 --warmup 3 --no-cpu > $O/r2_v0_ncu_p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 24 -c 8 \
    -o $O/r2_v0_c3n1_full -f python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $O/r2_v0_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:layers_fold -s 3 -c 1 \
    -o $O/r2_v0_c3n1_fold_full -f python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $O/r2_v0_ncu_c3f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 \
    -o $O/r2_v0_c4_full -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > $O/r2_v0_ncu_c4.log 2>&1
ls -la $O | tail -20
for f in $O/r2_v0_bench_*.json; do echo "== $f"; head -c 600 $f; echo; done
