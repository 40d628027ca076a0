mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r1_v6_pytest.log 2>&1; tail -15 $O/r1_v6_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 > $O/r1_v6_bench_c2.json 2> $O/r1_v6_bench_c2.err; cat $O/r1_v6_bench_c2.json; tail -3 $O/r1_v6_bench_c2.err
echo "== c3 n1"; timeout 200 python bench.py --workload c3 --steps 20 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'], d['composite_ms_per_frame'])"
echo "== c5"; timeout 200 python bench.py --workload c5 --steps 10 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'], d['composite_ms_per_frame'])"
echo "== c4"; timeout 300 python bench.py --workload c4 --steps 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
