mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r1_v7_pytest.log 2>&1; tail -15 $O/r1_v7_pytest.log
echo "== c4"; timeout 300 python bench.py --workload c4 --steps 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['modes_ms_per_step'], d['e2e']['modes_h2d_bytes'])"
echo "== c2"; timeout 300 python bench.py --steps 20 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['modes_ms_per_step'], d['e2e']['modes_h2d_bytes'])"
echo "== pcie"; timeout 300 ./profiles/experiments/pcie_gather 2>&1 | grep -i "64-byte\|dense\|variant 1" | tee $O/r1_v7_pcie_gather64.txt
