mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r1_v5_pytest.log 2>&1; tail -3 $O/r1_v5_pytest.log
for v in libvr_b200.so libvr_b200_mb8.so libvr_b200_old.so; do
  for S in 100 887; do
    echo "== $v S=$S"; VR_LIB_NAME=$v timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu --samples $S 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'])"
  done
done
echo "== c3 n1"; timeout 200 python bench.py --workload c3 --steps 20 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'], d['composite_ms_per_frame'])"
echo "== c5"; timeout 200 python bench.py --workload c5 --steps 10 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_frame'], d['composite_ms_per_frame'])"
echo "== pcie"; timeout 300 ./profiles/experiments/pcie_gather > $O/r1_v5_pcie_gather.txt 2>&1; cat $O/r1_v5_pcie_gather.txt
