#!/bin/bash
# round 2, call 3 (2 GPUs): multi-rank parity (ranks sharing one GPU, then one GPU per rank), c3 at N=2 (path B),
# and the 2-rank image-path diagnostic with the pushed and the pulled exchange
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -x -q --durations=5 > $O/r2_v2_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_v2_pytest_2gpu.log
tail -12 $O/r2_v2_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_v2_bench_c3_n2.json 2> $O/r2_v2_bench_c3_n2.err; tail -c 800 $O/r2_v2_bench_c3_n2.err; head -c 2500 $O/r2_v2_bench_c3_n2.json; echo
VR_PUSH=1 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --one-block-per-rank --no-cpu > $O/r2_v2_diag_push.json 2> $O/r2_v2_diag_push.err; tail -c 500 $O/r2_v2_diag_push.err; head -c 1800 $O/r2_v2_diag_push.json; echo
VR_PUSH=0 timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --one-block-per-rank --no-cpu > $O/r2_v2_diag_pull.json 2> $O/r2_v2_diag_pull.err; tail -c 500 $O/r2_v2_diag_pull.err; head -c 1800 $O/r2_v2_diag_pull.json; echo
