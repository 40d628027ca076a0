#!/bin/bash
# round 2, call 2 (1 GPU): whole -m gpu suite incl. ranks sharing one GPU and the full-size configs,
# smoke, the default bench line (c2 + c3_n1) and the torchrun N=1 line (c3)
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/r2_v1_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_v1_pytest.log
tail -30 $O/r2_v1_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/r2_v1_smoke.log 2>&1; tail -2 $O/r2_v1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_v1_bench_c2.json 2> $O/r2_v1_bench_c2.err; tail -c 1500 $O/r2_v1_bench_c2.err; head -c 3000 $O/r2_v1_bench_c2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_v1_bench_c3_torchrun1.json 2> $O/r2_v1_bench_c3_torchrun1.err; tail -c 1500 $O/r2_v1_bench_c3_torchrun1.err; head -c 2500 $O/r2_v1_bench_c3_torchrun1.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_v1_bench_ref.json 2> $O/r2_v1_bench_ref.err; cat $O/r2_v1_bench_ref.json
