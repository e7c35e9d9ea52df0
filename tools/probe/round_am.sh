#!/bin/bash
# GPU round AM of r2 (2 GPUs): multi-GPU parity test, full bench at N=2 launched as the driver does (streamed e2e figure)
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/ram_tests.log 2>&1
echo "tests rc=$?" >> $O/ram_tests.log; tail -3 $O/ram_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/ram_bench_n2.json 2> $O/ram_bench_n2.err
echo "bench rc=$?"; tail -c 600 $O/ram_bench_n2.err; cut -c1-300 $O/ram_bench_n2.json
echo done
