#!/bin/bash
# GPU round E of r2 (2 GPUs): multi-GPU parity tests, bench at N=2 launched as the driver does, CLI seam on 2 GPUs
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/re_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/re_tests.log 2>&1
echo "tests rc=$?" >> $O/re_tests.log
tail -5 $O/re_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/re_bench_n2.json 2> $O/re_bench_n2.err
echo "bench rc=$?"; tail -c 600 $O/re_bench_n2.err; cut -c1-400 $O/re_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/re_bench_ref_n2.json 2> $O/re_bench_ref_n2.err
echo "ref rc=$?"; cut -c1-200 $O/re_bench_ref_n2.json
echo done
