#!/bin/bash
# GPU round P of r2 (8 GPUs): the bench launched as the driver does at N=8; topology of the box
O=gpurun_out
mkdir -p $O
bash tools/probe/topo.sh > /dev/null 2>&1; cp $O/topo.txt $O/rp_topo_n8.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/rp_bench_n8.json 2> $O/rp_bench_n8.err
echo "bench rc=$?"; tail -c 400 $O/rp_bench_n8.err; cut -c1-300 $O/rp_bench_n8.json
echo done
