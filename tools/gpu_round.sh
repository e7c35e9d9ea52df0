#!/bin/bash
# One GPU-box pass: tests, bench lines, ncu launch list + full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; cut -c1-300 $O/bench_$TAG.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; cut -c1-200 $O/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-dense"
$CMD > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu1_$TAG.log 2>&1
$CMD > $O/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 5 -c 1 -o $O/prof_$TAG $CMD > $O/ncu2_$TAG.log 2>&1
tail -2 $O/ncu2_$TAG.log
