#!/bin/bash
# GPU round R of r2: forked small-unit kernels in untimed runs, link probe in the bench -- whole suite, full bench line
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/rr_smoke.log 2>&1; tail -1 $O/rr_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rr_tests.log 2>&1
echo "tests rc=$?" >> $O/rr_tests.log; tail -4 $O/rr_tests.log
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/rr_bench.json 2> $O/rr_bench.err
echo "bench rc=$?"; tail -c 600 $O/rr_bench.err; cut -c1-300 $O/rr_bench.json
echo done
