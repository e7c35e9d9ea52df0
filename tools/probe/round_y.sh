#!/bin/bash
# GPU round Y of r2 (final, 1 GPU): smoke, whole suite, the bench line as the driver runs it, reference arm, captures of the default mid-depth kernel and of k_pairs_fast
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/ry_smoke.log 2>&1; tail -1 $O/ry_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > $O/ry_tests.log 2>&1
echo "tests rc=$?" >> $O/ry_tests.log; tail -3 $O/ry_tests.log
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/ry_bench.json 2> $O/ry_bench.err
echo "bench rc=$?"; tail -c 300 $O/ry_bench.err; cut -c1-200 $O/ry_bench.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/ry_bench_ref.json 2> $O/ry_bench_ref.err; cut -c1-200 $O/ry_bench_ref.json
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-dense --no-strong --no-cfg5 --no-cfg1"
CMD4="python tools/time_cfg4.py 6000"
$CMD2 > $O/ry_plain2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ry_launches.csv $CMD2 > $O/ry_ncu_l2.log 2>&1
$CMD4 > $O/ry_cfg4.json 2> $O/ry_plain4.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ry_launches_cfg4.csv $CMD4 > $O/ry_ncu_l4.log 2>&1
cat $O/ry_cfg4.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 5 -c 1 -o $O/ry_prof_k_pairs_fast $CMD2 > $O/ry_ncu_a.log 2>&1; tail -1 $O/ry_ncu_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_gram_ws -s 2 -c 1 -o $O/ry_prof_k_tile_gram_ws $CMD4 > $O/ry_ncu_b.log 2>&1; tail -1 $O/ry_ncu_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gram_i8 -s 2 -c 1 -o $O/ry_prof_k_gram_i8 python bench.py --dense-only --steps 2 --warmup 1 > $O/ry_ncu_c.log 2>&1; tail -1 $O/ry_ncu_c.log
echo done
