#!/bin/bash
# GPU round D of r2 (first pass of the re-entered session): smoke, the whole -m gpu suite, one full bench line,
# the reference arm, cfg4 on the three mid-depth paths, launch lists, full ncu of the top kernels
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/rd_smoke.log 2>&1; tail -1 $O/rd_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rd_tests.log 2>&1
echo "tests rc=$?" >> $O/rd_tests.log
tail -8 $O/rd_tests.log
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/rd_bench.json 2> $O/rd_bench.err
echo "bench rc=$?"; tail -c 800 $O/rd_bench.err; cut -c1-600 $O/rd_bench.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/rd_bench_ref.json 2> $O/rd_bench_ref.err; cut -c1-300 $O/rd_bench_ref.json
for p in 2 1 0; do
  LGMI_TILE_PATH=$p timeout 600 python tools/time_cfg4.py 6000 > $O/rd_cfg4_path$p.json 2> $O/rd_cfg4_path$p.err; cat $O/rd_cfg4_path$p.json
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-dense --no-strong --no-cfg5 --no-cfg1"
$CMD > $O/rd_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/rd_launches.csv $CMD > $O/rd_ncu1.log 2>&1
for p in 2 1; do
  LGMI_TILE_PATH=$p timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rd_launches_cfg4_path$p.csv python tools/time_cfg4.py 6000 > $O/rd_ncu_path$p.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rd_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > $O/rd_ncu_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 5 -c 1 -o $O/rd_prof_pairs_fast $CMD > $O/rd_ncu2.log 2>&1
LGMI_TILE_PATH=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_gram_ws -s 2 -c 1 -o $O/rd_prof_tile_gram_ws python tools/time_cfg4.py 6000 > $O/rd_ncu3.log 2>&1
LGMI_TILE_PATH=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_finish -s 2 -c 1 -o $O/rd_prof_tile_finish python tools/time_cfg4.py 6000 > $O/rd_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gram_i8 -s 1 -c 1 -o $O/rd_prof_gram_i8 python bench.py --dense-only --steps 2 --warmup 1 > $O/rd_ncu5.log 2>&1
tail -2 $O/rd_ncu2.log $O/rd_ncu3.log $O/rd_ncu4.log $O/rd_ncu5.log
echo done
