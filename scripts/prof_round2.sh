#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): every command first runs plain and must exit 0, then under ncu.
#   launch list of the bench command, and --set full captures of the three dominant kernels at bench sizes
#   (tile kernel at the N=1 catalogue and at an N=8-sized shard).
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-dropin --no-from-file"
$BENCH > gpurun_out/r02_prof_bench_plain.json 2> gpurun_out/r02_prof_bench_plain.err || exit 1
# (the frame synthesis of the harness alone is > 600 torch launches: list the library's kernels only)
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:sad_|scene_select|match_|fragment_|upsert_|flush_l2|gather_wait' -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv $BENCH > gpurun_out/r02_prof_bench_ncu.log 2>&1
python scripts/prof_target.py match 6 || exit 1
ncu --set full --clock-control none --import-source on -k regex:match_tile -s 4 -c 1 -f -o gpurun_out/r02_tile_1m python scripts/prof_target.py match 6 > gpurun_out/ncu_a.log 2>&1
python scripts/prof_match_small.py 125000 || exit 1
ncu --set full --clock-control none --import-source on -k regex:match_tile -s 4 -c 1 -f -o gpurun_out/r02_tile_125k python scripts/prof_match_small.py 125000 > gpurun_out/ncu_b.log 2>&1
python scripts/prof_target.py score 4 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sad_bulk -s 2 -c 1 -f -o gpurun_out/r02_sad_bulk python scripts/prof_target.py score 4 > gpurun_out/ncu_c.log 2>&1
python scripts/prof_fragment.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:fragment_stream -s 2 -c 1 -f -o gpurun_out/r02_fragment_stream python scripts/prof_fragment.py > gpurun_out/ncu_d.log 2>&1
python scripts/prof_target.py batch 6 || exit 1
ncu --set full --clock-control none --import-source on -k regex:match_tile -s 4 -c 1 -f -o gpurun_out/r02_tile_batch8_1m python scripts/prof_target.py batch 6 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep
