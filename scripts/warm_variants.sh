#!/bin/bash
# dram bytes + duration of the tile kernel (single query and 8-query batch, 1M rows) for the three ways of
# warming a survivor's verification record (TVZ_WARM = 0 none, 1 prefetch.global.L2, 2 unused 4-byte load)
for w in 0 1 2; do
  export TVZ_LIB=$PWD/build/variants/libtvz_warm$w.so
  for what in match batch; do
    python scripts/prof_target.py $what 6 > /dev/null || exit 1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:match_tile -s 4 -c 1 --csv python scripts/prof_target.py $what 6 2>/dev/null | grep -E "match_tile" | awk -F'","' -v w=$w -v what=$what '{print "warm=" w, what, $(NF-2), $(NF-1), $NF}'
  done
done
