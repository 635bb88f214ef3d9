#!/bin/bash
# The 8-query pass in its four build variants: tile pairs (one wave of 148 CTAs) or single tiles (two waves), keys in
# the kernel parameters or staged through a copy.  Libraries are built beforehand (cross-compiled, no GPU needed):
#   for w in 0 1; do for p in 0 1; do TVZ_BUILD_OUT=$PWD/build/variants/libtvz_batch_w${w}p${p}.so \
#     TVZ_NVCC_EXTRA="-DTVZ_BATCH_WIDE=$w -DTVZ_BATCH_PARAMS=$p" python -m tvidz_b200.build --force; done; done
for w in 0 1; do for p in 0 1; do
  echo "== TVZ_BATCH_WIDE=$w TVZ_BATCH_PARAMS=$p"
  TVZ_LIB=$PWD/build/variants/libtvz_batch_w${w}p${p}.so python scripts/bench_batch.py "$@" || exit 1
done; done
