#!/bin/bash
# Bounds-checked variant of the library (-DTVZ_CHECKED: every index of the tile / upsert kernels is asserted,
# a violation traps with a message) and the matcher's tests + the small end-to-end pass against it.
# The pool's GPUs refuse compute-sanitizer, so this is the memory check there is.
#   bash scripts/checked_build.sh build     (here, no GPU)      bash scripts/checked_build.sh run   (GPU box)
set -e
OUT=build/libtvidz_b200_checked.so
if [ "$1" = "build" ]; then
    mkdir -p build
    TVZ_BUILD_OUT=$OUT TVZ_NVCC_EXTRA="-DTVZ_CHECKED" python -c "from tvidz_b200 import build; print(build.build(force=True))"
else
    export TVZ_LIB=$PWD/$OUT
    python scripts/sanitize_small.py
    python -m pytest tests/test_gpu_match.py -x -q
    echo "checked build: no bound violated"
fi
