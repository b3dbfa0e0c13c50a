#!/bin/bash
# Builds libaudiosync_cuda.so from the working tree (REV empty or -) or from a git revision into
# old-audiosync_b200/variants/NAME.so (git-ignored, travels to the GPU box) for A/B runs:
#   AUDIOSYNC_CUDA_LIB=old-audiosync_b200/variants/NAME.so python tools/sweep.py ...
# usage: tools/build_variant.sh NAME [REV] [extra nvcc flags...]
set -e
NAME=$1; REV=$2; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
if [ -n "$REV" ] && [ "$REV" != "-" ]; then
  git -C "$ROOT" archive "$REV" old-audiosync_b200/csrc old-audiosync_b200/Makefile include | tar -x -C "$TMP"
else
  mkdir -p "$TMP/old-audiosync_b200"; cp -r "$ROOT/old-audiosync_b200/csrc" "$ROOT/old-audiosync_b200/Makefile" "$TMP/old-audiosync_b200/"; cp -r "$ROOT/include" "$TMP/"
fi
mkdir -p "$ROOT/old-audiosync_b200/variants"
if grep -q "^BUILD" "$TMP/old-audiosync_b200/Makefile"; then
  make -C "$TMP/old-audiosync_b200" OUT="$ROOT/old-audiosync_b200/variants/$NAME.so" EXTRA="$*" > "$TMP/make.log" 2>&1 || (tail -30 "$TMP/make.log"; exit 1)
else   # revisions from before the library was split into translation units
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
    -Wno-deprecated-gpu-targets -Xcompiler -fPIC,-fno-finite-math-only,-fvisibility=hidden "$@" \
    -shared -o "$ROOT/old-audiosync_b200/variants/$NAME.so" "$TMP/old-audiosync_b200/csrc/audiosync_cuda.cu" \
    -Xlinker -Bsymbolic -lpthread
fi
rm -rf "$TMP"; echo "built variants/$NAME.so"
