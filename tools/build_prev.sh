#!/bin/bash
# Builds libev2b.so of another commit next to the current one (ev2gym_b200/csrc/libev2b_<name>.so, git-ignored) so that
# tools/ab_kernels.py can time both builds in ONE GPU session (EV2B_LIB=...): box-to-box noise is ~5 %.
#   [EXTRA="-DEV2B_EVL_MINB=6"] tools/build_prev.sh <commit | WORK> <name>
set -e
REV=${1:-HEAD}; NAME=${2:-prev}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
if [ "$REV" = WORK ]; then (cd "$ROOT" && tar -c ev2gym_b200/csrc/*.cu ev2gym_b200/csrc/*.cuh ev2gym_b200/csrc/*.h include) | tar -x -C "$TMP"
else git -C "$ROOT" archive "$REV" ev2gym_b200/csrc include | tar -x -C "$TMP"; fi
FLAGS=$(python -c "import sys; sys.path.insert(0, '$ROOT'); from ev2gym_b200 import _lib; print(' '.join(_lib.NVCC_FLAGS))" | sed "s#$ROOT/include#$TMP/include#g")
nvcc $FLAGS $EXTRA -o "$ROOT/ev2gym_b200/csrc/libev2b_$NAME.so" "$TMP/ev2gym_b200/csrc/ev2b.cu"
rm -rf "$TMP"
ls -la "$ROOT/ev2gym_b200/csrc/libev2b_$NAME.so"
