#!/bin/bash
# Same-box A/B runs: boxes of the GPU pool differ by ~5 % (profiles/r01i_march_variants.md), so two builds must be
# compared inside ONE gpurun call.  This builds the kernel library of another commit next to the current one:
#
#   bash tools/ab_build.sh HEAD~1                      # -> brats21_b200/libb21_prev.so (git-ignored, travels with gpurun)
#   gpurun -- 'for i in 1 2; do B21_LIB=$PWD/brats21_b200/libb21_prev.so python bench.py --no-cpu-baseline | cut -c1-120;
#                               python bench.py --no-cpu-baseline | cut -c1-120; done'
#
# B21_LIB (brats21_b200/_lib.py) selects the library; the Python side is the working tree's, so the two commits must
# share the C-ABI of the entry points the benchmark touches.
set -euo pipefail
ref=${1:-HEAD}
out=${2:-brats21_b200/libb21_prev.so}
root=$(git rev-parse --show-toplevel)
tmp=$(mktemp -d)
trap 'git -C "$root" worktree remove --force "$tmp" >/dev/null 2>&1 || true; rm -rf "$tmp"' EXIT
git -C "$root" worktree add -f "$tmp" "$ref" >/dev/null
(cd "$tmp" && python -c "
import sys; sys.path.insert(0, '.')
from brats21_b200 import build
print(build.build_lib(force=True))")
cp "$tmp/brats21_b200/libb21.so" "$root/$out"
echo "built $ref -> $out"
