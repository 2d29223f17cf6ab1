#!/usr/bin/env bash
# Are the kernels of the current tree the kernels a given commit had?  Builds that commit's csrc/ into a scratch directory with the
# product's flags and compares the SASS of every kernel with the current librio_cuda.so (cuobjdump -sass; the anonymous-namespace name
# hash nvcc derives from the source PATH is normalised).  No GPU needed.     usage: tools/sass_identity.sh <commit>
set -euo pipefail
commit="${1:?commit}"
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp="$(mktemp -d)"
git -C "$root" archive "$commit" rio_rs_b200/csrc include | tar -x -C "$tmp"
(cd "$tmp" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -o then.so \
    rio_rs_b200/csrc/k_assign.cu rio_rs_b200/csrc/k_trie.cu rio_rs_b200/csrc/k_affinity_umma.cu rio_rs_b200/csrc/k_directory.cu \
    rio_rs_b200/csrc/engine.cu rio_rs_b200/csrc/resolver.cu rio_rs_b200/csrc/durable.cu -ldl)
norm() { cuobjdump -sass "$1" | grep -v "^Fatbin\|^=====\|identifier\|^$" | sed -E 's/_GLOBAL__N__[0-9a-f]{8}_/_GLOBAL__N__X_/g'; }
python -c "from rio_rs_b200 import build as b; b.build()" >/dev/null
norm "$tmp/then.so" > "$tmp/then.sass"
norm "$root/rio_rs_b200/librio_cuda.so" > "$tmp/now.sass"
echo "kernels at $commit: $(grep -c 'Function :' "$tmp/then.sass"), now: $(grep -c 'Function :' "$tmp/now.sass"), SASS lines: $(wc -l < "$tmp/now.sass"), differing lines: $(diff "$tmp/then.sass" "$tmp/now.sass" | grep -c '^[<>]' || true)"
rm -rf "$tmp"
