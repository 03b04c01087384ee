#!/usr/bin/env bash
# Builds libcrvqa.so (sm_100a) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
       --expt-relaxed-constexpr)
OUT=../crvqa/libcrvqa.so
mkdir -p obj
pids=()
for f in gemm_sm100 elementwise select loss fused_ops attention fq_attention; do
  ( "$NVCC" "${FLAGS[@]}" -c "$f.cu" -o "obj/$f.o" > "obj/$f.log" 2>&1 || { cat "obj/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT" obj/gemm_sm100.o obj/elementwise.o obj/select.o obj/loss.o obj/fused_ops.o obj/attention.o obj/fq_attention.o -lcudart
echo "built $OUT"
