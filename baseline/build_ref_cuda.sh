#!/bin/bash
# Builds the reference's own CUDA alignment library for sm_100 into baseline/_ref/ (git-ignored; it travels
# to the GPU box with the snapshot).  The command is the reference's install.sh:25 plus an explicit -gencode
# (the reference passes no -arch at all).  Sources are compiled where they lie under
# /root/reference/cuda; nothing is copied into this repository.
# usage: baseline/build_ref_cuda.sh [reference root]   (default /root/reference)
set -e
REF=${1:-/root/reference}
OUT=$(cd "$(dirname "$0")" && pwd)/_ref
mkdir -p "$OUT"
[ -f "$REF/cuda/gpu_aln_noref.cu" ] || { echo "no reference sources under $REF (the GPU box uses the prebuilt $OUT/gpu_aln_pack.so)"; exit 0; }
nvcc "$REF/cuda/gpu_aln_common.cu" "$REF/cuda/gpu_aln_noref.cu" -o "$OUT/gpu_aln_pack.so" \
     -shared -Xcompiler -fPIC -lcufft -std=c++11 -gencode arch=compute_100,code=sm_100 -w
echo "$OUT/gpu_aln_pack.so"
