#!/usr/bin/env bash
# Builds libspfsplat.so (sm_100a only) in-tree.  project_fwd.cu is compiled with -fmad=false:
# the index-affecting projection path must match the oracle bit for bit.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
pids=()
for f in project_fwd binning blend project_bwd camera rope loss adapter ply allreduce capi; do
  EXTRA=""
  [ "$f" = project_fwd ] && EXTRA="-fmad=false"
  $NVCC $ARCH $FLAGS $EXTRA -c $f.cu -o build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC $ARCH -shared -o ../libspfsplat.so build/project_fwd.o build/binning.o build/blend.o build/project_bwd.o build/camera.o build/rope.o build/loss.o build/adapter.o build/ply.o build/allreduce.o build/capi.o -lcudart
echo "built $(cd .. && pwd)/libspfsplat.so"
