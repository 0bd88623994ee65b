#!/usr/bin/env bash
# Builds libfav_b200.so (C ABI, include/fav_b200.h) for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v)
mkdir -p build
pids=()
for f in api tables corrupt epilogue conv conv_flat conv_pair forward frame_stats trust comm; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ conv.cuh -nt build/$f.o ] || [ tc_ptx.cuh -nt build/$f.o ] || [ conv_dev.cuh -nt build/$f.o ] || [ tables.h -nt build/$f.o ] || [ ../../include/fav_b200.h -nt build/$f.o ]; then
    ( "$NVCC" "${FLAGS[@]}" -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o libfav_b200.so build/api.o build/tables.o build/corrupt.o build/epilogue.o build/conv.o build/conv_flat.o build/conv_pair.o build/forward.o build/frame_stats.o build/trust.o build/comm.o -cudart static -ldl
echo "built $(pwd)/libfav_b200.so"
# the C-only driver of the hot path (tests/c/c_abi_smoke.c): proves the ABI is usable without any host-language tables
ROOT_DIR="$(cd ../.. && pwd)"
mkdir -p "$ROOT_DIR/tests/bin"
gcc -O2 -std=c11 -Wall -I "$ROOT_DIR/include" -I /usr/local/cuda/include "$ROOT_DIR/tests/c/c_abi_smoke.c" -o "$ROOT_DIR/tests/bin/c_abi_smoke" \
  -L "$(pwd)" -l:libfav_b200.so -Wl,-rpath,'$ORIGIN/../../failure-aware-vision_b200/csrc' -L /usr/local/cuda/lib64 -lcudart
echo "built $ROOT_DIR/tests/bin/c_abi_smoke"
