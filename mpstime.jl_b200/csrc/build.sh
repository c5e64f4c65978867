#!/bin/bash
# Build libmpstime_b200.so in-tree for sm_100a.  cudart is linked statically; NCCL is dlopen'ed at run time.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p _obj
pids=()
for f in encode encode_table krao_gemm krao_slab bond_grad bond_grad_kr bond_misc gemm svd_jacobi svd_subspace sym_eig_reg impute api; do
  if [ ! -f _obj/$f.o ] || [ $f.cu -nt _obj/$f.o ] || [ mpst_common.cuh -nt _obj/$f.o ] || [ dmma.cuh -nt _obj/$f.o ] || [ streamk.h -nt _obj/$f.o ] || [ encode_device.cuh -nt _obj/$f.o ] || [ encode_table_device.cuh -nt _obj/$f.o ] || [ ../../include/mpstime_b200.h -nt _obj/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o _obj/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libmpstime_b200.so _obj/*.o -ldl
echo built ../libmpstime_b200.so
