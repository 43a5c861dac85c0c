#!/bin/bash
# Build an alternate libfhe_sign_cuda.so with extra nvcc flags (A/B experiments inside one gpurun call):
#   tools/build_alt.sh -DFSC_TORUS32_DADD   ->  fhe_sign_b200/lib/alt/libfhe_sign_cuda.so
set -e
cd "$(dirname "$0")/../fhe_sign_b200/csrc"
rm -rf /tmp/fsc_alt && mkdir -p /tmp/fsc_alt ../lib/alt
for f in fsc_api bsk_exact pbs_kernel pbs_stream_kernel pbs_split_kernel pbs_solo_kernel pbs_quad_kernel pbs_duo_kernel ks_kernel ks_mma_kernel ks_umma_kernel linear_kernels radix_cuda; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 "$@" -c $f.cu -o /tmp/fsc_alt/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/alt/libfhe_sign_cuda.so /tmp/fsc_alt/*.o radix.o radix_capi.o client.o keyfile.o -cudart shared -lpthread -ldl
ls -la ../lib/alt/libfhe_sign_cuda.so
