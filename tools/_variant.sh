# usage: tools/_variant.sh name "-DBZ_LS_THREADS=256 ..."   builds a variant library next to the product (development aid)
set -e
name=$1; flags=$2
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared $flags -o bzip2_opencl_b200/libbz2b200_$name.so bzip2_opencl_b200/csrc/bz2b200.cu
