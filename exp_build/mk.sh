#!/bin/bash
# usage: mk.sh name -DFOO=1 ...   -> exp_build/libb2r_name.so
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -shared "$@" cpu-raytracing-experiments_b200/csrc/b2r.cu cpu-raytracing-experiments_b200/csrc/b2r_host.cpp -o exp_build/libb2r_$name.so
