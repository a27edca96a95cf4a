#!/bin/bash
# dev: build libsuperman_b200.so variants that differ only in the LevelRyser kernel's compile-time knobs
# (level_reg.cuh: SPB_LV_SPLIT, SPB_LV_CHMAX, ...) into tools/_bin/<name>/ so that one GPU call can compare them.
#   tools/build_level_variants.sh name1 "-DSPB_LV_SPLIT=0" name2 "-DSPB_LV_SPLIT=1 -DSPB_LV_CHMAX=2" ...
set -e
cd "$(dirname "$0")/.."
make -j16 >/dev/null
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Isuperman_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p build/$name tools/_bin/$name
  for bs in 3s0 3s1 4s0 4s1; do
    b=${bs%s*}; s=${bs#*s}
    nvcc $NVF $flags -DSPB_LV_B=$b -DSPB_LV_SKIP=$s -c superman_b200/csrc/sp_level_inst.cu -o build/$name/sp_level_inst_b$bs.o &
  done
  nvcc $NVF $flags -c superman_b200/csrc/sp_sparse.cu -o build/$name/sp_sparse.o &
  wait
  others=$(ls build/*.o | grep -v "sp_level_inst_b\|/sp_sparse.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_bin/$name/libsuperman_b200.so $others build/$name/*.o -lpthread -lm
  echo "$name: $flags"
  cuobjdump --dump-resource-usage tools/_bin/$name/libsuperman_b200.so 2>/dev/null | grep -A1 "level_reg_kernel" | grep -o "REG:[0-9]* STACK:[0-9]*" | sort | uniq -c | sort -k2 | awk '$3!="STACK:0"'
done
