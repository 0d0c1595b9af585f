#!/bin/bash
# A/B builds of the symmetric stepper: tools/_build/variants/lib_<name>.so, selected with NB_LIB_PATH.
# usage: tools/build_sym_variants.sh name:"-DNB_SYM_I=3 ..." ...
set -e
cd "$(dirname "$0")/.."
PKG=nthu_ipc_nbody-simulation_b200
OUT=tools/_build/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags \
     -c -o $OUT/nb_sym_$name.o $PKG/csrc/nb_sym.cu 2> $OUT/nb_sym_$name.ptxas.log
  objs=$(ls $PKG/_build/*.o | grep -v nb_sym.o)
  nvcc $ARCH -shared -o $OUT/lib_$name.so $objs $OUT/nb_sym_$name.o -lcudart_static -lpthread -ldl -lrt
  echo "$name: $flags :: $(grep -A1 sym_accel $OUT/nb_sym_$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -B1 'Used' $OUT/nb_sym_$name.ptxas.log | grep -A1 -B0 'spill' | grep -o '[0-9]* bytes spill stores' | sort -u | tr '\n' ' ')"
done
