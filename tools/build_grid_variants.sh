#!/bin/bash
# A/B builds of the grid trajectory kernel: tools/_build/variants/lib_<name>.so, selected with NB_LIB_PATH.
# usage: tools/build_grid_variants.sh name:"-DNB_GRID_SWP=2 ..." ...
set -e
cd "$(dirname "$0")/.."
PKG=nthu_ipc_nbody-simulation_b200
OUT=tools/_build/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v $flags \
     -c -o $OUT/nb_grid_$name.o $PKG/csrc/nb_grid.cu 2> $OUT/nb_grid_$name.ptxas.log
  objs=$(ls $PKG/_build/*.o | grep -v nb_grid.o)
  nvcc $ARCH -shared -o $OUT/lib_$name.so $objs $OUT/nb_grid_$name.o -lcudart_static -lpthread -ldl -lrt
  echo "$name: $flags :: $(grep -A2 'grid_traj_kernelILi0ELi1ELi4ELb0ELb0' $OUT/nb_grid_$name.ptxas.log | grep -o 'Used [0-9]* registers') | T2: $(grep -A2 'grid_traj_kernelILi0ELi2ELi4ELb0ELb0' $OUT/nb_grid_$name.ptxas.log | grep -o 'Used [0-9]* registers') | spills: $(grep -o '[0-9]* bytes spill stores' $OUT/nb_grid_$name.ptxas.log | sort -u | tr '\n' ' ')"
done
