#!/bin/bash
# runs tools/sym_only.py once per A/B build of the library (tools/build_sym_variants.sh)
for so in tools/_build/variants/lib_*.so; do
  name=$(basename $so .so)
  echo "== $name: $(NB_LIB_PATH=$PWD/$so python tools/sym_only.py 65536 6 2>&1 | tail -1)"
done
