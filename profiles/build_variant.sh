#!/bin/bash
# Build a variant of libpgf_b200.so for kernel A/B runs: usage  build_variant.sh <name> "<extra nvcc flags>" [files...]
# Only the listed .cu files (default: pipeline_inst_probe.cu) are recompiled with the extra flags; the result is
# pg_fusion_b200/variants/libpgf_b200_<name>.so, selected at run time with PGF_B200_LIB.
set -e
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-pipeline_inst_probe.cu}
cd "$(dirname "$0")/../pg_fusion_b200/csrc"
make -s -j8 > /dev/null
mkdir -p build_w$NAME ../variants
OBJS=""
for f in build/*.o; do OBJS="$OBJS $f"; done
for f in $FILES; do
  o=build_w$NAME/${f%.cu}.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $FLAGS -c $f -o $o 2> build_w$NAME/${f%.cu}.ptxas.log
  OBJS=$(echo $OBJS | sed "s#build/${f%.cu}.o#$o#")
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libpgf_b200_$NAME.so $OBJS -cudart static -ldl
echo "built pg_fusion_b200/variants/libpgf_b200_$NAME.so"
