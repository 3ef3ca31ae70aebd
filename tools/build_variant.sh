#!/bin/sh
# A/B builds of the CUDA library: tools/build_variant.sh NAME "-DKNOB=VALUE ..."
# -> satrans_b200/lib/libsatrans_b200_NAME.so (objects in build_NAME/); select it at run time with SAT_LIB=<path>.
set -e
name=$1; shift
cd "$(dirname "$0")/../satrans_b200/csrc"
make -j"$(nproc)" B=../../build_$name LIB=../lib/libsatrans_b200_$name.so EXTRA="$*" >/dev/null
echo "built satrans_b200/lib/libsatrans_b200_$name.so"
