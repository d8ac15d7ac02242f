#!/bin/bash
# tools/build_variant.sh NAME 'sed-expression' FILE -- build ac_dsp_b200/lib/variants/libb200dsp_NAME.so from a copy of csrc
# with one file edited by a sed expression: kernel A/B runs on the GPU box pick it up with B2D_LIBRARY=<path>.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXPR=$2; FILE=$3
W=$(mktemp -d)
mkdir -p $W/ac_dsp_b200 $W/include $ROOT/ac_dsp_b200/lib/variants
cp -r $ROOT/ac_dsp_b200/csrc $W/ac_dsp_b200/csrc
cp $ROOT/include/b200dsp.h $W/include/
sed -i -e "$EXPR" $W/ac_dsp_b200/csrc/$FILE
cmp -s $W/ac_dsp_b200/csrc/$FILE $ROOT/ac_dsp_b200/csrc/$FILE && { echo "sed expression changed nothing"; exit 1; }
cd $W/ac_dsp_b200/csrc
SRCS=$(python3 -c "import sys; sys.path.insert(0,'$ROOT'); from ac_dsp_b200 import build as b; print(' '.join(b.SOURCES))")
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o $ROOT/ac_dsp_b200/lib/variants/libb200dsp_$NAME.so $SRCS -ldl
echo built $ROOT/ac_dsp_b200/lib/variants/libb200dsp_$NAME.so
rm -rf $W
