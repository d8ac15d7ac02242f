#!/bin/bash
# tools/facade_throughput.sh [build|run] -- the cost of the reference-signature (ac_channel) calls, facade vs reference.
#   build (dev container: needs /root/reference for the reference twin) -> oracle/_ref/facade_throughput_{b200,ref}
#   run   (GPU box)                                                      -> JSON lines on stdout
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${AC_DSP_REF:-/root/reference}
OUT=$ROOT/oracle/_ref
if [ "${1:-run}" = build ]; then
  mkdir -p $OUT
  g++ -std=c++11 -O2 -DIMPL='"facade (B200 engine)"' -I$ROOT/include/b200dsp -I$ROOT/oracle/ac_shim $ROOT/tests/cpp/facade_throughput.cpp \
      -L$ROOT/ac_dsp_b200/lib -lb200dsp -Wl,-rpath,'$ORIGIN/../../ac_dsp_b200/lib' -o $OUT/facade_throughput_b200 || exit 1
  if [ -d $REF/include/ac_dsp ]; then
    g++ -std=c++11 -O3 -march=native -DIMPL='"reference templates (CPU, 1 thread)"' -I$ROOT/oracle/ac_shim -I$REF/include $ROOT/tests/cpp/facade_throughput.cpp \
        -o $OUT/facade_throughput_ref || exit 1
  fi
  exit 0
fi
$OUT/facade_throughput_b200 ${2:-1048576} ${3:-4096}
[ -x $OUT/facade_throughput_ref ] && $OUT/facade_throughput_ref ${2:-1048576} ${3:-4096}
