#!/bin/bash
# Build a tuning variant of libdxb200.so:  tools/build_variant.sh NAME "-DDX_INGEST_U=4 ..."
# -> diffxpy_b200/variants/libdxb200_NAME.so  (objects under /tmp; the product library is untouched)
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${SRC:-$ROOT/diffxpy_b200/csrc}
OUT=$ROOT/diffxpy_b200/variants
OBJ=/tmp/dxvar_$NAME
mkdir -p $OUT $OBJ
for f in dx_api dx_ingest dx_fit_grouped dx_fit_lane dx_stats dx_overflow dx_exchange; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
    $EXTRA -c $SRC/$f.cu -o $OBJ/$f.o 2> $OBJ/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libdxb200_$NAME.so $OBJ/*.o -lcudart
grep -h -A1 "Compiling entry function.*\(ingest_kernel\|fit_grouped_kernel\|fit_tile\)" $OBJ/*.log | grep -v "^--" >/dev/null || true
for k in dx_ingest_kernel dx_fit_grouped_kernel; do
  grep -h -A3 "Compiling entry function.*$k" $OBJ/*.log | grep "Used" | sed "s/^/$NAME $k: /"
done
