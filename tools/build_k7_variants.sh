#!/usr/bin/env bash
# Builds sensor-kernel variants for tools/k7_ab.py: the library's own objects with sensor.cu
# recompiled under different macros (and, when /tmp/base/csrc/sensor.cu exists, the previous
# revision of the kernel).  Output: quickmcl_b200/variants/lib<name>.so
set -eu
cd "$(dirname "$0")/../quickmcl_b200/csrc"
make > /dev/null
mkdir -p ../variants build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC"
OTHERS=$(ls build/libqmcl_b200/*.o | grep -v sensor.o)
build() {  # name source macros...
  local name=$1 src=$2; shift 2
  nvcc $FLAGS "$@" -I. -c -o build/variants/sensor_$name.o $src
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib$name.so $OTHERS build/variants/sensor_$name.o -ldl -lpthread
}
pids=""
build cur sensor.cu & pids="$pids $!"
build nogather sensor.cu -DQMCL_DIAG=4 & pids="$pids $!"
build l1window sensor.cu -DQMCL_DIAG=5 & pids="$pids $!"
build blocks6 sensor.cu -DQMCL_SENSOR_MINBLOCKS=6 & pids="$pids $!"
build blocks4 sensor.cu -DQMCL_SENSOR_MINBLOCKS=4 & pids="$pids $!"
for p in $pids; do wait $p; done
ls -la ../variants
