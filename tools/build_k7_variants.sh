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
if [ -f /tmp/base/csrc/sensor.cu ]; then cp /tmp/base/csrc/sensor.cu build/variants/sensor_prev.cu; build prev build/variants/sensor_prev.cu & pids="$pids $!"; fi
build cur sensor.cu & pids="$pids $!"
build floor sensor.cu -DQMCL_SENSOR_FLOOR_FMA=1 & pids="$pids $!"
build wide sensor.cu -DQMCL_SENSOR_WIDE_ADDR=1 & pids="$pids $!"
build floorwide sensor.cu -DQMCL_SENSOR_FLOOR_FMA=1 -DQMCL_SENSOR_WIDE_ADDR=1 & pids="$pids $!"
for p in $pids; do wait $p; done
ls -la ../variants
