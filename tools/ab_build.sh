#!/bin/bash
# A/B builds of the tiled kernel: tools/ab_build.sh NAME "-DSTAR2_X=1 ..." [radii, default "1 2 3 4"]
# recompiles the k_star2 instantiation units (and kernel_star.cu) with the extra defines into /tmp/deo_ab/NAME/ and links
# them with the other objects of csrc/build/ into /tmp/deo_ab/NAME/libdeo_b200.so (copied to ab/NAME.so, which travels to the GPU box; select it with DEO_LIB_PATH).
set -e
cd "$(dirname "$0")/.."
NAME=$1; DEFS=$2; RADII=${3:-"1 2 3 4"}
OUT=/tmp/deo_ab/$NAME; mkdir -p $OUT
CS=diffeqoperators.jl_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC"
pids=()
for r in 1 2 3 4; do
  if [[ " $RADII " == *" $r "* ]]; then
    nvcc $FLAGS $DEFS -c -o $OUT/star2_inst_r$r.o $CS/star2_inst_r$r.cu & pids+=($!)
  else
    cp $CS/build/star2_inst_r$r.o $OUT/
  fi
done
nvcc $FLAGS $DEFS -c -o $OUT/kernel_star.o $CS/kernel_star.cu & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for o in $CS/build/*.o; do b=$(basename $o); [[ -f $OUT/$b ]] && OBJS="$OBJS $OUT/$b" || OBJS="$OBJS $o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $OUT/libdeo_b200.so $OBJS -ldl
mkdir -p ab && cp $OUT/libdeo_b200.so ab/$NAME.so
ls -la $OUT/libdeo_b200.so
