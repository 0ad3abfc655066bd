#!/bin/bash
# tools/ab_run.sh "LIB1 LIB2 ..." "SHAPE APPROX DTYPE; ..." [reps]  -- alternates the libraries (ab/LIB.so, or "main") over the workloads
LIBS=$1; IFS=';' read -ra WL <<< "$2"; REPS=${3:-2}
for rep in $(seq $REPS); do
  for w in "${WL[@]}"; do
    for lib in $LIBS; do
      if [ "$lib" = main ]; then unset DEO_LIB_PATH; else export DEO_LIB_PATH=$PWD/ab/$lib.so; fi
      echo -n "[$lib] "; python tools/sweep.py $w "DEO_STAR_V=2" 2>&1 | tail -1
    done
  done
done
