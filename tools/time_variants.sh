#!/bin/bash
# usage: tools/time_variants.sh lib1.so lib2.so ...   (GPU box) — kernel time of the lane kernel per library build
for lib in "$@"; do
  echo "== $lib"
  B200MPC_LIB=$PWD/$lib B=${B:-1048576} NREP=2 KIND=${KIND:-lane} VAR=${VAR:-B} python tools/prof_solve.py | grep "kernel ms" | tail -n 1
done
