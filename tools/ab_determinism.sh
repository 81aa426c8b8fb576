#!/bin/bash
# run-to-run reproducibility of the middle blocks for library variants.  usage: bash tools/ab_determinism.sh <run-tag> "<T n reps>" base nocf ...
run=$1; args=$2; shift; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset MOCODAD_B200_LIB; else export MOCODAD_B200_LIB=$PWD/mocodad_b200/libmocodad_b200_${tag}.so; fi
  timeout 600 python tools/determinism_taps.py $args > gpurun_out/${run}_${tag}_det.log 2>&1; echo "$tag rc=$?"; tail -12 gpurun_out/${run}_${tag}_det.log
done
