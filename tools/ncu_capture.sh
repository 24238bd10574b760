#!/bin/bash
# ncu captures of the dominant kernels on the bench workloads (run on the GPU box, one GPU; each program has already
# exited 0 without ncu in the same call).  Reports land in gpurun_out/; tools/make_calibration.py turns them into
# profiles/<tag>_calibration.json and the per-kernel summaries.
#   usage: bash tools/ncu_capture.sh <tag> [n1] [rhsjac] [n2] [n2wf]
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
for what in "$@"; do
  case $what in
    n1)
      B=1048576 METHOD=ros4 python tools/variants.py 384,RMT_REDUCED=1 > $OUT/${TAG}_n1_plain.log 2>&1 || { echo "n1 plain run failed"; tail -5 $OUT/${TAG}_n1_plain.log; continue; }
      B=1048576 METHOD=ros4 $NCU -k rmt_n1_solve -s 2 -c 1 -f -o $OUT/${TAG}_n1_solve python tools/variants.py 384,RMT_REDUCED=1 > $OUT/${TAG}_n1_ncu.log 2>&1
      ;;
    rhsjac)
      python tools/ncu_rhs_jac.py 4 > $OUT/${TAG}_rhsjac_plain.log 2>&1 || { echo "rhsjac plain failed"; continue; }
      $NCU -k regex:rmt_n1_\(rhs\|jac\) -f -o $OUT/${TAG}_n1_rhs_jac python tools/ncu_rhs_jac.py 2 > $OUT/${TAG}_rhsjac_ncu.log 2>&1
      ;;
    n2wf)
      B=9472 python tools/n2_lanes.py 0,256 > $OUT/${TAG}_n2wf_plain.log 2>&1 || { echo "n2wf plain failed"; continue; }
      B=9472 REPS=2 $NCU -k rmt_n2_solve -s 1 -c 1 -f -o $OUT/${TAG}_n2_solve_pipeline python tools/n2_lanes.py 0,256 > $OUT/${TAG}_n2wf_ncu.log 2>&1
      ;;
    n2)
      python tools/n2_lanes.py 8,64 > $OUT/${TAG}_n2_plain.log 2>&1 || { echo "n2 plain failed"; continue; }
      REPS=2 $NCU -k rmt_n2_solve -s 1 -c 1 -f -o $OUT/${TAG}_n2_solve python tools/n2_lanes.py 8,64 > $OUT/${TAG}_n2_ncu.log 2>&1
      ;;
  esac
done
ls -la $OUT/*.ncu-rep
