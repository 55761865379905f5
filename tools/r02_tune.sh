#!/bin/bash
# Team-engine constants on bounded 2^20 solves: pricing CTAs, replicas of the ENTER / CYC records (rebuilds libmcfgpu.so on the box).
run() { echo "$1: $(timeout 120 python tools/profile_team.py 20 300000 $2 2>&1 | tail -1 | cut -c1-60)"; }
for np in 4 8 12 16; do run "pricers=$np" $np; done
for v in "MCF_REP_CYC=3" "MCF_REP_CYC=4" "MCF_REP_ENT=2" "MCF_REP_ENT=1 -DMCF_REP_CYC=4"; do
  touch mincostflow_b200/csrc/mcf_team.cu; make -s -C mincostflow_b200/csrc EXTRA="-D$v" > /dev/null 2>&1 || echo "build failed $v"
  run "$v" ""
done
