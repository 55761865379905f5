#!/bin/bash
# resident slices vs flows in global memory ("team_spill"), and the number of pricing CTAs, on bounded Block Search solves
run() { echo "$1: $(timeout 120 python tools/profile_team.py $2 2>&1 | tail -1 | cut -c1-110)"; }
for k in "13 8000" "16 100000" "18 200000" "20 300000"; do for e in team team_spill; do run "$k $e" "$k 0 $e"; done; done
for np in 20 24 28 32; do run "2^20 spill pricers=$np" "20 300000 $np team_spill"; done
for np in 8 12 16 24; do run "2^18 spill pricers=$np" "18 200000 $np team_spill"; done
for np in 4 8 12; do run "2^16 spill pricers=$np" "16 100000 $np team_spill"; done
