#!/bin/bash
# Round-2 closing run on one B200 (under gpurun): the whole GPU suite, the bench line of every BASELINE config + the reference arm,
# the ncu launch list of the bench command and one `--set full` capture of the roofline kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final.log
./tools/r02_measure.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
if [ "${1:-}" = "sweep" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns_price_sweep_kernel -c 3 -f -o gpurun_out/r02_sweep \
    python tools/profile_cmd.py 20 300 > gpurun_out/r02_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
ncu -i gpurun_out/r02_sweep.ncu-rep --page raw --csv > gpurun_out/r02_sweep_ncu_raw.csv 2>/dev/null; echo "export rc=$?"
fi
