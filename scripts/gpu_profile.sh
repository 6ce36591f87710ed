#!/bin/bash
# ncu passes (1 GPU): launch list of one bench-like run, then --set full of every kernel of one step.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 26 -c 52 --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py 3 > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -s 26 -c 26 -f -o gpurun_out/prof_step \
    python scripts/profile_step.py 2 > gpurun_out/prof_step.log 2>&1
ls -la gpurun_out/
