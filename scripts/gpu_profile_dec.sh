#!/bin/bash
mkdir -p gpurun_out
P="python scripts/profile_step.py 3"
ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -s 2 -c 1 -f -o gpurun_out/prof_decoder $P > gpurun_out/prof_decoder.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bigru_kernel -s 5 -c 1 -f -o gpurun_out/prof_gru $P > gpurun_out/prof_gru.log 2>&1
ls -la gpurun_out/*.ncu-rep
