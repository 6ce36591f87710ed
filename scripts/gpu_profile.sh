#!/bin/bash
# ncu passes (1 GPU).  profile_step.py runs 3 eager steps; per kernel family we skip the launches of the first
# two steps (setup + warm) and capture the third.
mkdir -p gpurun_out
P="python scripts/profile_step.py 3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|bigru|decoder_kernel|maxpool|mask_rows|gather' \
    --csv --log-file gpurun_out/launches.csv $P > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -s 2 -c 1 -f -o gpurun_out/prof_decoder $P > gpurun_out/prof_decoder.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bigru_kernel -s 4 -c 2 -f -o gpurun_out/prof_gru $P > gpurun_out/prof_gru.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 53 -c 10 -f -o gpurun_out/prof_gemm $P > gpurun_out/prof_gemm.log 2>&1
ls -la gpurun_out/
