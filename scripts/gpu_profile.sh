#!/bin/bash
# ncu passes (1 GPU): launch list, decoder, bi-GRU, the 21 feed-forward contractions of a step, the tcgen05 weight-gradient kernel.
# profile_step.py runs 3 eager steps in the default precision (fp32x3); per kernel family the launches of
# the first two steps (setup + warm) are skipped and the third is captured.  Outputs: gpurun_out/${TAG}_*.
TAG=${1:-r02}
mkdir -p gpurun_out
P="python scripts/profile_step.py 3"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $P > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_decoder $P > gpurun_out/${TAG}_prof_decoder.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bigru_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_gru $P > gpurun_out/${TAG}_prof_gru.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 42 -c 21 -f -o gpurun_out/${TAG}_prof_gemm $P > gpurun_out/${TAG}_prof_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dw_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_dw python scripts/profile_dw.py 2 > gpurun_out/${TAG}_prof_dw.log 2>&1
ls -la gpurun_out/${TAG}_*
