#!/bin/bash
# ncu captures of the predict kernels (one launch each after warm-up), text exports
mkdir -p gpurun_out /tmp/ncu
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c 1 -f -o /tmp/ncu/$1 $2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page details > gpurun_out/r02b_ncu_$1_details.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv > gpurun_out/r02b_ncu_$1_source.csv 2>&1
}
cap rowwide "python tools/time_predict.py" "rowwide_umma_kernel" 4
cap moemma "python tools/time_predict.py" "moe_moments_mma_kernel" 2
cap gramswap "python tools/prof_driver.py cfg3 3" "gram_swap_kernel" 2
grep -h "Duration\|DRAM Throughput\|SM Frequency" gpurun_out/r02b_ncu_rowwide_details.txt gpurun_out/r02b_ncu_moemma_details.txt gpurun_out/r02b_ncu_gramswap_details.txt
