#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each, after warm-up launches), exported as text on the box
mkdir -p gpurun_out /tmp/ncu
cap() {  # name workload kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c 1 -f -o /tmp/ncu/$1 python tools/prof_driver.py $2 3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page details > gpurun_out/r02_ncu_$1_details.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>&1
  ls -la /tmp/ncu/$1.ncu-rep
}
cap estep cfg2 "estep_umma_kernel" 2
cap gram cfg2 "gram_umma_kernel<64, true, true, true>|gram_umma_kernel<64, 1, 1, 1>" 2
cap gramswap cfg3 "gram_swap_kernel" 2
cap diag iso "diag_estep_kernel" 2
cap hmm cfg4 "hmm_fb_lin_kernel" 2
# launch list (durations) of one cfg2 iteration incl. the small kernels, and of the zpack pre-pass
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"estep|gram|niw" -c 60 --csv --log-file gpurun_out/r02_launches_cfg2.csv python tools/prof_driver.py cfg2 3 > gpurun_out/ncu_ll.log 2>&1
cp /tmp/ncu/gram.ncu-rep gpurun_out/r02_gram.ncu-rep 2>/dev/null
grep -h "Duration\|DRAM Throughput\|dram__bytes_read.sum \|dram__bytes_write.sum " gpurun_out/r02_ncu_*_details.txt | head -40
