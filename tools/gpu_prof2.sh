#!/bin/bash
# ncu --set full captures (one launch each, after warm-up launches), text exports; second half of round 2
mkdir -p gpurun_out /tmp/ncu
cap() {  # name command kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c 1 -f -o /tmp/ncu/$1 $2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page details > gpurun_out/r02b_ncu_$1_details.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv > gpurun_out/r02b_ncu_$1_source.csv 2>&1
  ls -la /tmp/ncu/$1.ncu-rep
}
cap gram "python tools/prof_driver.py cfg2 3" "gram_umma_kernel<64, true, true, true>|gram_umma_kernel<64, 1, 1, 1>" 2
cap estep "python tools/prof_driver.py cfg2 3" "estep_umma_kernel" 2
cap rowterm "python tools/time_given.py" "rowterm_umma_kernel" 2
cap wsum "python tools/time_given.py" "gram_umma_kernel" 4
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"estep|gram|niw|wsum|rowterm" -c 60 --csv --log-file gpurun_out/r02b_launches_cfg2.csv python tools/prof_driver.py cfg2 3 > gpurun_out/ncu_ll.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"estep|gram|mnw|wsum|rowterm|rowgemm" -c 80 --csv --log-file gpurun_out/r02b_launches_given.csv python tools/time_given.py > gpurun_out/ncu_ll2.log 2>&1
grep -h "Duration\|DRAM Throughput\|Tensor.*Active\|Pipe Tensor" gpurun_out/r02b_ncu_*_details.txt | head -40
