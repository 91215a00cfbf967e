#!/bin/bash
# final tree of round 2: ncu --set full captures of the two contraction kernels at cfg2 (one launch each, after warm-up
# launches; text exports) and the launch list of the same workload
mkdir -p gpurun_out /tmp/ncu
python tools/prof_driver.py cfg2 3 > gpurun_out/prof_driver_plain.log 2>&1 || { echo "workload failed without ncu"; tail -5 gpurun_out/prof_driver_plain.log; exit 1; }
cap() {  # name command kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c 1 -f -o /tmp/ncu/$1 $2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page details > gpurun_out/r02c_ncu_$1_details.txt 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page source --csv > gpurun_out/r02c_ncu_$1_source.csv 2>&1
  python tools/ncu_source_top.py gpurun_out/r02c_ncu_$1_source.csv 30 > gpurun_out/r02c_ncu_$1_source_top.txt 2>&1
  ls -la /tmp/ncu/$1.ncu-rep
}
cap estep "python tools/prof_driver.py cfg2 3" "estep_umma_kernel" 2
cap gram "python tools/prof_driver.py cfg2 3" "gram_umma_kernel<64, true, true, true>|gram_umma_kernel<64, 1, 1, 1>" 2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"estep|gram|niw" -c 60 --csv --log-file gpurun_out/r02c_launches_cfg2.csv python tools/prof_driver.py cfg2 3 > gpurun_out/ncu_ll.log 2>&1
grep -h "Duration\|DRAM Throughput\|dram__bytes_read.sum \|dram__bytes_write.sum \|Tensor" gpurun_out/r02c_ncu_*_details.txt | head -30
