mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"estep_umma_kernel" --launch-skip 2 -c 1 -f -o /tmp/ncu/estep python tools/prof_driver.py cfg2 3 > gpurun_out/ncu_estep.log 2>&1
ncu -i /tmp/ncu/estep.ncu-rep --page details > gpurun_out/s3_ncu_estep_quad_details.txt 2>&1
ncu -i /tmp/ncu/estep.ncu-rep --page source --csv > gpurun_out/s3_ncu_estep_quad_source.csv 2>&1
grep -E "Duration|Issue Slots Busy|Executed Ipc|No Eligible|L1/TEX Hit|Mem Pipes Busy|shared" gpurun_out/s3_ncu_estep_quad_details.txt | head -20
