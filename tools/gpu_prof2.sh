tag=${1:-prof}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 16384 2 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cs_leg_kernel -c 18 -o gpurun_out/${tag}_cs python tools/profile_step.py 16384 1 > gpurun_out/${tag}_ncu_cs.log 2>&1
ncu -i gpurun_out/${tag}_cs.ncu-rep --page raw --csv > gpurun_out/${tag}_cs_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_cs.ncu-rep --page source --csv > gpurun_out/${tag}_cs_source.csv 2>/dev/null
rm -f gpurun_out/${tag}_cs.ncu-rep
ls -la gpurun_out | tail -4
