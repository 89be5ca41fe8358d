# profiles of the headline workload at HEAD: launch list with DRAM bytes, then `ncu --set full` of one launch of every dominant kernel
tag=${1:-r2y}
P="python tools/profile_step.py 16384"
$P 2 > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $P 2 > gpurun_out/${tag}_launches.log 2>&1
cap() {  # name, kernel regex, skip, count, source page?
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/${tag}_$1 $P 1 > gpurun_out/${tag}_ncu_$1.log 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  if [ "$5" = src ]; then ncu -i /tmp/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_$1_source.csv 2>/dev/null; fi
}
cap sl 'sl_(lap|down|up)_kernel' 0 8 src
cap csdown cs_leg_kernel 0 1 src
cap csup cs_leg_kernel 17 1 src
cap vec 'cg_p?update_tile_kernel' 0 2 nosrc
ls -la gpurun_out | grep ${tag}
