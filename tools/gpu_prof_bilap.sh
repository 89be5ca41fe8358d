# ncu --set full of the bilaplacian V-cycle kernels on the C2 pattern 4096^2 fp64: the level-0 launch of each kind
# (skip counts: first launch of a kind is on level 0 on the way down; prolongation reaches level 0 last, 8th of a cycle)
cap() {  # kernel, launches to skip
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o /tmp/bl_$1 python tools/profile_bilap.py 4096 fp64 1e-3 > gpurun_out/bl_ncu_$1.log 2>&1
  ncu -i /tmp/bl_$1.ncu-rep --page raw --csv > gpurun_out/bl_$1_raw.csv 2>/dev/null
}
cap gs25_rows_kernel 1
cap resid25_kernel 0
cap restrict25_kernel 0
cap prolong_add25_kernel 7
cap bilap_apply_kernel 1
ls -la gpurun_out | grep bl_
