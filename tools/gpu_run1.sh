set -x
./build/probes/cp_async_probe | tee gpurun_out/r2v_probe.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multigrid.py tests/test_gpu_baseline_sizes.py -x -q > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench.json'))
print(d['ms_per_step'], d['config']['iterations'], d['config']['relres_true'])
for k in d['kernels']: print(k['phase'], round(k['ms_total'],2), k['launches'], round(k.get('ms_per_launch',0),3), round(k.get('frac',0),3))
PY
ncu --set full --clock-control none --import-source on -k regex:cs_leg_kernel -c 2 -o gpurun_out/r2v_cs python tools/profile_step.py 16384 1 > gpurun_out/r2v_ncu_cs.log 2>&1
ncu -i gpurun_out/r2v_cs.ncu-rep --page raw --csv > gpurun_out/r2v_cs_raw.csv 2>/dev/null
ncu -i gpurun_out/r2v_cs.ncu-rep --page source --csv --print-kernel-base function > gpurun_out/r2v_cs_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
