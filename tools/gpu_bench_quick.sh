# quick A/B on the GPU box: the multigrid equivalence tests + the headline bench with its per-phase table
tag=${1:-quick}
python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_parity.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print(d['ms_per_step'], d['config']['iterations'], d['config']['relres_true'])
for k in d['kernels']: print(k['phase'], round(k['ms_total'],2), k['launches'], round(k.get('ms_per_launch',0),3), round(k.get('frac',0),3))
PY
