# 2-GPU validation: slab-mode parity test, then the bench at N=2 (weak C3) and the C5 strong / weak forms
python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/n2_pytest.log 2>&1; tail -3 gpurun_out/n2_pytest.log
for wl in c3 c5-strong c5-weak; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/n2_bench_$wl.json 2> gpurun_out/n2_bench_$wl.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n2_bench_$wl.json'))
    print('$wl', d['ms_per_step'], d['value'], d['config'].get('iterations'), d['config'].get('slab_check'), d.get('e2e',{}).get('value'))
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/n2_bench_$wl.err').read()[-1500:])
PY
done
