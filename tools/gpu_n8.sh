# 8-GPU validation: slab-mode parity at world 8, then the bench at N=8 (weak C3)
python -m pytest tests/test_gpu_distributed.py -x -q -k "8" > gpurun_out/n8_pytest.log 2>&1; tail -3 gpurun_out/n8_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n8_bench_c3.json 2> gpurun_out/n8_bench_c3.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/n8_bench_c3.json; grep -o '"slab_check": {[^}]*}' gpurun_out/n8_bench_c3.json; tail -3 gpurun_out/n8_bench_c3.err
