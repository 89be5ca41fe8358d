# 4-GPU validation: slab-mode parity tests at worlds 2, 3, 4, then the bench at N=4 (weak C3)
python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/n4_pytest.log 2>&1; tail -3 gpurun_out/n4_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n4_bench_c3.json 2> gpurun_out/n4_bench_c3.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/n4_bench_c3.json; grep -o '"slab_check": {[^}]*}' gpurun_out/n4_bench_c3.json; tail -3 gpurun_out/n4_bench_c3.err
