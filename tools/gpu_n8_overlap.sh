# A/B at 8 GPUs: the r exchange of the downward leg on a second stream (HGRID_OVERLAP=1)
export HGRID_OVERLAP=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-one-shot > gpurun_out/n8_overlap_bench_c3.json 2> gpurun_out/n8_overlap_bench_c3.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/n8_overlap_bench_c3.json; grep -o '"slab_check": {[^}]*}' gpurun_out/n8_overlap_bench_c3.json; tail -3 gpurun_out/n8_overlap_bench_c3.err
