# the driver's round-end sequence on one GPU (add `configs` for the full-size secondary configurations)
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 700 gpurun_out/final_bench.json
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
if [ "$1" = configs ]; then python tools/run_configs.py c2 c2b c4 > gpurun_out/final_configs.jsonl 2> gpurun_out/final_configs.err; cat gpurun_out/final_configs.jsonl | cut -c1-400; fi
