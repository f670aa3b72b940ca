timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err
timeout 300 python bench.py --steps 10 --warmup 3 --ties --no-cpu-baseline > gpurun_out/bench19t.json 2> gpurun_out/bench19t.err
