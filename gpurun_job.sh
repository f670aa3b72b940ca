timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err
