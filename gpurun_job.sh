timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mbd or ranks or band" 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err
timeout 300 python bench.py --steps 10 --warmup 3 --ties --no-cpu-baseline > gpurun_out/bench18t.json 2> gpurun_out/bench18t.err
