python -m pytest tests/test_gpu_split.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s9_tests.log
python bench.py --steps 50 --warmup 3 --repeats 3 --no-cpu-baseline --no-torch-gpu-baseline --single-precision --no-extras > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
