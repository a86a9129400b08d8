mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest.log 2>&1; echo "tests rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_config3.json 2> gpurun_out/r02c_bench_config3.err; echo "bench3 rc=$?"
