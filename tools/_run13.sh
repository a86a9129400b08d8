mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_graphed.py tests/test_gpu_rollout.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "tests rc=$?"
timeout 300 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_config2.json 2> gpurun_out/r02l_bench_config2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_config3.json 2> gpurun_out/r02l_bench_config3.err; echo "bench3 rc=$?"
CGNN_REORDER=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_config3_noreorder.json 2> gpurun_out/r02l_bench_config3_noreorder.err; echo "bench3 noreorder rc=$?"
