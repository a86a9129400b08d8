mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02h_pytest_tc.log 2>&1; echo "tc tests rc=$?"
echo "== bwd T1 k=16" >> gpurun_out/r02h_stamps.txt; timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02h_stamps.txt 2>&1
echo "== bwd noT1 k=16" >> gpurun_out/r02h_stamps.txt; CGNN_NO_T1=1 timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02h_stamps.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_config2.json 2> gpurun_out/r02h_bench_config2.err; echo "bench2 rc=$?"
CGNN_NO_T1=1 timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_config2_not1.json 2> gpurun_out/r02h_bench_config2_not1.err; echo "bench2 noT1 rc=$?"
