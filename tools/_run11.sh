mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02j_pytest_tc.log 2>&1; echo "tc tests rc=$?"
echo "== bwd gather4 k=16" >> gpurun_out/r02j_stamps.txt; timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02j_stamps.txt 2>&1
echo "== bwd no gather4 k=16" >> gpurun_out/r02j_stamps.txt; CGNN_GATHER4=0 timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02j_stamps.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_config2.json 2> gpurun_out/r02j_bench_config2.err; echo "bench2 rc=$?"
CGNN_GATHER4=0 timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_config2_nog4.json 2> gpurun_out/r02j_bench_config2_nog4.err; echo "bench2 nog4 rc=$?"
