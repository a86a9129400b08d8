mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_benched.py -m gpu -q > gpurun_out/r02g_pytest_tc.log 2>&1; echo "tc tests rc=$?"
for k in 16; do
 echo "== test_wait k=$k" >> gpurun_out/r02g_stamps.txt; timeout 120 python tools/chain_stamps.py --k $k >> gpurun_out/r02g_stamps.txt 2>&1
 echo "== test_wait split k=$k" >> gpurun_out/r02g_stamps.txt; CGNN_SPLIT=1 timeout 120 python tools/chain_stamps.py --k $k >> gpurun_out/r02g_stamps.txt 2>&1
done
echo "== bwd k=16" >> gpurun_out/r02g_stamps.txt; timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02g_stamps.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_config2.json 2> gpurun_out/r02g_bench_config2.err; echo "bench2 rc=$?"
CGNN_SPLIT=1 timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_config2_split.json 2> gpurun_out/r02g_bench_config2_split.err; echo "bench2 split rc=$?"
