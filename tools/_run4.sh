mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_optim.py -m gpu -x -q > gpurun_out/r02d_pytest_tc.log 2>&1; echo "tc tests rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graphed.py -m gpu -x -q > gpurun_out/r02d_pytest_parity.log 2>&1; echo "parity tests rc=$?"
for k in 16 32; do
 echo "== split k=$k" >> gpurun_out/r02d_stamps.txt; timeout 120 python tools/chain_stamps.py --k $k >> gpurun_out/r02d_stamps.txt 2>&1
 echo "== nosplit k=$k" >> gpurun_out/r02d_stamps.txt; CGNN_NO_SPLIT=1 timeout 120 python tools/chain_stamps.py --k $k >> gpurun_out/r02d_stamps.txt 2>&1
done
echo "== bwd k=16" >> gpurun_out/r02d_stamps.txt; timeout 120 python tools/chain_stamps.py --bwd >> gpurun_out/r02d_stamps.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_config2.json 2> gpurun_out/r02d_bench_config2.err; echo "bench2 rc=$?"
CGNN_NO_SPLIT=1 timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_config2_nosplit.json 2> gpurun_out/r02d_bench_config2_nosplit.err; echo "bench2 nosplit rc=$?"
