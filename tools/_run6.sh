mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q > gpurun_out/r02e_pytest_slab2.log 2>&1; echo "slab tests rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload config2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_config2_n2.json 2> gpurun_out/r02e_bench_config2_n2.err; echo "bench2 n2 rc=$?"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02e_bench_config3_n2.json 2> gpurun_out/r02e_bench_config3_n2.err; echo "bench3 n2 rc=$?"
