mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02i_bench_config3_n8.json 2> gpurun_out/r02i_bench_config3_n8.err; echo "bench3 n8 rc=$?"
