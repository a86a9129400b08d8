mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_optim.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "tests rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --rollout --workload config4 --steps 5 --warmup 3 > gpurun_out/r02m_rollout_config4_n2.json 2> gpurun_out/r02m_rollout_config4_n2.err; echo "rollout n2 rc=$?"
timeout 600 python bench.py --rollout --workload config4 --steps 5 --warmup 3 > gpurun_out/r02m_rollout_config4_n1.json 2> gpurun_out/r02m_rollout_config4_n1.err; echo "rollout n1 rc=$?"
