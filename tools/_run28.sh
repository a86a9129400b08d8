mkdir -p gpurun_out
T=r02g
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/${T}_pytest_slab.log 2>&1; echo "slab tests rc=$?"; tail -3 gpurun_out/${T}_pytest_slab.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_config3_slab2.json 2> gpurun_out/${T}_bench_config3_slab2.err; echo "bench slab2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02g_bench_config3_slab2.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d.get("parity"))
except Exception as e: print("ERR", e)
PY
tail -5 gpurun_out/${T}_bench_config3_slab2.err
