mkdir -p gpurun_out
T=r02k
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_slab.py -m gpu -x -q -k "padded_shapes or world or halo_row or narrow" > gpurun_out/${T}_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_config3_slab2.json 2> gpurun_out/${T}_bench_config3_slab2.err; echo "bench slab2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02k_bench_config3_slab2.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d.get("parity",{}).get("worst_gradient_rel_l2"), d.get("edge_stream",{}).get("buffers"))
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/${T}_bench_config3_slab2.err
