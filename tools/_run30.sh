mkdir -p gpurun_out
T=r02i
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/${T}_bench_config3_slab4.json 2> gpurun_out/${T}_bench_config3_slab4.err; echo "bench slab4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02i_bench_config3_slab4.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d.get("parity",{}).get("worst_gradient_rel_l2"), d.get("edge_stream"), d.get("peak_memory_gib"))
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/${T}_bench_config3_slab4.err
