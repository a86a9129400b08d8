mkdir -p gpurun_out
T=r02j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_graphed.py tests/test_gpu_slab.py -m gpu -x -q -k "checkpointing or reproducible or gradient_stream or graphed or world1 or halo_row or matches_oracle" > gpurun_out/${T}_pytest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02j_bench_config3.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("edge_stream",{}).get("buffers"), d.get("peak_memory_gib"))
except Exception as e: print("ERR", e)
PY
