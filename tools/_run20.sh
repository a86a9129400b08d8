mkdir -p gpurun_out
T=r02t
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "gradient_stream or edge_phase_backward or reproducible or rows_forward_backward or node_phase_backward" > gpurun_out/${T}_pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/${T}_pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_benched.py -m gpu -x -q -s -k config2_benched > gpurun_out/${T}_pytest_benched.log 2>&1; echo "benched rc=$?"; tail -3 gpurun_out/${T}_pytest_benched.log
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config2.json 2> gpurun_out/${T}_bench_config2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for c in ("config2","config3"):
    try:
        d=json.loads(open(f"gpurun_out/r02t_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d.get("peak_memory_gib"))
    except Exception as e: print(c, "ERR", e)
PY
