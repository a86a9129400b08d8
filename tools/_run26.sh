mkdir -p gpurun_out
T=r02z
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "gradient_stream or reproducible or rows_forward_backward or edge_phase_backward" > gpurun_out/${T}_pytest_tc.log 2>&1; echo "tc tests rc=$?"; grep "bf16 gradient stream" gpurun_out/${T}_pytest_tc.log | cut -c1-300; tail -2 gpurun_out/${T}_pytest_tc.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for c in ("config3",):
    try:
        d=json.loads(open(f"gpurun_out/r02z_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d.get("peak_memory_gib"))
    except Exception as e: print(c, "ERR", e)
PY
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/${T}_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -22 gpurun_out/${T}_pytest_all.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step3.csv python tools/profile_step.py --workload config3 > gpurun_out/${T}_ncu_step3.log 2>&1; echo "launches step3 rc=$?"
