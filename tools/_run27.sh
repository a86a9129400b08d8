mkdir -p gpurun_out
T=r02f
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "gradient_stream or reproducible or edge_phase_backward" > gpurun_out/${T}_pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -2 gpurun_out/${T}_pytest_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err; echo "bench3 rc=$?"
timeout 600 python bench.py --workload config2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_config2.json 2> gpurun_out/${T}_bench_config2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for c in ("config2","config3"):
    try:
        d=json.loads(open(f"gpurun_out/r02f_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["value"], d["e2e"]["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(c, "ERR", e)
PY
N="--clock-control none --profile-from-start off"
timeout 600 ncu --metrics gpu__time_duration.sum $N --csv --log-file gpurun_out/${T}_launches_config2.csv python tools/profile_step.py --workload config2 > gpurun_out/${T}_ncu_launches2.log 2>&1; echo "launches2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum $N --csv --log-file gpurun_out/${T}_launches_config3.csv python tools/profile_step.py --workload config3 > gpurun_out/${T}_ncu_launches3.log 2>&1; echo "launches3 rc=$?"
timeout 600 ncu --set full --import-source on $N -k regex:tc_wgrad -c 1 -o gpurun_out/${T}_wgrad_config2 python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --import-source on $N -k regex:scatter_chunk -c 1 -o gpurun_out/${T}_scatter_config2 python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_scatter.log 2>&1; echo "ncu scatter rc=$?"
