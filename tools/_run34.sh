mkdir -p gpurun_out
T=r02l
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/${T}_pytest_all.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 300 python bench.py --workload config1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_config1.json 2> gpurun_out/${T}_bench_config1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02l_bench_config1.json").read().strip().splitlines()[-1])
    print("config1", d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e: print("ERR", e)
PY
N="--clock-control none --profile-from-start off"
timeout 300 ncu --set full --import-source on $N -k regex:tc_chain_fwd -s 4 -c 1 -o gpurun_out/${T}_lnb_config2 python tools/profile_step.py --workload config2 --phase edge_bwd > gpurun_out/${T}_ncu_lnb.log 2>&1; echo "ncu lnb rc=$?"
