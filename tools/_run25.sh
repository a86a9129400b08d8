mkdir -p gpurun_out
T=r02y
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "gradient_stream or reproducible" > gpurun_out/${T}_pytest_tc.log 2>&1; echo "tc tests rc=$?"; grep "bf16 gradient stream" gpurun_out/${T}_pytest_tc.log; tail -3 gpurun_out/${T}_pytest_tc.log
O=gpurun_out/${T}_repro.log
: > $O
run() { echo "== $*" >> $O; timeout 600 "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run python tools/repro_diag.py --grad-stream bf16 --n 20000 --reps 8
run python tools/repro_diag.py --grad-stream bf16 --n 40000 --k 32 --M 2 --reps 6
grep -v "^rep .*tensors differ: \[" $O | tail -12
timeout 900 python -m pytest tests/test_gpu_benched.py -m gpu -x -q -s -k config2_benched > gpurun_out/${T}_pytest_benched.log 2>&1; echo "benched rc=$?"; tail -3 gpurun_out/${T}_pytest_benched.log
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config2.json 2> gpurun_out/${T}_bench_config2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config3.json 2> gpurun_out/${T}_bench_config3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for c in ("config2","config3"):
    try:
        d=json.loads(open(f"gpurun_out/r02y_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d.get("peak_memory_gib"))
    except Exception as e: print(c, "ERR", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step3.csv python tools/profile_step.py --workload config3 > gpurun_out/${T}_ncu_step3.log 2>&1; echo "launches step3 rc=$?"
