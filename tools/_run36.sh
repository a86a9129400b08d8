mkdir -p gpurun_out
timeout 300 python bench.py --message sender --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_config3_sender.json 2> gpurun_out/r02n_bench_config3_sender.err; echo "sender rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02n_bench_config3_sender.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d["config"]["message"], d.get("peak_memory_gib"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/r02n_bench_config3_sender.err
