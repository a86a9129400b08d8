mkdir -p gpurun_out
T=r02m
for v in "CGNN_SPLIT=1" "CGNN_SPLIT=1 CGNN_NO_RIN=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench_${tag}.json 2> gpurun_out/${T}_bench_${tag}.err; echo "$v rc=$?"
  python - "$tag" <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/r02m_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], d["ms_per_step"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
