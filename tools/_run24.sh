mkdir -p gpurun_out
O=gpurun_out/r02x_repro.log
: > $O
run() { echo "== $*" >> $O; timeout 600 "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3g --no-de-next
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3g
run python tools/g16_diag2.py bf16x3g none poison
run env CUDA_LAUNCH_BLOCKING=1 python tools/g16_diag2.py bf16x3g none
run python tools/repro_diag.py --grad-stream bf16
run env CGNN_T1_LNB=1 python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 30 --precision bf16x3g
run env CGNN_T1_LNB=1 python tools/stress_edge_bwd.py --n 3000 --halo 600 --k 16 --reps 40 --precision bf16x3
run env CGNN_T1_LNB=1 python tools/stress_edge_bwd.py --n 6000 --halo 0 --k 16 --reps 40 --precision bf16x3
run env CGNN_T1_LNB=1 python tools/repro_diag.py --grad-stream bf16 --reps 8
grep -v "^rep .*tensors differ: \[" $O | tail -40
timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r02x_pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r02x_pytest_tc.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_config3.json 2> gpurun_out/r02x_bench_config3.err; echo "bench3 rc=$?"
CGNN_T1_LNB=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_config3_t1lnb.json 2> gpurun_out/r02x_bench_config3_t1lnb.err; echo "bench3 t1lnb rc=$?"
python - <<'PY'
import json
for c in ("config3","config3_t1lnb"):
    try:
        d=json.loads(open(f"gpurun_out/r02x_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["value"], d["phases"]["forward"]["ms"], d["phases"]["backward"]["ms"])
    except Exception as e: print(c, "ERR", e)
PY
