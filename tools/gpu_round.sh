#!/bin/bash
# One GPU-box visit: tests, benches, launch list.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [steps...]
# Every step is bounded by its own timeout; outputs land in gpurun_out/<tag>_*.
tag=${1:-run}; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tests)     timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "tests rc=$?" ;;
    tests_new) timeout 900 python -m pytest tests/test_gpu_benched.py -m gpu -x -q -s > gpurun_out/${tag}_pytest_benched.log 2>&1; echo "tests_new rc=$?" ;;
    smoke)     timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench2)    timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err; echo "bench2 rc=$?" ;;
    bench3)    timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err; echo "bench3 rc=$?" ;;
    bench3nc)  timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err; echo "bench3 rc=$?" ;;
    ref)       timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?" ;;
    launches2) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches_config2.csv python bench.py --workload config2 --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline > gpurun_out/${tag}_ncu_config2.log 2>&1; echo "launches2 rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/${tag}_smi.txt 2>&1
