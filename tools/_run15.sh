mkdir -p gpurun_out
for v in default not1; do
  if [ $v = not1 ]; then export CGNN_NO_T1=1; fi
  timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q -k "world2" > gpurun_out/r02n_pytest_$v.log 2>&1; echo "$v rc=$?"
done
