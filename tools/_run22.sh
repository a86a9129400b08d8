mkdir -p gpurun_out
timeout 300 python tools/g16_diag.py 20000 > gpurun_out/r02v_g16diag.log 2>&1; echo rc=$?; cat gpurun_out/r02v_g16diag.log
timeout 300 python tools/g16_diag.py 6000 >> gpurun_out/r02v_g16diag.log 2>&1; echo rc=$?; tail -12 gpurun_out/r02v_g16diag.log
