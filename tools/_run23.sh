mkdir -p gpurun_out
O=gpurun_out/r02w_g16diag2.log
: > $O
run() { echo "== $*" >> $O; timeout 300 "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run python tools/g16_diag2.py bf16x3g none
run python tools/g16_diag2.py bf16x3g none poison
run python tools/g16_diag2.py bf16x3g de_next poison
run python tools/g16_diag2.py bf16x3 none poison
run env CUDA_LAUNCH_BLOCKING=1 python tools/g16_diag2.py bf16x3g none
cat $O
