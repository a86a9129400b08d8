mkdir -p gpurun_out
O=gpurun_out/r02u_repro.log
: > $O
run() { echo "== $*" >> $O; timeout 300 "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run python tools/repro_diag.py --grad-stream fp32
run python tools/repro_diag.py --grad-stream bf16
run python tools/repro_diag.py --grad-stream bf16 --M 1
run env CGNN_NO_T1=1 python tools/repro_diag.py --grad-stream bf16
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3g
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3g --no-de-next
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3
run python tools/stress_edge_bwd.py --n 20000 --halo 0 --k 16 --reps 20 --precision bf16x3 --no-de-next
cat $O
