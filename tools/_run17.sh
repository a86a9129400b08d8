mkdir -p gpurun_out
( for v in 0 2 3; do echo "== CGNN_NO_T1=$v"; CGNN_NO_T1=$v timeout 300 python tools/stress_edge_bwd.py --halo 0 --n 6000 --reps 40; done ) > gpurun_out/r02p_stress.txt 2>&1
echo done
