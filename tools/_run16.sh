mkdir -p gpurun_out
( echo "== default"; timeout 300 python tools/stress_edge_bwd.py; timeout 300 python tools/stress_edge_bwd.py --halo 0 --n 6000; echo "== CGNN_NO_T1=1"; CGNN_NO_T1=1 timeout 300 python tools/stress_edge_bwd.py; CGNN_NO_T1=1 timeout 300 python tools/stress_edge_bwd.py --halo 0 --n 6000 ) > gpurun_out/r02o_stress.txt 2>&1
echo done
