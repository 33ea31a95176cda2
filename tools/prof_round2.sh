set -x
cd $GRAFT_REPO_ROOT
# 1. launch list of the bench step (after the plain run exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_b.log 2>&1
# 2. pauli2 MLE inside the ordered fused bootstrap
python tools/prof_mle.py 1000 1e-6 > gpurun_out/p_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mle_rrr_pauli2 -s 1 -c 1 -o gpurun_out/r2b_pauli_tol -f python tools/prof_mle.py 1000 1e-6 > gpurun_out/ncu_p.log 2>&1
ncu -i gpurun_out/r2b_pauli_tol.ncu-rep --page raw --csv > gpurun_out/r2b_pauli_tol.csv
# 3. DMMA GEMM at C3 (linear inversion)
python tools/prof_gemm.py 3 100000 > gpurun_out/g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_counts_dmma -s 2 -c 1 -o gpurun_out/r2b_gemm_c3 -f python tools/prof_gemm.py 3 100000 > gpurun_out/ncu_g.log 2>&1
ncu -i gpurun_out/r2b_gemm_c3.ncu-rep --page raw --csv > gpurun_out/r2b_gemm_c3.csv
# 4. tiled general-POVM MLE at n = 4: both GEMM modes and the update kernel
python tools/prof_tiled_one.py 4 10000 4 > gpurun_out/t_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_counts_dmma|k_rrr_update_mat" -s 3 -c 3 -o gpurun_out/r2b_tiled_n4 -f python tools/prof_tiled_one.py 4 10000 4 > gpurun_out/ncu_t.log 2>&1
ncu -i gpurun_out/r2b_tiled_n4.ncu-rep --page raw --csv > gpurun_out/r2b_tiled_n4.csv
ls -la gpurun_out/*.ncu-rep
