set -x
cd $GRAFT_REPO_ROOT
# every profiled program first exits 0 without ncu
# 1. launch lists: bench step, end-to-end calls
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_b.log 2>&1
python tools/e2e_launches.py > gpurun_out/e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2c_e2e.csv python tools/e2e_launches.py > gpurun_out/ncu_e.log 2>&1
# 2. pauli2 MLE: uniform launch (thread-per-sample only, 100 iterations each: instruction counts), and inside the ordered fused step
QPB_NO_TAIL_MERGE=1 python tools/prof_mle.py 100 0 > gpurun_out/pu_plain.log 2>&1 && \
QPB_NO_TAIL_MERGE=1 ncu --set full --clock-control none --import-source on -k regex:k_mle_rrr_pauli2 -s 1 -c 1 -o gpurun_out/r2c_pauli_uniform -f python tools/prof_mle.py 100 0 > gpurun_out/ncu_pu.log 2>&1
ncu -i gpurun_out/r2c_pauli_uniform.ncu-rep --page raw --csv > gpurun_out/r2c_pauli_uniform.csv
python tools/prof_mle.py 1000 1e-6 > gpurun_out/p_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mle_rrr_pauli2 -s 1 -c 1 -o gpurun_out/r2c_pauli_tol -f python tools/prof_mle.py 1000 1e-6 > gpurun_out/ncu_p.log 2>&1
ncu -i gpurun_out/r2c_pauli_tol.ncu-rep --page raw --csv > gpurun_out/r2c_pauli_tol.csv
# 3. conditional-binomial sampler with the float32 prefilter
python tools/prof_sampler.py 2 100000 > gpurun_out/s_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_multinomial_binomial -s 3 -c 1 -o gpurun_out/r2c_btrs -f python tools/prof_sampler.py 2 100000 > gpurun_out/ncu_s.log 2>&1
ncu -i gpurun_out/r2c_btrs.ncu-rep --page raw --csv > gpurun_out/r2c_btrs.csv
# 4. sample sort of 1e5 distances (four kernels of the third end-to-end call)
ncu --set full --clock-control none --import-source on -k regex:k_ss_ -s 8 -c 4 -o gpurun_out/r2c_sort -f python tools/e2e_launches.py > gpurun_out/ncu_ss.log 2>&1
ncu -i gpurun_out/r2c_sort.ncu-rep --page raw --csv > gpurun_out/r2c_sort.csv
ls -la gpurun_out/*.ncu-rep
