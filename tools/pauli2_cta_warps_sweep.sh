# k_mle_rrr_pauli2: warps per CTA (8 thread-per-sample + 4 / 2 / 1 W workers) = register budget of the thread-per-sample
# mapping (168 / 204 / 224 registers).  Rebuilds mle_pauli2.o on the GPU box for every setting.
cd $GRAFT_REPO_ROOT
for w in 12 10 9; do
  (cd quantpy_b200/csrc && make -B mle_pauli2.o EXTRA="-DPAULI_CTA_WARPS=$w -Xptxas -v" 2>&1 | grep -A2 "k_mle_rrr_pauli2ILb1ELb0" | grep -E "registers|spill" | tr '\n' ' ' && make > /dev/null 2>&1)
  echo "warps per CTA $w"; python tools/pauli2_sweep_n.py 100000,12500 2>&1 | grep "^B="
  python bench.py --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench', d['value'], d['ms_per_step'], d['kernels_ms_alone'])"
done
(cd quantpy_b200/csrc && make -B mle_pauli2.o > /dev/null 2>&1 && make > /dev/null 2>&1)
