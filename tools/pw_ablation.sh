#!/bin/bash
# Ablation of pw_tc3_kernel<64> phases with the -DCNF_DEBUG library (make -C .../csrc debug; results are wrong by design).
# CNF_PW_DBG bits: 1 skip epilogue stores, 2 skip cp.async copies, 4 skip MMAs, 16 skip epilogue (tcgen05.ld + math), 32 skip transform
cp arl_conditional_normalizing_flows_b200/libcnf_dbg.so arl_conditional_normalizing_flows_b200/libcnf.so
for d in 0 1 2 4 16 17 32 34 36 55; do
  echo "CNF_PW_DBG=$d: $(CNF_PW_DBG=$d python tools/bench_pw.py 256 20 2>&1 | grep -E 'which|both' | tr '\n' ' ')"
done
for n in 2 3; do
  echo "CNF_PW_NST=$n: $(CNF_PW_NST=$n python tools/bench_pw.py 256 20 2>&1 | grep -E 'which|both' | tr '\n' ' ')"
done
