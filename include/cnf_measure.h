/* libcnf measurement hooks: used by bench.py and tools/ to time ONE stage of a coupling layer with CUDA events
 * (roofline of the dominant kernel).  Not part of the drop-in surface; replaces nothing in the reference. */
#ifndef CNF_MEASURE_H_
#define CNF_MEASURE_H_

#include "cnf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Launches ONLY one kernel of residual block 0 of the layer-per-kernel path (which = 0: pw1, X -> Y1 with LayerNorm-on-
 * load; which = 1: pw2, Y2 (+X) -> X; which = 2: the grouped dilated convs of all branches, Y1 -> Y2) on the workspace
 * state left by a previous cnf_coupling_nets call with the same batch.  The kernel-path bits of the descriptor apply. */
int cnf_measure_stage(const cnf_coupling* c, const DLManagedTensor* params, DLManagedTensor* workspace,
                      int64_t batch, int which, void* stream);

#ifdef CNF_DEBUG
/* -DCNF_DEBUG builds only: clock64() stamps of CTA 0 of the 1x1-conv kernel (CNF_PW_DBG=128), [role][chunk][8] */
int cnf_debug_read_clocks(long long* out, int n);
#endif

#ifdef __cplusplus
}
#endif
#endif /* CNF_MEASURE_H_ */
