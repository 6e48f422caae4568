// The LEAN instantiation of the tensor-core evaluator kernel (see the header comment of hsk_eval_tc.cu): the same source
// compiled with HSK_TC_LEAN = 1 — only the kernel, under the name eval_topk_tc_lean_kernel, and its launcher.
// Opt-in at run time with HSK_EVAL_TC=lean; not the default until it has been measured on a B200.
#define HSK_TC_LEAN 1
// the two __noinline__ device helpers have host-side stubs with external linkage: give this copy its own names
#define tc_cut_row tc_cut_row_lean
#define tc_final_sort tc_final_sort_lean
#include "hsk_eval_tc.cu"
