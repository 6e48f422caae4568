/* A plain-C host of the C-ABI (include/hassaku_b200.h): what a non-Python caller of the library looks like.
 * Without a GPU it only proves that the header is C99, that the library links and that the CUDA-free entry points work
 * (tests/test_cabi_and_host.py compiles and runs it):
 *     gcc -std=c99 -I include examples/c_abi_link_check.c -L hassaku_b200/lib -lhassaku_b200 -Wl,-rpath,$PWD/hassaku_b200/lib
 * With device pointers from cudaMalloc the same calls run the kernels (see INTEGRATION.md for the argument meaning). */
#include <stdio.h>
#include <string.h>

#include "hassaku_b200.h"

int main(void) {
    struct hsk_mf_tables t;
    memset(&t, 0, sizeof t);
    printf("hsk_version %d\n", hsk_version());
    /* struct sizes: tests/test_cabi_and_host.py compares them with the ctypes mirrors of hassaku_b200/_C.py */
    printf("sizeof hsk_mf_tables=%d hsk_row_segment=%d hsk_peer_items=%d hsk_peer_flags=%d\n", (int)sizeof(struct hsk_mf_tables),
           (int)sizeof(struct hsk_row_segment), (int)sizeof(struct hsk_peer_items), (int)sizeof(struct hsk_peer_flags));
    printf("kpad(d=256, bf16) = %d\n", hsk_eval_tc_kpad(256, HSK_PREC_BF16));
    printf("scratch(Be=8192, I=1000000, k=100) = %lld bytes\n", (long long)hsk_eval_topk_scratch_bytes(8192, 1000000, 100));
    /* argument validation happens before any CUDA call: a null table is rejected with a message, nothing is launched */
    int rc = hsk_mf_scores(&t, NULL, NULL, 4, 3, NULL, NULL, NULL);
    printf("hsk_mf_scores(null tables) -> %d: %s\n", rc, hsk_last_error());
    return rc < 0 ? 0 : 1;
}
