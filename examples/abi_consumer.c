/* A plain C consumer of libcmrl_b200.so: no Python, no torch.  It reads the ABI version, shows that argument validation
 * precedes any CUDA call (so it runs on a machine without a GPU) and prints the library's launch counter.
 *
 *   gcc -std=c99 -Iinclude examples/abi_consumer.c -o /tmp/abi_consumer \
 *       -Lcausal-mbrl_b200 -l:libcmrl_b200.so -Wl,-rpath,$PWD/causal-mbrl_b200 && /tmp/abi_consumer
 *
 * A real host (the reference binds through ctypes) passes device pointers it owns, e.g. for
 * cmrl/models/layers.py:60-65:  cmrl_ens_linear_fwd(x, 0, B*in, w, b, y, NULL, 1, E, B, in, out, NULL, 0, 0, 0, NULL,
 * CMRL_ACT_SILU, 0, stream).  tests/test_host_logic.py compiles and runs this file. */
#include <stdio.h>
#include <string.h>

#include "cmrl_b200.h"

int main(void) {
    if (cmrl_abi_version() != 1) return 1;
    /* null device pointers: < 0 = invalid argument, message through cmrl_last_error() */
    int rc = cmrl_gather_inputs(NULL, NULL, NULL, 4, 5, 1, NULL, NULL);
    if (rc >= 0) return 2;
    if (strstr(cmrl_last_error(), "null pointer") == NULL) return 3;
    printf("abi %d, launches %lld\n", cmrl_abi_version(), cmrl_launch_count());
    return 0;
}
