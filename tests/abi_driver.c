/* Plain-C consumer of include/deo_b200.h: proves the header is valid C11, the structs have the documented layout and
 * the library links and answers host-only entry points (no GPU needed).  Built and run by tests/test_abi_cpu.py. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "deo_b200.h"

int main(void) {
    char msg[256];
    int32_t n = -1;
    int64_t start = -1, count = -1;
    if (deo_abi_version() != DEO_ABI_VERSION) { printf("FAIL abi version\n"); return 1; }
    if (sizeof(deo_op_desc) != 10 * 4 + 4 * sizeof(void *)) { printf("FAIL sizeof(deo_op_desc)\n"); return 1; }
    if (sizeof(deo_bc_desc) != 4 * 4 + 4 * sizeof(void *)) { printf("FAIL sizeof(deo_bc_desc)\n"); return 1; }
    if (offsetof(deo_plan_desc, ops) != 56 || offsetof(deo_plan_desc, bc) != 64) { printf("FAIL deo_plan_desc layout\n"); return 1; }
    if (deo_dist_slab(1024, 8, 3, &start, &count) != DEO_OK || start != 384 || count != 128) { printf("FAIL deo_dist_slab\n"); return 1; }
    if (deo_dist_slab(3, 8, 0, &start, &count) != DEO_ERR_INVALID) { printf("FAIL deo_dist_slab error path\n"); return 1; }
    if (deo_last_error(msg, sizeof msg) != DEO_OK || strlen(msg) == 0) { printf("FAIL deo_last_error\n"); return 1; }
    if (deo_plan_create(NULL, NULL) == DEO_OK) { printf("FAIL null arguments accepted\n"); return 1; }
    (void)deo_device_count(&n);                /* DEO_ERR_CUDA without a driver is fine: it must not crash */
    printf("OK devices=%d last_error=\"%s\"\n", (int)n, msg);
    return 0;
}
