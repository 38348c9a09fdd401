/*
 * ddc_oracle_stub.c -- TEST INFRASTRUCTURE.  The handful of include/ddc.h entry points that
 * integration/reference_binding/CudaRcbPartitioner.cpp calls, answered by the CPU oracle
 * (ddc_oracle.c) instead of the CUDA library.  Linked ONLY into oracle/_ref/libref_binding_cpu.so,
 * so that the binding's own logic (assembling the mask from the ranks' blocks, handing boxes and
 * owners back to the reference's code) can be tested without a GPU; the product library has no
 * such path and is never linked with this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ddc.h"

int orc_partition(const int32_t* mask, int NX, int NY, int P, int use_hist, int32_t* boxes, int32_t* pid,
    int* changes_out, long* median_iters);

struct ddc_handle_s {
    int nx, ny, P;
    int32_t *mask, *boxes, *pid;
};

const char* ddc_last_error(ddc_handle_t h)
{
    (void)h;
    return "oracle stub";
}
int ddc_create(ddc_handle_t* h, int device, int rank, int nranks, const void* nccl_id)
{
    (void)device;
    (void)nccl_id;
    if (rank != 0 || nranks != 1)
        return DDC_ERR_ARG;
    *h = (ddc_handle_t)calloc(1, sizeof(struct ddc_handle_s));
    return *h ? DDC_OK : DDC_ERR_NOMEM;
}
int ddc_destroy(ddc_handle_t h)
{
    if (h) {
        free(h->mask);
        free(h->boxes);
        free(h->pid);
        free(h);
    }
    return DDC_OK;
}
int ddc_set_mask_host(ddc_handle_t h, const int32_t* rows, int nx, int ny, int y_begin, int y_count)
{
    if (y_begin != 0 || y_count != ny)
        return DDC_ERR_ARG;
    free(h->mask);
    h->mask = (int32_t*)malloc(sizeof(int32_t) * (size_t)nx * ny);
    memcpy(h->mask, rows, sizeof(int32_t) * (size_t)nx * ny);
    h->nx = nx;
    h->ny = ny;
    return DDC_OK;
}
int ddc_partition(ddc_handle_t h, int nparts, int px, int py, int flags)
{
    (void)px;
    (void)py;
    (void)flags;
    free(h->boxes);
    free(h->pid);
    h->P = nparts;
    h->boxes = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)nparts);
    h->pid = (int32_t*)malloc(sizeof(int32_t) * (size_t)h->nx * h->ny);
    int changes;
    long iters;
    return orc_partition(h->mask, h->nx, h->ny, nparts, 1, h->boxes, h->pid, &changes, &iters) == 0 ? DDC_OK : DDC_ERR_STATE;
}
int ddc_get_boxes(ddc_handle_t h, int32_t* x0, int32_t* y0, int32_t* ex, int32_t* ey)
{
    for (int p = 0; p < h->P; p++) {
        x0[p] = h->boxes[4 * p];
        y0[p] = h->boxes[4 * p + 1];
        ex[p] = h->boxes[4 * p + 2];
        ey[p] = h->boxes[4 * p + 3];
    }
    return DDC_OK;
}
int ddc_get_pid_host(ddc_handle_t h, int32_t* out)
{
    memcpy(out, h->pid, sizeof(int32_t) * (size_t)h->nx * h->ny);
    return DDC_OK;
}
