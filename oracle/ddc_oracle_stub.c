/*
 * ddc_oracle_stub.c -- TEST INFRASTRUCTURE.  The include/ddc.h entry points that the HOST code calls
 * (integration/reference_binding/CudaRcbPartitioner.cpp and domain_decomp_b200/host/), answered by
 * the CPU oracle (ddc_oracle.c) instead of the CUDA library.  Linked ONLY into test artefacts under
 * oracle/_ref/ (libref_binding_cpu.so, decomp_oracle, host_tests_oracle), so that host-side logic
 * -- assembling the mask, rank views, the writers, the CLI -- can be tested on a machine without a
 * GPU.  The product (libddc_cuda.so, libdomain_decomp.so, decomp) has no such path and is never
 * linked with this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ddc.h"

int orc_partition(const int32_t* mask, int NX, int NY, int P, int use_hist, int32_t* boxes, int32_t* pid,
    int* changes_out, long* median_iters);
void orc_neighbours(const int32_t* boxes, int P, int NX, int NY, int px, int py, int32_t* counts,
    const int64_t* offsets, int32_t* ids, int32_t* halos, int32_t* starts);
void orc_part_loads(const int32_t* pid, size_t ncell, int P, int64_t* loads);

struct ddc_handle_s {
    int nx, ny, P, px, py, changes, have_nbr;
    long iters;
    int32_t *mask, *boxes, *pid;
    int32_t *counts, *ids, *halos, *starts; /* neighbour tables, 8 lists */
    int64_t offsets[9];
};
static void free_tables(ddc_handle_t h)
{
    free(h->counts);
    free(h->ids);
    free(h->halos);
    free(h->starts);
    h->counts = h->ids = h->halos = h->starts = NULL;
    h->have_nbr = 0;
}

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
        free_tables(h);
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
    free(h->boxes);
    free(h->pid);
    free_tables(h);
    h->P = nparts;
    h->px = px;
    h->py = py;
    h->boxes = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)nparts);
    h->pid = (int32_t*)malloc(sizeof(int32_t) * (size_t)h->nx * h->ny);
    if (orc_partition(h->mask, h->nx, h->ny, nparts, 1, h->boxes, h->pid, &h->changes, &h->iters) != 0)
        return DDC_ERR_STATE;
    if ((flags & DDC_WANT_NEIGHBOURS) && nparts > 1) { /* P == 1 returns before neighbour discovery */
        const int P = nparts;
        h->counts = (int32_t*)calloc(8 * (size_t)P, sizeof(int32_t));
        orc_neighbours(h->boxes, P, h->nx, h->ny, px, py, h->counts, NULL, NULL, NULL, NULL);
        h->offsets[0] = 0;
        for (int l = 0; l < 8; l++) {
            int64_t t = 0;
            for (int p = 0; p < P; p++)
                t += h->counts[(size_t)l * P + p];
            h->offsets[l + 1] = h->offsets[l] + t;
        }
        const size_t n = (size_t)h->offsets[8] + 1;
        h->ids = (int32_t*)malloc(sizeof(int32_t) * n);
        h->halos = (int32_t*)malloc(sizeof(int32_t) * n);
        h->starts = (int32_t*)malloc(sizeof(int32_t) * n);
        orc_neighbours(h->boxes, P, h->nx, h->ny, px, py, h->counts, h->offsets, h->ids, h->halos, h->starts);
        h->have_nbr = 1;
    }
    return DDC_OK;
}
int ddc_get_neighbour_counts(ddc_handle_t h, int edge, int periodic, int32_t* counts)
{
    const int l = periodic * 4 + edge;
    if (!h->have_nbr)
        memset(counts, 0, sizeof(int32_t) * (size_t)h->P);
    else
        memcpy(counts, h->counts + (size_t)l * h->P, sizeof(int32_t) * (size_t)h->P);
    return DDC_OK;
}
int ddc_get_neighbour_total(ddc_handle_t h, int edge, int periodic, int64_t* total)
{
    const int l = periodic * 4 + edge;
    *total = h->have_nbr ? h->offsets[l + 1] - h->offsets[l] : 0;
    return DDC_OK;
}
int ddc_get_neighbours(ddc_handle_t h, int edge, int periodic, int32_t* ids, int32_t* halos, int32_t* starts)
{
    const int l = periodic * 4 + edge;
    if (!h->have_nbr)
        return DDC_OK;
    const size_t n = (size_t)(h->offsets[l + 1] - h->offsets[l]), o = (size_t)h->offsets[l];
    memcpy(ids, h->ids + o, sizeof(int32_t) * n);
    memcpy(halos, h->halos + o, sizeof(int32_t) * n);
    memcpy(starts, h->starts + o, sizeof(int32_t) * n);
    return DDC_OK;
}
int ddc_get_stats(ddc_handle_t h, ddc_stats* out)
{
    memset(out, 0, sizeof *out);
    out->nx = h->nx;
    out->ny = h->ny;
    out->nparts = h->P;
    out->changes = h->P > 1 ? h->changes : 0;
    out->median_iters = (int32_t)h->iters;
    int64_t* loads = (int64_t*)malloc(sizeof(int64_t) * (size_t)h->P);
    orc_part_loads(h->pid, (size_t)h->nx * h->ny, h->P, loads);
    out->load_min = out->load_max = loads[0];
    for (int p = 0; p < h->P; p++) {
        out->n_ocean += loads[p];
        if (loads[p] < out->load_min)
            out->load_min = loads[p];
        if (loads[p] > out->load_max)
            out->load_max = loads[p];
    }
    free(loads);
    for (int l = 0; h->have_nbr && l < 4; l++)
        for (int64_t k = h->offsets[l]; k < h->offsets[l + 1]; k++)
            out->edge_cut += h->halos[k];
    return DDC_OK;
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
/* host buffers: there is nothing to pin without a GPU */
int ddc_host_alloc(void** ptr, size_t bytes)
{
    *ptr = malloc(bytes ? bytes : 1);
    return *ptr ? DDC_OK : DDC_ERR_NOMEM;
}
int ddc_host_free(void* ptr)
{
    free(ptr);
    return DDC_OK;
}
/* one rank only behind the stub */
void ddc_shard_rows(int ny, int nranks, int rank, int* y_begin, int* y_count)
{
    const int rpr = (ny + nranks - 1) / nranks;
    int b = rank * rpr < ny ? rank * rpr : ny, e = b + rpr < ny ? b + rpr : ny;
    if (y_begin)
        *y_begin = b;
    if (y_count)
        *y_count = e - b;
}
int ddc_peer_connect(ddc_handle_t* handles, int n, int nx, int ny, int nparts)
{
    (void)handles;
    (void)n;
    (void)nx;
    (void)ny;
    (void)nparts;
    return DDC_ERR_ARG;
}
int ddc_peer_close(ddc_handle_t h)
{
    (void)h;
    return DDC_OK;
}
/* the halo exchange runs on the device only */
int ddc_halo_tile_offsets(ddc_handle_t h, int64_t* offsets)
{
    (void)h;
    (void)offsets;
    return DDC_ERR_STATE;
}
int ddc_halo_exchange_f64(ddc_handle_t h, double* tiles_dev, int periodic)
{
    (void)h;
    (void)tiles_dev;
    (void)periodic;
    return DDC_ERR_STATE;
}
