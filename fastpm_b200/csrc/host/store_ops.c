/* fastpm_b200 host layer -- the store utilities of api/fastpm/store.h that move whole particles between or inside stores with DEVICE
 * columns: copy / take / extend (libfastpm/store.c:925-966), the sub-sampling mask and the compaction it drives (store.c:967-1034,
 * what the command line's `particle_fraction` does to a snapshot, src/fastpm.c:1449-1461), the mask sum (store.c:289-299), permute and
 * the local sort by id (store.c:380-446), single positions (store.c:106-118).  Bulk work runs as kernels (csrc/particles.cu), a stable
 * compaction being an exclusive prefix sum of the mask followed by one scatter per column. */
#include "internal.h"

static void store_copy_range(FastPMStore *p, ptrdiff_t start, FastPMStore *po, ptrdiff_t offset, size_t ncopy)
{
    fpm_store_flush(p); fpm_store_flush(po);
    if (ncopy + start > p->np)
        fastpm_raise(-1, "Copy out of bounds from source FastPMStore: asking for %td but has %td\n", (ptrdiff_t) (ncopy + start), (ptrdiff_t) p->np);
    if (ncopy + offset > po->np_upper)
        fastpm_raise(-1, "Not enough storage in target FastPMStore: asking for %td but has %td\n", (ptrdiff_t) (ncopy + offset), (ptrdiff_t) po->np_upper);
    for (int c = 0; c < 32; c++) {
        if (!po->columns[c]) continue;
        if (!p->columns[c]) fastpm_raise(-1, "fastpm_store_copy: the source store has no %s column\n", po->_column_info[c].name);
        const size_t elsize = po->_column_info[c].elsize;
        if (ncopy) FPM_MUST(fpm_memcpy_d2d(po->columns[c] + offset * elsize, p->columns[c] + start * elsize, elsize * ncopy));
    }
    po->np = offset + ncopy;
}

void fastpm_store_copy(FastPMStore *p, FastPMStore *po) { store_copy_range(p, 0, po, 0, p->np); po->meta = p->meta; }
void fastpm_store_take(FastPMStore *p, ptrdiff_t i, FastPMStore *po, ptrdiff_t j) { store_copy_range(p, i, po, j, 1); po->meta = p->meta; }
void fastpm_store_extend(FastPMStore *p, FastPMStore *extra) { store_copy_range(extra, 0, p, p->np, extra->np); }

void fastpm_store_get_position(FastPMStore *p, ptrdiff_t index, double pos[3])
{
    fpm_store_flush(p);
    FPM_MUST(fpm_memcpy_d2h(pos, p->x + index, sizeof(p->x[0])));
}

void fastpm_store_get_lagrangian_position(FastPMStore *p, ptrdiff_t index, double pos[3])
{
    float q[3];
    FPM_MUST(fpm_memcpy_d2h(q, p->q + index, sizeof(q)));
    for (int d = 0; d < 3; d++) pos[d] = q[d];
}

size_t fastpm_store_get_mask_sum(FastPMStore *p, MPI_Comm comm)
{
    int64_t np = 0;
    FPM_MUST(fpm_mask_scan(p->mask, (int64_t) p->np, NULL, &np));
    fpm_comm_allreduce_i64(comm, &np, 1, 0);
    return (size_t) np;
}

/* mask[i] = fraction >= 1 || rand[i] <= fraction; `mask` is device memory of at least np entries (the reference's callers take
 * it from fastpm_memory_alloc, which hands out device memory here) */
void fastpm_store_fill_subsample_mask(FastPMStore *p, double fraction, FastPMParticleMaskType *mask)
{
    if (!p->rand) fastpm_raise(-1, "fastpm_store_fill_subsample_mask: the store has no rand column\n");
    FPM_MUST(fpm_subsample_mask(p->rand, NULL, fraction, (int64_t) p->np, mask));
}

/* the same with one fraction per particle, `fraction` being device memory of np doubles */
void fastpm_store_fill_subsample_mask_from_array(FastPMStore *p, double *fraction, FastPMParticleMaskType *mask)
{
    if (!p->rand) fastpm_raise(-1, "fastpm_store_fill_subsample_mask_from_array: the store has no rand column\n");
    FPM_MUST(fpm_subsample_mask(p->rand, fraction, 0.0, (int64_t) p->np, mask));
}

/* The particles with a non-zero mask, in their order, into po (columns po has; po == NULL: just the count; po == p: in place,
 * through a scratch column).  Returns the number kept. */
size_t fastpm_store_subsample(FastPMStore *p, FastPMParticleMaskType *mask, FastPMStore *po)
{
    fpm_store_flush(p);
    const int64_t n = (int64_t) p->np;
    int64_t kept = 0;
    if (po == NULL || n == 0) {
        FPM_MUST(fpm_mask_scan(mask, n, NULL, &kept));
        if (po) { po->np = 0; po->meta = p->meta; }
        return (size_t) kept;
    }
    fpm_store_flush(po);
    int64_t *dest = fpm_malloc(sizeof(int64_t) * (size_t) n);
    if (!dest) fastpm_raise(-1, "fastpm_store_subsample: %s\n", fpm_last_error());
    FPM_MUST(fpm_mask_scan(mask, n, dest, &kept));
    if ((size_t) kept > po->np_upper)
        fastpm_raise(-1, "Not enough storage in target FastPMStore: asking for %td but has %td\n", (ptrdiff_t) kept, (ptrdiff_t) po->np_upper);
    void *scratch = NULL;
    for (int c = 0; c < 32; c++) {
        if (!po->columns[c]) continue;
        if (!p->columns[c]) fastpm_raise(-1, "fastpm_store_subsample: the source store has no %s column\n", po->_column_info[c].name);
        const size_t elsize = po->_column_info[c].elsize;
        if (po->columns[c] != p->columns[c]) {
            FPM_MUST(fpm_compact_rows(po->columns[c], p->columns[c], mask, dest, n, (int) elsize));
            continue;
        }
        /* in place: a kept row may land where a later row has not been read yet */
        if (kept == n) continue;
        if (!scratch) {
            size_t maxel = 1;
            for (int k = 0; k < 32; k++) if (po->columns[k] && po->_column_info[k].elsize > maxel) maxel = po->_column_info[k].elsize;
            scratch = fpm_malloc(maxel * (size_t) (kept ? kept : 1));
            if (!scratch) fastpm_raise(-1, "fastpm_store_subsample: %s\n", fpm_last_error());
        }
        FPM_MUST(fpm_compact_rows(scratch, p->columns[c], mask, dest, n, (int) elsize));
        if (kept) FPM_MUST(fpm_memcpy_d2d(po->columns[c], scratch, elsize * (size_t) kept));
    }
    if (scratch) fpm_free(scratch);
    fpm_free(dest);
    po->np = (size_t) kept;
    po->meta = p->meta;
    return (size_t) kept;
}

/* row i of every column becomes row ind[i]; `ind` is a HOST array of np ints, as in the reference */
void fastpm_store_permute(FastPMStore *p, int *ind)
{
    fpm_store_flush(p);
    const size_t n = p->np;
    if (n == 0) return;
    size_t maxel = 1;
    for (int c = 0; c < 32; c++) if (p->columns[c] && p->_column_info[c].elsize > maxel) maxel = p->_column_info[c].elsize;
    int *ind_dev = fpm_malloc(sizeof(int) * n);
    void *scratch = fpm_malloc(maxel * n);
    if (!ind_dev || !scratch) fastpm_raise(-1, "No memory for permuting: %s\n", fpm_last_error());
    FPM_MUST(fpm_memcpy_h2d(ind_dev, ind, sizeof(int) * n));
    for (int c = 0; c < 32; c++) {
        if (!p->columns[c]) continue;
        const size_t elsize = p->_column_info[c].elsize;
        FPM_MUST(fpm_gather_rows(scratch, p->columns[c], ind_dev, (int64_t) n, (int) elsize));
        FPM_MUST(fpm_memcpy_d2d(p->columns[c], scratch, elsize * n));
    }
    fpm_free(scratch);
    fpm_free(ind_dev);
}

/* The comparator the reference ships (store.c:414-424).  It exists so that fastpm_store_sort(p, FastPMLocalSortByID) reads as in
 * the reference; called directly it compares through two small copies. */
int FastPMLocalSortByID(const int i1, const int i2, FastPMStore *p)
{
    uint64_t a, b;
    FPM_MUST(fpm_memcpy_d2h(&a, p->id + i1, sizeof(a)));
    FPM_MUST(fpm_memcpy_d2h(&b, p->id + i2, sizeof(b)));
    return (a > b) - (a < b);
}

/* store.c:426-446 sorts an index array with qsort and a comparator that reads host columns.  Device columns cannot be read by a
 * user's comparator, so the one ordering libfastpm itself uses -- by id -- is what this build sorts by: ids mirrored, a stable
 * radix sort of the index on the host, one gather per column on the device. */
void fastpm_store_sort(FastPMStore *p, int (*cmp_func)(const int i1, const int i2, FastPMStore *p))
{
    if (cmp_func != FastPMLocalSortByID)
        fastpm_raise(-1, "fastpm_b200: fastpm_store_sort orders device columns by FastPMLocalSortByID only\n");
    if (!p->id) fastpm_raise(-1, "fastpm_store_sort: the store has no id column\n");
    fpm_store_flush(p);
    const size_t n = p->np;
    if (n == 0) return;
    uint64_t *key = malloc(sizeof(uint64_t) * n), *perm = malloc(sizeof(uint64_t) * n);
    int *ind = malloc(sizeof(int) * n);
    FPM_MUST(fpm_memcpy_d2h(key, p->id, sizeof(uint64_t) * n));
    fastpm_b200_io_argsort_u64(key, n, perm);
    for (size_t i = 0; i < n; i++) ind[i] = (int) perm[i];
    fastpm_store_permute(p, ind);
    free(key); free(perm); free(ind);
}

/* for bindings that cannot lay out FastPMStore: a scratch store filled on pm's particle grid (q, rand and mask columns),
 * sub-sampled at `fraction` -- into a second store, or in place --, optionally reversed (fastpm_store_permute) and sorted back
 * (fastpm_store_sort), and the kept ids and positions mirrored to the host.  Returns the
 * number kept; *mask_sum receives fastpm_store_get_mask_sum of the filled store. */
int64_t fastpm_b200_subsample_probe(PM *pm, int64_t np_upper, double fraction, int in_place, int sort_back, uint64_t *id_host, double *x_host, int64_t *mask_sum)
{
    FastPMStore p[1], po[1];
    const FastPMColumnTags attrs = COLUMN_POS | COLUMN_ID | COLUMN_Q | COLUMN_RAND | COLUMN_MASK;
    fastpm_store_init(p, "probe", (size_t) np_upper, attrs, FASTPM_MEMORY_HEAP);
    fastpm_store_fill(p, pm, NULL, NULL);
    fastpm_store_fill_subsample_mask(p, fraction, p->mask);
    *mask_sum = (int64_t) fastpm_store_get_mask_sum(p, MPI_COMM_WORLD);
    FastPMStore *out = p;
    if (!in_place) {
        fastpm_store_init(po, "kept", fastpm_store_subsample(p, p->mask, NULL) + 1, attrs & ~COLUMN_MASK, FASTPM_MEMORY_HEAP);
        out = po;
    }
    fastpm_store_subsample(p, p->mask, out);
    if (sort_back && out->np) {                               /* reversed by fastpm_store_permute, put back by fastpm_store_sort */
        int *ind = malloc(sizeof(int) * out->np);
        for (size_t i = 0; i < out->np; i++) ind[i] = (int) (out->np - 1 - i);
        fastpm_store_permute(out, ind);
        free(ind);
        if (FastPMLocalSortByID(0, (int) out->np - 1, out) < 0) fastpm_raise(-1, "fastpm_b200_subsample_probe: the permutation did not reverse the store\n");
        fastpm_store_sort(out, FastPMLocalSortByID);
    }
    const int64_t kept = (int64_t) out->np;
    FPM_MUST(fpm_memcpy_d2h(id_host, out->id, sizeof(out->id[0]) * out->np));
    FPM_MUST(fpm_memcpy_d2h(x_host, out->x, sizeof(out->x[0]) * out->np));
    if (!in_place) fastpm_store_destroy(po);
    fastpm_store_destroy(p);
    return kept;
}
