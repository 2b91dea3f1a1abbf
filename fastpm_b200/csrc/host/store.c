/* fastpm_b200 host layer -- the column-oriented particle store with DEVICE columns
 * (reference: libfastpm/store.c; struct FastPMStore is public, api/fastpm/store.h:62-135).
 * Layout, column order, element sizes and the 1024-byte column alignment follow store.c:138-254; the
 * per-element accessors (pack/unpack/to_double/from_double) work on device memory through small copies,
 * bulk work (fill, wrap, summary, kick, drift, paint, readout) runs as kernels. */
#include "internal.h"

static void dev_pack(FastPMStore *p, ptrdiff_t index, int ci, void *packed)
{ fpm_store_flush(p); size_t es = p->_column_info[ci].elsize; FPM_MUST(fpm_memcpy_d2h(packed, p->columns[ci] + index * es, es)); }
static void dev_unpack(FastPMStore *p, ptrdiff_t index, int ci, void *packed)
{ fpm_store_flush(p); size_t es = p->_column_info[ci].elsize; FPM_MUST(fpm_memcpy_h2d(p->columns[ci] + index * es, packed, es)); }
static double dev_to_double_f4(FastPMStore *p, ptrdiff_t index, int ci, int memb)
{ fpm_store_flush(p); float v; FPM_MUST(fpm_memcpy_d2h(&v, p->columns[ci] + index * p->_column_info[ci].elsize + 4 * memb, 4)); return v; }
static double dev_to_double_f8(FastPMStore *p, ptrdiff_t index, int ci, int memb)
{ fpm_store_flush(p); double v; FPM_MUST(fpm_memcpy_d2h(&v, p->columns[ci] + index * p->_column_info[ci].elsize + 8 * memb, 8)); return v; }
static void dev_from_double_f4(FastPMStore *p, ptrdiff_t index, int ci, int memb, const double value)
{ fpm_store_flush(p); float v = (float) value; FPM_MUST(fpm_memcpy_h2d(p->columns[ci] + index * p->_column_info[ci].elsize + 4 * memb, &v, 4)); }

const char *fastpm_species_get_name(enum FastPMSpecies species)
{
    switch (species) {
        case FASTPM_SPECIES_BARYON: return "0";
        case FASTPM_SPECIES_CDM: return "1";
        case FASTPM_SPECIES_NCDM: return "2";
    }
    return "UNKNOWN";
}

double fastpm_store_get_mass(FastPMStore *p, ptrdiff_t index)
{
    if (p->mass) return p->meta.M0 + dev_to_double_f4(p, index, FASTPM_STORE_COLUMN_INDEX(mass), 0);
    return p->meta.M0;
}

static ptrdiff_t alignsize(ptrdiff_t size) { return ((size + 1024) / 1024) * 1024; }

#define DEF(column, attr_, dtype_, nmemb_) do { \
        int ci = FASTPM_STORE_COLUMN_INDEX(column); \
        if (attr_ != (1 << ci)) fastpm_raise(-1, "attr and column are out of order for %s\n", #column); \
        struct FastPMColumnInfo *c = &p->_column_info[ci]; \
        strcpy(c->dtype, dtype_); strcpy(c->name, #column); \
        c->elsize = sizeof(p->column[0]); c->nmemb = nmemb_; c->membsize = sizeof(p->column[0]) / nmemb_; \
        c->attribute = attr_; c->pack = dev_pack; c->unpack = dev_unpack; \
        c->to_double = (dtype_[0] == 'f') ? (c->membsize == 8 ? dev_to_double_f8 : dev_to_double_f4) : NULL; \
        c->from_double = (dtype_[0] == 'f' && c->membsize == 4) ? dev_from_double_f4 : NULL; \
    } while (0)

void fastpm_store_init_details(FastPMStore *p, const char *name, size_t np_upper, FastPMColumnTags attributes,
                               enum FastPMMemoryLocation loc, const char *file, const int line)
{
    p->mem = _libfastpm_get_gmem();
    if (name) strcpy(p->name, name);
    p->attributes = attributes;
    p->np = 0;
    p->np_upper = np_upper;
    memset(p->columns, 0, sizeof(p->columns));
    memset(p->_column_info, 0, sizeof(p->_column_info));
    memset(&p->meta, 0, sizeof(p->meta));
    p->_base = NULL;
    DEF(x, COLUMN_POS, "f8", 3);       DEF(q, COLUMN_Q, "f4", 3);          DEF(v, COLUMN_VEL, "f4", 3);
    DEF(acc, COLUMN_ACC, "f4", 3);     DEF(dx1, COLUMN_DX1, "f4", 3);      DEF(dx2, COLUMN_DX2, "f4", 3);
    DEF(dv1, COLUMN_DV1, "f4", 3);     DEF(aemit, COLUMN_AEMIT, "f4", 1);  DEF(rho, COLUMN_DENSITY, "f4", 1);
    DEF(potential, COLUMN_POTENTIAL, "f4", 1); DEF(tidal, COLUMN_TIDAL, "f4", 6); DEF(id, COLUMN_ID, "i8", 1);
    DEF(pgdc, COLUMN_PGDC, "f4", 3);   DEF(mask, COLUMN_MASK, "i1", 1);    DEF(minid, COLUMN_MINID, "i8", 1);
    DEF(task, COLUMN_TASK, "i4", 1);   DEF(length, COLUMN_LENGTH, "i4", 1); DEF(rdisp, COLUMN_RDISP, "f4", 6);
    DEF(vdisp, COLUMN_VDISP, "f4", 6); DEF(rvdisp, COLUMN_RVDISP, "f4", 9); DEF(mass, COLUMN_MASS, "f4", 1);
    DEF(rand, COLUMN_RAND, "f4", 1);   DEF(rmom, COLUMN_RMOM, "f4", 1);

    ptrdiff_t size = 0;
    for (int ci = 0; ci < 32; ci++)
        if (attributes & p->_column_info[ci].attribute) size += alignsize(p->_column_info[ci].elsize * np_upper);
    p->_base = fastpm_memory_alloc_details(p->mem, "FastPMStore", size, loc, file, line);
    FPM_MUST(fpm_memset(p->_base, 0, size));
    ptrdiff_t offset = 0;
    for (int ci = 0; ci < 32; ci++) {
        if (attributes & p->_column_info[ci].attribute) {
            p->columns[ci] = (char *) p->_base + offset;
            offset += alignsize(p->_column_info[ci].elsize * np_upper);
        }
    }
}

size_t fastpm_store_init_evenly_details(FastPMStore *p, const char *name, size_t np_total, FastPMColumnTags attributes,
                                        double alloc_factor, MPI_Comm comm, const char *file, const int line)
{
    int ntask = fpm_comm_size(comm);
    size_t np_upper = (size_t) (1.0 * np_total / ntask * alloc_factor);
    fastpm_store_init_details(p, name, np_upper, attributes, FASTPM_MEMORY_HEAP, file, line);
    return 0;
}

void fastpm_store_destroy(FastPMStore *p) { fpm_store_flush(p); fastpm_memory_free(p->mem, p->_base); p->_base = NULL; }

int fastpm_store_find_column_id(FastPMStore *p, FastPMColumnTags attribute)
{
    for (int ci = 0; ci < 32; ci++) if (p->_column_info[ci].attribute == attribute) return ci;
    return -1;
}

size_t fastpm_store_get_np_total(FastPMStore *p, MPI_Comm comm)
{
    int64_t np = p->np;
    fpm_comm_allreduce_i64(comm, &np, 1, 0);
    return (size_t) np;
}

/* store.c:723-806: one particle per cell of the Nc^3 grid, this rank's x-slab; id = i*Nc^2 + j*Nc + k.
 * The `rand` column (it feeds sub-sampling, store.c:967-997) is this rank's serial RANLUX stream, drawn on the host. */
void fastpm_store_fill(FastPMStore *p, PM *pm, double *shift, ptrdiff_t *Nc)
{
    ptrdiff_t nc[3];
    fpm_store_flush(p);
    for (int d = 0; d < 3; d++) nc[d] = Nc ? Nc[d] : pm->Nmesh[d];
    if (nc[0] != nc[1] || nc[0] != nc[2]) fastpm_raise(-1, "fastpm_b200: cubic particle grids only\n");
    ptrdiff_t start = pm->IRegion.start[0] * nc[0] / pm->Nmesh[0];
    ptrdiff_t end = (pm->IRegion.start[0] + pm->IRegion.size[0]) * nc[0] / pm->Nmesh[0];
    p->np = (size_t) (end - start) * nc[1] * nc[2];
    if (p->np > p->np_upper) fastpm_raise(-1, "Need %td particles; %td allocated\n", p->np, p->np_upper);
    for (int d = 0; d < 3; d++) {
        p->meta._q_shift[d] = shift ? shift[d] : 0;
        p->meta._q_scale[d] = pm->BoxSize[d] / nc[d];
    }
    p->meta._q_size = nc[0] * nc[1] * nc[2];
    p->meta._q_strides[0] = nc[1] * nc[2]; p->meta._q_strides[1] = nc[2]; p->meta._q_strides[2] = 1;
    FPM_MUST(fpm_fill_grid((double *) p->x, p->id, (float *) p->v, (int) nc[0], (int) start, (int64_t) p->np,
                           pm->BoxSize[0], p->meta._q_shift[0]));
    /* i-planes of nc x nc particles in id order: let paint / readout walk them in Lagrangian bricks */
    FPM_MUST(fpm_particle_grid_hint((int) nc[0]));
    if (p->q) FPM_MUST(fpm_cast_f64_to_f32((float *) p->q, (const double *) p->x, (int64_t) (3 * p->np)));      /* store.c:784-789 */
    if (p->mask) FPM_MUST(fpm_memset(p->mask, 0, sizeof(p->mask[0]) * p->np));
    if (p->rmom) FPM_MUST(fpm_memset(p->rmom, 0, sizeof(p->rmom[0]) * p->np));
    p->meta.a_x = p->meta.a_v = 0.;
    /* store.c:804: one serial RANLUX stream per rank over all np_upper entries */
    if (p->rand) FPM_MUST(fpm_fill_rand((float *) p->rand, (int64_t) p->np_upper, pm->ThisTask));
}

/* for bindings that cannot lay out FastPMStore: a scratch store with the q and rand columns, filled on pm's grid by fastpm_store_fill;
 * q_host [np][3] and rand_host [np_upper] receive the columns.  Returns np. */
int64_t fastpm_b200_fill_probe(PM *pm, int64_t np_upper, float *q_host, float *rand_host)
{
    FastPMStore p[1];
    fastpm_store_init(p, "probe", (size_t) np_upper, COLUMN_POS | COLUMN_ID | COLUMN_Q | COLUMN_RAND | COLUMN_MASK, FASTPM_MEMORY_HEAP);
    fastpm_store_fill(p, pm, NULL, NULL);
    const int64_t np = (int64_t) p->np;
    FPM_MUST(fpm_memcpy_d2h(q_host, p->q, sizeof(p->q[0]) * p->np));
    FPM_MUST(fpm_memcpy_d2h(rand_host, p->rand, sizeof(p->rand[0]) * p->np_upper));
    fastpm_store_destroy(p);
    return np;
}

void fastpm_store_wrap(FastPMStore *p, double BoxSize[3])
{
    fpm_store_flush(p);
    if (fpm_wrap((double *) p->x, (int64_t) p->np, BoxSize[0]) != 0)
        fastpm_raise(-1, "%s\n", fpm_last_error());
}

/* FastPMTargetPM (store.c:478-484): the owner of one particle; needs its position on the host */
int FastPMTargetPM(FastPMStore *p, ptrdiff_t i, PM *pm)
{
    double pos[3];
    fpm_store_flush(p);
    FPM_MUST(fpm_memcpy_d2h(pos, p->x + i, sizeof(pos)));
    return pm_pos_to_rank(pm, pos);
}

/* store.c:808-908.  fmt characters: '<' min, '>' max, '-' mean, 's' std, 'S' sample std, 'v' variance, 'V' sample variance */
void fastpm_store_summary(FastPMStore *p, FastPMColumnTags attribute, MPI_Comm comm, const char *fmt, ...)
{
    va_list va;
    va_start(va, fmt);
    fpm_store_flush(p);
    int ci = fastpm_store_find_column_id(p, attribute);
    if (ci < 0 || !p->columns[ci]) fastpm_raise(-1, "Column for attribute %d is not allocated\n", (int) attribute);
    struct FastPMColumnInfo *c = &p->_column_info[ci];
    if (c->to_double == NULL) fastpm_raise(-1, "Column %s didnot set to_double virtual function\n", c->name);
    int nmemb = (int) c->nmemb;
    double raw[9 * 4], rmin[9], rmax[9], rsum1[9], rsum2[9];
    FPM_MUST(fpm_summary(p->columns[ci], (int) c->membsize, nmemb, (int64_t) p->np, raw));
    for (int d = 0; d < nmemb; d++) { rmin[d] = raw[4 * d]; rmax[d] = raw[4 * d + 1]; rsum1[d] = raw[4 * d + 2]; rsum2[d] = raw[4 * d + 3]; }
    int64_t Ntot = p->np;
    fpm_comm_allreduce_double(comm, rsum1, nmemb, 0);
    fpm_comm_allreduce_double(comm, rsum2, nmemb, 0);
    fpm_comm_allreduce_double(comm, rmin, nmemb, 1);
    fpm_comm_allreduce_double(comm, rmax, nmemb, 2);
    fpm_comm_allreduce_i64(comm, &Ntot, 1, 0);
    for (size_t i = 0; i < strlen(fmt); i++) {
        double *dr = va_arg(va, double *);
        for (int d = 0; d < nmemb && d < 3; d++) {
            double mean = rsum1[d] / Ntot, var = rsum2[d] / Ntot - pow(rsum1[d] / Ntot, 2);
            switch (fmt[i]) {
                case '-': dr[d] = mean; break;
                case '<': dr[d] = rmin[d]; break;
                case '>': dr[d] = rmax[d]; break;
                case 's': dr[d] = sqrt(var); break;
                case 'S': dr[d] = sqrt(1.0 * Ntot / (Ntot - 1.)) * sqrt(var); break;
                case 'v': dr[d] = var; break;
                case 'V': dr[d] = (1.0 * Ntot / (Ntot - 1.)) * var; break;
                default: fastpm_raise(-1, "Unknown format str. Use '<->sSvV'\n");
            }
        }
    }
    va_end(va);
}

/* bulk host mirrors */
int fastpm_b200_store_get_column(FastPMStore *p, FastPMColumnTags attribute, void *host_dst, size_t first, size_t count)
{
    fpm_store_flush(p);
    int ci = fastpm_store_find_column_id(p, attribute);
    if (ci < 0 || !p->columns[ci]) return -1;
    size_t es = p->_column_info[ci].elsize;
    return fpm_memcpy_d2h(host_dst, p->columns[ci] + first * es, count * es);
}
int fastpm_b200_store_set_column(FastPMStore *p, FastPMColumnTags attribute, const void *host_src, size_t first, size_t count)
{
    fpm_store_flush(p);
    int ci = fastpm_store_find_column_id(p, attribute);
    if (ci < 0 || !p->columns[ci]) return -1;
    size_t es = p->_column_info[ci].elsize;
    return fpm_memcpy_h2d(p->columns[ci] + first * es, host_src, count * es);
}
