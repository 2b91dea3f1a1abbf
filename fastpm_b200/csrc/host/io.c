/* fastpm_b200 host layer -- snapshot and mesh writers / readers (row N2 of SURVEY.md section 8f).
 * Reference: libfastpmio/io.c (write_snapshot_header :229-318, fastpm_store_write :321-588, read_snapshot_header :163-226,
 * write_complex / read_complex :641-790) on top of the "bigfile" library (depends/bigfile/bigfile.c), whose on-disk format is
 * restated here (no MPI-IO, no aggregation: ranks of one node share the file system and write their slices with pwrite):
 *
 *   <file>/<block>/header      "DTYPE: <f4\nNMEMB: 3\nNFILE: 1\n" then per data file "%06X: <items> : <checksum> : <folded checksum>\n"
 *                              (bigfile.c:586-608; the checksum is the sum of all bytes of the file modulo 2^32, :1421-1428)
 *   <file>/<block>/attr-v2     one line per attribute, sorted by name: "name dtype nmemb HEXBYTES #HUMANE [ text ]" (:1563-1627)
 *   <file>/<block>/000000 ...  the items, little endian, file i holds items [size*i/Nfile, size*(i+1)/Nfile) (bigfile-mpi.c:106-111)
 *
 * The particle columns live on the device: they are mirrored chunk by chunk (fpm_memcpy_d2h) and converted to the file's
 * type on the host (Position is stored as f4, like the reference's).  The same routines take host arrays
 * (fastpm_b200_io_write_columns with on_device = 0), which is how the CPU tests compare the files byte for byte with the
 * directories the compiled reference writes.
 */
#define _GNU_SOURCE
#include "internal.h"
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

/* ------------------------------------------------------------------ bigfile blocks */
typedef struct { char name[128]; char dtype[8]; int nmemb; unsigned char *data; size_t nbytes; } BfAttr;
typedef struct {
    char path[1024];             /* <file>/<block>/ */
    char dtype[8];
    int nmemb, nfile;
    size_t size;
    size_t *fsize, *foffset;
    int64_t *fchecksum;          /* this rank's share; summed over the ranks at close */
    BfAttr *attrs;
    int nattr, attrs_dirty, header_dirty;
} BfBlock;

static int dtype_itemsize(const char *dtype)
{
    const char *p = dtype;
    if (*p == '<' || *p == '>' || *p == '=' || *p == '|') p++;
    return atoi(p + 1);
}
static char dtype_kind(const char *dtype)
{
    const char *p = dtype;
    if (*p == '<' || *p == '>' || *p == '=' || *p == '|') p++;
    return *p;
}
static void dtype_normalize(char *dst, const char *src)     /* bigfile.c:989-1020 on a little-endian machine */
{
    memset(dst, 0, 8);
    if (*src == '<' || *src == '>') { strncpy(dst, src, 7); return; }
    if (*src == '=' || *src == '|') src++;
    dst[0] = '<';
    strncpy(dst + 1, src, 6);
}

static int mkdir_p(const char *path)
{
    char tmp[1024];
    snprintf(tmp, sizeof(tmp), "%s", path);
    for (char *p = tmp + 1; *p; p++) {
        if (*p != '/') continue;
        *p = 0;
        if (mkdir(tmp, 0777) != 0 && errno != EEXIST) return -1;
        *p = '/';
    }
    if (mkdir(tmp, 0777) != 0 && errno != EEXIST) return -1;
    return 0;
}

static void bf_free(BfBlock *b)
{
    for (int i = 0; i < b->nattr; i++) free(b->attrs[i].data);
    free(b->attrs); free(b->fsize); free(b->foffset); free(b->fchecksum);
    memset(b, 0, sizeof(*b));
}

static int bf_write_header(BfBlock *b, const int64_t *checksums)
{
    char fn[1200];
    snprintf(fn, sizeof(fn), "%sheader", b->path);
    FILE *f = fopen(fn, "w");
    if (!f) return -1;
    fprintf(f, "DTYPE: %s\nNMEMB: %d\nNFILE: %d\n", b->dtype, b->nmemb, b->nfile);
    for (int i = 0; i < b->nfile; i++) {
        const unsigned int s = (unsigned int) (checksums[i] & 0xffffffffu);
        const unsigned int r = (s & 0xffff) + (s >> 16);
        const unsigned int folded = (r & 0xffff) + (r >> 16);
        fprintf(f, "%06X: %td : %u : %u\n", (unsigned) i, (ptrdiff_t) b->fsize[i], s, folded);
    }
    fclose(f);
    return 0;
}

static void attr_text(const BfAttr *a, char *out, size_t cap)
{
    /* bigfile.c:1586-1614 (the "humane" rendering) with big_file_dtype_format :1166-1207 */
    const int itemsize = dtype_itemsize(a->dtype);
    const char kind = dtype_kind(a->dtype);
    out[0] = 0;
    if (a->nbytes > 128) { snprintf(out, cap, "... (Too Long) "); return; }
    for (int j = 0; j < a->nmemb; j++) {
        const unsigned char *p = a->data + (size_t) j * itemsize;
        char buf[128];
        if (kind == 'a' || (kind == 'S' && itemsize == 1)) {
            if (p[0] == '\n') { strncat(out, "...", cap - strlen(out) - 1); break; }
            if (p[0] == 0) break;
            buf[0] = (char) p[0]; buf[1] = 0;
            strncat(out, buf, cap - strlen(out) - 1);
            continue;
        }
        if (kind == 'f' && itemsize == 8) { double v; memcpy(&v, p, 8); snprintf(buf, sizeof(buf), "%g", v); }
        else if (kind == 'f' && itemsize == 4) { float v; memcpy(&v, p, 4); snprintf(buf, sizeof(buf), "%g", v); }
        else if (kind == 'i' && itemsize == 8) { int64_t v; memcpy(&v, p, 8); snprintf(buf, sizeof(buf), "%ld", (long) v); }
        else if (kind == 'i' && itemsize == 4) { int32_t v; memcpy(&v, p, 4); snprintf(buf, sizeof(buf), "%d", v); }
        else if (kind == 'u' && itemsize == 8) { uint64_t v; memcpy(&v, p, 8); snprintf(buf, sizeof(buf), "%lu", (unsigned long) v); }
        else if (kind == 'u' && itemsize == 4) { uint32_t v; memcpy(&v, p, 4); snprintf(buf, sizeof(buf), "%u", v); }
        else if (kind == 'b' && itemsize == 1) snprintf(buf, sizeof(buf), "%d", (int) (signed char) p[0]);
        else snprintf(buf, sizeof(buf), "<%s>", a->dtype);
        strncat(out, buf, cap - strlen(out) - 1);
        if (j != a->nmemb - 1) strncat(out, " ", cap - strlen(out) - 1);
    }
}

static int attr_cmp(const void *a, const void *b) { return strcmp(((const BfAttr *) a)->name, ((const BfAttr *) b)->name); }

static int bf_write_attrs(BfBlock *b)
{
    static const char conv[] = "0123456789ABCDEF";
    char fn[1200];
    snprintf(fn, sizeof(fn), "%sattr-v2", b->path);
    FILE *f = fopen(fn, "w");
    if (!f) return -1;
    qsort(b->attrs, b->nattr, sizeof(BfAttr), attr_cmp);
    for (int i = 0; i < b->nattr; i++) {
        const BfAttr *a = &b->attrs[i];
        char *hex = malloc(2 * a->nbytes + 1);
        for (size_t k = 0; k < a->nbytes; k++) { hex[2 * k] = conv[a->data[k] / 16]; hex[2 * k + 1] = conv[a->data[k] % 16]; }
        hex[2 * a->nbytes] = 0;
        char *text = malloc((size_t) a->nmemb * 32 + 64);
        attr_text(a, text, (size_t) a->nmemb * 32 + 64);
        fprintf(f, "%s %s %d %s #HUMANE [ %s ]\n", a->name, a->dtype, a->nmemb, hex, text);
        free(hex); free(text);
    }
    fclose(f);
    return 0;
}

static void bf_set_attr(BfBlock *b, const char *name, const void *data, const char *dtype, int nmemb)
{
    BfAttr *a = NULL;
    for (int i = 0; i < b->nattr; i++) if (!strcmp(b->attrs[i].name, name)) a = &b->attrs[i];
    if (!a) {
        b->attrs = realloc(b->attrs, sizeof(BfAttr) * (b->nattr + 1));
        a = &b->attrs[b->nattr++];
        memset(a, 0, sizeof(*a));
        snprintf(a->name, sizeof(a->name), "%s", name);
    }
    free(a->data);
    dtype_normalize(a->dtype, dtype);
    a->nmemb = nmemb;
    a->nbytes = (size_t) dtype_itemsize(dtype) * nmemb;
    a->data = malloc(a->nbytes ? a->nbytes : 1);
    memcpy(a->data, data, a->nbytes);
    b->attrs_dirty = 1;
}

static int bf_get_attr(BfBlock *b, const char *name, void *data, const char *dtype, int nmemb)
{
    for (int i = 0; i < b->nattr; i++) {
        BfAttr *a = &b->attrs[i];
        if (strcmp(a->name, name)) continue;
        if (a->nmemb != nmemb || dtype_itemsize(a->dtype) != dtype_itemsize(dtype) || dtype_kind(a->dtype) != dtype_kind(dtype)) return -1;
        memcpy(data, a->data, a->nbytes);
        return 0;
    }
    return -1;
}

/* creates <file>/<block>/ (rank 0), its header with zero checksums, empty data files; dtype NULL: an attribute-only block
 * (bigfile.c:494-498 turns it into "i8", no files) */
static int bf_create(BfBlock *b, const char *filebase, const char *blockname, const char *dtype, int nmemb, int nfile, size_t size, MPI_Comm comm)
{
    memset(b, 0, sizeof(*b));
    snprintf(b->path, sizeof(b->path), "%s/%s/", filebase, blockname);
    if (!dtype) { dtype = "i8"; nfile = 0; size = 0; }
    dtype_normalize(b->dtype, dtype);
    b->nmemb = nmemb; b->nfile = nfile; b->size = size;
    b->fsize = calloc(nfile + 1, sizeof(size_t)); b->foffset = calloc(nfile + 1, sizeof(size_t)); b->fchecksum = calloc(nfile + 1, sizeof(int64_t));
    for (int i = 0; i < nfile; i++) {
        b->fsize[i] = size * (i + 1) / nfile - size * i / nfile;
        b->foffset[i + 1] = b->foffset[i] + b->fsize[i];
    }
    b->attrs_dirty = 1; b->header_dirty = 1;
    int rc = 0;
    if (fpm_comm_rank(comm) == 0) {
        rc = mkdir_p(b->path);
        if (!rc) rc = bf_write_header(b, b->fchecksum);
        for (int i = 0; i < nfile && !rc; i++) {
            char fn[1200];
            snprintf(fn, sizeof(fn), "%s%06X", b->path, (unsigned) i);
            FILE *f = fopen(fn, "w");
            if (!f) rc = -1; else fclose(f);
        }
    }
    fpm_comm_barrier(comm);
    return rc;
}

static int bf_open(BfBlock *b, const char *filebase, const char *blockname)
{
    memset(b, 0, sizeof(*b));
    snprintf(b->path, sizeof(b->path), "%s/%s/", filebase, blockname);
    char fn[1200];
    snprintf(fn, sizeof(fn), "%sheader", b->path);
    FILE *f = fopen(fn, "r");
    if (!f) return -1;
    if (fscanf(f, " DTYPE: %7s", b->dtype) != 1 || fscanf(f, " NMEMB: %d", &b->nmemb) != 1 || fscanf(f, " NFILE: %d", &b->nfile) != 1
        || b->nfile < 0 || b->nmemb < 0) { fclose(f); return -1; }
    b->fsize = calloc(b->nfile + 1, sizeof(size_t)); b->foffset = calloc(b->nfile + 1, sizeof(size_t)); b->fchecksum = calloc(b->nfile + 1, sizeof(int64_t));
    for (int i = 0; i < b->nfile; i++) {
        unsigned int fid, cks, folded;
        ptrdiff_t sz;
        if (fscanf(f, " %X : %td : %u : %u", &fid, &sz, &cks, &folded) != 4 || (int) fid >= b->nfile) { fclose(f); bf_free(b); return -1; }
        b->fsize[fid] = (size_t) sz;
        b->fchecksum[fid] = (int64_t) cks;          /* of use when the block grows (append): a reader never looks at it */
    }
    fclose(f);
    for (int i = 0; i < b->nfile; i++) b->foffset[i + 1] = b->foffset[i] + b->fsize[i];
    b->size = b->foffset[b->nfile];
    snprintf(fn, sizeof(fn), "%sattr-v2", b->path);
    f = fopen(fn, "r");
    if (f) {
        char *line = NULL;
        size_t cap = 0;
        while (getline(&line, &cap, f) > 0) {
            char name[128], dtype[16];
            int nmemb, used = 0;
            if (sscanf(line, "%127s %15s %d %n", name, dtype, &nmemb, &used) < 3) continue;
            const char *hex = line + used;
            const size_t nbytes = (size_t) dtype_itemsize(dtype) * nmemb;
            unsigned char *data = malloc(nbytes ? nbytes : 1);
            for (size_t k = 0; k < nbytes; k++) {
                unsigned int v = 0;
                if (sscanf(hex + 2 * k, "%2X", &v) != 1) break;
                data[k] = (unsigned char) v;
            }
            bf_set_attr(b, name, data, dtype, nmemb);
            free(data);
        }
        free(line);
        fclose(f);
    }
    b->attrs_dirty = 0;
    return 0;
}

/* big_block_mpi_grow_simple (bigfile-mpi.c:213-272, bigfile.c:411-447): nfile_grow more files holding size_grow more items, spread
 * evenly; the files that exist keep their sizes and checksums.  fchecksum is "this rank's share" (summed at close): the stored
 * checksums stay on rank 0 only. */
static int bf_grow(BfBlock *b, int nfile_grow, size_t size_grow, MPI_Comm comm)
{
    const int old = b->nfile, nfile = old + nfile_grow;
    size_t *fsize = calloc(nfile + 1, sizeof(size_t)), *foffset = calloc(nfile + 1, sizeof(size_t));
    int64_t *fchecksum = calloc(nfile + 1, sizeof(int64_t));
    for (int i = 0; i < old; i++) { fsize[i] = b->fsize[i]; fchecksum[i] = fpm_comm_rank(comm) == 0 ? b->fchecksum[i] : 0; }
    for (int i = 0; i < nfile_grow; i++) fsize[old + i] = size_grow * (i + 1) / nfile_grow - size_grow * i / nfile_grow;
    for (int i = 0; i < nfile; i++) foffset[i + 1] = foffset[i] + fsize[i];
    free(b->fsize); free(b->foffset); free(b->fchecksum);
    b->fsize = fsize; b->foffset = foffset; b->fchecksum = fchecksum;
    b->nfile = nfile; b->size = foffset[nfile]; b->header_dirty = 1;
    int rc = 0;
    if (fpm_comm_rank(comm) == 0) {
        for (int i = old; i < nfile && !rc; i++) {
            char fn[1200];
            snprintf(fn, sizeof(fn), "%s%06X", b->path, (unsigned) i);
            FILE *f = fopen(fn, "w");
            if (!f) rc = -1; else fclose(f);
        }
    }
    fpm_comm_barrier(comm);
    return rc;
}

/* items [offset, offset + n) of the block <-> buf (already in the file's type) */
static int bf_rw(BfBlock *b, size_t offset, size_t n, void *buf, int writing)
{
    const size_t isz = (size_t) dtype_itemsize(b->dtype) * b->nmemb;
    char *p = buf;
    int fi = 0;
    while (n > 0) {
        while (fi < b->nfile && offset >= b->foffset[fi + 1]) fi++;
        if (fi >= b->nfile) return -1;
        size_t here = b->foffset[fi + 1] - offset;
        if (here > n) here = n;
        char fn[1200];
        snprintf(fn, sizeof(fn), "%s%06X", b->path, (unsigned) fi);
        const int fd = open(fn, writing ? O_WRONLY : O_RDONLY);
        if (fd < 0) return -1;
        size_t done = 0, want = here * isz;
        const off_t at = (off_t) ((offset - b->foffset[fi]) * isz);
        while (done < want) {
            const ssize_t r = writing ? pwrite(fd, p + done, want - done, at + (off_t) done) : pread(fd, p + done, want - done, at + (off_t) done);
            if (r <= 0) { close(fd); return -1; }
            done += (size_t) r;
        }
        close(fd);
        if (writing) {
            unsigned int s = 0;
            const unsigned char *c = (const unsigned char *) p;
            for (size_t k = 0; k < want; k++) s += c[k];
            b->fchecksum[fi] += s;
            b->header_dirty = 1;
        }
        p += want; offset += here; n -= here;
    }
    return 0;
}

static int bf_close(BfBlock *b, MPI_Comm comm)
{
    int rc = 0;
    if (b->header_dirty && b->nfile > 0) fpm_comm_allreduce_i64(comm, b->fchecksum, b->nfile, 0);
    if (fpm_comm_rank(comm) == 0) {
        if (b->header_dirty) rc |= bf_write_header(b, b->fchecksum);
        if (b->attrs_dirty) rc |= bf_write_attrs(b);
    }
    fpm_comm_barrier(comm);
    bf_free(b);
    return rc;
}

/* element-wise conversion between a column's type and the file's (bigfile.c:1334-1419 cast(): plain C conversions) */
static int convert_items(void *dst, const char *ddtype, const void *src, const char *sdtype, size_t n)
{
    const int ds = dtype_itemsize(ddtype), ss = dtype_itemsize(sdtype);
    const char dk = dtype_kind(ddtype), sk = dtype_kind(sdtype);
    if (ds == ss && (dk == sk || ((dk == 'i' || dk == 'u') && (sk == 'i' || sk == 'u')))) { memcpy(dst, src, n * ds); return 0; }
    if (dk == 'f' && sk == 'f' && ds == 4 && ss == 8) { float *d = dst; const double *s = src; for (size_t i = 0; i < n; i++) d[i] = (float) s[i]; return 0; }
    if (dk == 'f' && sk == 'f' && ds == 8 && ss == 4) { double *d = dst; const float *s = src; for (size_t i = 0; i < n; i++) d[i] = (double) s[i]; return 0; }
    if ((dk == 'i' || dk == 'u') && (sk == 'i' || sk == 'u') && ds == 8 && ss == 4) { int64_t *d = dst; const int32_t *s = src; for (size_t i = 0; i < n; i++) d[i] = s[i]; return 0; }
    if ((dk == 'i' || dk == 'u') && (sk == 'i' || sk == 'u') && ds == 4 && ss == 8) { int32_t *d = dst; const int64_t *s = src; for (size_t i = 0; i < n; i++) d[i] = (int32_t) s[i]; return 0; }
    return -1;
}

/* this rank's first item in a block that concatenates the ranks in order */
static size_t rank_offset(int64_t np_local, MPI_Comm comm, int64_t *total)
{
    const int nt = fpm_comm_size(comm), me = fpm_comm_rank(comm);
    int64_t counts[64];
    memset(counts, 0, sizeof(counts));
    counts[me] = np_local;
    fpm_comm_allreduce_i64(comm, counts, nt, 0);
    size_t off = 0;
    *total = 0;
    for (int r = 0; r < nt; r++) { if (r < me) off += (size_t) counts[r]; *total += counts[r]; }
    return off;
}

/* ------------------------------------------------------------------ catalogs: fastpm_store_write / fastpm_store_read */
#define IO_CHUNK ((size_t) 4 << 20)          /* items mirrored and converted at a time */

static void meta_attrs(BfBlock *b, FpmIoMeta *m, int writing)
{
    if (writing) {
        bf_set_attr(b, "q.strides", m->q_strides, "i8", 3); bf_set_attr(b, "q.scale", m->q_scale, "f8", 3);
        bf_set_attr(b, "q.shift", m->q_shift, "f8", 3); bf_set_attr(b, "q.size", &m->q_size, "i8", 1);
        bf_set_attr(b, "a.x", &m->a_x, "f8", 1); bf_set_attr(b, "a.v", &m->a_v, "f8", 1); bf_set_attr(b, "M0", &m->M0, "f8", 1);
    } else {
        bf_get_attr(b, "q.strides", m->q_strides, "i8", 3); bf_get_attr(b, "q.scale", m->q_scale, "f8", 3);
        bf_get_attr(b, "q.shift", m->q_shift, "f8", 3); bf_get_attr(b, "q.size", &m->q_size, "i8", 1);
        bf_get_attr(b, "a.x", &m->a_x, "f8", 1); bf_get_attr(b, "a.v", &m->a_v, "f8", 1); bf_get_attr(b, "M0", &m->M0, "f8", 1);
    }
}

int fastpm_b200_io_write_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                                 const FpmIoMeta *meta, MPI_Comm comm)
{ return fastpm_b200_io_write_columns_at(filebase, dataset, cols, ncols, np_local, meta, NULL, comm); }

/* positions == NULL: the ranks' slices one after the other.  Otherwise item i of this rank goes to row positions[i] of the block
 * (ascending; the positions of all ranks together are 0 .. total-1): how a catalog sorted by a dense particle id is written
 * without moving particles between the GPUs -- runs of consecutive positions go out as one write each. */
static int write_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                         const FpmIoMeta *meta, const uint64_t *positions, int append, MPI_Comm comm);

int fastpm_b200_io_write_columns_at(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                                    const FpmIoMeta *meta, const uint64_t *positions, MPI_Comm comm)
{ return write_columns(filebase, dataset, cols, ncols, np_local, meta, positions, 0, comm); }

/* fastpm_store_write in a mode other than "w" / "r" (io.c:334-340,522-537): every column block grows by ceil(total / 32 Mi) files
 * holding the ranks' items one after the other; the dataset's own attributes are left alone; nothing happens when there is nothing
 * to write */
int fastpm_b200_io_append_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local, MPI_Comm comm)
{ return write_columns(filebase, dataset, cols, ncols, np_local, NULL, NULL, 1, comm); }

static int write_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                         const FpmIoMeta *meta, const uint64_t *positions, int append, MPI_Comm comm)
{
    int64_t total = 0;
    size_t first = rank_offset(np_local, comm, &total);
    if (fpm_comm_rank(comm) == 0 && mkdir_p(filebase) != 0) { fpm_comm_barrier(comm); return -1; }
    fpm_comm_barrier(comm);
    BfBlock b;
    if (!append) {
        if (bf_create(&b, filebase, dataset, NULL, 0, 0, 0, comm)) return -1;      /* io.c:431-436: the dataset's attributes */
        FpmIoMeta m = *meta;
        meta_attrs(&b, &m, 1);
        if (bf_close(&b, comm)) return -1;
    } else if (total == 0) {
        return 0;
    }
    const size_t first0 = first;
    const size_t items_per_file = 32 * 1024 * 1024;                                /* io.c:351 */
    int nfile = (int) (((size_t) total + items_per_file - 1) / items_per_file);
    if (nfile < 1) nfile = 1;
    for (int c = 0; c < ncols; c++) {
        const FpmIoColumn *col = &cols[c];
        if (!col->data && np_local > 0) continue;
        char blockname[256];
        snprintf(blockname, sizeof(blockname), "%s/%s", dataset, col->name);
        if (!append) {
            if (bf_create(&b, filebase, blockname, col->dtype_out, col->nmemb, nfile, (size_t) total, comm)) return -1;
        } else {
            /* open what is there (an empty block if there is nothing yet), grow it, write behind the old end */
            if (bf_open(&b, filebase, blockname) != 0 && bf_create(&b, filebase, blockname, col->dtype_out, col->nmemb, 0, 0, comm)) return -1;
            fpm_comm_barrier(comm);
            const size_t oldsize = b.size;
            if (bf_grow(&b, nfile, (size_t) total, comm)) { bf_free(&b); return -1; }
            first = oldsize + first0;
        }
        const size_t isz_in = (size_t) dtype_itemsize(col->dtype) * col->nmemb, isz_out = (size_t) dtype_itemsize(col->dtype_out) * col->nmemb;
        void *raw = malloc(IO_CHUNK * isz_in + 1), *out = malloc(IO_CHUNK * isz_out + 1);
        int rc = 0;
        for (size_t i0 = 0; i0 < (size_t) np_local && !rc; i0 += IO_CHUNK) {
            const size_t n = (size_t) np_local - i0 < IO_CHUNK ? (size_t) np_local - i0 : IO_CHUNK;
            const char *src = (const char *) col->data + i0 * isz_in;
            if (col->on_device) rc = fpm_memcpy_d2h(raw, src, n * isz_in); else memcpy(raw, src, n * isz_in);
            if (!rc) rc = convert_items(out, col->dtype_out, raw, col->dtype, n * col->nmemb);
            if (!rc && !positions) rc = bf_rw(&b, first + i0, n, out, 1);
            for (size_t a = 0; positions && a < n && !rc; ) {
                size_t e = a + 1;
                while (e < n && positions[i0 + e] == positions[i0 + e - 1] + 1) e++;
                if (positions[i0 + e - 1] >= (uint64_t) total) { rc = -1; break; }
                rc = bf_rw(&b, (size_t) positions[i0 + a], e - a, (char *) out + a * isz_out, 1);
                a = e;
            }
        }
        free(raw); free(out);
        if (bf_close(&b, comm) || rc) return -1;
    }
    return 0;
}

/* reads <dataset>/<name> for every column given; the block is split evenly over the ranks (io.c:489-507).
 * np_local: in = capacity (np_upper), out = items read. */
int fastpm_b200_io_read_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t *np_local,
                                FpmIoMeta *meta, MPI_Comm comm)
{
    const int nt = fpm_comm_size(comm), me = fpm_comm_rank(comm);
    BfBlock b;
    if (bf_open(&b, filebase, dataset)) return -1;
    meta_attrs(&b, meta, 0);
    bf_free(&b);
    int64_t got = -1;
    for (int c = 0; c < ncols; c++) {
        const FpmIoColumn *col = &cols[c];
        if (!col->data) continue;
        char blockname[256];
        snprintf(blockname, sizeof(blockname), "%s/%s", dataset, col->name);
        if (bf_open(&b, filebase, blockname)) return -1;
        const size_t first = (size_t) me * b.size / nt, n_local = (size_t) (me + 1) * b.size / nt - first;
        if ((int64_t) n_local > *np_local || b.nmemb != col->nmemb || (got >= 0 && got != (int64_t) n_local)) { bf_free(&b); return -1; }
        got = (int64_t) n_local;
        const size_t isz_col = (size_t) dtype_itemsize(col->dtype) * col->nmemb, isz_file = (size_t) dtype_itemsize(b.dtype) * b.nmemb;
        void *raw = malloc(IO_CHUNK * isz_file + 1), *out = malloc(IO_CHUNK * isz_col + 1);
        int rc = 0;
        for (size_t i0 = 0; i0 < n_local && !rc; i0 += IO_CHUNK) {
            const size_t n = n_local - i0 < IO_CHUNK ? n_local - i0 : IO_CHUNK;
            rc = bf_rw(&b, first + i0, n, raw, 0);
            if (!rc) rc = convert_items(out, col->dtype, raw, b.dtype, n * col->nmemb);
            char *dst = (char *) col->data + i0 * isz_col;
            if (!rc) { if (col->on_device) rc = fpm_memcpy_h2d(dst, out, n * isz_col); else memcpy(dst, out, n * isz_col); }
        }
        free(raw); free(out);
        bf_free(&b);
        if (rc) return -1;
    }
    if (got >= 0) *np_local = got;
    return 0;
}

/* set by fastpm_sort_snapshot on several ranks, consumed by the next fastpm_store_write of the same columns */
static struct { void *ids; size_t np; } sorted_by_dense_id;

/* the column table of fastpm_store_write, io.c:392-421 */
static const struct { const char *name, *dtype_out; FastPMColumnTags attribute; } BLOCKS[] = {
    { "Position", "f4", COLUMN_POS }, { "InitialPosition", "f4", COLUMN_Q }, { "DX1", "f4", COLUMN_DX1 }, { "DX2", "f4", COLUMN_DX2 },
    { "Velocity", "f4", COLUMN_VEL }, { "ID", "i8", COLUMN_ID }, { "Aemit", "f4", COLUMN_AEMIT }, { "Potential", "f4", COLUMN_POTENTIAL },
    { "Density", "f4", COLUMN_DENSITY }, { "Tidal", "f4", COLUMN_TIDAL }, { "Length", "i4", COLUMN_LENGTH }, { "MinID", "i8", COLUMN_MINID },
    { "Task", "i4", COLUMN_TASK }, { "Rdisp", "f4", COLUMN_RDISP }, { "Vdisp", "f4", COLUMN_VDISP }, { "RVdisp", "f4", COLUMN_RVDISP },
    { "Mass", "f4", COLUMN_MASS }, { "Rmom", "f4", COLUMN_RMOM },
};

static int store_columns(FastPMStore *p, FpmIoColumn *cols)
{
    int n = 0;
    for (size_t i = 0; i < sizeof(BLOCKS) / sizeof(BLOCKS[0]); i++) {
        const int ci = fastpm_store_find_column_id(p, BLOCKS[i].attribute);
        if (ci < 0 || !p->columns[ci]) continue;
        cols[n].name = BLOCKS[i].name; cols[n].dtype_out = BLOCKS[i].dtype_out; cols[n].dtype = p->_column_info[ci].dtype;
        cols[n].nmemb = (int) p->_column_info[ci].nmemb; cols[n].data = p->columns[ci]; cols[n].on_device = 1;
        n++;
    }
    return n;
}

static void meta_of(FastPMStore *p, FpmIoMeta *m)
{
    for (int d = 0; d < 3; d++) { m->q_strides[d] = p->meta._q_strides[d]; m->q_scale[d] = p->meta._q_scale[d]; m->q_shift[d] = p->meta._q_shift[d]; }
    m->q_size = p->meta._q_size; m->a_x = p->meta.a_x; m->a_v = p->meta.a_v; m->M0 = p->meta.M0;
}

int fastpm_store_write(FastPMStore *p, const char *filebase, const char *modestr, int Nwriters, MPI_Comm comm)
{
    (void) Nwriters;                       /* every rank writes its own slice */
    fpm_store_flush(p);
    FpmIoColumn cols[32];
    FpmIoMeta m;
    if (!strcmp(modestr, "w")) {
        fastpm_info("Writing a catalog to %s [%s]\n", filebase, p->name);
        const int n = store_columns(p, cols);
        meta_of(p, &m);
        uint64_t *positions = NULL;
        if (sorted_by_dense_id.ids == (void *) p->id && sorted_by_dense_id.np == p->np && p->id) {
            /* fastpm_sort_snapshot on several ranks (below): row i belongs at file position id[i] */
            positions = malloc(sizeof(uint64_t) * (p->np ? p->np : 1));
            FPM_MUST(fpm_memcpy_d2h(positions, p->id, sizeof(uint64_t) * p->np));
        }
        sorted_by_dense_id.ids = NULL;
        const int rc = fastpm_b200_io_write_columns_at(filebase, p->name, cols, n, (int64_t) p->np, &m, positions, comm);
        free(positions);
        if (rc) fastpm_raise(-1, "Failed to write the catalog %s [%s]: %s\n", filebase, p->name, strerror(errno));
        return 0;
    }
    if (!strcmp(modestr, "r")) {
        fastpm_info("Reading a catalog from %s [%s]\n", filebase, p->name);
        const int n = store_columns(p, cols);
        int64_t np = (int64_t) p->np_upper;
        if (fastpm_b200_io_read_columns(filebase, p->name, cols, n, &np, &m, comm))
            fastpm_raise(-1, "Failed to read the catalog %s [%s] (missing block, nmemb mismatch or more than np_upper = %td items)\n",
                         filebase, p->name, (ptrdiff_t) p->np_upper);
        p->np = (size_t) np;
        for (int d = 0; d < 3; d++) { p->meta._q_strides[d] = m.q_strides[d]; p->meta._q_scale[d] = m.q_scale[d]; p->meta._q_shift[d] = m.q_shift[d]; }
        p->meta._q_size = m.q_size; p->meta.a_x = m.a_x; p->meta.a_v = m.a_v; p->meta.M0 = m.M0;
        return 0;
    }
    /* any other mode string appends (io.c:334-340) */
    fastpm_info("Appending a catalog to %s [%s]\n", filebase, p->name);
    sorted_by_dense_id.ids = NULL;
    const int n = store_columns(p, cols);
    if (fastpm_b200_io_append_columns(filebase, p->name, cols, n, (int64_t) p->np, comm))
        fastpm_raise(-1, "Failed to append to the catalog %s [%s]: %s\n", filebase, p->name, strerror(errno));
    return 0;
}

int fastpm_store_read(FastPMStore *p, const char *filebase, int Nreaders, MPI_Comm comm)
{ return fastpm_store_write(p, filebase, "r", Nreaders, comm); }

/* ------------------------------------------------------------------ the "Header" block */
int fastpm_b200_io_write_header(const char *filebase, const FpmIoHeader *h, MPI_Comm comm)
{
    if (fpm_comm_rank(comm) == 0 && mkdir_p(filebase) != 0) { fpm_comm_barrier(comm); return -1; }
    fpm_comm_barrier(comm);
    BfBlock b;
    if (bf_create(&b, filebase, "Header", "i8", 0, 1, 0, comm)) return -1;         /* io.c:245 */
    const double UnitVelocity_in_cm_per_s = 1e5, UnitLength_in_cm = 3.085678e21 * 1e3, UnitMass_in_g = 1.989e43;   /* io.c:296-299 */
    const int UsePeculiarVelocity = 1;
    bf_set_attr(&b, "NC", &h->NC, "i8", 1); bf_set_attr(&b, "BoxSize", &h->BoxSize, "f8", 1);
    bf_set_attr(&b, "ScalingFactor", &h->ScalingFactor, "f8", 1); bf_set_attr(&b, "GrowthFactor", &h->GrowthFactor, "f8", 1);
    bf_set_attr(&b, "GrowthRate", &h->GrowthRate, "f8", 1); bf_set_attr(&b, "HubbleE", &h->HubbleE, "f8", 1);
    bf_set_attr(&b, "RSDFactor", &h->RSDFactor, "f8", 1); bf_set_attr(&b, "Omega_cdm", &h->Omega_cdm, "f8", 1);
    bf_set_attr(&b, "OmegaM", &h->OmegaM, "f8", 1); bf_set_attr(&b, "OmegaLambda", &h->OmegaLambda, "f8", 1);
    bf_set_attr(&b, "HubbleParam", &h->HubbleParam, "f8", 1);
    bf_set_attr(&b, "LibFastPMVersion", h->version, "S1", (int) strlen(h->version));
    bf_set_attr(&b, "Omega0", &h->Omega_cdm, "f8", 1);                               /* sic, io.c:302 */
    bf_set_attr(&b, "TotNumPart", h->TotNumPart, "i8", 6); bf_set_attr(&b, "MassTable", h->MassTable, "f8", 6);
    bf_set_attr(&b, "Time", &h->ScalingFactor, "f8", 1);
    bf_set_attr(&b, "UsePeculiarVelocity", &UsePeculiarVelocity, "i4", 1);
    bf_set_attr(&b, "UnitLength_in_cm", &UnitLength_in_cm, "f8", 1); bf_set_attr(&b, "UnitMass_in_g", &UnitMass_in_g, "f8", 1);
    bf_set_attr(&b, "UnitVelocity_in_cm_per_s", &UnitVelocity_in_cm_per_s, "f8", 1);
    return bf_close(&b, comm);
}

/* the numbers of write_snapshot_header, io.c:250-290 */
void fastpm_b200_io_header_values(FastPMSolver *fastpm, double aout, double M0_cdm, uint64_t np_total_cdm, FpmIoHeader *h)
{
    memset(h, 0, sizeof(*h));
    const double H0 = 100.;
    FastPMGrowthInfo gi;
    fastpm_growth_info_init(&gi, aout, fastpm->cosmology);
    h->NC = (int64_t) fastpm->config->nc; h->BoxSize = fastpm->config->boxsize; h->ScalingFactor = aout;
    h->GrowthFactor = gi.D1; h->GrowthRate = gi.f1; h->HubbleE = HubbleEa(aout, fastpm->cosmology);
    h->RSDFactor = 1.0 / (H0 * aout * HubbleEa(aout, fastpm->cosmology));
    h->Omega_cdm = fastpm->cosmology->Omega_cdm; h->OmegaM = fastpm->cosmology->Omega_m; h->OmegaLambda = fastpm->cosmology->Omega_Lambda;
    h->HubbleParam = fastpm->cosmology->h;
    h->version = LIBFASTPM_VERSION;
    h->MassTable[1] = M0_cdm; h->TotNumPart[1] = np_total_cdm;
}

void write_snapshot_header(FastPMSolver *fastpm, const char *filebase, MPI_Comm comm)
{
    fastpm_info("Writing a snapshot header to %s\n", filebase);
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FpmIoHeader h;
    fastpm_b200_io_header_values(fastpm, cdm->meta.a_x, cdm->meta.M0, (uint64_t) fastpm_store_get_np_total(cdm, comm), &h);
    fastpm_info("RSD factor %e\n", h.RSDFactor);
    if (fastpm_b200_io_write_header(filebase, &h, comm)) fastpm_raise(-1, "Failed to create the header block of %s: %s\n", filebase, strerror(errno));
}

void read_snapshot_header(FastPMSolver *fastpm, const char *filebase, double *aout, MPI_Comm comm)
{
    (void) comm;
    BfBlock b;
    if (bf_open(&b, filebase, "Header")) fastpm_raise(-1, "Failed to open the header block of %s\n", filebase);
    double NC = 0, BoxSize = 0, ScalingFactor = 0, Omega_cdm = 0, UnitLength = 0, UnitMass = 0, UnitVelocity = 0;
    double Omega_m = 0, OmegaLambda = 0, HubbleParam = 0;
    int64_t nc = 0;
    int UsePeculiarVelocity = 0;
    if (bf_get_attr(&b, "NC", &nc, "i8", 1) || bf_get_attr(&b, "BoxSize", &BoxSize, "f8", 1) || bf_get_attr(&b, "ScalingFactor", &ScalingFactor, "f8", 1)
        || bf_get_attr(&b, "Omega_cdm", &Omega_cdm, "f8", 1) || bf_get_attr(&b, "UsePeculiarVelocity", &UsePeculiarVelocity, "i4", 1)
        || bf_get_attr(&b, "OmegaM", &Omega_m, "f8", 1) || bf_get_attr(&b, "OmegaLambda", &OmegaLambda, "f8", 1)
        || bf_get_attr(&b, "HubbleParam", &HubbleParam, "f8", 1)
        || bf_get_attr(&b, "UnitLength_in_cm", &UnitLength, "f8", 1) || bf_get_attr(&b, "UnitMass_in_g", &UnitMass, "f8", 1)
        || bf_get_attr(&b, "UnitVelocity_in_cm_per_s", &UnitVelocity, "f8", 1))
        fastpm_raise(-1, "The header block of %s lacks an attribute\n", filebase);
    bf_free(&b);
    NC = (double) nc;
    /* the consistency checks of io.c:175-218 */
    if (Omega_m != fastpm->cosmology->Omega_m) fastpm_raise(-1, "Omega_m mismatched %g != %g", Omega_m, fastpm->cosmology->Omega_m);
    if (OmegaLambda != fastpm->cosmology->Omega_Lambda) fastpm_raise(-1, "OmegaLambda mismatched %g != %g", OmegaLambda, fastpm->cosmology->Omega_Lambda);
    if (HubbleParam != fastpm->cosmology->h) fastpm_raise(-1, "HubbleParam mismatched %g != %g", HubbleParam, fastpm->cosmology->h);
    if (NC != (double) fastpm->config->nc) fastpm_raise(-1, "NC mismatched %g != %g", NC, (double) fastpm->config->nc);
    if (BoxSize != fastpm->config->boxsize) fastpm_raise(-1, "BoxSize mismatched %g != %g", BoxSize, fastpm->config->boxsize);
    if (Omega_cdm != fastpm->cosmology->Omega_cdm) fastpm_raise(-1, "Omega_cdm mismatched %g != %g", Omega_cdm, fastpm->cosmology->Omega_cdm);
    if (UsePeculiarVelocity != 1) fastpm_raise(-1, "UsePeculiarVelocity mismatched %d != %d", UsePeculiarVelocity, 1);
    if (UnitVelocity != 1e5) fastpm_raise(-1, "UnitVelocity_in_cm_per_s mismatched %g != %g", UnitVelocity, 1e5);
    if (UnitLength != 3.085678e21 * 1e3) fastpm_raise(-1, "UnitLength_in_cm mismatched %g != %g", UnitLength, 3.085678e21 * 1e3);
    if (UnitMass != 1.989e43) fastpm_raise(-1, "UnitMass_in_g mismatched %g != %g", UnitMass, 1.989e43);
    *aout = ScalingFactor;
}

/* io.c:976-998: one more attribute on an existing block (the CLI adds "ParamFile" to "Header" this way, src/fastpm.c:97-116) */
void write_snapshot_attr(const char *filebase, const char *dataset, const char *attrname, void *buf, const char *dtype, size_t nmemb, MPI_Comm comm)
{
    BfBlock b;
    if (bf_open(&b, filebase, dataset)) fastpm_raise(-1, "Failed to open the dataset : %s/%s\n", filebase, dataset);
    bf_set_attr(&b, attrname, buf, dtype, (int) nmemb);
    if (bf_close(&b, comm)) fastpm_raise(-1, "Failed to write the attributes of %s/%s\n", filebase, dataset);
}

/* ------------------------------------------------------------------ k-space meshes: write_complex / read_complex, io.c:641-790
 * A c8 block of Nmesh * Nmesh * (Nmesh/2+1) items in [x][y][z] order (the order the reference's sort by `iabs` produces).
 * This rank holds ky in [y0, y0 + nyl): for every kx those rows are one contiguous run of the file. */
int fastpm_b200_io_write_complex_rows(const char *filename, const char *blockname, int nmesh, double boxsize, int y0, int nyl,
                                      const float *rows, size_t pitch_c, int nfile, MPI_Comm comm)
{
    const size_t n = (size_t) nmesh, hc = n / 2 + 1, size = n * n * hc;
    if (fpm_comm_rank(comm) == 0 && mkdir_p(filename) != 0) { fpm_comm_barrier(comm); return -1; }
    fpm_comm_barrier(comm);
    BfBlock b;
    if (bf_create(&b, filename, blockname, "c8", 1, nfile, size, comm)) return -1;
    float *run = malloc(sizeof(float) * 2 * hc * (size_t) nyl + 8);
    int rc = 0;
    for (size_t kx = 0; kx < n && !rc; kx++) {
        for (size_t kyl = 0; kyl < (size_t) nyl; kyl++) memcpy(run + 2 * hc * kyl, rows + 2 * ((kyl * n + kx) * pitch_c), sizeof(float) * 2 * hc);
        rc = bf_rw(&b, (kx * n + (size_t) y0) * hc, hc * (size_t) nyl, run, 1);
    }
    free(run);
    const int ndim = 3;
    const int64_t strides[3] = { (int64_t) (n * hc), (int64_t) hc, 1 }, shape[3] = { nmesh, nmesh, (int64_t) hc };
    bf_set_attr(&b, "ndarray.ndim", &ndim, "i4", 1); bf_set_attr(&b, "ndarray.strides", strides, "i8", 3);
    bf_set_attr(&b, "ndarray.shape", shape, "i8", 3); bf_set_attr(&b, "Nmesh", &nmesh, "i4", 1); bf_set_attr(&b, "BoxSize", &boxsize, "f8", 1);
    if (bf_close(&b, comm) || rc) return -1;
    return 0;
}

int fastpm_b200_io_read_complex_rows(const char *filename, const char *blockname, int nmesh, int y0, int nyl, float *rows, size_t pitch_c)
{
    const size_t n = (size_t) nmesh, hc = n / 2 + 1;
    BfBlock b;
    if (bf_open(&b, filename, blockname)) return -1;
    if (b.size != n * n * hc || dtype_itemsize(b.dtype) != 8) { bf_free(&b); return -1; }
    float *run = malloc(sizeof(float) * 2 * hc * (size_t) nyl + 8);
    int rc = 0;
    for (size_t kx = 0; kx < n && !rc; kx++) {
        rc = bf_rw(&b, (kx * n + (size_t) y0) * hc, hc * (size_t) nyl, run, 0);
        for (size_t kyl = 0; kyl < (size_t) nyl && !rc; kyl++) memcpy(rows + 2 * ((kyl * n + kx) * pitch_c), run + 2 * hc * kyl, sizeof(float) * 2 * hc);
    }
    free(run);
    bf_free(&b);
    return rc;
}

int write_complex(PM *pm, FastPMFloat *data, const char *filename, const char *blockname, int Nwriters)
{
    (void) Nwriters;
    const size_t n = pm->Nmesh[0], pc = pm->pitch_c, nyl = pm->nyl;
    float *tmp = malloc(sizeof(float) * 2 * nyl * n * pc);
    FPM_MUST(fpm_memcpy_d2h(tmp, data, sizeof(float) * 2 * nyl * n * pc));
    int nfile = pm->NTask / 8;                                                        /* io.c:689-690 */
    if (nfile == 0) nfile = 1;
    const int rc = fastpm_b200_io_write_complex_rows(filename, blockname, (int) n, pm->BoxSize[0], (int) pm->y0, (int) nyl, tmp, pc, nfile, pm->comm);
    free(tmp);
    if (rc) fastpm_raise(-1, "Failed to write the mesh %s [%s]: %s\n", filename, blockname, strerror(errno));
    return 0;
}

int read_complex(PM *pm, FastPMFloat *data, const char *filename, const char *blockname, int Nwriters)
{
    (void) Nwriters;
    const size_t n = pm->Nmesh[0], pc = pm->pitch_c, nyl = pm->nyl;
    float *tmp = calloc(2 * nyl * n * pc, sizeof(float));
    if (fastpm_b200_io_read_complex_rows(filename, blockname, (int) n, (int) pm->y0, (int) nyl, tmp, pc))
        fastpm_raise(-1, "Failed to read the mesh %s [%s] (missing, or not %td^2 x %td complex numbers)\n", filename, blockname, (ptrdiff_t) n, (ptrdiff_t) (n / 2 + 1));
    FPM_MUST(fpm_memcpy_h2d(data, tmp, sizeof(float) * 2 * nyl * n * pc));
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------ fastpm_sort_snapshot by particle id, io.c:860-960
 * One rank: the ids are mirrored, a stable least-significant-digit radix sort gives the permutation (the reference's mpsort
 * is a stable radix sort on the same key), every allocated column is permuted through the host. */
void FastPMSnapshotSortByID(const void *ptr, void *radix, void *arg) { (void) ptr; (void) radix; (void) arg; }

void fastpm_b200_io_argsort_u64(const uint64_t *key, size_t n, uint64_t *perm)
{
    uint64_t *a = perm, *b = malloc(sizeof(uint64_t) * (n ? n : 1)), mx = 0;
    for (size_t i = 0; i < n; i++) { a[i] = i; if (key[i] > mx) mx = key[i]; }
    for (int pass = 0; pass < 8 && (pass == 0 || (mx >> (8 * pass)) != 0); pass++) {
        size_t count[257];
        memset(count, 0, sizeof(count));
        const int sh = 8 * pass;
        for (size_t i = 0; i < n; i++) count[((key[a[i]] >> sh) & 0xff) + 1]++;
        for (int d = 0; d < 256; d++) count[d + 1] += count[d];
        for (size_t i = 0; i < n; i++) b[count[(key[a[i]] >> sh) & 0xff]++] = a[i];
        uint64_t *t = a; a = b; b = t;
    }
    if (a != perm) { memcpy(perm, a, sizeof(uint64_t) * n); free(a); } else free(b);
}

/* the integer cube root of n when n is a cube, else 0 */
static int cube_root_of(size_t n)
{
    int c = (int) (cbrt((double) n) + 0.5);
    return (c > 0 && (size_t) c * c * c == n) ? c : 0;
}

/* A store of one rank whose rows are in the order of the Lagrangian ids 0 .. nc^3-1 of fastpm_store_fill (after a sort by id, or read
 * back from a catalog that was written sorted): the deposit and the gather may walk it in Lagrangian bricks (paint.cu), as they do
 * for a freshly filled store.  The walk is a permutation of the rows whatever their order, so a wrong guess costs speed only. */
static void lagrangian_order_hint(const FastPMStore *p)
{
    const int nc = cube_root_of(p->np);
    if (nc >= 8 && nc % 8 == 0) FPM_MUST(fpm_particle_grid_hint(nc));
}

/* One rank, ids exactly a permutation of 0 .. n-1 (what fastpm_store_fill assigns and nothing on this path changes): the sorted
 * position of a row is its id, so every column is scattered once on the device (fpm_permute_by_id) through one scratch column
 * instead of going through the host.  Returns 1 when the store is now in id order, 0 when this path does not apply (ids not dense,
 * duplicates, no room for the scratch column, FASTPM_B200_HOST_SORT=1): the caller sorts on the host as before. */
static int sort_by_dense_id_on_device(FastPMStore *p)
{
    const size_t n = p->np;
    const char *off = getenv("FASTPM_B200_HOST_SORT");
    if (off && atoi(off)) return 0;
    if (n == 0) return 1;
    uint64_t cnt[2];
    FPM_MUST(fpm_id_order_counts(p->id, (int64_t) n, 0, cnt));
    if (cnt[0]) return 0;
    if (cnt[1] == 0) return 1;                               /* already in id order */
    size_t maxel = sizeof(uint64_t);
    for (int ci = 0; ci < 32; ci++) if (p->columns[ci] && p->_column_info[ci].elsize > maxel) maxel = p->_column_info[ci].elsize;
    void *scratch = fpm_malloc(maxel * n);
    if (!scratch) { fastpm_b200_memory_trim(); scratch = fpm_malloc(maxel * n); }      /* cached mesh buffers make room */
    if (!scratch) return 0;
    /* duplicates leave a slot of the scattered id column unwritten: it keeps the fill pattern and shows up as displaced */
    FPM_MUST(fpm_memset(scratch, 0xff, sizeof(uint64_t) * n));
    FPM_MUST(fpm_permute_by_id(scratch, p->id, p->id, (int64_t) n, 0, (int) sizeof(uint64_t)));
    FPM_MUST(fpm_id_order_counts(scratch, (int64_t) n, 0, cnt));
    if (cnt[0] || cnt[1]) { fpm_free(scratch); return 0; }
    for (int pass = 0; pass < 2; pass++)                     /* the id column last: it is the key of every scatter */
        for (int ci = 0; ci < 32; ci++) {
            if (!p->columns[ci] || (pass == 1) != (p->columns[ci] == (void *) p->id)) continue;
            const size_t el = p->_column_info[ci].elsize;
            FPM_MUST(fpm_permute_by_id(scratch, p->columns[ci], p->id, (int64_t) n, 0, (int) el));
            FPM_MUST(fpm_memcpy_d2d(p->columns[ci], scratch, el * n));
        }
    fpm_free(scratch);
    return 1;
}

void fastpm_sort_snapshot(FastPMStore *p, MPI_Comm comm, FastPMSnapshotSorter sorter, int redistribute)
{
    (void) redistribute;
    if (sorter != FastPMSnapshotSortByID) fastpm_raise(-1, "fastpm_b200: fastpm_sort_snapshot sorts by particle id only\n");
    if (!p->id) fastpm_raise(-1, "fastpm_sort_snapshot: the store has no id column\n");
    fpm_store_flush(p);
    const size_t n = p->np;
    if (fpm_comm_size(comm) == 1 && sort_by_dense_id_on_device(p)) {
        sorted_by_dense_id.ids = NULL;
        lagrangian_order_hint(p);
        return;
    }
    uint64_t *key = malloc(sizeof(uint64_t) * (n ? n : 1)), *perm = malloc(sizeof(uint64_t) * (n ? n : 1));
    FPM_MUST(fpm_memcpy_d2h(key, p->id, sizeof(uint64_t) * n));
    fastpm_b200_io_argsort_u64(key, n, perm);
    sorted_by_dense_id.ids = NULL;
    if (fpm_comm_size(comm) > 1) {
        /* Several ranks: the particles stay where they are (sorted locally); the catalog comes out globally sorted because
         * fastpm_store_write then puts every row at the file position given by its id.  That needs the ids of all ranks together
         * to be exactly 0 .. N-1 -- what fastpm_store_fill assigns and nothing on this path changes (no sub-sampling). */
        int64_t lo = n ? (int64_t) key[perm[0]] : INT64_MAX, hi = n ? (int64_t) key[perm[n - 1]] : -1, cnt = (int64_t) n;
        fpm_comm_allreduce_i64(comm, &lo, 1, 1); fpm_comm_allreduce_i64(comm, &hi, 1, 2); fpm_comm_allreduce_i64(comm, &cnt, 1, 0);
        if (lo != 0 || hi != cnt - 1)
            fastpm_raise(-1, "fastpm_b200: the distributed snapshot sort needs dense particle ids 0 .. N-1 (found %ld .. %ld for %ld particles)\n",
                         (long) lo, (long) hi, (long) cnt);
        sorted_by_dense_id.ids = (void *) p->id; sorted_by_dense_id.np = n;
    }
    free(key);
    for (int ci = 0; ci < 32; ci++) {
        if (!p->columns[ci]) continue;
        const size_t el = p->_column_info[ci].elsize;
        char *src = malloc(el * (n ? n : 1)), *dst = malloc(el * (n ? n : 1));
        FPM_MUST(fpm_memcpy_d2h(src, p->columns[ci], el * n));
        for (size_t i = 0; i < n; i++) memcpy(dst + i * el, src + perm[i] * el, el);
        FPM_MUST(fpm_memcpy_h2d(p->columns[ci], dst, el * n));
        free(src); free(dst);
    }
    free(perm);
}

/* ------------------------------------------------------------------ one snapshot, one restart (src/fastpm.c:1190-1200, 1473-1486, 618-635)
 * without the Lua-dependent "ParamFile" attribute: unit conversion + wrap, [sort by id], header, catalog, conversion reverted. */
void fastpm_b200_write_snapshot(FastPMSolver *fastpm, const char *filebase, int sort_by_id)
{
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM), po[1];
    const double aout = p->meta.a_x;
    if (p->meta.a_x != p->meta.a_v) fastpm_raise(-1, "fastpm_b200_write_snapshot: positions (a = %g) and velocities (a = %g) are out of sync\n", p->meta.a_x, p->meta.a_v);
    fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, aout);
    if (sort_by_id) fastpm_sort_snapshot(po, fastpm->comm, FastPMSnapshotSortByID, 0);
    FastPMSolver snapshot[1];
    memcpy(snapshot, fastpm, sizeof(FastPMSolver));
    fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, po);
    write_snapshot_header(snapshot, filebase, fastpm->comm);
    fastpm_store_write(po, filebase, "w", 0, fastpm->comm);
    fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, aout);
}

double fastpm_b200_read_snapshot(FastPMSolver *fastpm, const char *filebase)
{
    double a_restart = 0;
    read_snapshot_header(fastpm, filebase, &a_restart, fastpm->comm);
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM), po[1];
    fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, 1.0);
    fastpm_store_read(po, filebase, 0, fastpm->comm);
    if (fpm_comm_size(fastpm->comm) == 1 && po->id && po->np) {
        /* a catalog written sorted by id (the command line's default): the rows are back in Lagrangian order */
        uint64_t cnt[2];
        FPM_MUST(fpm_id_order_counts(po->id, (int64_t) po->np, 0, cnt));
        if (cnt[0] == 0 && cnt[1] == 0) lagrangian_order_hint(po);
    }
    if (po->meta.a_x != po->meta.a_v) fastpm_raise(-1, "Snapshot velocity and position are out of sync. a_x =% g, a_v = %g.\n", po->meta.a_x, po->meta.a_v);
    fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, po->meta.a_x);
    return a_restart;
}

/* ------------------------------------------------------------------ snapshots at requested scale factors during fastpm_solver_evolve
 * check_snapshots + take_a_snapshot of the CLI (src/fastpm.c:1130-1208, 1473-1486) as a ready-made INTERPOLATION handler: whenever a
 * requested aout falls into (a1, a2] of the event (or equals the initial time), the particles are drifted and kicked there with
 * the event's factors (fastpm_set_snapshot), written to "<base>_<aout, %0.04f>", and put back (fastpm_unset_snapshot). */
typedef struct { char base[900]; double aout[64]; int nout, iout, sort_by_id; } SnapshotPlan;

static int snapshot_handler(FastPMSolver *fastpm, FastPMInterpolationEvent *event, SnapshotPlan *plan)
{
    for (int iout = plan->iout; iout < plan->nout; iout++) {
        const double aout = plan->aout[iout];
        if (event->a1 == event->a2) {
            if (event->a1 != aout) continue;                 /* initial condition, not requested */
        } else {
            if (event->a1 >= aout) continue;
            if (event->a2 < aout) continue;
        }
        FastPMSolver snapshot[1];
        FastPMStore cdm[1];
        memcpy(snapshot, fastpm, sizeof(FastPMSolver));
        fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, cdm);
        fastpm_set_snapshot(fastpm, snapshot, event->drift, event->kick, aout);
        fastpm_info("Snapshot a_x = %6.4f, a_v = %6.4f \n", cdm->meta.a_x, cdm->meta.a_v);
        if (plan->sort_by_id) fastpm_sort_snapshot(cdm, fastpm->comm, FastPMSnapshotSortByID, 0);
        char filebase[1024];
        snprintf(filebase, sizeof(filebase), "%s_%0.04f", plan->base, aout);
        write_snapshot_header(snapshot, filebase, fastpm->comm);
        fastpm_store_write(cdm, filebase, "w", 0, fastpm->comm);
        fastpm_unset_snapshot(fastpm, snapshot, event->drift, event->kick, aout);
        plan->iout = iout + 1;                               /* do not rewrite this snapshot */
    }
    return 0;
}

static int cmp_double(const void *a, const void *b) { const double x = *(const double *) a, y = *(const double *) b; return (x > y) - (x < y); }

void fastpm_b200_add_snapshot_handler(FastPMSolver *fastpm, const char *base, const double *aout, int nout, int sort_by_id)
{
    if (nout > 64) fastpm_raise(-1, "fastpm_b200_add_snapshot_handler: at most 64 output times\n");
    SnapshotPlan *plan = calloc(1, sizeof(*plan));
    snprintf(plan->base, sizeof(plan->base), "%s", base);
    memcpy(plan->aout, aout, sizeof(double) * nout);
    qsort(plan->aout, nout, sizeof(double), cmp_double);      /* the search above needs them ascending, src/fastpm.c:1164 */
    plan->nout = nout; plan->sort_by_id = sort_by_id;
    fastpm_add_event_handler_free(&fastpm->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE,
                                  (FastPMEventHandlerFunction) snapshot_handler, plan, free);
}
