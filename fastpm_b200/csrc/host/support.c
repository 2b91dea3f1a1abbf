/* fastpm_b200 host layer -- logging, events, tagged device allocator, named clocks.
 * Mirrors libfastpm/logging.c, events.c, memory.c, prof.c, libfastpm.c of the reference in behaviour
 * (message format, LIFO checks, abort on error); the allocator hands out DEVICE memory. */
#include "internal.h"
#include <sys/time.h>
#include <time.h>

const char *LIBFASTPM_VERSION = "fastpm_b200-0.1 (libfastpm API mirror, sm_100a)";

/* ------------------------------------------------------------------ logging (logging.c) */
typedef struct MsgHandler { fastpm_msg_handler handler; void *userdata; MPI_Comm comm; struct MsgHandler *prev; } MsgHandler;
static MsgHandler handler_data = { NULL, NULL, 0, NULL };

void fastpm_void_msg_handler(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                             const char *message, MPI_Comm comm, void *userdata)
{
    (void) errcode; (void) userdata;
    int rank = fpm_comm_rank(comm);
    if (level == ERROR) {
        if (type == COLLECTIVE) { if (rank == 0) { fprintf(stdout, "%s", message); fflush(stdout); } }
        else { fprintf(stdout, "ThisTask = %d %s", rank, message); fflush(stdout); }
        abort();
    }
}

void fastpm_default_msg_handler(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                                const char *message, MPI_Comm comm, void *userdata)
{
    (void) errcode; (void) userdata;
    int rank = fpm_comm_rank(comm);
    if (type == COLLECTIVE) {
        if (rank == 0) { fprintf(stdout, "%s", message); fflush(stdout); }
    } else {
        if (level == ERROR) fprintf(stdout, "ThisTask = %d %s", rank, message);
        else fprintf(stdout, "%s", message);
        fflush(stdout);
    }
    if (rank != 0 && type == COLLECTIVE) return;
    if (level == ERROR) abort();
}

void fastpm_set_msg_handler(fastpm_msg_handler handler, MPI_Comm comm, void *userdata)
{ handler_data.handler = handler; handler_data.userdata = userdata; handler_data.comm = comm; }

void fastpm_push_msg_handler(fastpm_msg_handler handler, MPI_Comm comm, void *userdata)
{
    MsgHandler *prev = malloc(sizeof(*prev));
    *prev = handler_data;
    handler_data.prev = prev;
    fastpm_set_msg_handler(handler, comm, userdata);
}

void fastpm_pop_msg_handler(void)
{
    MsgHandler *prev = handler_data.prev;
    if (!prev) return;
    handler_data = *prev;
    free(prev);
}

static double wallclock(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* "[ seconds ]: text [ file:line ]" on every line, as logging.c:137-185 formats it */
static void emit(const char *file, int line, enum FastPMLogLevel level, enum FastPMLogType type, int code, const char *fmt, va_list ap)
{
    if (!handler_data.handler) fastpm_set_msg_handler(fastpm_default_msg_handler, MPI_COMM_WORLD, NULL);
    static double t0 = -1;
    if (t0 < 0) t0 = wallclock();
    char body[4096];
    vsnprintf(body, sizeof(body), fmt, ap);
    char head[64], tail[64];
    snprintf(head, sizeof(head), "[ %012.04f ]: ", wallclock() - t0);
    snprintf(tail, sizeof(tail), " [ %.20s:%d ]", file, line);
    size_t cap = strlen(body) * 2 + 4096;
    char *out = malloc(cap), *q = out;
    q += sprintf(q, "%s", head);
    size_t len = strlen(body);
    for (size_t i = 0; i < len; i++) {
        if (body[i] == '\n' && i + 1 == len) q += sprintf(q, "%s", tail);
        *q++ = body[i];
        if (body[i] == '\n' && i + 1 != len) q += sprintf(q, "%s", head);
        if ((size_t) (q - out) > cap - 256) break;
    }
    if (len > 0 && body[len - 1] != '\n') { q += sprintf(q, "%s", tail); *q++ = '\n'; }
    *q = 0;
    handler_data.handler(level, type, code, out, handler_data.comm, handler_data.userdata);
    free(out);
}

void fastpm_info_(const char *file, int line, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); emit(file, line, INFO, COLLECTIVE, 0, fmt, ap); va_end(ap); }
void fastpm_ilog_(const char *file, int line, const enum FastPMLogLevel level, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); emit(file, line, level, INDIVIDUAL, 0, fmt, ap); va_end(ap); }
void fastpm_raise_(const char *file, int line, const int code, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); emit(file, line, ERROR, INDIVIDUAL, code, fmt, ap); va_end(ap); abort(); }

/* ------------------------------------------------------------------ events (events.c) */
void fastpm_add_event_handler_free(FastPMEventHandler **handlers, const char *type, enum FastPMEventStage stage,
                                   FastPMEventHandlerFunction function, void *userdata, void (*freefn)(void *))
{
    FastPMEventHandler *nh = malloc(sizeof(*nh));
    strncpy(nh->type, type, 31); nh->type[31] = 0;
    nh->stage = stage; nh->function = function; nh->userdata = userdata; nh->free = freefn; nh->next = NULL;
    /* append: handlers run in registration order */
    if (!*handlers) { *handlers = nh; return; }
    FastPMEventHandler *h = *handlers;
    while (h->next) h = h->next;
    h->next = nh;
}
void fastpm_add_event_handler(FastPMEventHandler **handlers, const char *type, enum FastPMEventStage stage,
                              FastPMEventHandlerFunction function, void *userdata)
{ fastpm_add_event_handler_free(handlers, type, stage, function, userdata, NULL); }

void fastpm_remove_event_handler(FastPMEventHandler **handlers, const char *type, enum FastPMEventStage stage,
                                 FastPMEventHandlerFunction function, void *userdata)
{
    FastPMEventHandler **pp = handlers;
    while (*pp) {
        FastPMEventHandler *h = *pp;
        if (!strcmp(h->type, type) && h->stage == stage && h->function == function && h->userdata == userdata) {
            *pp = h->next;
            if (h->free) h->free(h->userdata);
            free(h);
            return;
        }
        pp = &h->next;
    }
    fastpm_raise(-1, "removing an event handler that was never added.\n");
}

void fastpm_destroy_event_handlers(FastPMEventHandler **handlers)
{
    FastPMEventHandler *h = *handlers;
    while (h) { FastPMEventHandler *n = h->next; if (h->free) h->free(h->userdata); free(h); h = n; }
    *handlers = NULL;
}

/* Handlers that only look at the event (a progress line, say) and never at the particles: emitting an event to them does not
 * apply the queued kicks and drifts, so the fused particle update between two force evaluations survives (an ordinary
 * TRANSITION handler costs it: 18 kick + 18 drift passes instead of 1 + 9 fused ones over ten steps). */
#define NPASSIVE 16
static FastPMEventHandlerFunction passive_handlers[NPASSIVE];
static int npassive = 0;
void fastpm_b200_mark_handler_passive(FastPMEventHandlerFunction function)
{
    for (int i = 0; i < npassive; i++) if (passive_handlers[i] == function) return;
    if (npassive < NPASSIVE) passive_handlers[npassive++] = function;
}
static int handler_is_passive(FastPMEventHandlerFunction function)
{
    for (int i = 0; i < npassive; i++) if (passive_handlers[i] == function) return 1;
    return 0;
}

void fastpm_emit_event(FastPMEventHandler *handlers, const char *type, enum FastPMEventStage stage, FastPMEvent *event, void *context)
{
    strncpy(event->type, type, 31); event->type[31] = 0;
    event->stage = stage;
    for (FastPMEventHandler *h = handlers; h; h = h->next)
        if (h->stage == stage && !strcmp(h->type, type)) {
            /* a handler may look at the particles: no queued kick / drift may be outstanding (unless it promised not to) */
            if (!handler_is_passive(h->function)) fpm_store_flush(NULL);
            h->function(context, event, h->userdata);
        }
}

/* ------------------------------------------------------------------ memory (memory.c): tagged DEVICE allocator */
struct MemoryBlock { void *p; size_t size; MemoryBlock *prev; char tag[128]; int loc; };
static FastPMMemory GMEM;
static MemoryBlock *blocks = NULL;           /* most recent first */
/* freed blocks kept for reuse: cudaMalloc/cudaFree of tens of GB per force step would serialise the stream */
#define NCACHE 8
static struct { void *p; size_t size; } cache[NCACHE];

void fastpm_memory_init(FastPMMemory *m, size_t total_bytes)
{
    memset(m, 0, sizeof(*m));
    m->alignment = 4096;
    m->total_bytes = total_bytes ? total_bytes : (size_t) 1 << 62;
    m->free_bytes = m->total_bytes;
}
void fastpm_memory_set_handlers(FastPMMemory *m, fastpm_memory_func abortfunc, fastpm_memory_func peakfunc, void *userdata)
{ m->abortfunc = abortfunc; m->peakfunc = peakfunc; m->userdata = userdata; }

void fastpm_memory_dump_status_str(FastPMMemory *m, char *buf, int n)
{
    int off = snprintf(buf, n, "device memory in use: %zu bytes, peak %zu bytes\n", m->used_bytes, m->peak_bytes);
    for (MemoryBlock *b = blocks; b && off < n - 160; b = b->prev)
        off += snprintf(buf + off, n - off, "  %p %012zu : %s\n", b->p, b->size, b->tag);
}

/* ---- symmetric arena (several GPUs): one device allocation per process, mapped once into every peer with CUDA IPC
 * (host/comm.c).  Every rank performs the same sequence of allocations with the same sizes (meshes, particle columns,
 * exchange buffers), and the first-fit policy below is deterministic, so a buffer sits at the SAME offset in every
 * rank's arena: the address of a peer's copy is peer_base + (p - my_base), no handle exchange per call, and no mapping
 * can go stale because the arena is never freed while the communicator lives. */
static char *arena_base = NULL;
static size_t arena_size = 0;
typedef struct { size_t off, size; } ArenaBlock;
static ArenaBlock *arena_blocks = NULL;     /* sorted by offset */
static int arena_nblocks = 0;
#define ARENA_ALIGN ((size_t) 1 << 20)

int fastpm_b200_arena_init(size_t bytes)
{
    if (arena_base) return 0;
    bytes = bytes / ARENA_ALIGN * ARENA_ALIGN;
    arena_base = fpm_malloc(bytes);
    if (!arena_base) return -1;
    arena_size = bytes;
    return 0;
}
void *fastpm_b200_arena_base(void) { return arena_base; }
size_t fastpm_b200_arena_size(void) { return arena_size; }
int fastpm_b200_arena_contains(const void *p) { return arena_base && (const char *) p >= arena_base && (const char *) p < arena_base + arena_size; }

/* the largest block arena_alloc could still hand out */
size_t fastpm_b200_arena_largest_free(void)
{
    if (!arena_base) return 0;
    size_t best = 0, off = 0;
    for (int i = 0; i < arena_nblocks; i++) {
        if (arena_blocks[i].off - off > best) best = arena_blocks[i].off - off;
        off = arena_blocks[i].off + arena_blocks[i].size;
    }
    if (arena_size - off > best) best = arena_size - off;
    return best;
}

static void *arena_alloc(size_t s)
{
    s = (s + ARENA_ALIGN - 1) / ARENA_ALIGN * ARENA_ALIGN;
    size_t off = 0;
    int at = 0;
    for (; at < arena_nblocks; at++) {
        if (arena_blocks[at].off - off >= s) break;
        off = arena_blocks[at].off + arena_blocks[at].size;
    }
    if (at == arena_nblocks && arena_size - off < s) return NULL;
    arena_blocks = realloc(arena_blocks, sizeof(ArenaBlock) * (arena_nblocks + 1));
    memmove(arena_blocks + at + 1, arena_blocks + at, sizeof(ArenaBlock) * (arena_nblocks - at));
    arena_blocks[at].off = off; arena_blocks[at].size = s;
    arena_nblocks++;
    return arena_base + off;
}
static void arena_free(void *p)
{
    const size_t off = (size_t) ((char *) p - arena_base);
    for (int i = 0; i < arena_nblocks; i++) {
        if (arena_blocks[i].off != off) continue;
        memmove(arena_blocks + i, arena_blocks + i + 1, sizeof(ArenaBlock) * (arena_nblocks - i - 1));
        arena_nblocks--;
        return;
    }
    fastpm_raise(-1, "arena_free: %p is not the start of an arena block\n", p);
}
void fastpm_b200_arena_destroy(void)
{
    if (arena_nblocks) fastpm_raise(-1, "arena destroyed with %d live blocks\n", arena_nblocks);
    if (arena_base) fpm_free(arena_base);
    free(arena_blocks);
    arena_base = NULL; arena_size = 0; arena_blocks = NULL; arena_nblocks = 0;
}

/* Host-only self test of the arena's placement policy (no device memory is touched: a fake base address is used).  The
 * multi-GPU design rests on it being deterministic -- the same sequence of sizes must give the same offsets on every rank --
 * first-fit, 1 MiB granular, reusing freed gaps and failing cleanly when full.  Returns 0 on success, else the failed check. */
int fastpm_b200_arena_selftest(void)
{
    if (arena_base) return -1;                         /* only before a real arena exists */
    int bad = 0;
    const size_t MiB = ARENA_ALIGN;
    for (int pass = 0; pass < 2 && !bad; pass++) {     /* two identical passes: identical offsets (determinism) */
        static size_t first[6];
        arena_base = (char *) (uintptr_t) 0x100000000ull; arena_size = 64 * MiB; arena_blocks = NULL; arena_nblocks = 0;
        char *a = arena_alloc(1), *b = arena_alloc(10 * MiB), *c = arena_alloc(3 * MiB + 5), *d = arena_alloc(20 * MiB);
        size_t off[6] = { (size_t) (a - arena_base), (size_t) (b - arena_base), (size_t) (c - arena_base), (size_t) (d - arena_base), 0, 0 };
        if (off[0] != 0 || off[1] != MiB || off[2] != 11 * MiB || off[3] != 15 * MiB) bad = 1;      /* packed, rounded up to 1 MiB */
        arena_free(b);                                   /* a 10 MiB gap at 1 MiB */
        char *e = arena_alloc(12 * MiB);                 /* does not fit the gap: goes behind d */
        char *f = arena_alloc(4 * MiB);                  /* first fit: into the gap */
        off[4] = (size_t) (e - arena_base); off[5] = (size_t) (f - arena_base);
        if (!bad && (off[4] != 35 * MiB || off[5] != MiB)) bad = 2;
        if (!bad && fastpm_b200_arena_largest_free() != 17 * MiB) bad = 8;     /* the tail; the inner gap is 6 MiB */
        if (!bad && arena_alloc(30 * MiB) != NULL) bad = 3;                    /* 47 MiB used at the top: 17 left */
        if (!bad && arena_alloc(17 * MiB) == NULL) bad = 4;                    /* exactly the tail */
        if (!bad && !fastpm_b200_arena_contains(f)) bad = 5;
        if (!bad && fastpm_b200_arena_contains(arena_base + arena_size)) bad = 6;
        if (pass == 0) memcpy(first, off, sizeof(off));
        else if (!bad && memcmp(first, off, sizeof(off))) bad = 7;
        free(arena_blocks);
    }
    arena_base = NULL; arena_size = 0; arena_blocks = NULL; arena_nblocks = 0;
    return bad;
}

/* Is there room for an optional block of `need` bytes (plus `slack`)?  Several GPUs: the largest free block of the symmetric arena
 * (identical on every rank).  One GPU: free device memory plus what the buffer cache would give back when asked. */
int fastpm_b200_device_room(size_t need, size_t slack)
{
    if (arena_base) return fastpm_b200_arena_largest_free() >= need + slack;
    size_t fr = 0, tot = 0, cached = 0;
    if (fpm_device_mem_info(&fr, &tot) != 0) return 0;
    for (int i = 0; i < NCACHE; i++) if (cache[i].p) cached += cache[i].size;
    return fr + cached >= need + slack;
}

void *fastpm_memory_alloc_details(FastPMMemory *m, const char *name, size_t s, enum FastPMMemoryLocation loc, const char *file, const int line)
{
    if (s == 0) s = 1;
    if (m->used_bytes + s > m->total_bytes) {
        if (m->abortfunc) m->abortfunc(m, m->userdata);
        fastpm_raise(-1, "Out of memory bound allocating %zu bytes for %s at %s:%d\n", s, name, file, line);
    }
    void *p = NULL;
    if (arena_base) {
        p = arena_alloc(s);
        if (!p) {
            if (m->abortfunc) m->abortfunc(m, m->userdata);
            fastpm_raise(-1, "Out of device memory: the %zu-byte arena cannot hold %zu more bytes for %s at %s:%d (FASTPM_B200_ARENA_GB)\n",
                         arena_size, s, name, file, line);
        }
    }
    if (!p) for (int i = 0; i < NCACHE; i++) if (cache[i].p && cache[i].size == s) { p = cache[i].p; cache[i].p = NULL; break; }
    if (!p) p = fpm_malloc(s);
    if (!p) {
        /* drop the cache and retry once */
        for (int i = 0; i < NCACHE; i++) if (cache[i].p) { fpm_free(cache[i].p); cache[i].p = NULL; }
        p = fpm_malloc(s);
    }
    if (!p) {
        if (m->abortfunc) m->abortfunc(m, m->userdata);
        fastpm_raise(-1, "Out of device memory allocating %zu bytes for %s at %s:%d: %s\n", s, name, file, line, fpm_last_error());
    }
    MemoryBlock *b = malloc(sizeof(*b));
    b->p = p; b->size = s; b->prev = blocks; b->loc = loc;
    snprintf(b->tag, sizeof(b->tag), "%s:%.80s:%d", name, file, line);
    blocks = b;
    m->used_bytes += s;
    if (m->used_bytes > m->peak_bytes) { m->peak_bytes = m->used_bytes; if (m->peakfunc) m->peakfunc(m, m->userdata); }
    return p;
}

void fastpm_memory_free(FastPMMemory *m, void *p)
{
    MemoryBlock **pp = &blocks, *b = NULL;
    /* HEAP and STACK blocks are freed last-in first-out like the reference (memory.c:268-270) */
    for (; *pp; pp = &(*pp)->prev) if ((*pp)->p == p) { b = *pp; break; }
    if (!b) fastpm_raise(-1, "Freeing a pointer that was not allocated by fastpm_memory_alloc: %p\n", p);
    if (b->loc != FASTPM_MEMORY_FLOATING) {
        for (MemoryBlock *q = blocks; q != b; q = q->prev)
            if (q->loc == b->loc) fastpm_raise(-1, "Freeing %s out of order; %s was allocated later and is still alive.\n", b->tag, q->tag);
    }
    *pp = b->prev;
    m->used_bytes -= b->size;
    if (fastpm_b200_arena_contains(p)) { arena_free(p); free(b); return; }
    int slot = -1;
    for (int i = 0; i < NCACHE; i++) if (!cache[i].p) { slot = i; break; }
    if (slot < 0) {                     /* evict the smallest */
        slot = 0;
        for (int i = 1; i < NCACHE; i++) if (cache[i].size < cache[slot].size) slot = i;
        fpm_free(cache[slot].p);
    }
    cache[slot].p = p; cache[slot].size = b->size;
    free(b);
}

void fastpm_memory_destroy(FastPMMemory *m)
{
    if (blocks) {
        char buf[4096];
        fastpm_memory_dump_status_str(m, buf, sizeof(buf));
        fastpm_raise(-1, "Memory leak: blocks are still allocated at destroy (memory.c:118-130)\n%s", buf);
    }
    for (int i = 0; i < NCACHE; i++) if (cache[i].p) { fpm_free(cache[i].p); cache[i].p = NULL; }
}

void fastpm_b200_memory_trim(void)
{
    for (int i = 0; i < NCACHE; i++) if (cache[i].p) { fpm_free(cache[i].p); cache[i].p = NULL; }
}

FastPMMemory *_libfastpm_get_gmem(void) { return &GMEM; }

static int lib_inited = 0;
void libfastpm_init(void)
{
    if (lib_inited) return;
    if (fpm_device_count() <= 0) {
        fprintf(stderr, "libfastpm_init: fastpm_b200 needs a CUDA device (sm_100a); none is visible. There is no CPU path.\n");
        abort();
    }
    const char *dev = getenv("FASTPM_B200_DEVICE");
    const char *lr = getenv("LOCAL_RANK");
    int device = dev ? atoi(dev) : (lr ? atoi(lr) : 0);
    if (fpm_device_init(device % fpm_device_count()) != 0) { fprintf(stderr, "libfastpm_init: %s\n", fpm_last_error()); abort(); }
    fastpm_set_msg_handler(fastpm_void_msg_handler, MPI_COMM_WORLD, NULL);
    fastpm_memory_init(&GMEM, 0);
    lib_inited = 1;
}
void libfastpm_cleanup(void) { if (lib_inited) { fastpm_memory_destroy(&GMEM); lib_inited = 0; } }
void libfastpm_set_memory_bound(size_t size) { GMEM.total_bytes = size ? size : (size_t) 1 << 62; }

/* ------------------------------------------------------------------ clocks (prof.c) */
struct FastPMClock { double tcum, t0; char file[128], func[128], name[128]; struct FastPMClock *next; };
static FastPMClock *clock_head = NULL;
static int clocks_sync = -1;

FastPMClock *fastpm_clock_find(const char *file, const char *func, const char *name)
{
    for (FastPMClock *p = clock_head; p; p = p->next)
        if (!strcmp(p->file, file) && !strcmp(p->func, func) && !strcmp(p->name, name)) return p;
    FastPMClock *p = calloc(1, sizeof(*p));
    strncpy(p->file, file, 120); strncpy(p->func, func, 120); strncpy(p->name, name, 120);
    p->next = clock_head; clock_head = p;
    return p;
}
/* kernels are asynchronous: with FASTPM_B200_SYNC_CLOCKS=1 the stream is drained at each clock edge so that
 * the named clocks (decompose, paint, r2c, transfer, c2r, readout, kick, drift ...) measure device time */
static void clock_edge(void)
{
    if (clocks_sync < 0) { const char *e = getenv("FASTPM_B200_SYNC_CLOCKS"); clocks_sync = e ? atoi(e) : 0; }
    if (clocks_sync) fpm_sync();
}
void fastpm_clock_in(FastPMClock *clock) { clock_edge(); clock->t0 = wallclock(); }
void fastpm_clock_out(FastPMClock *clock) { clock_edge(); clock->tcum += wallclock() - clock->t0; }

void fastpm_clock_stat(MPI_Comm comm)
{
    int n = 0;
    for (FastPMClock *p = clock_head; p; p = p->next) n++;
    fastpm_info("%8s %8s %8s %16s\n", "min", "max", "mean", "name");
    for (FastPMClock *p = clock_head; p; p = p->next) {
        double v[3] = { p->tcum, p->tcum, p->tcum };
        fpm_comm_allreduce_double(comm, &v[0], 1, 1);
        fpm_comm_allreduce_double(comm, &v[1], 1, 2);
        fpm_comm_allreduce_double(comm, &v[2], 1, 0);
        fastpm_info("%8.2f %8.2f %8.2f %16s %s\n", v[0], v[1], v[2] / fpm_comm_size(comm), p->name, p->func);
    }
}
int fastpm_b200_clock_get(const char *name, double *seconds)
{
    double t = 0; int found = 0;
    for (FastPMClock *p = clock_head; p; p = p->next) if (!strcmp(p->name, name)) { t += p->tcum; found = 1; }
    *seconds = t;
    return found ? 0 : -1;
}
void fastpm_b200_clock_reset(void) { for (FastPMClock *p = clock_head; p; p = p->next) p->tcum = 0; }
