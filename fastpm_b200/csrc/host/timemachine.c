/* fastpm_b200 host layer -- the kick-drift-kick state table (reference: libfastpm/timemachine.c).
 * A state is the triple of time indices (force, x, v); indices advance by cycle_len per step, odd
 * indices are half steps at the geometric mean of the neighbouring step times (timemachine.c:69-87). */
#include "internal.h"

FastPMStates *fastpm_tevo_generate_states(FastPMStates *states, int cycles, FastPMState *templ, double *ts)
{
    int len = 0;
    while (templ[len].force != -1) len++;
    const int nrows = len * cycles;
    FastPMState *tab = malloc(sizeof(FastPMState) * (nrows + 3));
    tab[0] = (FastPMState) { -2, 0, 0 };            /* initial condition */
    tab[1] = (FastPMState) { 0, 0, 0 };             /* first force */
    for (int c = 0; c < cycles; c++) {
        const FastPMState origin = tab[c * len + 1];
        for (int j = 0; j < len; j++) {
            FastPMState *row = &tab[c * len + j + 2];
            row->force = origin.force + templ[j].force;
            row->x = origin.x + templ[j].x;
            row->v = origin.v + templ[j].v;
        }
    }
    tab[nrows + 2] = (FastPMState) { -1, -1, -1 };
    states->table = tab;
    states->cycle_len = templ[len - 1].force;
    states->cycles = cycles;
    states->timesteps = malloc(sizeof(double) * (cycles + 1));
    memcpy(states->timesteps, ts, sizeof(double) * (cycles + 1));
    return states;
}

void fastpm_tevo_destroy_states(FastPMStates *states) { free(states->table); free(states->timesteps); }

static double index_to_time(FastPMStates *st, int i)
{
    const int step = i / st->cycle_len;
    const double frac = (i - st->cycle_len * step) / (1.0 * st->cycle_len);
    if (step >= st->cycles) return st->timesteps[st->cycles];
    if (step < 0) return st->timesteps[0];
    if (frac == 0.0) return st->timesteps[step];       /* exact table value keeps == comparisons valid */
    return exp((1 - frac) * log(st->timesteps[step]) + frac * log(st->timesteps[step + 1]));
}

void fastpm_tevo_transition_init(FastPMTransition *tr, FastPMStates *states, int istart, int iend)
{
    FastPMState *s = &states->table[istart], *e = &states->table[iend];
    tr->states = states; tr->istart = istart; tr->iend = iend; tr->start = s; tr->end = e;
    int from = 0, to = 0, ref = 0;
    if (s->force != e->force) {
        if (s->x != e->x) fastpm_raise(-1, "A force action must have identical x stamp\n");
        tr->action = FASTPM_ACTION_FORCE; from = s->force; to = e->force; ref = e->x;
    }
    if (s->v != e->v) {
        if (s->force != e->force) fastpm_raise(-1, "A kick action must have identical a stamp\n");
        tr->action = FASTPM_ACTION_KICK; from = s->v; to = e->v; ref = e->force;
    }
    if (s->x != e->x) {
        if (s->v != e->v) fastpm_raise(-1, "A drift action must have identical v stamp\n");
        tr->action = FASTPM_ACTION_DRIFT; from = s->x; to = e->x; ref = e->v;
    }
    tr->i.i = from; tr->i.f = to; tr->i.r = ref;
    tr->a.i = index_to_time(states, from); tr->a.f = index_to_time(states, to); tr->a.r = index_to_time(states, ref);
}

int fastpm_tevo_transition_find_dual(FastPMTransition *tr, FastPMTransition *dual)
{
    if (tr->end->x != tr->end->v) fastpm_raise(-1, "Only transitions towards a synced x and v has a dual.\n");
    enum FastPMAction want;
    if (tr->action == FASTPM_ACTION_DRIFT) want = FASTPM_ACTION_KICK;
    else if (tr->action == FASTPM_ACTION_KICK) want = FASTPM_ACTION_DRIFT;
    else { fastpm_raise(-1, "Only Kick and Drift has dual transitions\n"); return 0; }
    int i;
    for (i = tr->istart; i >= 1; i--) {
        fastpm_tevo_transition_init(dual, tr->states, i - 1, i);
        if (dual->action == want) break;
    }
    if (i < 1) return 0;
    fastpm_tevo_transition_init(dual, tr->states, i, i - 1);       /* reversed: the reference lies in the future */
    if (dual->a.r != tr->a.i) fastpm_raise(-1, "dual transition reference is not the same as my initial state.\n");
    return 1;
}

int fastpm_tevo_transition_find_next(FastPMTransition *tr, FastPMTransition *next)
{
    FastPMStates *st = tr->states;
    for (int i = tr->iend; st->table[i + 1].force != -1; i++) {
        fastpm_tevo_transition_init(next, st, i, i + 1);
        if (next->action == tr->action) return 1;
    }
    return 0;
}
