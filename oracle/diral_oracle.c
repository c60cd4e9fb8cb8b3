/*
 * diral_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see diral_oracle.h).
 *
 * Sequential, literal restatement of the reference algorithm, one env at a time, in the
 * reference's own loop order.  Every function cites the reference lines it follows
 * (paths are relative to the reference checkout).  Arithmetic is float64 with the same
 * libm entry points CPython uses (`x ** 2` on floats is libm pow(x, 2.0); `math.sqrt`,
 * `math.exp`; float `%` is fmod for non-negative operands; builtin `sum` over floats is
 * Neumaier-compensated since CPython 3.12, which is what the golden vectors were made with).
 *
 * Built by oracle/Makefile into oracle/_build/libdiral_oracle.so (gcc, no -ffast-math, no FMA
 * contraction).
 */
#include "diral_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static int g_threads = 1;

/* minimal static-partition parallel-for over envs (libgomp is not in the image) */
typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; int64_t lo, hi; void *ctx; } par_job;
static void *par_trampoline(void *p) { par_job *j = (par_job *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void par_for(int64_t n, range_fn fn, void *ctx)
{
    int nt = g_threads;
    if (nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    par_job *jobs = (par_job *)malloc(sizeof(par_job) * nt);
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / nt; jobs[t].hi = n * (t + 1) / nt;
        pthread_create(&th[t], NULL, par_trampoline, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int  orc_get_threads(void) { return g_threads; }

/* ---- per-env view -------------------------------------------------------------------- */
typedef struct {
    double *x, *y, *v;
    double *tx, *ty;
    int32_t *seq, *lu, *lat;
} env_t;

static env_t env_at(const orc_cfg *c, const orc_batch *b, int64_t e)
{
    env_t v;
    int64_t n = c->N, nn = (int64_t)c->N * c->N;
    v.x = b->pos_x + e * n;  v.y = b->pos_y + e * n;  v.v = b->vel + e * n;
    v.tx = b->tab_x + e * nn; v.ty = b->tab_y + e * nn;
    v.seq = b->tab_seq + e * nn; v.lu = b->tab_lu + e * nn;
    v.lat = b->lat ? b->lat + e * nn : NULL;
    return v;
}

/* test_env.py:49-85 -- state-space size from the State flags */
int orc_state_space(const orc_cfg *c)
{
    int s = 0;
    if (c->add_action) s += c->action_binary ? c->R : 1;
    if (c->add_channel_obs) s += c->R;
    if (c->add_reward) s += 1;
    if (c->add_index) s += 1;
    if (c->add_velocity) s += 1;
    if (c->add_position) s += 2;
    if (c->add_positional_dist) s += c->N - 1;
    if (c->fingerprint) s += 2;
    if (c->add_piggy) s += c->B;
    return s;
}

/* network.py:39-42, vehicle.py:24-33 -- zero tables, last_arrival_time = -1 */
void orc_reset(const orc_cfg *c, orc_batch *b, const double *x0, const double *y0, const double *v0)
{
    int64_t n = c->N, nn = n * n;
    memcpy(b->pos_x, x0, sizeof(double) * b->E * n);
    memcpy(b->pos_y, y0, sizeof(double) * b->E * n);
    memcpy(b->vel, v0, sizeof(double) * b->E * n);
    memset(b->tab_x, 0, sizeof(double) * b->E * nn);
    memset(b->tab_y, 0, sizeof(double) * b->E * nn);
    memset(b->tab_seq, 0, sizeof(int32_t) * b->E * nn);
    memset(b->tab_lu, 0, sizeof(int32_t) * b->E * nn);
    if (b->lat)
        for (int64_t i = 0; i < b->E * nn; ++i) b->lat[i] = -1;
}

/* network.py:318-332 -- Network.dist: sqrt((x2-x1)**2 + (y2-y1)**2) */
static double dist(const env_t *v, int p1, int p2)
{
    double dx = v->x[p2] - v->x[p1];
    double dy = v->y[p2] - v->y[p1];
    return sqrt(pow(dx, 2.0) + pow(dy, 2.0));
}

/* CPython >= 3.12 builtin sum() over floats (Python/bltinmodule.c, Neumaier variant) */
typedef struct { double f, c; int first; } pysum_t;
static void pysum_init(pysum_t *s) { s->f = 0.0; s->c = 0.0; s->first = 1; }
static void pysum_add(pysum_t *s, double x)
{
    if (s->first) { s->f = 0.0 + x; s->first = 0; return; } /* int 0 + float */
    double t = s->f + x;
    if (fabs(s->f) >= fabs(x)) s->c += (s->f - t) + x;
    else                       s->c += (x - t) + s->f;
    s->f = t;
}
static double pysum_result(const pysum_t *s)
{
    double r = s->f;
    if (s->c != 0.0 && isfinite(s->c)) r += s->c;
    return r;
}

/* network.py:307-316 -- mean pairwise distance over itertools.combinations(users, 2) */
static double calculate_avg_distance(const env_t *v, const int *users, int cnt)
{
    pysum_t s; pysum_init(&s);
    int pairs = 0;
    for (int i = 0; i < cnt; ++i)
        for (int j = i + 1; j < cnt; ++j) { pysum_add(&s, dist(v, users[i], users[j])); ++pairs; }
    return pysum_result(&s) / (double)pairs;
}

/* network.py:225-246 -- distance between first-min-x and first-max-x vehicle */
static double calculate_norm(const orc_cfg *c, const env_t *v)
{
    double x_min = c->L + 1, x_max = -c->L - 1;
    int imin = -1, imax = -1;
    for (int u = 0; u < c->N; ++u) {
        if (v->x[u] < x_min) { x_min = v->x[u]; imin = u; }
        if (v->x[u] > x_max) { x_max = v->x[u]; imax = u; }
    }
    return dist(v, imin, imax);
}

/* network.py:273-300 -- w in {0,1} */
static int calculate_reward_weights(const orc_cfg *c, const env_t *v, const int *tx, int cnt)
{
    double m = calculate_avg_distance(v, tx, cnt);
    if (c->toy) return m == calculate_norm(c, v);
    return m > c->C;
}

/* network.py:378-398 -- nearest in-range transmitter; resets last_arrival_time out of range */
static double find_closest_tx(const orc_cfg *c, env_t *v, const int *tx, int cnt, int rx, int *tx_id)
{
    double min_dist = c->sentinel;
    int best = -1;
    for (int k = 0; k < cnt; ++k) {
        double d = dist(v, tx[k], rx);
        if (d < c->C) {
            if (d < min_dist) { min_dist = d; best = tx[k]; }
        } else if (v->lat) {
            v->lat[(int64_t)tx[k] * c->N + rx] = -1;
        }
    }
    *tx_id = best;
    return min_dist;
}

/* vehicle.py:56-70 via network.py:587-593 -- every vehicle ticks its own table */
static void periodic_update(const orc_cfg *c, env_t *v)
{
    int n = c->N;
    for (int i = 0; i < n; ++i) {
        int32_t *seq = v->seq + (int64_t)i * n, *lu = v->lu + (int64_t)i * n;
        seq[i] += 1;
        v->tx[(int64_t)i * n + i] = v->x[i];
        v->ty[(int64_t)i * n + i] = v->y[i];
        for (int j = 0; j < n; ++j) {
            if (j == i) lu[j] = 0; else lu[j] += 1;
        }
    }
}

/* vehicle.py:35-47 via network.py:576-585 -- merge the transmitter's LIVE row (vehicle.py:61
 * aliases rather than copies) into the receiver's row by strictly greater sequence number */
static void received_update(const orc_cfg *c, env_t *v, int rx, int tx)
{
    int n = c->N;
    int64_t r = (int64_t)rx * n, t = (int64_t)tx * n;
    for (int j = 0; j < n; ++j) {
        if (v->seq[t + j] > v->seq[r + j]) {
            v->tx[r + j] = v->tx[t + j];
            v->ty[r + j] = v->ty[t + j];
            v->seq[r + j] = v->seq[t + j];
            v->lu[r + j] = 0;
        }
    }
}

/* network.py:189-206,302-305 -- mobility or trace replay */
static void update_mobility(const orc_cfg *c, const orc_batch *b, env_t *v, int64_t timestep)
{
    if (!c->mobility) return;
    if (b->trace) {
        int64_t t = timestep % b->trace_len;
        if (t < 0) t += b->trace_len;      /* python % */
        for (int u = 0; u < c->N; ++u) v->x[u] = b->trace[t * c->N + u];
    } else {
        for (int u = 0; u < c->N; ++u) {
            double a = (v->x[u] + v->v[u]) + c->L;
            double m = fmod(a, c->L);
            if (m != 0.0 && ((c->L < 0) != (m < 0))) m += c->L;  /* python float % */
            v->x[u] = m;
        }
    }
}

/* test_env.py:319-349 (+ network.py:122-157) -- reward for one collided tx under my_step_design */
static double calculate_reward_design(const orc_cfg *c, const env_t *v, int tx_user, const int *tx, int cnt,
                                      int *scratch)
{
    int k = 0;
    scratch[k++] = tx_user;
    for (int i = 0; i < cnt; ++i) {
        if (tx[i] == tx_user) continue;
        if (dist(v, tx_user, tx[i]) < 2 * c->C) scratch[k++] = tx[i];
    }
    if (k == 1) return 1.0;
    if (k == 2) {
        double m = calculate_avg_distance(v, scratch, k);
        int w = m > c->C * 2;
        return w == 1 ? 0.0 : -(double)k;
    }
    return -(double)k;
}

static void step_env(const orc_cfg *c, const orc_batch *b, env_t *v, int mode, const int32_t *a,
                     int64_t timestep, double *obs, double *rews, int64_t *counts, int *tx, int *scratch,
                     double *prr)
{
    const int n = c->N, R = c->R;
    int64_t n_recv = 0, n_pairs = 0;
    for (int u = 0; u < n; ++u) rews[u] = 0.0;
    for (int i = 0; i < n * R; ++i) obs[i] = 0.0;
    if (c->add_piggy) periodic_update(c, v);               /* test_env.py:138-139 */

    for (int i = 0; i < R; ++i) {                          /* test_env.py:147 */
        int tot = 0;
        for (int u = 0; u < n; ++u)
            if (a[u] == i) tx[tot++] = u;                  /* :153-157 */

        /* statistics for the episode metrics (not part of the reference's return values) */
        if (counts)
            for (int k = 0; k < tot; ++k)
                for (int u = 0; u < n; ++u)
                    if (a[u] != i && dist(v, tx[k], u) < c->C) ++n_pairs;

        double reward = tot == 1 ? 1.0 : 0.0, rewards = 0.0;
        if (mode == ORC_MY_STEP && tot > 1) {              /* :163-199 */
            switch (c->reward_design) {
            case 1: {
                int w = calculate_reward_weights(c, v, tx, tot);
                double Rr = (double)w / (double)tot;
                rewards = -1 * (1 - Rr);
                break; }
            case 2:
                if (tot == 2) rewards = 2 * calculate_reward_weights(c, v, tx, tot) - (double)tot;
                else          rewards = 0 - (double)tot;
                break;
            case 3: rewards = -1 * exp(1 - 1 / (double)tot); break;
            case 4: rewards = 1 / (double)tot; break;
            case 5:
                if (tot == 2) rewards = calculate_reward_weights(c, v, tx, tot) == 1 ? 0.0 : -1.0;
                else          rewards = -1.0;
                break;
            default: break;
            }
        }
        if (mode == ORC_MY_STEP_CH && tot > 1) {           /* :384-405 */
            for (int k = 0; k < tot; ++k) {
                int t = tx[k], received = 0, in_range = 0;
                for (int rx = 0; rx < n; ++rx) {
                    if (a[rx] == i) continue;              /* half duplex */
                    if (!(dist(v, t, rx) < c->C)) continue;  /* network.py:595-607 */
                    ++in_range;
                    int nearest; find_closest_tx(c, v, tx, tot, rx, &nearest);
                    if (nearest == t) ++received;
                }
                /* :402-405 -- the no-receiver branch assigns the *int* 1, so design 2's
                 * -1*(1-R) yields integer 0 (+0.0) there but -0.0 when received/in_range == 1.0;
                 * encode the int case as a negative sentinel and resolve it below */
                prr[t] = in_range > 0 ? (double)received / (double)in_range : -1.0;
            }
        }

        for (int u = 0; u < n; ++u) {                      /* :203-254 / :294-311 / :408-439 */
            if (a[u] == i) {
                obs[u * R + i] = 0.0;
                if (mode == ORC_MY_STEP) {
                    rews[u] = tot > 1 ? rewards : reward;
                } else if (mode == ORC_MY_STEP_DESIGN) {
                    rews[u] = tot == 1 ? 1.0 : calculate_reward_design(c, v, u, tx, tot, scratch);
                } else {
                    if (tot > 1) {
                        int int_one = prr[u] < 0.0;
                        double Rr = int_one ? 1.0 : prr[u];
                        if (c->reward_design == 3)      rews[u] = 1 - exp(1 - Rr);
                        else if (c->reward_design == 4) rews[u] = -1 * exp(1 - Rr);
                        else if (c->reward_design == 2) rews[u] = int_one ? 0.0 : -1 * (1 - Rr);
                    } else {
                        if (c->reward_design == 3)      rews[u] = 1;
                        else if (c->reward_design == 4) rews[u] = exp(1);
                        else if (c->reward_design == 2) rews[u] = 1;
                    }
                }
            } else if (tot > 0) {
                int tx_id;
                double d = find_closest_tx(c, v, tx, tot, u, &tx_id);
                if (tx_id >= 0) ++n_recv;
                if (mode == ORC_MY_STEP) {
                    if (c->state_type == 1) {              /* :226-232 (reference raises on None) */
                        obs[u * R + i] = 1.0;
                        if (c->add_piggy && tx_id >= 0) received_update(c, v, u, tx_id);
                    } else if (c->state_type == 2) {       /* :234-240 */
                        if (c->add_piggy && tx_id >= 0) received_update(c, v, u, tx_id);
                        obs[u * R + i] = d;
                    }
                } else if (mode == ORC_MY_STEP_DESIGN) {   /* :305-311 */
                    obs[u * R + i] = 1.0;
                    if (c->add_piggy && tx_id >= 0) received_update(c, v, u, tx_id);
                } else {                                   /* :431-439 */
                    obs[u * R + i] = 1.0;
                    if (tx_id >= 0) {
                        if (v->lat) v->lat[(int64_t)tx_id * n + u] = (int32_t)timestep;
                        if (c->add_piggy) received_update(c, v, u, tx_id);
                    }
                }
            }
        }
    }
    update_mobility(c, b, v, timestep);                    /* :259 / :314 / :441 */
    if (counts) { counts[0] = n_recv; counts[1] = n_pairs; }
}

typedef struct {
    const orc_cfg *c; orc_batch *b; int mode; const int32_t *actions; int64_t timestep;
    double *obs, *rews; int64_t *counts;
} step_ctx;

static void step_range(int64_t lo, int64_t hi, void *p)
{
    step_ctx *s = (step_ctx *)p;
    const orc_cfg *c = s->c;
    const int n = c->N, R = c->R;
    int *tx = (int *)malloc(sizeof(int) * n);
    int *scratch = (int *)malloc(sizeof(int) * n);
    double *prr = (double *)malloc(sizeof(double) * n);
    for (int64_t e = lo; e < hi; ++e) {
        env_t v = env_at(c, s->b, e);
        step_env(c, s->b, &v, s->mode, s->actions + e * n, s->timestep, s->obs + e * n * R, s->rews + e * n,
                 s->counts ? s->counts + e * 2 : NULL, tx, scratch, prr);
    }
    free(tx); free(scratch); free(prr);
}

void orc_step(const orc_cfg *c, orc_batch *b, int mode, const int32_t *actions, int64_t timestep,
              double *obs, double *rews, int64_t *counts)
{
    step_ctx s = { c, b, mode, actions, timestep, obs, rews, counts };
    par_for(b->E, step_range, &s);
}

/* ---- observation build ----------------------------------------------------------------- */

/* network.py:538-558 -- Network.dist_piggy(rx_id=j, tx_id=i) on observer i's table row */
static int dist_piggy(const orc_cfg *c, const env_t *v, int j, int i, double *d, int *sign)
{
    int64_t k = (int64_t)i * c->N + j;
    if (!(c->mobility || c->design_topology)) return -1;   /* reference returns None */
    if (v->lu[k] < c->age_threshold) {
        double x1 = v->tx[k], y1 = v->ty[k];
        double x2 = v->x[i], y2 = v->y[i];
        *d = sqrt(pow(x2 - x1, 2.0) + pow(y2 - y1, 2.0));
        *sign = (x1 - x2 > 0.0) ? 1 : -1;
        return 1;
    }
    return 0;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* numpy.linspace(start, stop, num) as numpy/_core/function_base.py computes it */
static void np_linspace(double start, double stop, int num, double *y)
{
    int div = num - 1;
    double delta = stop - start;
    double step = delta / div;
    for (int k = 0; k < num; ++k) {
        double t = (double)k;
        if (step == 0) { t = t / div; t = t * delta; } else { t = t * step; }
        y[k] = t + start;
    }
    if (num > 1) y[num - 1] = stop;
}

/* network.py:473-513 -- VPD type 2: np.histogram(sorted(s), B, range=(-W, W))[0] / len(s)
 * with NumPy's equal-width fast path (numpy/lib/_histograms_impl.py, "Fast algorithm for
 * equal bins"): k = int((s-first)/(last-first)*B); k==B -> B-1; s<edge[k] -> k-1;
 * s>=edge[k+1] and k!=B-1 -> k+1. */
static void vpd_type2(const orc_cfg *c, const env_t *v, int i, const double *edges, double *s, double *out)
{
    int B = c->B, m = 0;
    for (int j = 0; j < c->N; ++j) {
        if (j == i) continue;
        double d; int sg;
        if (dist_piggy(c, v, j, i, &d, &sg) == 1 && d < c->W) s[m++] = d * sg;
    }
    for (int k = 0; k < B; ++k) out[k] = 0.0;
    if (m == 0) return;
    double first = -c->W, last = c->W, denom = last - first;
    for (int q = 0; q < m; ++q) {
        double a = s[q];
        if (!(a >= first && a <= last)) continue;
        double f = ((a - first) / denom) * (double)B;
        long k = (long)f;
        if (k == B) k -= 1;
        if (a < edges[k]) k -= 1;
        if (a >= edges[k + 1] && k != B - 1) k += 1;
        out[k] += 1.0;
    }
    for (int k = 0; k < B; ++k) out[k] = out[k] / (double)m;
}

/* network.py:432-471 -- VPD type 1: np.histogram(s/max|s|, linspace(-1,1,B+1), weights=same)
 * through NumPy's cumulative path (explicit edges): sort, sequential cumsum of the weights,
 * searchsorted(left) for every edge but the last (right), difference of the cumulative sums. */
static void vpd_type1(const orc_cfg *c, const env_t *v, int i, const double *edges1, double *s, double *cw,
                      double *out)
{
    int B = c->B, m = 0;
    double norm = 0.0;
    for (int j = 0; j < c->N; ++j) {
        if (j == i) continue;
        double d; int sg;
        if (dist_piggy(c, v, j, i, &d, &sg) == 1) { s[m++] = d * sg; }
    }
    for (int k = 0; k < B; ++k) out[k] = 0.0;
    if (m == 0) return;
    qsort(s, m, sizeof(double), cmp_double);
    for (int q = 0; q < m; ++q) if (fabs(s[q]) > norm) norm = fabs(s[q]);   /* LA.norm(., inf) */
    for (int q = 0; q < m; ++q) s[q] = s[q] / norm;
    cw[0] = 0.0;
    for (int q = 0; q < m; ++q) cw[q + 1] = cw[q] + s[q];
    double prev = 0.0;
    for (int k = 0; k <= B; ++k) {
        int idx = 0;
        if (k < B) { while (idx < m && s[idx] < edges1[k]) ++idx; }     /* side='left'  */
        else       { while (idx < m && s[idx] <= edges1[k]) ++idx; }    /* side='right' */
        double cum = cw[idx];
        if (k > 0) out[k - 1] = cum - prev;
        prev = cum;
    }
}

/* network.py:409-430 -- true signed distances to everyone else, sorted, over the max */
static void positional_dist(const orc_cfg *c, const env_t *v, int i, double *out)
{
    int m = 0; double max_dist = 0.0;
    for (int u = 0; u < c->N; ++u) {
        if (u == i) continue;
        double d = dist(v, u, i);
        if (d > max_dist) max_dist = d;
        int sg = (v->x[u] - v->x[i] > 0.0) ? 1 : -1;     /* dist_sign(user, tx_user), :334-349 */
        out[m++] = d * sg;
    }
    qsort(out, m, sizeof(double), cmp_double);
    for (int q = 0; q < m; ++q) out[q] = out[q] / max_dist;
}

/* test_env.py:527-583 */
typedef struct {
    const orc_cfg *c; const orc_batch *b; const double *obs; const int32_t *acts; const double *rews;
    double episode, epsilon; double *out; const double *edges, *edges1;
} state_ctx;

static void state_range(int64_t lo, int64_t hi, void *p)
{
    state_ctx *q = (state_ctx *)p;
    const orc_cfg *c = q->c;
    const int n = c->N, R = c->R, B = c->B, S = orc_state_space(c);
    double *s = (double *)malloc(sizeof(double) * (n + 1));
    double *cw = (double *)malloc(sizeof(double) * (n + 2));
    for (int64_t e = lo; e < hi; ++e) {
        env_t v = env_at(c, q->b, e);
        for (int u = 0; u < n; ++u) {
            double *o = q->out + (e * n + u) * S;
            int k = 0;
            if (c->add_action) {
                if (c->action_binary) { for (int r = 0; r < R; ++r) o[k++] = (q->acts[e * n + u] == r); }
                else o[k++] = (double)q->acts[e * n + u];
            }
            if (c->add_channel_obs) for (int r = 0; r < R; ++r) o[k++] = q->obs[(e * n + u) * R + r];
            if (c->add_positional_dist) { positional_dist(c, &v, u, o + k); k += n - 1; }
            if (c->add_piggy) {
                if (c->pos_dist_type == 1) vpd_type1(c, &v, u, q->edges1, s, cw, o + k);
                else                       vpd_type2(c, &v, u, q->edges, s, o + k);
                k += B;
            }
            if (c->add_reward) o[k++] = q->rews[e * n + u];
            if (c->add_index) o[k++] = (double)(u + 1);
            if (c->add_position) { o[k++] = v.x[u] / c->L; o[k++] = v.y[u] / 2.0; } /* network.py:403-407 */
            if (c->add_velocity) o[k++] = v.v[u];
            if (c->fingerprint) { o[k++] = q->episode; o[k++] = q->epsilon; }
        }
    }
    free(s); free(cw);
}

void orc_obtain_state(const orc_cfg *c, const orc_batch *b, const double *obs, const int32_t *acts,
                      const double *rews, double episode, double epsilon, double *out)
{
    const int B = c->B;
    double *edges = (double *)malloc(sizeof(double) * (B + 1));
    double *edges1 = (double *)malloc(sizeof(double) * (B + 1));
    np_linspace(-c->W, c->W, B + 1, edges);
    np_linspace(-1.0, 1.0, B + 1, edges1);
    state_ctx q = { c, b, obs, acts, rews, episode, epsilon, out, edges, edges1 };
    par_for(b->E, state_range, &q);
    free(edges); free(edges1);
}

/* network.py:560-574 */
void orc_information_age(const orc_cfg *c, const orc_batch *b, int64_t timestep, int32_t *out)
{
    const int n = c->N;
    for (int64_t e = 0; e < b->E; ++e) {
        env_t v = env_at(c, b, e);
        int32_t *h = out + e * 100;
        for (int k = 0; k < 100; ++k) h[k] = 0;
        for (int t = 0; t < n; ++t)
            for (int r = 0; r < n; ++r) {
                if (t == r) continue;
                int32_t at = v.lat[(int64_t)t * n + r];
                if (at != -1) {
                    int64_t ia = timestep - at;
                    if (ia < 100) {
                        if (ia < 0) ia += 100;   /* python negative index */
                        if (ia >= 0) h[ia] += 1;
                    }
                }
            }
    }
}

/* network.py:208-222 */
void orc_update_velocity(const orc_cfg *c, orc_batch *b, const int8_t *draws)
{
    if (!c->mobility_vary) return;                          /* test_env.py:498-504 */
    for (int64_t i = 0; i < b->E * c->N; ++i) {
        if (draws[i] == 1) { b->vel[i] += 0.55; if (b->vel[i] > 2.77) b->vel[i] = 2.77; }
        else if (draws[i] == 2) { b->vel[i] -= 0.55; if (b->vel[i] < 1.1) b->vel[i] = 1.1; }
    }
}

/* ---- Philox4x32-10 (Salmon et al., SC'11), published constants ------------------------------ */
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { STREAM_ACTIONS = 1, STREAM_TOPOLOGY = 2, STREAM_VELOCITY = 3 };

/* TestEnv.sample (test_env.py:116-122) restated on the counter-based generator */
void orc_philox_actions(uint64_t seed, int64_t env0, int64_t E, int32_t N, int32_t R, int64_t t, int32_t *out)
{
    for (int64_t e = 0; e < E; ++e)
        for (int u = 0; u < N; ++u) {
            uint32_t o[4];
            orc_philox((uint32_t)u, (uint32_t)(env0 + e), (uint32_t)t, (uint32_t)((uint64_t)t >> 32),
                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ STREAM_ACTIONS, o);
            out[e * N + u] = (int32_t)(((uint64_t)o[0] * (uint32_t)R) >> 32);
        }
}

/* Network.initialize_mobility_topology (network.py:92-112): x integer-valued uniform on [0,L),
 * y = 0, v ~ U(1.1, 2.7) as a + (b-a)*random() (or 1.7 under mobility_vary) */
void orc_philox_topology(const orc_cfg *c, uint64_t seed, int64_t env0, int64_t E, double *x0, double *y0, double *v0)
{
    for (int64_t e = 0; e < E; ++e)
        for (int u = 0; u < c->N; ++u) {
            uint32_t o[4];
            orc_philox((uint32_t)u, (uint32_t)(env0 + e), 0, 0,
                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ STREAM_TOPOLOGY, o);
            uint32_t Li = (uint32_t)c->L;
            x0[e * c->N + u] = (double)(uint32_t)(((uint64_t)o[0] * Li) >> 32);
            y0[e * c->N + u] = 0.0;
            double r53 = ((double)(o[1] >> 5) * 67108864.0 + (double)(o[2] >> 6)) / 9007199254740992.0;
            v0[e * c->N + u] = c->mobility_vary ? 1.7 : 1.1 + (2.7 - 1.1) * r53;
        }
}

/* random.randrange(1, 4) per vehicle (network.py:214) */
void orc_philox_draws(uint64_t seed, int64_t env0, int64_t E, int32_t N, int64_t episode, int8_t *out)
{
    for (int64_t e = 0; e < E; ++e)
        for (int u = 0; u < N; ++u) {
            uint32_t o[4];
            orc_philox((uint32_t)u, (uint32_t)(env0 + e), (uint32_t)episode, (uint32_t)((uint64_t)episode >> 32),
                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ STREAM_VELOCITY, o);
            out[e * N + u] = (int8_t)(1 + (((uint64_t)o[0] * 3u) >> 32));
        }
}
