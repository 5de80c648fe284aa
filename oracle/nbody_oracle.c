/*
 * nbody_oracle.c — CPU restatement of the nbodygo per-cycle compute path.
 *
 * TEST INFRASTRUCTURE ONLY (see nbody_oracle.h).  PARITY UNPINNED: pinned
 * against hand-derived KATs and an independent Python restatement only.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 */
#define _GNU_SOURCE
#include "nbody_oracle.h"
#include "gomath.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* cmd/body/body.go:19 */
static const double G = 6.673e-11;
/* cmd/body/body.go:23 */
static const double MAX_FRAGS = 2000;

static inline int exists(const orc_bodies *bc, int64_t i) { return bc->flags[i] & ORC_F_EXISTS; }
static inline int fragmenting(const orc_bodies *bc, int64_t i) { return bc->flags[i] & ORC_F_FRAGMENTING; }
static inline int collided(const orc_bodies *bc, int64_t i) { return bc->flags[i] & ORC_F_COLLIDED; }
static inline int elastic_or_fragment(uint8_t b) { return b == ORC_ELASTIC || b == ORC_FRAGMENT; }

/* ---- event sink ------------------------------------------------------- */
typedef struct {
    orc_event *ev;
    int64_t cap, n;
    int grow; /* realloc on demand (pool slices) */
} sink;

static void sink_push(sink *s, int32_t kind, int64_t a, int64_t b, double dist)
{
    if (s->n >= s->cap) {
        if (s->grow) {
            int64_t ncap = s->cap ? s->cap * 2 : 256;
            orc_event *p = (orc_event *)realloc(s->ev, (size_t)ncap * sizeof(orc_event));
            if (!p) { s->n++; return; }
            s->ev = p;
            s->cap = ncap;
        } else {
            s->n++;
            return;
        }
    }
    orc_event *e = &s->ev[s->n++];
    e->kind = kind;
    e->a = (int32_t)a;
    e->b = (int32_t)b;
    e->dist = dist;
    e->f1 = e->f2 = 0;
}

/* ---- Body.calcForceFrom, cmd/body/body.go:214-225 ---------------------- */
static inline void calc_force_from(const orc_bodies *bc, int64_t i, int64_t j,
                                   double *fx, double *fy, double *fz)
{
    double dx = bc->x[j] - bc->x[i];
    double dy = bc->y[j] - bc->y[i];
    double dz = bc->z[j] - bc->z[i];
    double dist = sqrt(dx * dx + dy * dy + dz * dz);
    if (collided(bc, i) || dist > bc->radius[i] + bc->radius[j]) {
        double force = G * bc->mass[i] * bc->mass[j] / (dist * dist);
        *fx += force * dx / dist;
        *fy += force * dy / dist;
        *fz += force * dz / dist;
    }
}

/* ---- Body.Collided, cmd/body/body.go:192-208 --------------------------- */
static inline int body_collided(const orc_bodies *bc, int64_t i, int64_t j, double *dist_out)
{
    if (collided(bc, i)) return 0;
    double dx = bc->x[j] - bc->x[i];
    double dy = bc->y[j] - bc->y[i];
    double dz = bc->z[j] - bc->z[i];
    double dist = sqrt(dx * dx + dy * dy + dz * dz);
    double s = bc->radius[i] + bc->radius[j];
    if (dist > s) return 0;
    if (dist <= s) { *dist_out = dist; return 1; }
    return 0; /* NaN */
}

/* collision sweep body for one (i,j): cmd/body/body.go:172-186 */
static inline void collision_pair(const orc_bodies *bc, int64_t i, int64_t j, uint32_t opts, sink *s)
{
    double dist;
    if (!body_collided(bc, i, j, &dist)) return;
    if (i == j && !(opts & ORC_OPT_SELF_PAIRS)) return;
    const int dead_j = !exists(bc, j);
    if (dead_j && !(opts & (ORC_OPT_DEAD_J | ORC_OPT_DEAD_J_SUBSUME))) return;
    uint8_t bi = bc->behavior[i], bj = bc->behavior[j];
    if (elastic_or_fragment(bi) && elastic_or_fragment(bj)) {
        if (dead_j && !(opts & ORC_OPT_DEAD_J)) return; /* a no-op in ResolveCollision */
        sink_push(s, ORC_EV_COLLISION, i, j, dist);
    } else if (bi == ORC_SUBSUME || bj == ORC_SUBSUME) {
        if (bc->radius[i] > bc->radius[j] && dist <= bc->radius[i])
            sink_push(s, ORC_EV_SUBSUME, i, j, dist);
        else if (bc->radius[j] > bc->radius[i] && dist <= bc->radius[j])
            sink_push(s, ORC_EV_SUBSUME, j, i, dist);
    }
}

/* ---- Body.Compute, cmd/body/body.go:148-187 ---------------------------- */
static void compute_one(const orc_bodies *bc, int64_t i, uint32_t opts, sink *s)
{
    if (!exists(bc, i)) return;
    if (fragmenting(bc, i)) return; /* b.fragment(bc): host-side, out of scope */
    const int64_t n = bc->n;
    double fx = 0, fy = 0, fz = 0;
    if (opts & ORC_OPT_SINGLE_SWEEP) {
        for (int64_t j = 0; j < n; j++) {
            if (exists(bc, j) && j != i && !fragmenting(bc, j)) calc_force_from(bc, i, j, &fx, &fy, &fz);
            collision_pair(bc, i, j, opts, s);
        }
    } else {
        /* force sweep: body.go:161-169 */
        for (int64_t j = 0; j < n; j++) {
            if (!exists(bc, j)) continue;
            if (j != i && !fragmenting(bc, j)) calc_force_from(bc, i, j, &fx, &fy, &fz);
        }
        /* collision sweep: body.go:172-186 (no self / exists filter) */
        for (int64_t j = 0; j < n; j++) collision_pair(bc, i, j, opts, s);
    }
    bc->fx[i] = fx;
    bc->fy[i] = fy;
    bc->fz[i] = fz;
}

int orc_compute(const orc_bodies *bc, int64_t i0, int64_t i1, uint32_t opts,
                orc_event *ev, int64_t ev_cap, int64_t *n_ev)
{
    sink s = { ev, ev_cap, n_ev ? *n_ev : 0, 0 };
    if (i0 < 0) i0 = 0;
    if (i1 > bc->n) i1 = bc->n;
    for (int64_t i = i0; i < i1; i++) compute_one(bc, i, opts, &s);
    if (n_ev) *n_ev = s.n;
    return s.n > ev_cap ? -1 : 0;
}

/* ---- exact adjudicator -------------------------------------------------- */
int orc_compute_exact(const orc_bodies *bc, int64_t i0, int64_t i1,
                      double *fx, double *fy, double *fz, double *fnorm)
{
    const int64_t n = bc->n;
    if (i0 < 0) i0 = 0;
    if (i1 > n) i1 = n;
    for (int64_t i = i0; i < i1; i++) {
        __float128 ax = 0, ay = 0, az = 0, an = 0;
        if (exists(bc, i) && !fragmenting(bc, i)) {
            for (int64_t j = 0; j < n; j++) {
                if (!exists(bc, j) || j == i || fragmenting(bc, j)) continue;
                double dx = bc->x[j] - bc->x[i];
                double dy = bc->y[j] - bc->y[i];
                double dz = bc->z[j] - bc->z[i];
                double dist = sqrt(dx * dx + dy * dy + dz * dz);
                if (!(collided(bc, i) || dist > bc->radius[i] + bc->radius[j])) continue;
                /* exact differences of the double inputs */
                __float128 qx = (__float128)bc->x[j] - (__float128)bc->x[i];
                __float128 qy = (__float128)bc->y[j] - (__float128)bc->y[i];
                __float128 qz = (__float128)bc->z[j] - (__float128)bc->z[i];
                __float128 d2 = qx * qx + qy * qy + qz * qz;
                /* d^-3 via double seed + two Newton steps in quad (no libquadmath needed) */
                __float128 y = 1.0 / sqrt((double)d2);
                y = y * ((__float128)1.5 - (__float128)0.5 * d2 * y * y);
                y = y * ((__float128)1.5 - (__float128)0.5 * d2 * y * y);
                __float128 w = (__float128)G * (__float128)bc->mass[i] * (__float128)bc->mass[j] * y * y * y;
                ax += w * qx;
                ay += w * qy;
                az += w * qz;
                __float128 t = w * d2 * y; /* |f_ij| = w * |d| */
                an += t < 0 ? -t : t;
            }
        }
        fx[i] = (double)ax;
        fy[i] = (double)ay;
        fz[i] = (double)az;
        if (fnorm) fnorm[i] = (double)an;
    }
    return 0;
}

/* ---- work pool, cmd/runner/computation-runner.go:286-311 --------------- */
typedef struct {
    const orc_bodies *bc;
    uint32_t opts;
    int worker, workers;
    int64_t base, n_total, size; /* slices cover [base, base+n_total) */
    int64_t n_slices;
    sink *slice_sinks; /* one per slice */
    int count_only;
} pool_arg;

static void *pool_worker(void *p)
{
    pool_arg *a = (pool_arg *)p;
    /* round-robin slice assignment: cmd/runner/workpool.go:212-220 */
    for (int64_t sl = a->worker; sl < a->n_slices; sl += a->workers) {
        int64_t off = a->base + sl * a->size;
        int64_t end = off + a->size;
        if (end > a->base + a->n_total) end = a->base + a->n_total;
        sink *s = &a->slice_sinks[sl];
        /* worker loop: cmd/runner/workpool.go:103-110 */
        for (int64_t i = off; i < end; i++) compute_one(a->bc, i, a->opts, s);
    }
    return NULL;
}

static int run_pool(const orc_bodies *bc, int64_t base, int64_t n_total, int workers, uint32_t opts,
                    int count_only, sink *out, int64_t *count)
{
    if (workers < 1) workers = 1;
    int64_t size = n_total / workers;
    if (n_total < 100) size = n_total; /* computation-runner.go:290-293 */
    if (size < 1) size = 1;
    int64_t n_slices = (n_total + size - 1) / size;
    if (n_total == 0) n_slices = 0;
    sink *sinks = (sink *)calloc((size_t)(n_slices ? n_slices : 1), sizeof(sink));
    pthread_t *th = (pthread_t *)calloc((size_t)workers, sizeof(pthread_t));
    pool_arg *args = (pool_arg *)calloc((size_t)workers, sizeof(pool_arg));
    if (!sinks || !th || !args) { free(sinks); free(th); free(args); return -2; }
    for (int64_t s = 0; s < n_slices; s++) sinks[s].grow = count_only ? 0 : 1;
    for (int w = 0; w < workers; w++) {
        args[w] = (pool_arg){ bc, opts, w, workers, base, n_total, size, n_slices, sinks, count_only };
        pthread_create(&th[w], NULL, pool_worker, &args[w]);
    }
    for (int w = 0; w < workers; w++) pthread_join(th[w], NULL);
    int64_t total = 0;
    for (int64_t s = 0; s < n_slices; s++) {
        if (!count_only && out) {
            for (int64_t k = 0; k < sinks[s].n; k++) {
                if (out->n < out->cap) out->ev[out->n] = sinks[s].ev[k];
                out->n++;
            }
        }
        total += sinks[s].n;
        free(sinks[s].ev);
    }
    if (count) *count = total;
    free(sinks); free(th); free(args);
    return 0;
}

int orc_compute_pool(const orc_bodies *bc, int workers, uint32_t opts,
                     orc_event *ev, int64_t ev_cap, int64_t *n_ev)
{
    sink out = { ev, ev_cap, n_ev ? *n_ev : 0, 0 };
    int rc = run_pool(bc, 0, bc->n, workers, opts, 0, &out, NULL);
    if (rc) return rc;
    if (n_ev) *n_ev = out.n;
    return out.n > ev_cap ? -1 : 0;
}

int64_t orc_compute_slice_timed(const orc_bodies *bc, int64_t i0, int64_t i1, int workers, uint32_t opts)
{
    int64_t count = 0;
    if (i0 < 0) i0 = 0;
    if (i1 > bc->n) i1 = bc->n;
    if (i1 < i0) i1 = i0;
    int rc = run_pool(bc, i0, i1 - i0, workers, opts, 1, NULL, &count);
    return rc ? rc : count;
}

/* ---- transcendental backend of calcElasticCollision --------------------- */
/* ORC_MATH_LIBM (default): glibc.  ORC_MATH_GO: the Go standard library's algorithms restated in
 * gomath.c — what the reference itself executes on amd64.  Process-wide; set before use. */
typedef struct {
    double (*acos_)(double), (*asin_)(double), (*sin_)(double), (*cos_)(double), (*tan_)(double);
    double (*atan2_)(double, double);
} math_backend;
static const math_backend MATH_LIBM = { acos, asin, sin, cos, tan, atan2 };
static const math_backend MATH_GO = { go_acos, go_asin, go_sin, go_cos, go_tan, go_atan2 };
static const math_backend *g_math = &MATH_LIBM;

int orc_set_math(int which)
{
    if (which == ORC_MATH_LIBM) g_math = &MATH_LIBM;
    else if (which == ORC_MATH_GO) g_math = &MATH_GO;
    else return -1;
    return 0;
}
int orc_get_math(void) { return g_math == &MATH_GO ? ORC_MATH_GO : ORC_MATH_LIBM; }

/* ---- calcElasticCollision, cmd/body/collisioncalc.go:42-186 ------------ */
typedef struct {
    int collided;
    double vx1, vy1, vz1, vx2, vy2, vz2, vx_cm, vy_cm, vz_cm;
} coll_result;

static coll_result calc_elastic(const orc_bodies *bc, int64_t a, int64_t b)
{
    coll_result res;
    memset(&res, 0, sizeof res);
    double m1 = bc->mass[a], m2 = bc->mass[b];
    double r1 = bc->radius[a], r2 = bc->radius[b];
    double x1 = bc->x[a], y1 = bc->y[a], z1 = bc->z[a];
    double x2 = bc->x[b], y2 = bc->y[b], z2 = bc->z[b];
    double vx1 = bc->vx[a], vy1 = bc->vy[a], vz1 = bc->vz[a];
    double vx2 = bc->vx[b], vy2 = bc->vy[b], vz2 = bc->vz[b];

    double r12 = r1 + r2;
    double m21 = m2 / m1;
    double x21 = x2 - x1, y21 = y2 - y1, z21 = z2 - z1;
    double vx21 = vx2 - vx1, vy21 = vy2 - vy1, vz21 = vz2 - vz1;

    double vx_cm = (m1 * vx1 + m2 * vx2) / (m1 + m2);
    double vy_cm = (m1 * vy1 + m2 * vy2) / (m1 + m2);
    double vz_cm = (m1 * vz1 + m2 * vz2) / (m1 + m2);

    double d = sqrt(x21 * x21 + y21 * y21 + z21 * z21);
    double v = sqrt(vx21 * vx21 + vy21 * vy21 + vz21 * vz21);

    if (v == 0) return res; /* :89-92 */

    x2 = x21; y2 = y21; z2 = z21;
    vx1 = -vx21; vy1 = -vy21; vz1 = -vz21;

    const math_backend *M = g_math;
    double theta2 = M->acos_(z2 / d);
    double phi2 = (x2 == 0 && y2 == 0) ? 0 : M->atan2_(y2, x2);
    double st = M->sin_(theta2), ct = M->cos_(theta2), sp = M->sin_(phi2), cp = M->cos_(phi2);

    double vx1r = ct * cp * vx1 + ct * sp * vy1 - st * vz1;
    double vy1r = cp * vy1 - sp * vx1;
    double vz1r = st * cp * vx1 + st * sp * vy1 + ct * vz1;
    double fvz1r = vz1r / v;
    if (fvz1r > 1) fvz1r = 1;
    else if (fvz1r < -1) fvz1r = -1;
    double thetav = M->acos_(fvz1r);
    double phiv = (vx1r == 0 && vy1r == 0) ? 0 : M->atan2_(vy1r, vx1r);

    double dr = d * M->sin_(thetav) / r12;

    if (thetav > M_PI / 2 || fabs(dr) > 1) return res; /* :137-140 */

    double alpha = M->asin_(-dr);
    double beta = phiv;
    double sbeta = M->sin_(beta), cbeta = M->cos_(beta);

    double a_ = M->tan_(thetav + alpha);
    double dvz2 = 2 * (vz1r + a_ * (cbeta * vx1r + sbeta * vy1r)) / ((1 + a_ * a_) * (1 + m21));

    double vz2r = dvz2;
    double vx2r = a_ * cbeta * dvz2;
    double vy2r = a_ * sbeta * dvz2;
    vz1r = vz1r - m21 * vz2r;
    vx1r = vx1r - m21 * vx2r;
    vy1r = vy1r - m21 * vy2r;

    res.collided = 1;
    res.vx1 = ct * cp * vx1r - sp * vy1r + st * cp * vz1r + vx2;
    res.vy1 = ct * sp * vx1r + cp * vy1r + st * sp * vz1r + vy2;
    res.vz1 = ct * vz1r - st * vx1r + vz2;
    res.vx2 = ct * cp * vx2r - sp * vy2r + st * cp * vz2r + vx2;
    res.vy2 = ct * sp * vx2r + cp * vy2r + st * sp * vz2r + vy2;
    res.vz2 = ct * vz2r - st * vx2r + vz2;
    res.vx_cm = vx_cm; res.vy_cm = vy_cm; res.vz_cm = vz_cm;
    return res;
}

void orc_calc_elastic(const orc_bodies *bc, int64_t a, int64_t b, double out[10])
{
    coll_result r = calc_elastic(bc, a, b);
    out[0] = r.collided;
    out[1] = r.vx1; out[2] = r.vy1; out[3] = r.vz1;
    out[4] = r.vx2; out[5] = r.vy2; out[6] = r.vz2;
    out[7] = r.vx_cm; out[8] = r.vy_cm; out[9] = r.vz_cm;
}

/* doElastic, cmd/body/collisioncalc.go:26-35 */
static void do_elastic(orc_bodies *bc, int64_t a, int64_t b, const coll_result *r)
{
    double br = bc->rest[a];
    bc->vx[a] = (r->vx1 - r->vx_cm) * br + r->vx_cm;
    bc->vy[a] = (r->vy1 - r->vy_cm) * br + r->vy_cm;
    bc->vz[a] = (r->vz1 - r->vz_cm) * br + r->vz_cm;
    bc->vx[b] = (r->vx2 - r->vx_cm) * br + r->vx_cm;
    bc->vy[b] = (r->vy2 - r->vy_cm) * br + r->vy_cm;
    bc->vz[b] = (r->vz2 - r->vz_cm) * br + r->vz_cm;
    bc->flags[a] |= ORC_F_COLLIDED;
    bc->flags[b] |= ORC_F_COLLIDED;
}

/* shouldFragment, cmd/body/fragcalc.go:24-49 */
static int should_fragment(const orc_bodies *bc, int64_t a, int64_t b, const coll_result *r,
                           double *this_factor, double *other_factor)
{
    *this_factor = *other_factor = 0;
    if (!(bc->behavior[a] == ORC_FRAGMENT || bc->behavior[b] == ORC_FRAGMENT)) return 0;
    double br = bc->rest[a];
    double ffa = bc->frag_factor ? bc->frag_factor[a] : 0;
    double ffb = bc->frag_factor ? bc->frag_factor[b] : 0;
    double vThis = bc->vx[a] + bc->vy[a] + bc->vz[a];
    double dvThis = fabs(bc->vx[a] - ((r->vx1 - r->vx_cm) * br + r->vx_cm)) +
                    fabs(bc->vy[a] - ((r->vy1 - r->vy_cm) * br + r->vy_cm)) +
                    fabs(bc->vz[a] - ((r->vz1 - r->vz_cm) * br + r->vz_cm));
    double thisFactor = dvThis / fabs(vThis);
    double vOther = bc->vx[b] + bc->vy[b] + bc->vz[b];
    double dvOther = fabs(bc->vx[b] - ((r->vx2 - r->vx_cm) * br + r->vx_cm)) +
                     fabs(bc->vy[b] - ((r->vy2 - r->vy_cm) * br + r->vy_cm)) +
                     fabs(bc->vz[b] - ((r->vz2 - r->vz_cm) * br + r->vz_cm));
    double otherFactor = dvOther / fabs(vOther);
    if ((bc->behavior[a] == ORC_FRAGMENT && thisFactor > ffa) ||
        (bc->behavior[b] == ORC_FRAGMENT && otherFactor > ffb)) {
        *this_factor = thisFactor;
        *other_factor = otherFactor;
        return 1;
    }
    return 0;
}

/* initiateFragmentation, cmd/body/fragcalc.go:66-83 — flag part only; the
 * fragInfo bookkeeping and fragment() spawning stay host-side (out of scope) */
static void initiate_fragmentation(orc_bodies *bc, int64_t i, double frag_factor)
{
    double ff = bc->frag_factor ? bc->frag_factor[i] : 0;
    double fs = bc->frag_step ? bc->frag_step[i] : 0;
    double fragDelta = ff > 10 ? 10 : frag_factor - ff;
    double fragments = fmin(fragDelta * fs, MAX_FRAGS);
    if (fragments <= 1) {
        bc->behavior[i] = ORC_FRAGMENT;
        return;
    }
    bc->flags[i] |= ORC_F_FRAGMENTING;
}

/* ResolveCollision, cmd/body/body.go:248-264 */
static void resolve_collision(orc_bodies *bc, int64_t a, int64_t b, sink *out)
{
    if (!exists(bc, a) || !exists(bc, b)) return;
    if (bc->behavior[a] == ORC_ELASTIC && elastic_or_fragment(bc->behavior[b])) {
        coll_result r = calc_elastic(bc, a, b);
        if (r.collided) {
            double tf, of;
            if (should_fragment(bc, a, b, &r, &tf, &of)) {
                /* doFragment, cmd/body/fragcalc.go:54-61 */
                double ffa = bc->frag_factor ? bc->frag_factor[a] : 0;
                double ffb = bc->frag_factor ? bc->frag_factor[b] : 0;
                if (bc->behavior[a] == ORC_FRAGMENT && tf > ffa) {
                    if (out) {
                        sink_push(out, ORC_EV_FRAG_INIT, a, b, bc->mass[a]);
                        if (out->n <= out->cap) { out->ev[out->n - 1].f1 = tf; out->ev[out->n - 1].f2 = 1; }
                    }
                    initiate_fragmentation(bc, a, tf);
                }
                if (bc->behavior[b] == ORC_FRAGMENT && of > ffb) {
                    if (out) {
                        sink_push(out, ORC_EV_FRAG_INIT, b, a, bc->mass[b]);
                        if (out->n <= out->cap) { out->ev[out->n - 1].f1 = of; out->ev[out->n - 1].f2 = 2; }
                    }
                    initiate_fragmentation(bc, b, of);
                }
                if (out) {
                    sink_push(out, ORC_EV_FRAGMENT, a, b, 0);
                    if (out->n <= out->cap) { out->ev[out->n - 1].f1 = tf; out->ev[out->n - 1].f2 = of; }
                }
            } else {
                do_elastic(bc, a, b, &r);
            }
        }
    }
}

/* ResolveSubsume, cmd/body/body.go:228-244; SetNotExists :93-96 */
static void resolve_subsume(orc_bodies *bc, int64_t a, int64_t b)
{
    double thisMass = bc->mass[a], otherMass = bc->mass[b];
    bc->mass[a] = thisMass + otherMass;
    bc->mass[b] = 0;
    bc->flags[b] &= (uint8_t)~ORC_F_EXISTS;
}

int orc_process_mods(orc_bodies *bc, const orc_event *ev, int64_t n_ev,
                     orc_event *out_ev, int64_t out_cap, int64_t *n_out)
{
    sink out = { out_ev, out_cap, n_out ? *n_out : 0, 0 };
    /* PushFront + Front→Next == reverse arrival order */
    for (int64_t k = n_ev - 1; k >= 0; k--) {
        const orc_event *e = &ev[k];
        if (e->kind == ORC_EV_COLLISION) resolve_collision(bc, e->a, e->b, out_ev ? &out : NULL);
        else if (e->kind == ORC_EV_SUBSUME) resolve_subsume(bc, e->a, e->b);
    }
    if (n_out) *n_out = out.n;
    return out.n > out_cap && out_ev ? -1 : 0;
}

/* ---- Body.Update, cmd/body/body.go:114-139 ------------------------------ */
int orc_update(orc_bodies *bc, int64_t i0, int64_t i1, double ts, double R,
               float *render_xyz, uint8_t *render_exists)
{
    if (i0 < 0) i0 = 0;
    if (i1 > bc->n) i1 = bc->n;
    for (int64_t i = i0; i < i1; i++) {
        if (exists(bc, i)) {
            if (!collided(bc, i)) {
                bc->vx[i] += ts * bc->fx[i] / bc->mass[i];
                bc->vy[i] += ts * bc->fy[i] / bc->mass[i];
                bc->vz[i] += ts * bc->fz[i] / bc->mass[i];
            }
            bc->x[i] += ts * bc->vx[i];
            bc->y[i] += ts * bc->vy[i];
            bc->z[i] += ts * bc->vz[i];
            bc->flags[i] &= (uint8_t)~ORC_F_COLLIDED;
            bc->rest[i] = R;
            if (isnan(bc->x[i]) || isnan(bc->y[i]) || isnan(bc->z[i]))
                bc->flags[i] &= (uint8_t)~ORC_F_EXISTS;
        }
        /* NewRenderable, cmd/body/renderable.go:22-40 */
        int ex = exists(bc, i) != 0;
        if (render_exists) render_exists[i] = (uint8_t)ex;
        if (render_xyz) {
            render_xyz[3 * i + 0] = ex ? (float)bc->x[i] : 0.0f;
            render_xyz[3 * i + 1] = ex ? (float)bc->y[i] : 0.0f;
            render_xyz[3 * i + 2] = ex ? (float)bc->z[i] : 0.0f;
        }
    }
    return 0;
}

/* ---- BodyCollection.Cycle (compaction), body_collection.go:253-272 ------ */
int64_t orc_cycle_compact(orc_bodies *bc, int64_t *map_out)
{
    int64_t j = 0;
    for (int64_t i = 0; i < bc->n; i++) {
        if (!exists(bc, i)) continue;
        if (j != i) {
            bc->x[j] = bc->x[i]; bc->y[j] = bc->y[i]; bc->z[j] = bc->z[i];
            bc->vx[j] = bc->vx[i]; bc->vy[j] = bc->vy[i]; bc->vz[j] = bc->vz[i];
            bc->mass[j] = bc->mass[i]; bc->radius[j] = bc->radius[i];
            bc->rest[j] = bc->rest[i];
            if (bc->frag_factor) bc->frag_factor[j] = bc->frag_factor[i];
            if (bc->frag_step) bc->frag_step[j] = bc->frag_step[i];
            bc->fx[j] = bc->fx[i]; bc->fy[j] = bc->fy[i]; bc->fz[j] = bc->fz[i];
            bc->behavior[j] = bc->behavior[i];
            bc->flags[j] = bc->flags[i];
        }
        if (map_out) map_out[j] = i;
        j++;
    }
    bc->n = j;
    return j;
}

/* ---- one cycle, cmd/runner/computation-runner.go:297-320 ---------------- */
int orc_step(orc_bodies *bc, double ts, double R, uint32_t opts,
             orc_event *ev, int64_t ev_cap, int64_t *n_ev,
             orc_event *out_ev, int64_t out_cap, int64_t *n_out,
             float *render_xyz, uint8_t *render_exists)
{
    int64_t n = 0;
    int rc = orc_compute(bc, 0, bc->n, opts, ev, ev_cap, &n);
    if (n_ev) *n_ev = n;
    if (rc) return rc;
    rc = orc_process_mods(bc, ev, n, out_ev, out_cap, n_out);
    if (rc) return rc;
    return orc_update(bc, 0, bc->n, ts, R, render_xyz, render_exists);
}
