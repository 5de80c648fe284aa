/*
 * nbody_oracle.h — CPU oracle for the nbodygo per-cycle compute path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped GPU
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.
 *
 * PARITY UNPINNED: the reference (aceeric/nbodygo, Go) cannot be built in this
 * image (no Go toolchain) and its own tests hold no numeric golden vector for
 * the hot path (cmd/runner/workpool_test.go:41-56 asserts only Vx != 0,
 * cmd/body/body_collection_test.go:319-344 asserts only collided == true).
 * The oracle is therefore pinned against (1) hand-derived known-answer tests
 * KAT-1..6 of SURVEY.md §8c and (2) an independent pure-Python restatement
 * (tests/golden/make_golden.py), not against reference output.
 *
 * Every function cites the reference lines it restates (paths relative to
 * the reference repository root).
 *
 * FP model: Go gc on amd64 — no FMA contraction, strict left-to-right
 * evaluation, math.Sqrt == SQRTSD.  Build with -ffp-contract=off, no
 * -ffast-math (see oracle/Makefile).  The Go math library transcendentals
 * (Acos/Atan2/Sin/Cos/Asin/Tan, cmd/body/collisioncalc.go:104-160) come from
 * one of two backends (orc_set_math): glibc libm (default; the committed
 * golden fixtures were generated with it) or a restatement of the Go standard
 * library's own algorithms (gomath.c — what the reference executes on amd64).
 * The two differ in the last ulp; tests/test_gomath.py measures what that does
 * to post-collision velocities, which is why those carry a tolerance.
 */
#ifndef NBODY_ORACLE_H
#define NBODY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cmd/globals/globals.go:11-16 */
enum { ORC_NONE = 0, ORC_SUBSUME = 1, ORC_ELASTIC = 2, ORC_FRAGMENT = 3 };

/* flag bits (same values as include/nbody_b200.h) */
#define ORC_F_EXISTS 0x01u      /* Body.Exists      cmd/body/body.go:43 */
#define ORC_F_FRAGMENTING 0x02u /* Body.fragmenting cmd/body/body.go:47 */
#define ORC_F_PINNED 0x04u      /* Body.Pinned      cmd/body/body.go:45 */
#define ORC_F_SUN 0x08u         /* Body.IsSun       cmd/body/body.go:42 */
#define ORC_F_TELEMETRY 0x10u   /* Body.WithTelemetry */
#define ORC_F_COLLIDED 0x20u    /* Body.collided    cmd/body/body.go:51 */

/* event kinds, cmd/body/event.go:20-24 (+ fragment hand-off record) */
enum { ORC_EV_COLLISION = 0, ORC_EV_SUBSUME = 1, ORC_EV_FRAGMENT = 2,
       /* initiateFragmentation (fragcalc.go:66-83) ran for body a (partner b): dist = a's mass at that point of the
        * queue, f1 = the fragFactor it was called with, f2 = 1 if a is the event's b1 else 2.  Pushed in handling
        * order, before the ORC_EV_FRAGMENT record of the same event. */
       ORC_EV_FRAG_INIT = 3 };

/* Structure-of-arrays view of []*Body (cmd/body/body.go:34-52). All arrays
 * have at least n entries; the oracle never allocates or frees them. */
typedef struct {
    int64_t n;
    double *x, *y, *z, *vx, *vy, *vz, *mass, *radius;
    double *rest;        /* Body.r — restitution applied by doElastic     */
    double *frag_factor; /* Body.FragFactor (may be NULL → 0)             */
    double *frag_step;   /* Body.FragStep   (may be NULL → 0)             */
    double *fx, *fy, *fz;
    uint8_t *behavior;   /* ORC_NONE..ORC_FRAGMENT                        */
    uint8_t *flags;      /* ORC_F_*                                       */
} orc_bodies;

/* One deferred event (cmd/body/event.go:29-33). For ORC_EV_FRAGMENT the two
 * factors are shouldFragment's thisFactor/otherFactor. */
typedef struct {
    int32_t kind;
    int32_t a, b;    /* array indices of b1, b2 */
    double dist;     /* centre distance computed by Collided */
    double f1, f2;
} orc_event;

/* Options for orc_compute */
#define ORC_OPT_SELF_PAIRS 0x1u   /* keep (i,i) events like the reference (F5)   */
#define ORC_OPT_DEAD_J 0x2u       /* keep events whose j does not exist (A2)      */
#define ORC_OPT_SINGLE_SWEEP 0x4u /* fuse force+collision sweep (timing variant)  */
#define ORC_OPT_DEAD_J_SUBSUME 0x8u /* keep the SUBSUME events whose j does not exist: ResolveSubsume
                                     * (body.go:228-244) has no Exists gate, so they change state; the
                                     * collision events with a dead j stay dropped (ResolveCollision's gate,
                                     * body.go:249-251, makes them no-ops).  This is the canonical stream the
                                     * device emits. */

/* Body.Compute for i in [i0,i1): force sweep then collision sweep
 * (cmd/body/body.go:148-187, 192-225).  Events are appended to ev (capacity
 * ev_cap) in single-worker arrival order (i asc, j asc); *n_ev is read as the
 * current length and updated.  Returns 0, or -1 if ev overflowed (the excess
 * is counted in *n_ev but not stored). */
int orc_compute(const orc_bodies *bc, int64_t i0, int64_t i1, uint32_t opts,
                orc_event *ev, int64_t ev_cap, int64_t *n_ev);

/* Same force definition evaluated term-by-term in __float128 and summed in
 * __float128 (rounded to double on output); the inclusion predicate is the
 * reference's double-precision one.  Also returns sum_j |f_ij| per body in
 * fnorm (may be NULL).  Used to adjudicate the force tolerance. */
int orc_compute_exact(const orc_bodies *bc, int64_t i0, int64_t i1,
                      double *fx, double *fy, double *fz, double *fnorm);

/* The work-pool fan-out of ComputationRunner.runOneComputation
 * (cmd/runner/computation-runner.go:286-311, cmd/runner/workpool.go:103-110):
 * contiguous slices of n/workers bodies (one slice if n < 100), one pthread
 * per worker, round-robin slice assignment.  Events from every slice are
 * concatenated in slice order, i.e. identical to the single-worker order. */
int orc_compute_pool(const orc_bodies *bc, int workers, uint32_t opts,
                     orc_event *ev, int64_t ev_cap, int64_t *n_ev);

/* Timing-only variant: computes the slice [i0,i1) with `workers` threads and
 * discards events beyond counting them. Returns number of events or <0. */
int64_t orc_compute_slice_timed(const orc_bodies *bc, int64_t i0, int64_t i1,
                                int workers, uint32_t opts);

/* BodyCollection.ProcessMods idealised: handle events in
 * reverse arrival order (PushFront + Front→Next,
 * cmd/body/body_collection.go:82-104,212-233) through event.Handle
 * (cmd/body/event.go:53-62) → ResolveCollision / ResolveSubsume
 * (cmd/body/body.go:228-264).  Fragment decisions (shouldFragment,
 * cmd/body/fragcalc.go:24-49) are appended to out_ev (may be NULL). */
int orc_process_mods(orc_bodies *bc, const orc_event *ev, int64_t n_ev,
                     orc_event *out_ev, int64_t out_cap, int64_t *n_out);

/* Transcendental backend used by calcElasticCollision (process-wide). */
enum { ORC_MATH_LIBM = 0, ORC_MATH_GO = 1 };
int orc_set_math(int which); /* 0, or -1 for an unknown backend */
int orc_get_math(void);

/* calcElasticCollision (cmd/body/collisioncalc.go:42-186). out[0]=collided
 * (0/1), out[1..3]=v1', out[4..6]=v2', out[7..9]=v_cm. */
void orc_calc_elastic(const orc_bodies *bc, int64_t a, int64_t b, double out[10]);

/* Body.Update over i in [i0,i1) (cmd/body/body.go:114-139) +
 * NewRenderable (cmd/body/renderable.go:22-40).  render_xyz (3 floats per
 * body, may be NULL) and render_exists (may be NULL) are indexed by i. */
int orc_update(orc_bodies *bc, int64_t i0, int64_t i1, double time_scaling,
               double R, float *render_xyz, uint8_t *render_exists);

/* BodyCollection.Cycle compaction half (cmd/body/body_collection.go:253-296):
 * stable removal of !Exists.  Returns the new n (also stored in bc->n).
 * If map_out != NULL, map_out[new_index] = old_index. */
int64_t orc_cycle_compact(orc_bodies *bc, int64_t *map_out);

/* One full cycle, steps 5-7 of runOneComputation
 * (cmd/runner/computation-runner.go:297-320): compute (single worker) →
 * ProcessMods → Update.  Events (arrival order) are left in ev. */
int orc_step(orc_bodies *bc, double time_scaling, double R, uint32_t opts,
             orc_event *ev, int64_t ev_cap, int64_t *n_ev,
             orc_event *out_ev, int64_t out_cap, int64_t *n_out,
             float *render_xyz, uint8_t *render_exists);

#ifdef __cplusplus
}
#endif
#endif
