/*
 * nbody_b200.h — C ABI of libnbody_b200.so: the B200 (sm_100a) replacement for
 * the goroutine work pool of aceeric/nbodygo.
 *
 * The reference has no FFI seam on this path; the boundary is the block
 *   cmd/runner/computation-runner.go:285-320
 * inside ComputationRunner.runOneComputation (partition → submitSlice → wait →
 * ProcessMods → Update loop).  One nb_step() call replaces that block; the
 * control drain above it (:268-279) and the queue publish / Cycle below it
 * (:321-325) stay in the host application.  INTEGRATION.md shows the cgo
 * binding that a maintainer would add to cmd/runner.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; caller-allocated buffers only; the
 *     library never keeps a host pointer past the call (cgo pointer rules).
 *   - every function returns 0 (NB_OK) or a negative NB_ERR_* code;
 *     nb_last_error() returns a message for the last failing call.
 *   - a handle is not thread-safe: call from one OS thread
 *     (runtime.LockOSThread in the runner goroutine).
 *   - body order is array order of BodyCollection.arr
 *     (cmd/body/body_collection.go:16); indices in pair/event lists are array
 *     indices at the time of the step.
 *   - there is no CPU fallback: without a CUDA device nb_create fails.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------- */
#define NB_OK 0
#define NB_ERR_INVALID (-1)       /* bad argument / handle                        */
#define NB_ERR_CUDA (-2)          /* CUDA runtime failure (see nb_last_error)     */
#define NB_ERR_CAPACITY (-3)      /* body capacity exceeded                       */
#define NB_ERR_PAIR_OVERFLOW (-4) /* collision list overflowed; step NOT applied  */
#define NB_ERR_COMM (-5)          /* NCCL failure / not initialised               */
#define NB_ERR_NO_DEVICE (-6)     /* no usable CUDA device (no CPU fallback)      */

/* ---- CollisionBehavior, cmd/globals/globals.go:11-16 ------------------ */
#define NB_NONE 0
#define NB_SUBSUME 1
#define NB_ELASTIC 2
#define NB_FRAGMENT 3

/* ---- per-body flag byte (Body bool fields, cmd/body/body.go:42-51) ---- */
#define NB_F_EXISTS 0x01u      /* Body.Exists                                    */
#define NB_F_FRAGMENTING 0x02u /* Body.fragmenting: skipped as i and as j        */
#define NB_F_PINNED 0x04u      /* Body.Pinned (carried, not interpreted)         */
#define NB_F_SUN 0x08u         /* Body.IsSun  (carried, not interpreted)         */
#define NB_F_TELEMETRY 0x10u   /* Body.WithTelemetry (carried)                   */
#define NB_F_COLLIDED 0x20u    /* Body.collided: set by resolve, cleared by step */

/* ---- step options ------------------------------------------------------ */
#define NB_STEP_COLLISIONS 0x01u   /* detect (dist <= r1+r2) and emit events        */
#define NB_STEP_NO_RESOLVE 0x02u   /* leave elastic events unresolved (host does)   */
#define NB_STEP_NO_INTEGRATE 0x04u /* forces + events only; state is not advanced   */
#define NB_STEP_ASYNC 0x08u        /* enqueue only; nb_sync() collects the result   */
#define NB_STEP_PHASE_TIMINGS 0x10u /* CUDA events between the kernels: ms_prep ... ms_integrate
                                       (every event is a node between two kernels and costs
                                       2-3 us of a small cycle; without it only ms_total)   */
#define NB_STEP_DEFAULT (NB_STEP_COLLISIONS)

typedef struct nb_sim *nb_handle;

/* Result of one step (filled by nb_step, or by nb_sync after an async step). */
typedef struct {
    int64_t n_bodies;       /* bodies in the array                                  */
    int64_t n_pairs;        /* ordered elastic pairs (i,j) detected, all ranks      */
    int64_t n_host_events;  /* subsume / fragment records reported to the host      */
    int64_t n_resolved;     /* events for which doElastic ran                       */
    int64_t n_culled;       /* bodies whose Exists was cleared by the NaN cull      */
    int64_t n_dead;         /* bodies with Exists == false after the step           */
    int64_t n_subsumed;     /* bodies swallowed by ResolveSubsume in this step      */
    int32_t resolve_rounds; /* dependency rounds the resolve kernel needed          */
    int32_t pair_overflow;  /* 1: pair/event capacity exceeded, state NOT advanced  */
    /* device timings of the last step, milliseconds (CUDA events); the per-phase
     * values are 0 unless the step ran with NB_STEP_PHASE_TIMINGS */
    float ms_total, ms_prep, ms_force, ms_exchange, ms_resolve, ms_integrate;
} nb_step_result;

/* Event kinds reported to the host (cmd/body/event.go:20-24).  The device resolves
 * the whole event queue itself, in the reference's order (BodyCollection.ProcessMods,
 * cmd/body/body_collection.go:212-233): elastic collisions, ResolveSubsume (mass,
 * Exists) and the `fragmenting` flag of initiateFragmentation.  The records tell the
 * host what happened so that it can log, keep its own Body objects in step and do
 * the host-only part (fragInfo + spawning fragments, cmd/body/fragcalc.go:84-117). */
#define NB_EV_COLLISION 0 /* reserved                                               */
#define NB_EV_SUBSUME 1   /* a subsumed b (larger radius first), body.go:178-184,228-244 */
#define NB_EV_FRAGMENT 2  /* shouldFragment said yes: f1/f2 = thisFactor/otherFactor */
#define NB_EV_FRAG_INIT 3 /* initiateFragmentation ran for body a (partner b) while the queue was processed
                           * (fragcalc.go:54-83): dist = a's Mass AT THAT POINT of the queue (earlier subsumes
                           * included, later ones not), f1 = the fragFactor it was called with, applied = 1 if a
                           * is the event's b1, 2 if it is b2.  One record per call, reported in the reference's
                           * handling order, so that a host replays them in sequence (a body named twice keeps the
                           * fragInfo of the later call, as in the reference).  The position the reference records
                           * (fragcalc.go:77) is the one before Update: nb_get_cycle_top_positions. */

typedef struct {
    int32_t kind;
    int32_t a, b;    /* array indices of b1, b2 */
    int32_t applied; /* 1: the device applied it; 0: reported only (NB_STEP_NO_RESOLVE / NO_INTEGRATE) */
    double dist; /* centre distance (subsume / collision) */
    double f1, f2;
} nb_event;

/* ---- lifetime ---------------------------------------------------------- */

/* Creates a simulation on CUDA device `device` able to hold `capacity` bodies
 * and `pair_capacity` collision events per step (0 → default 4*capacity+65536).
 * Replaces NewWorkPool (cmd/runner/workpool.go:129-142). */
int nb_create(int device, int64_t capacity, int64_t pair_capacity, nb_handle *out);
int nb_destroy(nb_handle h);
/* Message for the last error on this handle (h may be NULL: last nb_create). */
const char *nb_last_error(nb_handle h);
int nb_abi_version(void);

/* ---- state sync (cycle top, computation-runner.go:268-273) ------------- */

/* Replaces the whole body array (NewSimBodyCollection,
 * cmd/body/body_collection.go:45-68).  restitution/frag_factor/frag_step/flags
 * may be NULL (→ 1, 0, 0, NB_F_EXISTS); behavior may be NULL (→ NB_ELASTIC). */
int nb_upload(nb_handle h, int64_t n,
              const double *x, const double *y, const double *z,
              const double *vx, const double *vy, const double *vz,
              const double *mass, const double *radius,
              const double *restitution, const double *frag_factor, const double *frag_step,
              const uint8_t *behavior, const uint8_t *flags);

/* nb_upload for several GPUs without n bodies crossing every host link: each handle passes only
 * ITS slice — bodies [first, first+count) must be the handle's i-range for n (nb_plan), the array
 * pointers address that slice — and the slices reach the other handles' replicas over NVLink
 * (peer stores; NCCL all-gathers in the fallback mode).  Collective: every handle of the
 * communicator must call it with the same n and the same set of non-NULL arrays.  On a single-GPU
 * handle it is nb_upload.  The partition it serves is the slice rule of
 * cmd/runner/computation-runner.go:286-293. */
int nb_upload_shard(nb_handle h, int64_t n, int64_t first, int64_t count,
                    const double *x, const double *y, const double *z,
                    const double *vx, const double *vy, const double *vz,
                    const double *mass, const double *radius,
                    const double *restitution, const double *frag_factor, const double *frag_step,
                    const uint8_t *behavior, const uint8_t *flags);

/* Overwrites bodies [first, first+count) — Body.ApplyMods
 * (cmd/body/body.go:274-313) and SetNotExists (:93-96) for a dirty range.
 * Any array pointer may be NULL (field unchanged). */
int nb_patch(nb_handle h, int64_t first, int64_t count,
             const double *x, const double *y, const double *z,
             const double *vx, const double *vy, const double *vz,
             const double *mass, const double *radius,
             const double *restitution, const double *frag_factor, const double *frag_step,
             const uint8_t *behavior, const uint8_t *flags);

/* Appends `count` bodies at the end of the array with r = R — the add half of
 * BodyCollection.Cycle (cmd/body/body_collection.go:273-291). */
int nb_append(nb_handle h, int64_t count, double R,
              const double *x, const double *y, const double *z,
              const double *vx, const double *vy, const double *vz,
              const double *mass, const double *radius,
              const double *frag_factor, const double *frag_step,
              const uint8_t *behavior, const uint8_t *flags);

/* Stable removal of bodies whose Exists is false — the delete half of
 * BodyCollection.Cycle (cmd/body/body_collection.go:253-272).  *n_out = new
 * count; if old_index != NULL (capacity >= new count) old_index[k] = previous
 * array index of the body now at k. */
int nb_compact(nb_handle h, int64_t *n_out, int64_t *old_index, int64_t old_index_cap);

int nb_count(nb_handle h, int64_t *n);

/* ---- the compute cycle ------------------------------------------------- */

/* One cycle: Body.Compute for every body (force accumulation + collision
 * detection, cmd/body/body.go:148-225), BodyCollection.ProcessMods for the
 * elastic events (cmd/body/body_collection.go:212-233 → body.go:248-264 →
 * collisioncalc.go:26-186) and Body.Update (body.go:114-139).
 * time_scaling and R are ComputationRunner.timeScaling / .R. */
int nb_step(nb_handle h, double time_scaling, double R, uint32_t opts, nb_step_result *out);
/* Waits for an NB_STEP_ASYNC step (or any pending work) and fills *out (may be NULL). */
int nb_sync(nb_handle h, nb_step_result *out);

/* ---- results ----------------------------------------------------------- */

/* Current state, array order. Any pointer may be NULL. */
int nb_download_state(nb_handle h,
                      double *x, double *y, double *z, double *vx, double *vy, double *vz,
                      double *mass, double *radius, double *restitution,
                      uint8_t *behavior, uint8_t *flags);
/* The same for bodies [first, first+count) only.  After a cycle every handle of a communicator
 * holds the whole state, so a host driving several GPUs lets each handle return its own i-range
 * (n/P bodies per host link instead of n). */
int nb_download_state_range(nb_handle h, int64_t first, int64_t count,
                            double *x, double *y, double *z, double *vx, double *vy, double *vz,
                            double *mass, double *radius, double *restitution,
                            uint8_t *behavior, uint8_t *flags);
int nb_download_render_range(nb_handle h, int64_t first, int64_t count, float *xyz, uint8_t *exists);
/* Renderable snapshot of the last step (cmd/body/renderable.go:22-40):
 * xyz = 3 floats per body (zeros for !Exists stubs), exists = 1 byte per body. */
int nb_download_render(nb_handle h, float *xyz, uint8_t *exists);
/* Zero-copy variant for the per-cycle render path: returns two library-owned pinned
 * host buffers (capacity-sized; C memory, so no Go pointer is retained) that EVERY
 * subsequent nb_step fills with the snapshot as part of the cycle's own stream —
 * valid from the return of nb_step / nb_sync until the next step.  Replaces the
 * per-body Renderable allocation of computation-runner.go:317-320. */
int nb_render_buffers(nb_handle h, float **xyz, uint8_t **exists);
/* Force accumulated on each body in the last step (Body.fx,fy,fz). */
int nb_get_forces(nb_handle h, double *fx, double *fy, double *fz);
/* Overwrites Body.fx,fy,fz of bodies [first, first+count).  Only a body that does not compute
 * (fragmenting, body.go:152-155) ever reads them back: Update keeps applying the force of its last
 * Compute.  nb_upload starts every body at 0 (NewBody, body.go:73-75); a host that re-uploads a
 * running collection restores the forces it read with nb_get_forces through this call. */
int nb_set_forces(nb_handle h, int64_t first, int64_t count, const double *fx, const double *fy, const double *fz);
/* Elastic collision events of the last step as ordered pairs, sorted by
 * (i asc, j asc) — the single-worker arrival order of the reference.
 * *n receives the total; at most cap are written. */
int nb_get_pairs(nb_handle h, int32_t *i, int32_t *j, int64_t cap, int64_t *n);
/* Positions the last nb_step started from — what Body.X,Y,Z held while ProcessMods ran (Update moves them
 * afterwards).  Valid for bodies that existed at finite positions at the top of that cycle, until the next step.
 * The host-only half of initiateFragmentation (fragInfo.curPos, fragcalc.go:77) reads them. */
int nb_get_cycle_top_positions(nb_handle h, int64_t first, int64_t count, double *x, double *y, double *z);
/* Subsume / fragment records of the last step, sorted by (kind, a, b); NB_EV_FRAG_INIT records in handling order.  (With
 * NB_STEP_NO_RESOLVE the unresolved collision events are the pair list of
 * nb_get_pairs.)  On several GPUs the subsume records cover the handle's own
 * i-range — merge the handles' lists — while the fragment records come from the
 * replicated resolve and are identical on every handle: take them from one. */
int nb_get_host_events(nb_handle h, nb_event *ev, int64_t cap, int64_t *n);

/* ---- multi-GPU (one handle per GPU, one process per GPU or one process) - */

/* 128-byte NCCL unique id; create on rank 0 and distribute by any means. */
int nb_comm_unique_id(void *id128);
/* Joins a communicator of `nranks` handles. After this the i-bodies are
 * sharded in contiguous ranges of ceil(n/nranks) (the slice rule of
 * computation-runner.go:286-293); every rank must make the same sequence of
 * upload/patch/append/compact/step calls with the same arguments. */
int nb_comm_init(nb_handle h, int rank, int nranks, const void *id128);
/* How this handle exchanges data with the other handles of its communicator each cycle:
 * NB_COMM_SINGLE (no communicator), NB_COMM_PEER_PUSH (the kernels store into the peers'
 * replicas over NVLink: no collective in steady state) or NB_COMM_NCCL (NCCL all-gathers — the
 * fallback when a peer mapping failed, or NB_PEER_PUSH=0; nb_comm_init also prints a [WARN]
 * line to stderr when it falls back on its own).  Results are bit-identical in all three. */
#define NB_COMM_SINGLE 0
#define NB_COMM_PEER_PUSH 1
#define NB_COMM_NCCL 2
int nb_comm_mode(nb_handle h, int *mode);
/* The i-range [i0,i1) this handle computes. */
int nb_shard_range(nb_handle h, int64_t *i0, int64_t *i1);

/* Pure planning function, usable without a device: the i-range [i0,i1) of `rank`
 * out of `nranks` for n bodies (contiguous ceil(n/nranks) slices,
 * computation-runner.go:286-293) and the j-chunking (partial-sum slots per body,
 * j-tiles per chunk; a tile holds 64 bodies below 16,384 bodies, 256 up to 786,431,
 * 512 from there on) — all functions of n only. */
int nb_plan(int64_t n, int rank, int nranks, int64_t *i0, int64_t *i1, int32_t *n_chunks,
            int32_t *tiles_per_chunk);

/* ---- diagnostics -------------------------------------------------------- */

/* FP64 FMA throughput of the device: runs `iters` dependent-chain DFMA
 * batches on every SM and returns achieved TFLOP/s (2 flops per FMA). Used by
 * bench.py to state the roofline denominator. */
int nb_measure_fp64_peak(int device, int iters, double *tflops, float *ms);
/* A cycle whose parameters repeat (same body count, options, time scaling, R) is
 * captured once into a CUDA graph and replayed with a single launch: for small
 * collections the CPU cost of issuing the ~17 stream calls of a cycle exceeds the
 * kernels' run time.  Single-GPU handles only; NB_GRAPH=0 in the environment turns
 * it off.  Returns how many graphs were captured and how many cycles were replays. */
int nb_graph_stats(nb_handle h, int64_t *captures, int64_t *replays);

/* Number of kernel launches issued by this handle since creation. */
int nb_launch_count(nb_handle h, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
