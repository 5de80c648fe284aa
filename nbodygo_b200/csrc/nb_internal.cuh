// nb_internal.cuh — shared device/host declarations of libnbody_b200.so.
//
// Data layout in HBM (all arrays sized to `cap_pad`, a multiple of TJ):
//   fp64 SoA state   x y z vx vy vz mass radius rest frag_factor frag_step
//   fp64 derived     jm   (effective j-mass: 0 for !Exists / fragmenting bodies)
//                    m0   (mass at the top of the cycle; a subsume changes mass before Update)
//   fp64 output      fx fy fz
//   u8               behavior flags
//   per-tile         tile_rmax[n_tiles]  (max radius of the live bodies of a j-tile)
//                    tile_muni[n_tiles]  (the mass every live body of the tile has, or 0 if they differ)
//                    tile_dead[n_tiles]  (1 if a !Exists body still sits in the tile at a finite position)
//   partial sums     px py pz [S][n_pad_local]  (one slot per j-chunk, summed in
//                    ascending chunk order by the integrate kernel — the result
//                    for body i never depends on the grid or the rank count)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nbody_b200.h"

namespace nb {

// bodies per j-tile (one TMA bulk copy per field and tile).  Small collections use small tiles so
// that the sweep has enough independent (i-block, j-chunk) work items; the choice is a function of
// n only (chunking() in nb_api.cu), like everything that fixes the per-body summation order.
// Three tile sizes, chosen from n only: 64 below 16,384 bodies (enough independent work items for small
// collections), 256 up to 786,431, 512 from there on — at 1 M bodies the per-tile work of K1 (barrier wait, per-body
// screen thresholds, commit, __syncthreads) is 0.85 % of the sweep with 256-body tiles; measured -0.45 % with 512
// (profiles/r2_k1_variants.txt).  The redo of a screened (body, tile) costs twice as much, which only matters for
// dense small collections, and those use the small tiles.
constexpr int TJ_HUGE = 512, TJ_LARGE = 256, TJ_SMALL = 64;
constexpr long long TJ_SMALL_BELOW = 16384;  // n < this uses TJ_SMALL; n >= LARGE_N_BELOW uses TJ_HUGE
constexpr int TJ = TJ_HUGE;                  // allocation granularity
// smem stages of the j pipeline: 3 for the 64- and 256-body tiles (measured -0.1 % / -0.6 % on the uniform /
// per-body-mass sweep against 2), 2 for the 512-body tiles (3 x 16 KB would pass the 48 KB of static shared memory)
__host__ __device__ constexpr int stages_for(int tj) { return tj >= 512 ? 2 : 3; }
#ifndef NB_EXP_NSTAGE
#define NB_EXP_NSTAGE 0   // development knob (tools/k1_hw_variants.py): force a stage count (0: stages_for)
#endif
// j-chunks per body (= partial-sum slots; a function of n only, see chunking() in nb_api.cu)
constexpr int MIN_CHUNKS = 32;
constexpr int MAX_CHUNKS = 64;  // more slots cost K4 more than they gain K1 (tools/chunk_sweep.py: n = 32 k cycle 1114 -> 1092 us)
constexpr long long LARGE_N_BELOW = 786432;  // n >= this always uses MAX_CHUNKS slots (multi-GPU wave tail, chunking())
constexpr long long CHUNK_TARGET_CTAS = 148 * 2 * 40;  // (n/512) * chunks >= this when possible
constexpr int MAX_RANKS = 16;
// One event list for collisions and subsumes (event.go:20-24 kinds share one queue in the reference):
// entry = (i, j | kind bit), i = the body whose sweep raised it, j = the other body.
constexpr int EV_SUBSUME_BIT = 1 << 30;
constexpr int EV_INDEX_MASK = EV_SUBSUME_BIT - 1;
// fragcalc.go: maxFrags
constexpr double MAX_FRAGS = 2000.0;

// cmd/body/body.go:19
__host__ __device__ constexpr double G_CONST() { return 6.673e-11; }

struct Counters {
    unsigned long long n_pairs;     // local pairs appended this step
    unsigned long long n_hev;       // host events appended this step
    unsigned long long n_resolved;  // doElastic applications
    unsigned long long n_culled;    // NaN-culled bodies (local shard)
    unsigned long long n_dead;      // !Exists bodies after the step (all bodies)
    unsigned long long n_sub_events;  // subsume events in the event list (all ranks)
    unsigned long long n_subsumed;    // bodies whose Exists was cleared by ResolveSubsume
    int overflow;                   // pair / host-event capacity exceeded
    int rounds;                     // resolve rounds
    int total_pairs;                // pairs over all ranks (set by resolve)
    int peer_timeout;               // a peer flag wait ran into its time limit
};

// The part of calcElasticCollision that depends on positions and radii only (fixed during
// ProcessMods): computed for every collision event up front, in parallel, so that it is off the
// dependent chain of the resolve rounds.
struct __align__(16) ElasticGeo {
    double d, st, ct, sp, cp, r12;
};

struct DevState {
    double *x, *y, *z, *vx, *vy, *vz, *mass, *radius, *rest, *ff, *fs;
    double *jx, *jy, *jz, *jm;  // j-stream built by K0 (sanitised positions + effective mass)
    double *m0;                 // Body.Mass at the top of the cycle (the m_i of this cycle's force)
    uint8_t *computes0;         // 1 if Body.Compute ran for the body this cycle (Exists && !fragmenting at the top)
    double *fx, *fy, *fz;
    uint8_t *behavior, *flags;
    double *tile_rmax;
    double *tile_muni;          // > 0: every live j-body of the tile has exactly this mass (slots without a live
                                // body match any mass); 0: mixed (K0)
    uint8_t *tile_dead;         // 1: the tile holds a body that does not exist but still sits in the array at a
                                // finite position (visited by the collision sweep only: dead_j_sweep in K1)
    double *px, *py, *pz;
    float *render;
    uint8_t *render_exists;
    int2 *pairs;               // events appended by the local K1 (capacity seg_cap)
    int2 *pairs_all;           // [nranks][seg_stride] after the exchange (== pairs on one GPU)
    unsigned long long *pair_counts;  // [nranks] after the exchange
    nb_event *hev;
    unsigned long long *head;  // [cap_pad] resolve scheduling: largest pending event key per body (0 between steps)
    int *adj_off, *adj_cnt;    // [cap_pad] resolve scheduling: the body's slice of rs_list (0 between steps)
    // resolve scratch, capacity = events of all ranks (seg_cap * nranks)
    int2 *rs_ev;               // [E] the event list, contiguous
    int *rs_list;              // [2E] event indices grouped by body (unsorted CSR)
    unsigned long long *rs_lkey;  // [2E] their keys; 0 once the event is resolved
    int2 *rs_pos;              // [E] where the event sits in its two bodies' slices
    unsigned long long *rs_candkey;  // [2E] keys of the new-head candidates of a round
    int *rs_state;             // [E] 0 pending, 1 queued, 2 done
    int *rs_queue;             // [E] ready events in wavefront order
    int *rs_cand;              // [2E] new-head candidates of a round
    int *rs_active;            // [2E] bodies that have events this step
    ElasticGeo *rs_geo;        // [E] geometry of the collision events
    int *rs_ctl;               // [8] scheduling counters of the resolve cluster (active bodies, cursor, queue begin/end/tail)
    Counters *ctr;
    unsigned *zeros;           // 1024 zeros (opaque low words for the rsqrt seeds in K1)
};

// Peer replicas of the state arrays K4 updates (mapped through CUDA IPC, or plain UVA pointers when
// all handles live in one process).  K4 stores the new shard state into every peer over NVLink:
// the "all-gather" is fused into the integrate kernel.  sync[q] points at rank q's flag block:
// slots [0,MAX_RANKS) = "arrived(step)" written by each peer, [MAX_RANKS,2*MAX_RANKS) = "done reading(step)".
struct PeerTable {
    double *x[MAX_RANKS], *y[MAX_RANKS], *z[MAX_RANKS], *vx[MAX_RANKS], *vy[MAX_RANKS], *vz[MAX_RANKS],
        *rest[MAX_RANKS];
    uint8_t *flags[MAX_RANKS];
    unsigned long long *sync[MAX_RANKS];
    int2 *pairs_all[MAX_RANKS];               // base of rank q's gathered pair buffer: [2 parities][nranks][seg_cap]
    unsigned long long *pair_counts[MAX_RANKS];  // base of rank q's counts: [2 parities][MAX_RANKS]
    // Body.fx,fy,fz of the bodies that do not compute (fragmenting: Update keeps applying the force of
    // their last Compute, body.go:152-155) — pushed by K4 so that a shard boundary moved by Cycle finds
    // them on the new owner
    double *fx[MAX_RANKS], *fy[MAX_RANKS], *fz[MAX_RANKS];
    // the fields only a sharded upload writes (nb_upload_shard)
    double *mass[MAX_RANKS], *radius[MAX_RANKS], *ff[MAX_RANKS], *fs[MAX_RANKS];
    uint8_t *behavior[MAX_RANKS];
};
constexpr int PEER_ARRAYS = 19;  // mapped arrays per rank (see setup_peer_push)
// flag slots of a rank's sync block; every slot holds the last id its writer published (monotonic)
constexpr int PEER_SLOT_ARRIVED = 0, PEER_SLOT_DONE = MAX_RANKS, PEER_SLOT_PAIRS = 2 * MAX_RANKS;
constexpr int PEER_SLOT_UP_READY = 3 * MAX_RANKS, PEER_SLOT_UP_ARRIVED = 4 * MAX_RANKS;
constexpr int PEER_SYNC_SLOTS = 5 * MAX_RANKS;

// which arrays a sharded upload pushes to the peers (bit k = field k of write_range's order:
// x y z vx vy vz mass radius rest ff fs, then behavior (11) and flags (12))
constexpr unsigned PUSH_BEHAVIOR = 1u << 11, PUSH_FLAGS = 1u << 12;

struct StepParams {
    DevState s;
    const PeerTable *peers;      // device pointer, or nullptr (single GPU / NCCL fallback)
    unsigned long long step_id;  // cycle counter, identical on every rank
    long long n;         // bodies
    long long i0, i1;    // local i-shard
    long long n_pad_local;  // stride of partial-sum slots
    int tj;              // tile size of this cycle: TJ_SMALL, TJ_LARGE or TJ_HUGE
    int n_tiles;         // ceil(n / tj)
    int n_chunks;        // S
    int tiles_per_chunk;
    int rank, nranks;
    long long seg_cap;   // pair capacity per rank
    long long seg_stride;  // stride of the per-rank segments in pairs_all
    long long hev_cap;
    unsigned opts;
    int uniform_tiles;   // 1: K0 marks uniform-mass tiles and K1 hoists the mass out of their pair loop
    int res_cluster;     // CTAs of the resolve cluster (1..8)
    int res_fast;        // 1: event lists of up to 64 entries are scheduled in shared memory (NB_RES_FAST=0: never)
    double ts, R;
};

// launchers (each returns the number of kernel launches it issued)
int launch_prep(const StepParams &p, cudaStream_t st);
int launch_force(const StepParams &p, cudaStream_t st, int force_R);
int launch_resolve(const StepParams &p, cudaStream_t st);
int launch_integrate(const StepParams &p, cudaStream_t st);
int launch_count_dead(const StepParams &p, cudaStream_t st);
// peer flag protocol: slot_base 0 = arrived, MAX_RANKS = done reading
int launch_peer_signal(const StepParams &p, int slot_base, cudaStream_t st);
int launch_peer_wait(const StepParams &p, int slot_base, cudaStream_t st);
// copies this rank's pair list and count into every rank's gathered buffer (this cycle's parity)
int launch_push_pairs(const StepParams &p, cudaStream_t st);
// stores the rank's own slice [i0,i1) of the arrays in `mask` into every peer's replica (nb_upload_shard)
int launch_push_shard(const StepParams &p, unsigned mask, cudaStream_t st);
// stable compaction of !Exists bodies; returns launches. d_map[k] = old index of new body k.
int launch_compact_map(const DevState &s, long long n, long long *d_map, unsigned *d_block_sums,
                       long long *d_new_n, cudaStream_t st);
int launch_gather_f64(const double *src, double *dst, const long long *d_map, const long long *d_new_n,
                      long long n_old, cudaStream_t st);
int launch_gather_u8(const uint8_t *src, uint8_t *dst, const long long *d_map, const long long *d_new_n,
                     long long n_old, cudaStream_t st);
int launch_fill_f64(double *dst, double v, long long n, cudaStream_t st);
int launch_fill_u8(uint8_t *dst, uint8_t v, long long n, cudaStream_t st);
int launch_fp64_peak(int iters, int blocks, double *d_out, cudaStream_t st);
int launch_fp64_mix(int kind, int iters, int blocks, double *d_out, cudaStream_t st);

}  // namespace nb
