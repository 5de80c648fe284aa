// nb_resolve.cu — K3: deterministic parallel restatement of BodyCollection.ProcessMods
// for collision events (cmd/body/body_collection.go:212-233 → event.Handle,
// cmd/body/event.go:53-62 → Body.ResolveCollision, cmd/body/body.go:248-264 →
// calcElasticCollision / doElastic, cmd/body/collisioncalc.go:26-186, and
// shouldFragment, cmd/body/fragcalc.go:24-49).
//
// Subsume events (Body.ResolveSubsume, cmd/body/body.go:228-244) travel in the same list with
// their arrival position, exactly as they share the reference's one event queue, and are applied
// here in that serial order; the flag half of initiateFragmentation (cmd/body/fragcalc.go:66-83)
// is applied too.  The host only receives notification records (nb_get_host_events).
//
// Reference semantics (idealised): events are handled serially in
// reverse arrival order; single-worker arrival order is (i asc, j asc), so the
// resolve order is descending key = (i << 32 | j).  Two events commute unless they
// share a body, so the kernel runs rounds: an event is ready when it holds the
// largest key among the unprocessed events of BOTH its bodies (head[] built with
// 64-bit atomicMax); ready events of one round touch disjoint bodies and run in
// parallel.  The outcome equals the serial order exactly, independent of the order
// in which K1 appended the events.
//
// Compiled with -fmad=false: arithmetic is the unfused left-to-right sequence of
// the reference; only the libm calls (acos/atan2/sin/cos/asin/tan) can differ from
// Go's / glibc's in the last ulp.
#include <cstdlib>

#include <cooperative_groups.h>

#include "nb_internal.cuh"

namespace cg = cooperative_groups;

namespace nb {

constexpr int RES_THREADS_DEFAULT = 512;  // 128 registers per thread: the libm chains of calcElasticCollision overlap, no spills
constexpr double PI_D = 3.14159265358979323846;

__device__ __forceinline__ bool ef(unsigned b) { return b == NB_ELASTIC || b == NB_FRAGMENT; }

struct CollResult {
    bool collided;
    double vx1, vy1, vz1, vx2, vy2, vz2, vx_cm, vy_cm, vz_cm;
};

// calcElasticCollision, cmd/body/collisioncalc.go:42-186 — the position-only part (:79,104-113):
// distance, polar angles of r21 and their sines / cosines, r12.  Same operations, same order.
__device__ ElasticGeo elastic_geometry(const DevState &s, int a, int b)
{
    ElasticGeo g;
    g.r12 = s.radius[a] + s.radius[b];
    const double x21 = s.x[b] - s.x[a], y21 = s.y[b] - s.y[a], z21 = s.z[b] - s.z[a];
    g.d = sqrt(x21 * x21 + y21 * y21 + z21 * z21);
    const double theta2 = acos(z21 / g.d);
    const double phi2 = (x21 == 0 && y21 == 0) ? 0.0 : atan2(y21, x21);
    // sincos shares the argument reduction of a sin/cos pair
    sincos(theta2, &g.st, &g.ct);
    sincos(phi2, &g.sp, &g.cp);
    return g;
}

// calcElasticCollision, cmd/body/collisioncalc.go:42-186 — the part that depends on the bodies'
// current velocities and masses (both change while the queue is processed)
__device__ CollResult calc_elastic(const DevState &s, int a, int b, const ElasticGeo &g)
{
    CollResult res;
    res.collided = false;
    const double m1 = s.mass[a], m2 = s.mass[b];
    double vx1 = s.vx[a], vy1 = s.vy[a], vz1 = s.vz[a];
    const double vx2 = s.vx[b], vy2 = s.vy[b], vz2 = s.vz[b];

    const double r12 = g.r12;
    const double m21 = m2 / m1;
    const double vx21 = vx2 - vx1, vy21 = vy2 - vy1, vz21 = vz2 - vz1;

    const double vx_cm = (m1 * vx1 + m2 * vx2) / (m1 + m2);
    const double vy_cm = (m1 * vy1 + m2 * vy2) / (m1 + m2);
    const double vz_cm = (m1 * vz1 + m2 * vz2) / (m1 + m2);

    const double d = g.d;
    const double v = sqrt(vx21 * vx21 + vy21 * vy21 + vz21 * vz21);
    if (v == 0) return res;

    vx1 = -vx21; vy1 = -vy21; vz1 = -vz21;
    const double st = g.st, ct = g.ct, sp = g.sp, cp = g.cp;

    double vx1r = ct * cp * vx1 + ct * sp * vy1 - st * vz1;
    double vy1r = cp * vy1 - sp * vx1;
    double vz1r = st * cp * vx1 + st * sp * vy1 + ct * vz1;
    double fvz1r = vz1r / v;
    if (fvz1r > 1) fvz1r = 1;
    else if (fvz1r < -1) fvz1r = -1;
    const double thetav = acos(fvz1r);
    const double phiv = (vx1r == 0 && vy1r == 0) ? 0.0 : atan2(vy1r, vx1r);

    const double dr = d * sin(thetav) / r12;
    if (thetav > PI_D / 2 || fabs(dr) > 1) return res;

    const double alpha = asin(-dr);
    const double beta = phiv;
    double sbeta, cbeta;
    sincos(beta, &sbeta, &cbeta);
    const double a_ = tan(thetav + alpha);
    const double dvz2 = 2 * (vz1r + a_ * (cbeta * vx1r + sbeta * vy1r)) / ((1 + a_ * a_) * (1 + m21));

    const double vz2r = dvz2;
    const double vx2r = a_ * cbeta * dvz2;
    const double vy2r = a_ * sbeta * dvz2;
    vz1r = vz1r - m21 * vz2r;
    vx1r = vx1r - m21 * vx2r;
    vy1r = vy1r - m21 * vy2r;

    res.collided = true;
    res.vx1 = ct * cp * vx1r - sp * vy1r + st * cp * vz1r + vx2;
    res.vy1 = ct * sp * vx1r + cp * vy1r + st * sp * vz1r + vy2;
    res.vz1 = ct * vz1r - st * vx1r + vz2;
    res.vx2 = ct * cp * vx2r - sp * vy2r + st * cp * vz2r + vx2;
    res.vy2 = ct * sp * vx2r + cp * vy2r + st * sp * vz2r + vy2;
    res.vz2 = ct * vz2r - st * vx2r + vz2;
    res.vx_cm = vx_cm; res.vy_cm = vy_cm; res.vz_cm = vz_cm;
    return res;
}

__device__ void push_host_event(const StepParams &p, int kind, int a, int b, int applied, double dist, double f1,
                                double f2)
{
    const unsigned long long k = atomicAdd(&p.s.ctr->n_hev, 1ull);
    if (k < (unsigned long long)p.hev_cap) {
        nb_event e;
        e.kind = kind; e.a = a; e.b = b; e.applied = applied; e.dist = dist; e.f1 = f1; e.f2 = f2;
        p.s.hev[k] = e;
    }  // beyond the capacity the records are lost, the count is not (pair_overflow == 2)
}

// initiateFragmentation, cmd/body/fragcalc.go:66-83 — the state the next Compute sees (the
// `fragmenting` flag); fragInfo stays with the host.  math.Min propagates NaN, fmin does not.
__device__ void initiate_fragmentation(const DevState &s, int i, double fragFactor)
{
    const double ff = s.ff[i];
    const double fragDelta = ff > 10 ? 10.0 : fragFactor - ff;
    const double t = fragDelta * s.fs[i];
    const double fragments = isnan(t) ? t : fmin(t, MAX_FRAGS);
    if (fragments <= 1) {
        s.behavior[i] = NB_FRAGMENT;
        return;
    }
    s.flags[i] |= NB_F_FRAGMENTING;
}

// The subsume event raised by body i's sweep at j (body.go:178-184): the body with the larger
// radius subsumes the other.  K1 only queues it when the centre distance is within that radius.
// ResolveSubsume has no Exists gate: a second event for the same couple adds the 0 of SetNotExists.
__device__ void resolve_subsume(const StepParams &p, int i, int j, bool apply)
{
    const DevState &s = p.s;
    const double ri = s.radius[i], rj = s.radius[j];
    const int a = ri > rj ? i : j, b = ri > rj ? j : i;
    const double dx = s.x[j] - s.x[i], dy = s.y[j] - s.y[i], dz = s.z[j] - s.z[i];
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);  // unfused, as K1 computed it
    if (apply) {
        const double thisMass = s.mass[a], otherMass = s.mass[b];
        s.mass[a] = thisMass + otherMass;
        s.mass[b] = 0.0;  // SetNotExists, body.go:93-96
        if (s.flags[b] & NB_F_EXISTS) {
            s.flags[b] &= (uint8_t)~NB_F_EXISTS;
            atomicAdd(&s.ctr->n_subsumed, 1ull);
        }
    }
    push_host_event(p, NB_EV_SUBSUME, a, b, apply ? 1 : 0, dist, 0.0, 0.0);
}

// Body.ResolveCollision for the ready event (a,b); the caller guarantees no other
// thread touches a or b in this round.
__device__ void resolve_one(const StepParams &p, int a, int b, const ElasticGeo &geo)
{
    const DevState &s = p.s;
    if (!(s.flags[a] & NB_F_EXISTS) || !(s.flags[b] & NB_F_EXISTS)) return;  // body.go:249-251
    const unsigned ba = s.behavior[a], bb = s.behavior[b];
    if (!(ba == NB_ELASTIC && ef(bb))) return;  // body.go:252-253
    const CollResult r = calc_elastic(s, a, b, geo);
    if (!r.collided) return;
    const double br = s.rest[a];
    const double nvx1 = (r.vx1 - r.vx_cm) * br + r.vx_cm;
    const double nvy1 = (r.vy1 - r.vy_cm) * br + r.vy_cm;
    const double nvz1 = (r.vz1 - r.vz_cm) * br + r.vz_cm;
    const double nvx2 = (r.vx2 - r.vx_cm) * br + r.vx_cm;
    const double nvy2 = (r.vy2 - r.vy_cm) * br + r.vy_cm;
    const double nvz2 = (r.vz2 - r.vz_cm) * br + r.vz_cm;
    if (ba == NB_FRAGMENT || bb == NB_FRAGMENT) {
        // shouldFragment, fragcalc.go:24-49
        const double vThis = s.vx[a] + s.vy[a] + s.vz[a];
        const double dvThis = fabs(s.vx[a] - nvx1) + fabs(s.vy[a] - nvy1) + fabs(s.vz[a] - nvz1);
        const double thisFactor = dvThis / fabs(vThis);
        const double vOther = s.vx[b] + s.vy[b] + s.vz[b];
        const double dvOther = fabs(s.vx[b] - nvx2) + fabs(s.vy[b] - nvy2) + fabs(s.vz[b] - nvz2);
        const double otherFactor = dvOther / fabs(vOther);
        if ((ba == NB_FRAGMENT && thisFactor > s.ff[a]) || (bb == NB_FRAGMENT && otherFactor > s.ff[b])) {
            // doFragment, fragcalc.go:54-61: the flag half happens here, in event order; the fragInfo
            // bookkeeping and the spawning of fragments are host work fed by this record
            // one NB_EV_FRAG_INIT record per initiateFragmentation call, with the mass the body has at this point of
            // the queue (fragInfo.mass = Mass / fragments, fragcalc.go:80-82)
            if (ba == NB_FRAGMENT && thisFactor > s.ff[a]) {
                push_host_event(p, NB_EV_FRAG_INIT, a, b, 1, s.mass[a], thisFactor, 0.0);
                initiate_fragmentation(s, a, thisFactor);
            }
            if (bb == NB_FRAGMENT && otherFactor > s.ff[b]) {
                push_host_event(p, NB_EV_FRAG_INIT, b, a, 2, s.mass[b], otherFactor, 0.0);
                initiate_fragmentation(s, b, otherFactor);
            }
            push_host_event(p, NB_EV_FRAGMENT, a, b, 1, 0.0, thisFactor, otherFactor);
            return;
        }
    }
    // doElastic, collisioncalc.go:26-35
    s.vx[a] = nvx1; s.vy[a] = nvy1; s.vz[a] = nvz1;
    s.vx[b] = nvx2; s.vy[b] = nvy2; s.vz[b] = nvz2;
    s.flags[a] |= NB_F_COLLIDED;
    s.flags[b] |= NB_F_COLLIDED;
    atomicAdd(&s.ctr->n_resolved, 1ull);
}

// One cluster of CTAs (see below).  Event e of the concatenated per-rank segments lives at
// pairs_all[rank*seg_stride + k].
//
// Work-efficient wavefront: the events are first linked to their two bodies (an unsorted CSR built
// with counting atomics), every body publishes the largest key of its list (head[]), and the events
// that head both of their bodies form the first wavefront.  A round resolves the wavefront in
// parallel, recomputes the heads of just the bodies it touched (a scan of their own lists) and
// tests just those new heads for readiness — the cost of a round follows the width of the wavefront,
// not the length of the event list (the first version rescanned the whole list three times per round:
// 27.5 k events in 359 rounds, Sim3 geometry with 3001 bodies, took 22.6 ms).
__device__ __forceinline__ unsigned long long ev_key(int2 pr)
{
    return ((unsigned long long)(unsigned)pr.x << 32) | (unsigned)(pr.y & EV_INDEX_MASK);
}

// Round 2: the kernel runs as ONE THREAD-BLOCK CLUSTER of up to 8 CTAs (8 SMs).  A wavefront wider than a CTA (C3
// with 4x radii: ~900 events per round, 11.8 k events per cycle) used to be bound by one SM's FP64 / libm rate
// (480 us of an 11 ms cycle on one GPU, a third of the cycle on eight, where K3 is replicated); with the cluster
// the independent events of a round are spread over 4096 threads.  The round barrier is the hardware cluster
// barrier (barrier.cluster, release/acquire at cluster scope: what a CTA wrote to global memory before it is
// visible to every CTA after it), the five scheduling counters live in global memory (ctl[]).  Short lists
// (fewer events than one CTA has threads) run on the first CTA alone with __syncthreads(), the others leave at
// once: the small cycles keep their latency.  Which events share a round, and therefore every result bit, does not
// depend on how many CTAs take part.
enum { CTL_N_ACTIVE = 0, CTL_CURSOR, CTL_Q_BEGIN, CTL_Q_END, CTL_Q_TAIL, CTL_WORDS };

template <int RES_THREADS>
__global__ void __launch_bounds__(RES_THREADS) k_resolve(const __grid_constant__ StepParams p)
{
    __shared__ int seg_start[MAX_RANKS + 1];
    __shared__ int overflow;
    const DevState &s = p.s;
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int tid_local = threadIdx.x;
    volatile int *ctl = s.rs_ctl;
    if (tid_local == 0) {
        int acc = 0, ov = s.ctr->overflow;
        for (int r = 0; r < p.nranks; ++r) {
            seg_start[r] = acc;
            const unsigned long long cu = p.nranks == 1 ? s.ctr->n_pairs : s.pair_counts[r];
            int c = (int)(cu > (unsigned long long)p.seg_cap ? (unsigned long long)p.seg_cap : cu);
            if (cu > (unsigned long long)p.seg_cap) ov = 1;
            acc += c;
        }
        seg_start[p.nranks] = acc;
        overflow = ov;
        if (cta == 0) {
            s.ctr->total_pairs = acc;
            if (ov) s.ctr->overflow = 1;
            for (int k = 0; k < CTL_WORDS; ++k) ctl[k] = 0;
        }
    }
    __syncthreads();
    const int total = seg_start[p.nranks];
    if (overflow || total == 0) return;   // uniform over the cluster: every CTA computed the same numbers

    // ---- short lists (most cycles of a sparse cloud: a handful of events): the whole schedule lives in shared
    //      memory of the first CTA.  Same rule as below — an event is ready when no pending event with a larger key
    //      shares one of its bodies — so the same events share a round and `rounds` comes out the same; what is
    //      saved is the dozen dependent trips to global memory that build and walk the per-body lists (18 us of a
    //      40 us cycle at 1,000 bodies).  GROUP threads test one event's readiness, the group's first thread
    //      resolves it; up to 4 events per warp, so the libm chains of a round run side by side.
    constexpr int FAST_MAX = 64, GROUP = RES_THREADS / FAST_MAX;
    static_assert(GROUP >= 1 && GROUP <= 32 && (GROUP & (GROUP - 1)) == 0, "threads per event");
    if (p.res_fast && total <= FAST_MAX) {
        if (cta != 0) return;
        __shared__ int2 f_ev[FAST_MAX];
        __shared__ unsigned long long f_key[FAST_MAX];
        __shared__ int f_pend[FAST_MAX];
        __shared__ ElasticGeo f_geo[FAST_MAX];
        const bool report = (p.opts & (NB_STEP_NO_RESOLVE | NB_STEP_NO_INTEGRATE)) != 0;
        if (tid_local < total) {
            int r = 0;
            while (tid_local >= seg_start[r + 1]) ++r;
            const int2 pr = s.pairs_all[(long long)r * p.seg_stride + (tid_local - seg_start[r])];
            f_ev[tid_local] = pr;
            f_key[tid_local] = ev_key(pr);
            f_pend[tid_local] = 1;
            if (pr.y & EV_SUBSUME_BIT) {
                atomicAdd(&s.ctr->n_sub_events, 1ull);
                if (report) resolve_subsume(p, pr.x, pr.y & EV_INDEX_MASK, false);
            } else if (!report) {
                f_geo[tid_local] = elastic_geometry(s, pr.x, pr.y & EV_INDEX_MASK);  // positions are fixed during ProcessMods
            }
        }
        if (report) return;
        __syncthreads();
        const int e = tid_local / GROUP, sub = tid_local % GROUP;
        bool pending = e < total;
        const int2 pr = pending ? f_ev[e] : make_int2(-1, -1);
        const int bi = pr.x, bj = pr.y & EV_INDEX_MASK;
        const unsigned long long key = pending ? f_key[e] : 0ull;
        int rounds = 0;
        for (;;) {
            int blocked = 0;
            if (pending) {
                for (int k = sub; k < total; k += GROUP) {
                    if (!f_pend[k] || !(f_key[k] > key)) continue;
                    const int2 q = f_ev[k];
                    const int qj = q.y & EV_INDEX_MASK;
                    if (q.x == bi || q.x == bj || qj == bi || qj == bj) blocked = 1;
                }
            }
#pragma unroll
            for (int m = GROUP / 2; m > 0; m >>= 1) blocked |= __shfl_xor_sync(0xffffffffu, blocked, m);
            const bool ready = pending && !blocked;
            if (!__syncthreads_or(ready)) break;  // (also: every read of f_pend precedes this round's writes)
            if (ready) {
                if (sub == 0) {
                    if (pr.y & EV_SUBSUME_BIT) resolve_subsume(p, bi, bj, true);
                    else resolve_one(p, bi, bj, f_geo[e]);
                    f_pend[e] = 0;
                }
                pending = false;
            }
            ++rounds;
            __syncthreads();  // what the round wrote (bodies, f_pend) is visible to the next one
        }
        if (tid_local == 0) s.ctr->rounds = rounds;
        return;
    }

    // how many CTAs of the cluster work on this list
    const int nct = total >= RES_THREADS ? (int)cluster.num_blocks() : 1;
    if (cta >= nct) return;               // before any barrier: an exited CTA counts as arrived
    const int tid = cta * RES_THREADS + tid_local;
    const int nthreads = nct * RES_THREADS;
    auto sync_all = [&]() {
        if (nct > 1) cluster.sync();
        else __syncthreads();
    };
    sync_all();   // ctl[] is zero for everybody

    auto slot = [&](int e) -> long long {
        int r = 0;
        while (e >= seg_start[r + 1]) ++r;
        return (long long)r * p.seg_stride + (e - seg_start[r]);
    };

    // subsume events in the list; when nothing is resolved in this step they are only reported
    const bool report_only = (p.opts & (NB_STEP_NO_RESOLVE | NB_STEP_NO_INTEGRATE)) != 0;
    {
        int subs = 0;
        for (int e = tid; e < total; e += nthreads) {
            const int2 pr = s.pairs_all[slot(e)];
            if (pr.y & EV_SUBSUME_BIT) {
                ++subs;
                if (report_only) resolve_subsume(p, pr.x, pr.y & EV_INDEX_MASK, false);
            }
        }
        if (subs) atomicAdd(&s.ctr->n_sub_events, (unsigned long long)subs);
    }
    if (report_only) return;

    // ---- link the events to their bodies: count, place, fill (adj_cnt / adj_off are zero between steps).
    //      A body's slice holds the keys of its events (rs_lkey) and their indices (rs_list); a resolved
    //      event is struck out of both of its slices (key 0), so finding a body's next head is one pass
    //      over contiguous keys.
    for (int e = tid; e < total; e += nthreads) {
        const int2 pr = s.pairs_all[slot(e)];
        s.rs_ev[e] = pr;
        s.rs_state[e] = 0;
        const int i = pr.x, j = pr.y & EV_INDEX_MASK;
        if (!(pr.y & EV_SUBSUME_BIT)) s.rs_geo[e] = elastic_geometry(s, i, j);  // positions are fixed during ProcessMods
        if (atomicAdd(&s.adj_cnt[i], 1) == 0) s.rs_active[atomicAdd((int *)&ctl[CTL_N_ACTIVE], 1)] = i;
        if (atomicAdd(&s.adj_cnt[j], 1) == 0) s.rs_active[atomicAdd((int *)&ctl[CTL_N_ACTIVE], 1)] = j;
    }
    sync_all();
    const int na = ctl[CTL_N_ACTIVE];
    for (int a = tid; a < na; a += nthreads) {
        const int b = s.rs_active[a];
        s.adj_off[b] = atomicAdd((int *)&ctl[CTL_CURSOR], s.adj_cnt[b]);
        s.adj_cnt[b] = 0;  // refilled below
    }
    sync_all();
    for (int e = tid; e < total; e += nthreads) {
        const int2 pr = s.rs_ev[e];
        const int i = pr.x, j = pr.y & EV_INDEX_MASK;
        const unsigned long long key = ev_key(pr);
        const int pi = s.adj_off[i] + atomicAdd(&s.adj_cnt[i], 1);
        const int pj = s.adj_off[j] + atomicAdd(&s.adj_cnt[j], 1);
        s.rs_list[pi] = e; s.rs_lkey[pi] = key;
        s.rs_list[pj] = e; s.rs_lkey[pj] = key;
        s.rs_pos[e] = make_int2(pi, pj);
    }
    sync_all();

    // largest pending key of body b's slice and the event that holds it (-1: none left).  Every writer is a
    // thread of this cluster and a barrier lies in between.  Eight independent loads per trip keep the pass at
    // one or two memory latencies for usual slice lengths.
    auto scan_head = [&](int b, int &arg) -> unsigned long long {
        const int off = s.adj_off[b], cnt = s.adj_cnt[b];
        const unsigned long long *keys = s.rs_lkey + off;
        unsigned long long best = 0ull;
        int at = -1;
        for (int k0 = 0; k0 < cnt; k0 += 8) {
            unsigned long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = k0 + u < cnt ? keys[k0 + u] : 0ull;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (v[u] > best) { best = v[u]; at = k0 + u; }
        }
        arg = at >= 0 ? s.rs_list[off + at] : -1;
        return best;
    };
    // does the event with this key head both of its bodies?  (the bodies are in the key)
    auto heads_both = [&](unsigned long long key) -> bool {
        return s.head[(int)(key >> 32)] == key && s.head[(int)(key & 0xFFFFFFFFu)] == key;
    };

    // ---- first wavefront
    for (int a = tid; a < na; a += nthreads) {
        const int b = s.rs_active[a];
        int arg;
        s.head[b] = scan_head(b, arg);
    }
    sync_all();
    for (int e = tid; e < total; e += nthreads) {
        if (heads_both(ev_key(s.rs_ev[e]))) {
            s.rs_state[e] = 1;
            s.rs_queue[atomicAdd((int *)&ctl[CTL_Q_TAIL], 1)] = e;
        }
    }
    sync_all();
    if (tid == 0) ctl[CTL_Q_END] = ctl[CTL_Q_TAIL];
    sync_all();

    // ---- rounds: every event enters the queue exactly once, in wavefront order
    int rounds = 0;
    while (ctl[CTL_Q_BEGIN] < ctl[CTL_Q_END]) {
        const int qb = ctl[CTL_Q_BEGIN], qe = ctl[CTL_Q_END];
        // 1. the events of a wavefront touch disjoint bodies: resolve them in parallel, strike them out
        for (int f = qb + tid; f < qe; f += nthreads) {
            const int e = s.rs_queue[f];
            const int2 pr = s.rs_ev[e];
            const int2 pos = s.rs_pos[e];
            if (pr.y & EV_SUBSUME_BIT) resolve_subsume(p, pr.x, pr.y & EV_INDEX_MASK, true);
            else resolve_one(p, pr.x, pr.y & EV_INDEX_MASK, s.rs_geo[e]);
            s.rs_lkey[pos.x] = 0ull;
            s.rs_lkey[pos.y] = 0ull;
        }
        sync_all();
        // 2. new heads for the bodies of the resolved events (two tasks per event); a thread keeps its
        //    first task's candidate in registers, further tasks (wavefronts wider than the cluster) go
        //    through scratch
        const int tasks = 2 * (qe - qb);
        int c0 = -1;
        unsigned long long ck0 = 0ull;
        for (int t = tid; t < tasks; t += nthreads) {
            const int2 pr = s.rs_ev[s.rs_queue[qb + (t >> 1)]];
            const int b = (t & 1) ? (pr.y & EV_INDEX_MASK) : pr.x;
            int arg;
            const unsigned long long hk = scan_head(b, arg);
            s.head[b] = hk;
            if (t == tid) { c0 = arg; ck0 = hk; }
            else { s.rs_cand[2 * qb + t] = arg; s.rs_candkey[2 * qb + t] = hk; }
        }
        sync_all();
        // 3. a new head that now heads both of its bodies joins the next wavefront (once)
        for (int t = tid; t < tasks; t += nthreads) {
            const int c = t == tid ? c0 : s.rs_cand[2 * qb + t];
            if (c < 0) continue;
            const unsigned long long ck = t == tid ? ck0 : s.rs_candkey[2 * qb + t];
            if (heads_both(ck) && atomicCAS(&s.rs_state[c], 0, 1) == 0)
                s.rs_queue[atomicAdd((int *)&ctl[CTL_Q_TAIL], 1)] = c;
        }
        ++rounds;
        sync_all();
        if (tid == 0) { ctl[CTL_Q_BEGIN] = qe; ctl[CTL_Q_END] = ctl[CTL_Q_TAIL]; }
        sync_all();
    }

    // ---- leave the per-body scratch zeroed for the next step
    for (int a = tid; a < na; a += nthreads) {
        const int b = s.rs_active[a];
        s.head[b] = 0ull;
        s.adj_cnt[b] = 0;
        s.adj_off[b] = 0;
    }
    if (tid == 0) s.ctr->rounds = rounds;
}

// Pair exchange fused over peer memory: every rank stores its list (and count) into the segment
// [rank] of every rank's gathered buffer — no collective, no host round trip for the counts.
__global__ void __launch_bounds__(256) k_push_pairs(const __grid_constant__ StepParams p)
{
    const PeerTable &pt = *p.peers;
    const unsigned long long cnt_raw = p.s.ctr->n_pairs;
    const long long cnt = (long long)(cnt_raw > (unsigned long long)p.seg_cap ? (unsigned long long)p.seg_cap : cnt_raw);
    const long long par = (long long)(p.step_id & 1ull);
    const long long seg = (par * p.nranks + p.rank) * p.seg_cap;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < cnt;
         k += (long long)gridDim.x * blockDim.x) {
        const int2 v = p.s.pairs[k];
        for (int q = 0; q < p.nranks; ++q) pt.pairs_all[q][seg + k] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < p.nranks)
        pt.pair_counts[threadIdx.x][par * MAX_RANKS + p.rank] = cnt_raw;  // raw: overflow stays visible to all
}

int launch_push_pairs(const StepParams &p, cudaStream_t st)
{
    k_push_pairs<<<64, 256, 0, st>>>(p);
    return 1;
}

template <int T>
static void launch_resolve_cluster(const StepParams &p, cudaStream_t st, int cluster)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cluster);
    cfg.blockDim = dim3(T);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_resolve<T>, p);
}

int launch_resolve(const StepParams &p, cudaStream_t st)
{
    static const int threads = [] {  // development override
        const char *e = getenv("NB_RES_THREADS");
        return e ? atoi(e) : RES_THREADS_DEFAULT;
    }();
    const int cluster = p.res_cluster < 1 ? 1 : (p.res_cluster > 8 ? 8 : p.res_cluster);
    switch (threads) {
        case 1024: launch_resolve_cluster<1024>(p, st, cluster); break;
        case 256: launch_resolve_cluster<256>(p, st, cluster); break;
        default: launch_resolve_cluster<RES_THREADS_DEFAULT>(p, st, cluster); break;
    }
    return 1;
}

}  // namespace nb
