// nb_force.cu — K0 (prep) and K1 (tiled all-pairs force + collision detection).
//
// Replaces Body.Compute (cmd/body/body.go:148-187) with calcForceFrom (:214-225)
// and Collided (:192-208) for every body of the local i-shard, i.e. the work the
// reference fans out over goroutines (cmd/runner/workpool.go:103-110).
//
// K1 layout: one CTA = NT threads x R register-blocked i-bodies, one j-chunk.
// j-tiles (TJ = 64 / 256 / 512 bodies of jx,jy,jz,jm) are staged into shared memory with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), double buffered, and
// read two bodies at a time with broadcast LDS.128.
//
// Fast pass (branch-free), per pair 16 FP64-pipe instructions:
//   3 DADD (dx,dy,dz)  1 DMUL + 2 DFMA (d2)
//   MUFU.RSQ64H seed (XU pipe) + 7 DMUL/DFMA (cubic refinement of mj*d^-3)
//   3 DFMA accumulate
// The overlap predicate of the reference (dist > r_i + r_j, body.go:219) is screened with the
// running minimum of hi(d2) per (body, tile) against a conservative integer threshold (one
// VIMNMX3 per two pairs, no branch, no select).  The fast pass is speculative: it accumulates a
// tile's contribution in per-tile temporaries; if the minimum stayed above the threshold the
// temporaries are committed, otherwise (rare) the tile is redone for that body: unscreened groups
// {(i,j),(i,j+1)} with the same fast formula, screened groups through exact_pair, which restates
// the reference's unfused arithmetic bit for bit (sqrt_rn(fl(fl(dx*dx+dy*dy)+dz*dz)) vs
// fl(r_i+r_j)) and emits collision events.  Which path a (body, tile) takes depends on the bodies
// only — never on the launch shape — so results are bit-identical for every R / grid / rank count.
//
// Measured issue model on B200 (tools/issue_probe.py): an FP64 instruction holds the
// SMSP for max(2, #distinct 64-bit register operands) cycles and every other instruction
// costs about one more, so the loop is written to keep register reads and non-FP64
// instructions per pair minimal.
//
// Uniform-mass tiles: K0 notes, per j-tile, whether every body of the tile is live with exactly the
// same mass (the generated sims and clouds of the reference give whole clusters one mass).  For
// such a tile the mass is the same factor in every term of the tile's sum, so the fast pass
// accumulates sum_j d_ij^-3 (x_j - x_i) — 15 FP64 instructions per pair and no LDS of the masses —
// and the tile's mass multiplies the tile sum once at the commit.  The choice is made per j-chunk
// (every tile of the chunk uniform, each with its own mass) and the two passes are two
// instantiations of the kernel launched over the same grid: a CTA whose chunk is of the other kind
// exits at once.  (Both passes in one kernel cost the hot loops their MOV-free register pairing.)
// A property of the bodies only, like the screen: results stay independent of launch shape and
// rank count.
//
// This file is compiled with -fmad=false: every FMA below is explicit.
#if defined(NB_EXP_CONV_PROBE) && NB_EXP_CONV_PROBE
#include <cstdio>
#endif
#include "nb_internal.cuh"

namespace nb {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NB_DONE;\n"
        "bra NB_WAIT;\n"
        "NB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// MUFU.RSQ64H: ~2^-22 relative seed of 1/sqrt(x) from the high word of x
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// Seed whose low word is `lo` instead of the canonical 0.  MUFU.RSQ64H writes only the high word;
// the low word of the seed is irrelevant to the result's accuracy (it perturbs y0 by < 2^-20, and
// the cubic refinement absorbs e up to ~2^-19 with an O(e^3) ~ 1e-17 residue), so pairing the MUFU
// result with a long-lived register spares one MOV per pair.  `lo` is always 0 at run time (an
// opaque zero, so the compiler cannot rematerialise it) which keeps results reproducible.
__device__ __forceinline__ double rsqrt_seed_lo(double x, unsigned lo)
{
    double y;
    asm("{\n"
        ".reg .b32 tl, th;\n"
        ".reg .f64 t;\n"
        "rsqrt.approx.ftz.f64 t, %1;\n"
        "mov.b64 {tl, th}, t;\n"
        "mov.b64 %0, {%2, th};\n"
        "}\n"
        : "=d"(y)
        : "d"(x), "r"(lo));
    return y;
}

// Compile-time knobs of the hot loop (tools/k1_variants.py ranks builds offline by instruction counts,
// tools/k1_hw_variants.py times them on the GPU; the defaults are the production code).  Round 2 measured ten
// more source variants that are no longer in this file — accumulate order, polynomial form, where the running
// minimum is updated, three loop forms, a padding instruction that moves ptxas' yield hints off the accumulate
// triples, a stage hand-back through "empty" mbarriers — all of them slower or equal
// (profiles/r2_k1_variants.txt; the code is in the history up to commit 61898ce).
//   NB_EXP_KREG   1: the constant 15/8 lives in a register pair built from an opaque zero, so ptxas cannot
//                 rematerialise it with two IMAD.MOV per loop trip: one instruction less per 8 pairs, measured
//                 -0.55 % on the uniform-mass launch (60.03 -> 59.70 ms at n = 256 k)
//   NB_EXP_KZ_GEN / NB_EXP_KZ_UNI   1: the constant gets an opaque zero of its own instead of sharing the first
//                 seed's (one IMAD.MOV less per trip).  Measured at n = 256 k: the per-body-mass loop gains 0.4 %
//                 with it, the uniform-mass loop LOSES 0.7 % — one instruction less, slower: at this level the
//                 outcome is decided by how ptxas' schedule falls, not by counts.
//   NB_EXP_UNR4   unroll of the j-group loop of the production shape (R = 4); 2 is rated 2.5 % faster by ptxas'
//                 static schedule and measures 2 % slower (255 registers, nine extra MOVs)
#ifndef NB_EXP_KREG
#define NB_EXP_KREG 1
#endif
#ifndef NB_EXP_KZ_GEN
#define NB_EXP_KZ_GEN 1
#endif
#ifndef NB_EXP_KZ_UNI
#define NB_EXP_KZ_UNI 0
#endif
#ifndef NB_EXP_CONV_PROBE
#define NB_EXP_CONV_PROBE 0
#endif
#ifndef NB_EXP_UNR4
#define NB_EXP_UNR4 1
#endif

// w = mj * d2^(-3/2) from the seed y0: with e = 1 - d2*y0^2 (|e| <~ 2^-21),
// d2^(-3/2) = y0^3 (1-e)^(-3/2) = y0^3 (1 + e(3/2 + 15/8 e) + O(e^3)).  7 FP64 ops, each reading at
// most two distinct registers.
__device__ __forceinline__ double w_from_seed(double y0, double d2, double mj, double k1875 = 1.875)
{
    const double u = __dmul_rn(y0, y0);
    const double e = __fma_rn(-d2, u, 1.0);
    const double pp = __fma_rn(k1875, e, 1.5);
    const double t = __dmul_rn(mj, y0);
    const double tu = __dmul_rn(t, u);
    const double c = __fma_rn(e, pp, 1.0);
    return __dmul_rn(tu, c);
}

// The same without the mass: d2^(-3/2) (6 FP64 ops), for tiles whose bodies all have one mass.
__device__ __forceinline__ double w_from_seed_uni(double y0, double d2, double k1875 = 1.875)
{
    const double u = __dmul_rn(y0, y0);
    const double e = __fma_rn(-d2, u, 1.0);
    const double pp = __fma_rn(k1875, e, 1.5);
    const double s = __dmul_rn(y0, u);
    const double c = __fma_rn(e, pp, 1.0);
    return __dmul_rn(s, c);
}

// ---------------------------------------------------------------- K0: prep
// Builds the j-stream (jx,jy,jz,jm) K1 sweeps: jm = Exists && !fragmenting ? mass : 0 (the
// j-filter of body.go:162-165 folded into the mass); bodies that do not exist — and the tail of
// the last tile — are parked massless at a far, finite position so they are never screened.
// Also the per-tile max radius of live bodies.
template <int TJ>
__global__ void __launch_bounds__(TJ) k_prep(StepParams p)
{
    const long long j = (long long)blockIdx.x * TJ + threadIdx.x;
    // the step counters start from zero (nothing else runs on the stream while K0 does)
    if (blockIdx.x == 0 && threadIdx.x < sizeof(Counters) / sizeof(unsigned))
        reinterpret_cast<unsigned *>(p.s.ctr)[threadIdx.x] = 0u;
    double r = 0.0, x = 1e150, y = 1e150, z = 1e150, m = 0.0;
    bool live = false, dead_here = false;
    if (j < p.n) {
        const unsigned fl = p.s.flags[j];
        // calcForceFrom multiplies by the mass the body has during Compute; a subsume handled in
        // ProcessMods changes Body.Mass before Update divides by it (body.go:120-122,241)
        p.s.m0[j] = p.s.mass[j];
        // ... and whether Compute runs at all is decided before ProcessMods may clear Exists (subsume) or
        // set `fragmenting` (body.go:149-155): Update then applies the force of THIS cycle
        p.s.computes0[j] = (fl & NB_F_EXISTS) && !(fl & NB_F_FRAGMENTING);
        const double px = p.s.x[j], py = p.s.y[j], pz = p.s.z[j];
        // a live body at a non-finite (or absurd, |coord| >= 1e150: d2 would overflow) position is
        // inert as a j-body (the reference would poison every force with NaN); as an i-body it still
        // receives NaN and is culled by K4
        const bool finite = fabs(px) < 1e150 && fabs(py) < 1e150 && fabs(pz) < 1e150;
        live = (fl & NB_F_EXISTS) != 0 && finite;
        // a body that does not exist but is still in the array (SetNotExists at the cycle top, removed by
        // the next Cycle) is skipped by the force sweep yet visited by the collision sweep
        // (body.go:172-186 has no Exists filter): K1 looks at those few bodies separately (dead_j_sweep)
        dead_here = !(fl & NB_F_EXISTS) && finite;
        if (live) {
            x = px; y = py; z = pz;
            r = p.s.radius[j];
            if (isnan(r)) r = INFINITY;  // NaN radius: the reference's predicate is never true; screen its whole tile
            if (!(fl & NB_F_FRAGMENTING)) m = p.s.mass[j];
        }
    }
    p.s.jx[j] = x;
    p.s.jy[j] = y;
    p.s.jz[j] = z;
    p.s.jm[j] = m;
    // block max of r (NaN radii are ignored by fmax); block min / max of the live masses
    __shared__ double red[TJ / 32], red_lo[TJ / 32], red_hi[TJ / 32];
    // A tile is "uniform" if every LIVE body in it has the same positive finite mass.  Slots without a
    // live body (bodies that do not exist, the tail of the last tile) are parked at 1e150 and add an
    // exact 0 to the mass-free sums of K1's uniform pass (y0^3 underflows to 0), so they match any
    // mass; a fragmenting body stays at its position with m = 0 (it is still a collision partner) and
    // breaks uniformity.
    const int odd = __syncthreads_or(live && !(m > 0.0 && m < INFINITY));
    const int any_dead = __syncthreads_or(dead_here);
    double mlo = live ? m : INFINITY, mhi = live ? m : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
        mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = r;
        red_lo[threadIdx.x >> 5] = mlo;
        red_hi[threadIdx.x >> 5] = mhi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = red[0], lo = red_lo[0], hi = red_hi[0];
#pragma unroll
        for (int w = 1; w < TJ / 32; ++w) {
            mx = fmax(mx, red[w]);
            lo = fmin(lo, red_lo[w]);
            hi = fmax(hi, red_hi[w]);
        }
        p.s.tile_rmax[blockIdx.x] = mx;
        double mu = 0.0;
        if (p.uniform_tiles && !odd) {
            if (lo == hi) mu = lo;            // every live body has this mass
            else if (hi == 0.0) mu = 1.0;     // no live body at all: the tile's sums are exact zeros
        }
        p.s.tile_muni[blockIdx.x] = mu;
        p.s.tile_dead[blockIdx.x] = any_dead ? 1 : 0;
    }
}

int launch_prep(const StepParams &p, cudaStream_t st)
{
    if (p.n_tiles <= 0) return 0;
    if (p.tj == TJ_SMALL) k_prep<TJ_SMALL><<<p.n_tiles, TJ_SMALL, 0, st>>>(p);
    else if (p.tj == TJ_HUGE) k_prep<TJ_HUGE><<<p.n_tiles, TJ_HUGE, 0, st>>>(p);
    else k_prep<TJ_LARGE><<<p.n_tiles, TJ_LARGE, 0, st>>>(p);
    return 1;
}

// ---------------------------------------------------------------- K1: exact path
__device__ __forceinline__ bool elastic_or_fragment(unsigned b) { return b == NB_ELASTIC || b == NB_FRAGMENT; }

// The j-side facts an exact pair needs from global memory (radius, Exists, behaviour): loaded by the
// caller for both pairs of a group at once, so that a trip of the redo costs one memory latency, not a
// chain of them.
struct JFacts {
    double r;
    unsigned flags, behavior;
};
__device__ __forceinline__ JFacts load_jfacts(const StepParams &p, long long j)
{
    JFacts f;  // j < n_tiles * tj <= cap_pad: in bounds even past n
    f.r = p.s.radius[j];
    f.flags = p.s.flags[j];
    f.behavior = p.s.behavior[j];
    return f;
}

__device__ __noinline__ void emit_event(const StepParams &p, long long i, long long j, double dist, double ri,
                                        double rj, unsigned bi, unsigned bj)
{
    if (elastic_or_fragment(bi) && elastic_or_fragment(bj)) {
        // newCollision(b, otherBody), body.go:175-177
        const unsigned long long k = atomicAdd(&p.s.ctr->n_pairs, 1ull);
        if (k < (unsigned long long)p.seg_cap)
            p.s.pairs[k] = make_int2((int)i, (int)j);
        else
            p.s.ctr->overflow = 1;
    } else if (bi == NB_SUBSUME || bj == NB_SUBSUME) {
        // body.go:178-184: the larger radius subsumes, only if the centre is inside it.  The event
        // keeps its arrival position (i,j) in the one event list (K3 resolves collisions and
        // subsumes in the reference's serial order); who subsumes whom follows from the radii.
        if ((ri > rj && dist <= ri) || (rj > ri && dist <= rj)) {
            const unsigned long long k = atomicAdd(&p.s.ctr->n_pairs, 1ull);
            if (k < (unsigned long long)p.seg_cap)
                p.s.pairs[k] = make_int2((int)i, (int)j | EV_SUBSUME_BIT);
            else
                p.s.ctr->overflow = 1;
        }
    }
}

// Exact restatement of calcForceFrom / Collided for one screened pair. Returns the
// weight w = mj / dist^3 to accumulate (0 when the pair exerts no force).
__device__ __noinline__ double exact_pair(const StepParams &p, long long i, long long j, bool alive_i, double xi,
                                          double yi, double zi, double ri, unsigned bi, double xj, double yj,
                                          double zj, double mj, JFacts fj)
{
    if (!alive_i || j >= p.n || j == i) return 0.0;
    const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi), dz = __dsub_rn(zj, zi);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double dist = __dsqrt_rn(d2);
    const double rj = fj.r;
    const double s = __dadd_rn(ri, rj);
    if (dist > s) {
        if (mj == 0.0) return 0.0;
        return __ddiv_rn(__ddiv_rn(mj, __dmul_rn(dist, dist)), dist);
    }
    if (dist <= s) {
        if ((p.opts & NB_STEP_COLLISIONS) && (fj.flags & NB_F_EXISTS)) emit_event(p, i, j, dist, ri, rj, bi, fj.behavior);
    }
    return 0.0;
}

// The collision sweep of Body.Compute has no Exists filter (body.go:172-186): a body that was set not
// to exist at the cycle top (RemoveBodies / mod-body exists=false, computation-runner.go:176-216) but is
// still in the array until the next Cycle is visited as `otherBody`.  Collision events with such a body
// are no-ops (ResolveCollision's Exists gate, body.go:249-251) and are not queued; subsume events are
// NOT gated (ResolveSubsume, body.go:228-244): a dead body with the larger radius still swallows a live
// one, a dead body inside a live subsumer adds its mass.  Those bodies are parked far away in the
// j-stream, so the tiled sweep never sees them; K0 flags the tiles that hold one and the CTAs of chunk 0
// walk the few flagged tiles here (rare, and a handful of bodies when it happens).  One warp per call;
// the body's own facts are re-read from global memory so that the hot kernel's registers stay its own.
__device__ __noinline__ void dead_j_sweep(const StepParams &p, long long ibase, int stride, int R)
{
    const int lane = threadIdx.x & 31;
    int any = 0;
    for (int t = lane; t < p.n_tiles; t += 32) any |= p.s.tile_dead[t];
    if (!__any_sync(0xffffffffu, any)) return;
    for (int r = 0; r < R; ++r) {
        const long long i = ibase + (long long)r * stride + threadIdx.x;
        unsigned fl = 0;
        if (i < p.i1) fl = p.s.flags[i];
        const bool alive = (fl & NB_F_EXISTS) && !(fl & NB_F_FRAGMENTING);  // body.go:149-155
        const double xi = alive ? p.s.x[i] : 0.0, yi = alive ? p.s.y[i] : 0.0, zi = alive ? p.s.z[i] : 0.0;
        const double ri = alive ? p.s.radius[i] : 0.0;
        const unsigned bi = alive ? p.s.behavior[i] : 0u;
        for (int t = 0; t < p.n_tiles; ++t) {
            if (!p.s.tile_dead[t]) continue;  // warp-uniform
            const long long j0 = (long long)t * p.tj;
            for (int jj = 0; jj < p.tj && j0 + jj < p.n; ++jj) {
                const long long j = j0 + jj;
                if (p.s.flags[j] & NB_F_EXISTS) continue;  // warp-uniform
                const double xj = p.s.x[j], yj = p.s.y[j], zj = p.s.z[j];
                if (!(fabs(xj) < 1e150 && fabs(yj) < 1e150 && fabs(zj) < 1e150)) continue;
                const double rj = p.s.radius[j];
                const unsigned bj = p.s.behavior[j];
                if (!alive || j == i) continue;
                if (elastic_or_fragment(bi) && elastic_or_fragment(bj)) continue;  // a gated no-op in the reference
                if (!(bi == NB_SUBSUME || bj == NB_SUBSUME)) continue;
                // Collided, body.go:192-208 (unfused)
                const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi), dz = __dsub_rn(zj, zi);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const double dist = __dsqrt_rn(d2);
                const double s = __dadd_rn(ri, rj);
                if (dist > s) continue;
                if (dist <= s) emit_event(p, i, j, dist, ri, rj, bi, bj);
            }
        }
    }
}

// Conservative integer screen: a pair whose radii sum to sr can only overlap (or be degenerate) if
// hi(d2) < hi(sr^2 (1+2^-18)) + 2.  inf / NaN radii screen every finite pair.
__device__ __forceinline__ unsigned screen_threshold(double sr)
{
    const double t2 = __dmul_rn(__dmul_rn(sr, sr), 1.0 + 1.0 / 262144.0);
    unsigned v = (unsigned)__double2hiint(t2) + 2u;
    if (v > 0x7FF00000u) v = 0x7FF00000u;
    return v;
}

// ---------------------------------------------------------------- K1: fast pass over one tile
// Two j-bodies per iteration (LDS.128), R i-bodies per thread.  SELF: the tile holds bodies of this
// CTA; the pair (i,i) gets the seed 0 (hence w == 0 exactly) and is left out of the minimum.
// UNI: every body of the tile has the same mass; the sums are taken without it (the caller scales).
template <int R, int UNR, bool SELF, bool UNI, int TJ>
__device__ __forceinline__ void fast_tile(const double *sx, const double *sy, const double *sz, const double *sj,
                                          const double (&xi)[R], const double (&yi)[R], const double (&zi)[R],
                                          const unsigned (&zlo)[2 * R + 1], const int (&self_j)[R], double (&tx)[R],
                                          double (&ty)[R], double (&tz)[R], unsigned (&lo)[R])
{
#if NB_EXP_KREG
    // 15/8 with an opaque (always zero) low word: a value, not a literal, so it stays in a register pair
    const double k1875 = __hiloint2double(0x3ffe0000, (int)zlo[(UNI ? NB_EXP_KZ_UNI : NB_EXP_KZ_GEN) ? 2 * R : 0]);
#else
    const double k1875 = 1.875;
#endif
#pragma unroll(UNR)
    for (int jj = 0; jj < TJ; jj += 2) {
        const double2 vx = *reinterpret_cast<const double2 *>(sx + jj);
        const double2 vy = *reinterpret_cast<const double2 *>(sy + jj);
        const double2 vz = *reinterpret_cast<const double2 *>(sz + jj);
        double2 vm = make_double2(0.0, 0.0);
        if (!UNI) vm = *reinterpret_cast<const double2 *>(sj + jj);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double dxa = __dsub_rn(vx.x, xi[r]), dxb = __dsub_rn(vx.y, xi[r]);
            const double dya = __dsub_rn(vy.x, yi[r]), dyb = __dsub_rn(vy.y, yi[r]);
            const double dza = __dsub_rn(vz.x, zi[r]), dzb = __dsub_rn(vz.y, zi[r]);
            const double d2a = __fma_rn(dza, dza, __fma_rn(dya, dya, __dmul_rn(dxa, dxa)));
            const double d2b = __fma_rn(dzb, dzb, __fma_rn(dyb, dyb, __dmul_rn(dxb, dxb)));
            unsigned ha = (unsigned)__double2hiint(d2a), hb = (unsigned)__double2hiint(d2b);
            double ya = rsqrt_seed_lo(d2a, zlo[2 * r]);
            double yb = rsqrt_seed_lo(d2b, zlo[2 * r + 1]);
            if (SELF) {
                const bool sa = jj == self_j[r], sb = jj + 1 == self_j[r];
                ha = sa ? 0xFFFFFFFFu : ha;
                hb = sb ? 0xFFFFFFFFu : hb;
                ya = __hiloint2double(sa ? 0 : __double2hiint(ya), __double2loint(ya));
                yb = __hiloint2double(sb ? 0 : __double2hiint(yb), __double2loint(yb));
            }
            lo[r] = min(lo[r], min(ha, hb));
            const double wa = UNI ? w_from_seed_uni(ya, d2a, k1875) : w_from_seed(ya, d2a, vm.x, k1875);
            const double wb = UNI ? w_from_seed_uni(yb, d2b, k1875) : w_from_seed(yb, d2b, vm.y, k1875);
            tx[r] = __fma_rn(wa, dxa, tx[r]);
            ty[r] = __fma_rn(wa, dya, ty[r]);
            tz[r] = __fma_rn(wa, dza, tz[r]);
            tx[r] = __fma_rn(wb, dxb, tx[r]);
            ty[r] = __fma_rn(wb, dyb, ty[r]);
            tz[r] = __fma_rn(wb, dzb, tz[r]);
        }
    }
}

// ---------------------------------------------------------------- K1: kernel
// UNR = unroll of the j-group loop (each group is two j-bodies).
// MODE: FORCE_ALL = every chunk, per-body masses; FORCE_MIXED = the same, but only the chunks that
// hold a non-uniform tile; FORCE_UNI = only the chunks whose tiles are all uniform, masses hoisted.
enum { FORCE_ALL = 0, FORCE_MIXED = 1, FORCE_UNI = 2 };

template <int R, int NT, int MINB, int UNR, int TJ, int MODE>
__global__ void __launch_bounds__(NT, MINB) k_force(const __grid_constant__ StepParams p)
{
    static_assert(TJ % 2 == 0 && R <= 16, "tile of j-pairs");
    constexpr int NSTAGE = NB_EXP_NSTAGE ? NB_EXP_NSTAGE : stages_for(TJ);
    __shared__ __align__(128) double sm[NSTAGE][4][TJ];
    __shared__ __align__(8) uint64_t bar[NSTAGE];
    constexpr int LOOK = NSTAGE - 1;  // tiles in flight ahead of the one being consumed
    constexpr bool UNI = MODE == FORCE_UNI;

    const int tid = threadIdx.x;
    const long long ibase = p.i0 + (long long)blockIdx.x * (NT * R);
    const int chunk = blockIdx.y;
    const int t0 = chunk * p.tiles_per_chunk;
    int t1 = t0 + p.tiles_per_chunk;
    if (t1 > p.n_tiles) t1 = p.n_tiles;
    const int nt = t1 - t0;

    if (MODE != FORCE_ALL) {  // is this chunk mine?  (block-uniform: every thread sees the same answer)
        int mixed = 0;
        for (int t = tid; t < nt; t += NT) mixed |= !(p.s.tile_muni[t0 + t] > 0.0);
        const bool chunk_uni = __syncthreads_or(mixed) == 0;
        if (chunk_uni != UNI) return;
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only: stage tile t0+t of the j-stream
        const int s = t % NSTAGE;
        const long long j0 = (long long)(t0 + t) * TJ;
        mbar_expect_tx(&bar[s], 4u * TJ * sizeof(double));
        bulk_g2s(&sm[s][0][0], p.s.jx + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][1][0], p.s.jy + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][2][0], p.s.jz + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][3][0], p.s.jm + j0, TJ * sizeof(double), &bar[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < LOOK && t < nt; ++t) issue(t);
    }

    double xi[R], yi[R], zi[R], ri[R];
    double ax[R], ay[R], az[R];
    bool alive[R];
    unsigned zlo[2 * R + 1];  // opaque zeros: low words of the rsqrt seeds (see rsqrt_seed_lo) and of the constant 15/8
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) zlo[k] = __ldcg(&p.s.zeros[(tid + 32 * k) & 1023]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long i = ibase + (long long)r * NT + tid;
        unsigned fl = 0;
        if (i < p.i1) fl = p.s.flags[i];
        alive[r] = (fl & NB_F_EXISTS) && !(fl & NB_F_FRAGMENTING);  // body.go:149-155
        xi[r] = alive[r] ? p.s.x[i] : 0.0;
        yi[r] = alive[r] ? p.s.y[i] : 0.0;
        zi[r] = alive[r] ? p.s.z[i] : 0.0;
        ri[r] = alive[r] ? p.s.radius[i] : 0.0;
        ax[r] = ay[r] = az[r] = 0.0;
    }

    for (int t = 0; t < nt; ++t) {
        const int s = t % NSTAGE;
        if (tid == 0 && t + LOOK < nt) issue(t + LOOK);

        // Conservative per-body screen for this tile: a pair (i,j) can only overlap (or be
        // degenerate) if hi(d2) < thr_i, thr_i = hi((r_i + rmax_tile)^2 (1+2^-18)) + 2.
        unsigned thr[R], lo[R];
        double tx[R], ty[R], tz[R];  // this tile's contribution, committed only if nothing was screened
        const double rm = p.s.tile_rmax[t0 + t];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            thr[r] = 0u;
            lo[r] = 0xFFFFFFFFu;
            tx[r] = ty[r] = tz[r] = 0.0;
            if (alive[r]) thr[r] = screen_threshold(__dadd_rn(ri[r], rm));
        }

        mbar_wait(&bar[s], (unsigned)((t / NSTAGE) & 1));
        const double *sx = sm[s][0], *sy = sm[s][1], *sz = sm[s][2], *sj = sm[s][3];
        const long long jt0 = (long long)(t0 + t) * TJ;

        // ---- fast pass: branch-free and unmasked (speculative); only the running minimum of
        //      hi(d2) is kept per body.  A tile that contains bodies of this CTA uses the SELF
        //      variant, which masks the pair (i,i) (d2 = 0) so that it neither poisons the sums nor
        //      forces a redo.
        if (jt0 < ibase + (long long)NT * R && jt0 + TJ > ibase) {
            int self_j[R];
#pragma unroll
            for (int r = 0; r < R; ++r) self_j[r] = (int)(ibase + (long long)r * NT + tid - jt0);
            fast_tile<R, UNR, true, UNI, TJ>(sx, sy, sz, sj, xi, yi, zi, zlo, self_j, tx, ty, tz, lo);
        } else {
            int self_j[R];
#pragma unroll
            for (int r = 0; r < R; ++r) self_j[r] = -1;
            fast_tile<R, UNR, false, UNI, TJ>(sx, sy, sz, sj, xi, yi, zi, zlo, self_j, tx, ty, tz, lo);
        }

        // ---- commit, or (rare) redo the tile carefully for a body that saw a screened pair.
        //      The redo walks the tile in blocks of 64 j-bodies: first the unscreened groups {j,j+1}
        //      with the fast formula (branch-free; the screened ones are only noted in a bit mask), then
        //      the screened groups through exact_pair, one per trip — lanes of a warp whose screened
        //      pairs sit at different j's make those calls together instead of one after the other
        //      (dense clusters: the Sim3 geometry redoes most of its tiles).  Which path a body takes,
        //      and the order of its sums, depend on (i, tile) only.
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (lo[r] < thr[r]) {
                double cx = 0.0, cy = 0.0, cz = 0.0;
                const unsigned bi = p.s.behavior[ibase + (long long)r * NT + tid];
                const bool fine = !(rm <= __dmul_rn(2.0, ri[r]));  // also for inf (a NaN radius in the tile)
#pragma unroll 1
                for (int jb = 0; jb < TJ; jb += 64) {
                    unsigned mask = 0u;
#pragma unroll(R == 1 ? 4 : (R == 2 ? 2 : 1))
                    for (int q = 0; q < 32; ++q) {
                        const int jj = jb + 2 * q;
                        double dx[2], dy[2], dz[2], d2[2];
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            dx[g] = __dsub_rn(sx[jj + g], xi[r]);
                            dy[g] = __dsub_rn(sy[jj + g], yi[r]);
                            dz[g] = __dsub_rn(sz[jj + g], zi[r]);
                            d2[g] = __fma_rn(dz[g], dz[g], __fma_rn(dy[g], dy[g], __dmul_rn(dx[g], dx[g])));
                        }
                        // When the tile's largest radius dwarfs this body's (one big body — a sun — in the
                        // tile) the redo screens each pair with its own radius, so that the 63 tile-mates
                        // of the sun stay on the fast formula; the loads are warp-uniform.  Otherwise the
                        // tile-level threshold is as tight and cheaper.
                        bool screened;
                        if (fine)
                            screened = (unsigned)__double2hiint(d2[0]) <
                                           screen_threshold(__dadd_rn(ri[r], p.s.radius[jt0 + jj])) ||
                                       (unsigned)__double2hiint(d2[1]) <
                                           screen_threshold(__dadd_rn(ri[r], p.s.radius[jt0 + jj + 1]));
                        else
                            screened = min((unsigned)__double2hiint(d2[0]), (unsigned)__double2hiint(d2[1])) < thr[r];
                        mask |= (screened ? 1u : 0u) << q;
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            // a screened group's d2 may be 0: its garbage weight is masked, not branched on
                            const double w = screened ? 0.0 : w_from_seed(rsqrt_seed(d2[g]), d2[g], sj[jj + g]);
                            if (w != 0.0) {
                                cx = __fma_rn(w, dx[g], cx);
                                cy = __fma_rn(w, dy[g], cy);
                                cz = __fma_rn(w, dz[g], cz);
                            }
                        }
                    }
                    while (mask) {
                        const int jj = jb + 2 * (__ffs((int)mask) - 1);
                        mask &= mask - 1u;
                        const JFacts fj[2] = {load_jfacts(p, jt0 + jj), load_jfacts(p, jt0 + jj + 1)};
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const double w = exact_pair(p, ibase + (long long)r * NT + tid, jt0 + jj + g, alive[r], xi[r],
                                                        yi[r], zi[r], ri[r], bi, sx[jj + g], sy[jj + g],
                                                        sz[jj + g], sj[jj + g], fj[g]);
                            if (w != 0.0) {
                                cx = __fma_rn(w, __dsub_rn(sx[jj + g], xi[r]), cx);
                                cy = __fma_rn(w, __dsub_rn(sy[jj + g], yi[r]), cy);
                                cz = __fma_rn(w, __dsub_rn(sz[jj + g], zi[r]), cz);
                            }
                        }
                    }
                }
                tx[r] = cx; ty[r] = cy; tz[r] = cz;  // the redo always carries the masses
            } else if (UNI) {
                const double mu = p.s.tile_muni[t0 + t];  // the one mass of every body of this tile
                tx[r] = __dmul_rn(mu, tx[r]);
                ty[r] = __dmul_rn(mu, ty[r]);
                tz[r] = __dmul_rn(mu, tz[r]);
            }
            ax[r] = __dadd_rn(ax[r], tx[r]);
            ay[r] = __dadd_rn(ay[r], ty[r]);
            az[r] = __dadd_rn(az[r], tz[r]);
        }
        __syncthreads();  // every warp is done with stage s before it is refilled
#if NB_EXP_CONV_PROBE  // development: is the warp converged when it leaves the barrier after a redo?
        {
            const unsigned am = __activemask();
            bool redo = false;
#pragma unroll
            for (int r = 0; r < R; ++r) redo |= lo[r] < thr[r];
            const unsigned vote = __ballot_sync(am, redo);
            if (vote && (tid & 31) == __ffs(am) - 1)
                printf("CONV %s mask=%08x redo=%08x\n", am == 0xffffffffu ? "full" : "PARTIAL", am, vote);
        }
#endif
    }

    // one partial-sum slot per (chunk, body); G*m_i is applied by the integrate kernel
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long i = ibase + (long long)r * NT + tid;
        if (i < p.i1) {
            const long long o = (long long)chunk * p.n_pad_local + (i - p.i0);
            p.s.px[o] = ax[r];
            p.s.py[o] = ay[r];
            p.s.pz[o] = az[r];
        }
    }
    // bodies that do not exist but are still in the array: collision sweep only (see dead_j_sweep)
    if (chunk == 0 && (p.opts & NB_STEP_COLLISIONS)) dead_j_sweep(p, ibase, NT, R);
}

template <int R, int NT, int MINB, int UNR, bool SPLIT = false>
static int launch_force_t(const StepParams &p, cudaStream_t st)
{
    const long long n_local = p.i1 - p.i0;
    const long long per = (long long)NT * R;
    dim3 grid((unsigned)((n_local + per - 1) / per), (unsigned)p.n_chunks);
    if (p.tj == TJ_SMALL) {
        // small collections: one kernel, per-body masses (a cycle is a few launches' worth of latency)
        k_force<R, NT, MINB, UNR, TJ_SMALL, FORCE_ALL><<<grid, NT, 0, st>>>(p);
        return 1;
    }
    if (SPLIT && p.uniform_tiles) {
        // same grid twice: each CTA runs in the instantiation that matches its j-chunk and exits in the other
        if (p.tj == TJ_HUGE) {
            k_force<R, NT, MINB, UNR, TJ_HUGE, FORCE_UNI><<<grid, NT, 0, st>>>(p);
            k_force<R, NT, MINB, UNR, TJ_HUGE, FORCE_MIXED><<<grid, NT, 0, st>>>(p);
        } else {
            k_force<R, NT, MINB, UNR, TJ_LARGE, FORCE_UNI><<<grid, NT, 0, st>>>(p);
            k_force<R, NT, MINB, UNR, TJ_LARGE, FORCE_MIXED><<<grid, NT, 0, st>>>(p);
        }
        return 2;
    }
    if (p.tj == TJ_HUGE) k_force<R, NT, MINB, UNR, TJ_HUGE, FORCE_ALL><<<grid, NT, 0, st>>>(p);
    else k_force<R, NT, MINB, UNR, TJ_LARGE, FORCE_ALL><<<grid, NT, 0, st>>>(p);
    return 1;
}

// R is chosen from the local shard size only; it never changes the result bits
// (each body's j-order, chunking and screening are functions of the bodies alone).
int launch_force(const StepParams &p, cudaStream_t st, int force_R)
{
    const long long n_local = p.i1 - p.i0;
    if (n_local <= 0 || p.n_tiles <= 0) return 0;
    int R = force_R;
    if (R <= 0) {
        const long long ctas4 = ((n_local + 511) / 512) * p.n_chunks;
        const long long ctas2 = ((n_local + 255) / 256) * p.n_chunks;
        if (ctas4 >= 148 * 4 * 4) R = 4;
        else if (ctas2 >= 148 * 4 * 2) R = 2;
        else R = 1;
    }
    // Values above 9 (NB_FORCE_R) select alternative launch shapes for tools/kbench.py:
    // 1000*UNR + 100*MINB + 10*(NT==256) + R.  Production shapes: R in {4,2,1}, NT 128.
    switch (R) {
        case 4: return launch_force_t<4, 128, 1, NB_EXP_UNR4, true>(p, st);
        case 2: return launch_force_t<2, 128, 1, 1, true>(p, st);
        case 1: return launch_force_t<1, 128, 1, 2, true>(p, st);
        case 3: return launch_force_t<3, 128, 1, 1>(p, st);
        case 2004: return launch_force_t<4, 128, 1, 2>(p, st);
        case 2002: return launch_force_t<2, 128, 1, 2>(p, st);
        case 1012: return launch_force_t<2, 256, 1, 1>(p, st);
        case 1022: return launch_force_t<2, 64, 1, 1>(p, st);
        case 304: return launch_force_t<4, 128, 3, 1>(p, st);
        default: return launch_force_t<1, 128, 1, 2>(p, st);
    }
}

// ---------------------------------------------------------------- FP64 peak probe
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double *out)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

int launch_fp64_peak(int iters, int blocks, double *d_out, cudaStream_t st)
{
    k_fp64_peak<<<blocks, 256, 0, st>>>(iters, d_out);
    return 1;
}

// ---------------------------------------------------------------- issue-model probes
// 8 independent DFMA chains with NI integer (ALU) op pairs and NM MUFU.RSQ64H per 8 DFMA.
template <int NI, int NM>
__global__ void __launch_bounds__(256) k_fp64_mix(int iters, double *out)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    unsigned x0 = threadIdx.x, x1 = blockIdx.x, x2 = 7u, x3 = 11u;
    double q = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if ((i & 3) == 0) x0 = __funnelshift_l(x0, x0, 3) ^ x1;
                if ((i & 3) == 1) x1 = __funnelshift_l(x1, x1, 5) ^ x2;
                if ((i & 3) == 2) x2 = __funnelshift_l(x2, x2, 7) ^ x3;
                if ((i & 3) == 3) x3 = __funnelshift_l(x3, x3, 9) ^ x0;
            }
#pragma unroll
            for (int i = 0; i < NM; ++i) q = rsqrt_seed(q + 1.5);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + q;
    if (s == 123.456 || (x0 ^ x1 ^ x2 ^ x3) == 0x12345u) out[0] = s;
}

// every DFMA reads three distinct register pairs (no operand reuse)
__global__ void __launch_bounds__(256) k_fp64_rf(int iters, double *out)
{
    double a[8], b[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = 1.0 + threadIdx.x * 1e-9 + k * 1e-3;
        b[k] = 0.999999 + k * 1e-9 + threadIdx.x * 1e-12;
        c[k] = 1e-9 * (k + 1) + threadIdx.x * 1e-15;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __fma_rn(b[k], c[(k + u) & 7], a[k]);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) out[0] = s;
}

int launch_fp64_mix(int kind, int iters, int blocks, double *d_out, cudaStream_t st)
{
    switch (kind) {
        case 0: k_fp64_mix<0, 0><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 1: k_fp64_mix<2, 0><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 2: k_fp64_mix<4, 0><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 3: k_fp64_mix<8, 0><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 4: k_fp64_mix<16, 0><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 5: k_fp64_mix<0, 1><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 6: k_fp64_mix<4, 1><<<blocks, 256, 0, st>>>(iters, d_out); break;
        case 7: k_fp64_rf<<<blocks, 256, 0, st>>>(iters, d_out); break;
        default: return 0;
    }
    return 1;
}

}  // namespace nb
