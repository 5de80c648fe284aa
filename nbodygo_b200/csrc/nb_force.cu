// nb_force.cu — K0 (prep) and K1 (tiled all-pairs force + collision detection).
//
// Replaces Body.Compute (cmd/body/body.go:148-187) with calcForceFrom (:214-225)
// and Collided (:192-208) for every body of the local i-shard, i.e. the work the
// reference fans out over goroutines (cmd/runner/workpool.go:103-110).
//
// K1 layout: one CTA = NT threads x R register-blocked i-bodies, one j-chunk.
// j-tiles (TJ bodies of x,y,z,jm = 4 x 2 KB) are staged into shared memory with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), double buffered.
// Per pair the fast path issues 16 FP64-pipe instructions:
//   3 DADD (dx,dy,dz)  1 DMUL + 2 DFMA (d2)
//   MUFU.RSQ64H seed (XU pipe)  + 7 DMUL/DFMA (cubic refinement of mj*d^-3)
//   3 DFMA accumulate
// The overlap predicate of the reference (dist > r_i + r_j, body.go:219) is
// screened with an integer compare on the high word of d2 against a per-(i,tile)
// conservative threshold (ALU pipe, no FP64 issue slot); only screened pairs take
// the exact path, which restates the reference's unfused arithmetic bit for bit
// (sqrt_rn(fl(fl(dx*dx+dy*dy)+dz*dz)) vs fl(r_i+r_j)) and emits collision events.
//
// This file is compiled with -fmad=false: every FMA below is explicit.
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

// mj * d2^(-3/2) from the seed y0: with e = 1 - d2*y0^2 (|e| <~ 2^-21),
// d2^(-3/2) = y0^3 (1-e)^(-3/2) = y0^3 (1 + e(3/2 + 15/8 e) + O(e^3)).  7 FP64 ops.
__device__ __forceinline__ double fast_w(double d2, double mj)
{
    const double y0 = rsqrt_seed(d2);
    const double u = __dmul_rn(y0, y0);
    const double e = __fma_rn(-d2, u, 1.0);
    const double pp = __fma_rn(1.875, e, 1.5);
    const double t = __dmul_rn(mj, y0);
    const double tu = __dmul_rn(t, u);
    const double q = __dmul_rn(tu, e);
    return __fma_rn(q, pp, tu);
}

// ---------------------------------------------------------------- K0: prep
// jm[j] = Exists && !fragmenting ? mass : 0 (the j-filter of body.go:162-165 folded
// into the mass), per-tile max radius of live bodies, benign tail beyond n.
__global__ void __launch_bounds__(TJ) k_prep(StepParams p)
{
    const long long j = (long long)blockIdx.x * TJ + threadIdx.x;
    double r = 0.0;
    if (j < p.n) {
        const unsigned fl = p.s.flags[j];
        const bool live = (fl & NB_F_EXISTS) != 0;
        const bool src = live && !(fl & NB_F_FRAGMENTING);
        p.s.jm[j] = src ? p.s.mass[j] : 0.0;
        if (live) r = p.s.radius[j];
    } else {
        // tail of the last tile: massless, far away, finite
        p.s.x[j] = 1e150;
        p.s.y[j] = 1e150;
        p.s.z[j] = 1e150;
        p.s.jm[j] = 0.0;
    }
    // block max of r (NaN radii are ignored by fmax)
    __shared__ double red[TJ / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = red[0];
#pragma unroll
        for (int w = 1; w < TJ / 32; ++w) m = fmax(m, red[w]);
        p.s.tile_rmax[blockIdx.x] = m;
    }
}

int launch_prep(const StepParams &p, cudaStream_t st)
{
    if (p.n_tiles <= 0) return 0;
    k_prep<<<p.n_tiles, TJ, 0, st>>>(p);
    return 1;
}

// ---------------------------------------------------------------- K1: exact path
__device__ __forceinline__ bool elastic_or_fragment(unsigned b) { return b == NB_ELASTIC || b == NB_FRAGMENT; }

__device__ __noinline__ void emit_event(const StepParams &p, long long i, long long j, double dist, double ri,
                                        double rj)
{
    const unsigned bi = p.s.behavior[i], bj = p.s.behavior[j];
    if (elastic_or_fragment(bi) && elastic_or_fragment(bj)) {
        // newCollision(b, otherBody), body.go:175-177
        const unsigned long long k = atomicAdd(&p.s.ctr->n_pairs, 1ull);
        if (k < (unsigned long long)p.seg_cap)
            p.s.pairs[k] = make_int2((int)i, (int)j);
        else
            p.s.ctr->overflow = 1;
    } else if (bi == NB_SUBSUME || bj == NB_SUBSUME) {
        // body.go:178-184: the larger radius subsumes, only if the centre is inside it
        int a = -1, b = -1;
        if (ri > rj && dist <= ri) { a = (int)i; b = (int)j; }
        else if (rj > ri && dist <= rj) { a = (int)j; b = (int)i; }
        if (a >= 0) {
            const unsigned long long k = atomicAdd(&p.s.ctr->n_hev, 1ull);
            if (k < (unsigned long long)p.hev_cap) {
                nb_event e;
                e.kind = NB_EV_SUBSUME; e.a = a; e.b = b; e._pad = 0; e.dist = dist; e.f1 = 0; e.f2 = 0;
                p.s.hev[k] = e;
            } else {
                p.s.ctr->overflow = 1;
            }
        }
    }
}

// Exact restatement of calcForceFrom / Collided for one screened pair. Returns the
// weight w = mj / dist^3 to accumulate (0 when the pair exerts no force).
__device__ __noinline__ double exact_pair(const StepParams &p, long long i, long long j, bool alive_i, double xi,
                                          double yi, double zi, double ri, double xj, double yj, double zj,
                                          double mj)
{
    if (!alive_i || j >= p.n || j == i) return 0.0;
    const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi), dz = __dsub_rn(zj, zi);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double dist = __dsqrt_rn(d2);
    const double rj = p.s.radius[j];
    const double s = __dadd_rn(ri, rj);
    if (dist > s) {
        if (mj == 0.0) return 0.0;
        return __ddiv_rn(__ddiv_rn(mj, __dmul_rn(dist, dist)), dist);
    }
    if (dist <= s) {
        if ((p.opts & NB_STEP_COLLISIONS) && (p.s.flags[j] & NB_F_EXISTS)) emit_event(p, i, j, dist, ri, rj);
    }
    return 0.0;
}

// ---------------------------------------------------------------- K1: kernel
template <int R, int NT>
__global__ void __launch_bounds__(NT) k_force(const __grid_constant__ StepParams p)
{
    __shared__ __align__(128) double sm[NSTAGE][4][TJ];
    __shared__ __align__(8) uint64_t bar[NSTAGE];

    const int tid = threadIdx.x;
    const long long ibase = p.i0 + (long long)blockIdx.x * (NT * R);
    const int chunk = blockIdx.y;
    const int t0 = chunk * p.tiles_per_chunk;
    int t1 = t0 + p.tiles_per_chunk;
    if (t1 > p.n_tiles) t1 = p.n_tiles;
    const int nt = t1 - t0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only: stage tile t0+t
        const int s = t % NSTAGE;
        const long long j0 = (long long)(t0 + t) * TJ;
        mbar_expect_tx(&bar[s], 4u * TJ * sizeof(double));
        bulk_g2s(&sm[s][0][0], p.s.x + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][1][0], p.s.y + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][2][0], p.s.z + j0, TJ * sizeof(double), &bar[s]);
        bulk_g2s(&sm[s][3][0], p.s.jm + j0, TJ * sizeof(double), &bar[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < NSTAGE - 1 && t < nt; ++t) issue(t);
    }

    double xi[R], yi[R], zi[R], ri[R];
    double ax[R], ay[R], az[R];
    bool alive[R];
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
        if (tid == 0 && t + NSTAGE - 1 < nt) issue(t + NSTAGE - 1);

        // conservative screen for this tile: not screened  <=>  thr+1 <= hi(d2) < 0x7FF00000
        unsigned thrp1[R], lim[R];
        const double rm = p.s.tile_rmax[t0 + t];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            unsigned thr = 0;
            if (alive[r]) {
                const double sr = __dadd_rn(ri[r], rm);
                const double t2 = __dmul_rn(__dmul_rn(sr, sr), 1.0 + 1.0 / 262144.0);
                thr = (unsigned)__double2hiint(t2) + 1u;
                if (thr > 0x7FEFFFFFu) thr = 0x7FEFFFFFu;
            }
            thrp1[r] = thr + 1u;
            lim[r] = 0x7FF00000u - thrp1[r];
        }

        mbar_wait(&bar[s], (unsigned)((t / NSTAGE) & 1));
        const double *sx = sm[s][0], *sy = sm[s][1], *sz = sm[s][2], *sj = sm[s][3];
        const long long jt0 = (long long)(t0 + t) * TJ;

#pragma unroll 2
        for (int jj = 0; jj < TJ; ++jj) {
            const double xj = sx[jj], yj = sy[jj], zj = sz[jj], mj = sj[jj];
            double dx[R], dy[R], dz[R], d2[R];
            bool screened = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                dx[r] = __dsub_rn(xj, xi[r]);
                dy[r] = __dsub_rn(yj, yi[r]);
                dz[r] = __dsub_rn(zj, zi[r]);
                d2[r] = __fma_rn(dz[r], dz[r], __fma_rn(dy[r], dy[r], __dmul_rn(dx[r], dx[r])));
                screened |= ((unsigned)__double2hiint(d2[r]) - thrp1[r]) >= lim[r];
            }
            if (__any_sync(0xffffffffu, screened)) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double w;
                    if (((unsigned)__double2hiint(d2[r]) - thrp1[r]) >= lim[r])
                        w = exact_pair(p, ibase + (long long)r * NT + tid, jt0 + jj, alive[r], xi[r], yi[r], zi[r],
                                       ri[r], xj, yj, zj, mj);
                    else
                        w = fast_w(d2[r], mj);
                    if (w != 0.0) {  // a forceless pair may carry non-finite dx (dead NaN-culled j)
                        ax[r] = __fma_rn(w, dx[r], ax[r]);
                        ay[r] = __fma_rn(w, dy[r], ay[r]);
                        az[r] = __fma_rn(w, dz[r], az[r]);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double w = fast_w(d2[r], mj);
                    ax[r] = __fma_rn(w, dx[r], ax[r]);
                    ay[r] = __fma_rn(w, dy[r], ay[r]);
                    az[r] = __fma_rn(w, dz[r], az[r]);
                }
            }
        }
        __syncthreads();  // every warp is done with stage s before it is refilled
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
}

template <int R, int NT>
static int launch_force_t(const StepParams &p, cudaStream_t st)
{
    const long long n_local = p.i1 - p.i0;
    const long long per = (long long)NT * R;
    dim3 grid((unsigned)((n_local + per - 1) / per), (unsigned)p.n_chunks);
    k_force<R, NT><<<grid, NT, 0, st>>>(p);
    return 1;
}

// R is chosen from the local shard size only; it never changes the result bits
// (each body's j-order and chunking are functions of n alone).
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
    switch (R) {
        case 4: return launch_force_t<4, 128>(p, st);
        case 2: return launch_force_t<2, 128>(p, st);
        default: return launch_force_t<1, 128>(p, st);
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

}  // namespace nb
