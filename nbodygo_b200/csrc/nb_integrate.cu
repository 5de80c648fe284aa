// nb_integrate.cu — K4 (integrate + NaN cull + render snapshot) and K5 (stable
// compaction), plus small utility kernels.
//
// K4 restates Body.Update (cmd/body/body.go:114-139) and NewRenderable
// (cmd/body/renderable.go:22-40); K5 restates the delete half of
// BodyCollection.Cycle (cmd/body/body_collection.go:253-272).
// Compiled with -fmad=false so that  v += ts*f/m ; x += ts*v  round exactly like the
// reference's unfused expressions.
#include "nb_internal.cuh"

namespace nb {

constexpr int INT_THREADS = 256;

__global__ void __launch_bounds__(INT_THREADS, 2) k_integrate(const __grid_constant__ StepParams p)
{
    const DevState &s = p.s;
    const long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 64 or INT_THREADS threads, see launch_integrate
    const long long i = p.i0 + il;
    int culled = 0, dead = 0;
    const bool apply = !(p.opts & NB_STEP_NO_INTEGRATE) && !s.ctr->overflow;
    if (i < p.i1) {
        // partial sums in ascending chunk order, then F = (G*m_i) * sum
        // (the adds stay one chain in chunk order; the loads of a batch are issued together — with a few thousand
        // bodies the kernel is a handful of CTAs waiting on L2, and 63 dependent trips were 16 us of a 65 us cycle;
        // the (INT_THREADS, 2) launch bound is what lets ptxas keep a batch's 48 loads in registers instead of
        // trimming to 32 registers for occupancy)
        double sx = 0.0, sy = 0.0, sz = 0.0;
        const double *px = s.px + il, *py = s.py + il, *pz = s.pz + il;
        const long long stride = p.n_pad_local;
        int c = 0;
        for (; c + 16 <= p.n_chunks; c += 16) {
            double bx[16], by[16], bz[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const long long o = (long long)(c + u) * stride;
                bx[u] = px[o]; by[u] = py[o]; bz[u] = pz[o];
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) { sx += bx[u]; sy += by[u]; sz += bz[u]; }
        }
        for (; c + 4 <= p.n_chunks; c += 4) {
            double bx[4], by[4], bz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long o = (long long)(c + u) * stride;
                bx[u] = px[o]; by[u] = py[o]; bz[u] = pz[o];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { sx += bx[u]; sy += by[u]; sz += bz[u]; }
        }
        for (; c < p.n_chunks; ++c) {
            const long long o = (long long)c * stride;
            sx += px[o];
            sy += py[o];
            sz += pz[o];
        }
        unsigned fl = s.flags[i];
        const double m = s.mass[i];          // after ProcessMods (a subsume adds the swallowed mass)
        const double gm = G_CONST() * s.m0[i];  // calcForceFrom ran with the mass at the top of the cycle
        // Body.Compute returns early for !Exists and for fragmenting bodies (body.go:149-155): their
        // fx,fy,fz keep the value of the last cycle they computed, and Update still applies it
        const bool computes = s.computes0[i] != 0;  // as of the top of the cycle (K0), not after ProcessMods
        const double fx = computes ? gm * sx : s.fx[i];
        const double fy = computes ? gm * sy : s.fy[i];
        const double fz = computes ? gm * sz : s.fz[i];
        s.fx[i] = fx;
        s.fy[i] = fy;
        s.fz[i] = fz;
        if (apply && (fl & NB_F_EXISTS)) {
            double vx = s.vx[i], vy = s.vy[i], vz = s.vz[i];
            if (!(fl & NB_F_COLLIDED)) {
                vx += p.ts * fx / m;  // body.go:120-122
                vy += p.ts * fy / m;
                vz += p.ts * fz / m;
                s.vx[i] = vx;
                s.vy[i] = vy;
                s.vz[i] = vz;
            }
            const double x = s.x[i] + p.ts * vx;  // body.go:124-126
            const double y = s.y[i] + p.ts * vy;
            const double z = s.z[i] + p.ts * vz;
            s.x[i] = x;
            s.y[i] = y;
            s.z[i] = z;
            fl &= ~NB_F_COLLIDED;
            s.rest[i] = p.R;
            if (isnan(x) || isnan(y) || isnan(z)) {  // body.go:134-137
                fl &= ~NB_F_EXISTS;
                culled = 1;
            }
            s.flags[i] = (uint8_t)fl;
            // fused exchange: the shard's new state goes straight into every peer's replica (NVLink
            // peer stores, coalesced over i) instead of a separate all-gather
            if (p.peers) {
                const PeerTable &pt = *p.peers;
                for (int q = 0; q < p.nranks; ++q) {
                    if (q == p.rank) continue;
                    pt.x[q][i] = x; pt.y[q][i] = y; pt.z[q][i] = z;
                    pt.vx[q][i] = vx; pt.vy[q][i] = vy; pt.vz[q][i] = vz;
                    pt.rest[q][i] = p.R;
                    pt.flags[q][i] = (uint8_t)fl;
                }
            }
        }
        // A body that does not compute (or will not, from the next cycle on: K3 just set `fragmenting`)
        // keeps being updated with this stored force (body.go:152-155).  Cycle may move the shard
        // boundaries before then, so every replica gets it.
        if (p.peers && apply && (!computes || (fl & NB_F_FRAGMENTING))) {
            const PeerTable &pt = *p.peers;
            for (int q = 0; q < p.nranks; ++q) {
                if (q == p.rank) continue;
                pt.fx[q][i] = fx; pt.fy[q][i] = fy; pt.fz[q][i] = fz;
            }
        }
        // NewRenderable
        const bool ex = (fl & NB_F_EXISTS) != 0;
        dead = ex ? 0 : 1;
        s.render_exists[i] = ex ? 1 : 0;
        s.render[3 * i + 0] = ex ? (float)s.x[i] : 0.0f;
        s.render[3 * i + 1] = ex ? (float)s.y[i] : 0.0f;
        s.render[3 * i + 2] = ex ? (float)s.z[i] : 0.0f;
    }
    // block-level counts
    const unsigned cm = __ballot_sync(0xffffffffu, culled), dm = __ballot_sync(0xffffffffu, dead);
    if ((threadIdx.x & 31) == 0) {
        if (cm) atomicAdd(&s.ctr->n_culled, (unsigned long long)__popc(cm));
        if (dm && p.nranks == 1) atomicAdd(&s.ctr->n_dead, (unsigned long long)__popc(dm));
    }
}

int launch_integrate(const StepParams &p, cudaStream_t st)
{
    const long long n_local = p.i1 - p.i0;
    if (n_local <= 0) return 0;
    // a few thousand bodies: the partial sums (24 B x chunks per body) come out of L2 at the rate of the SMs that
    // ask for them — 64-thread CTAs put the same threads on four times as many SMs
    const int threads = n_local < 32768 ? 64 : INT_THREADS;
    k_integrate<<<(unsigned)((n_local + threads - 1) / threads), threads, 0, st>>>(p);
    return 1;
}

// multi-GPU: after the state exchange every rank holds the whole new state; it counts the dead bodies
// and takes the Renderable snapshot (NewRenderable) of ALL bodies, not only of its own shard
__global__ void __launch_bounds__(INT_THREADS) k_count_dead(const __grid_constant__ StepParams p)
{
    const long long i = (long long)blockIdx.x * INT_THREADS + threadIdx.x;
    const int dead = (i < p.n && !(p.s.flags[i] & NB_F_EXISTS)) ? 1 : 0;
    if (i < p.n) {
        p.s.render_exists[i] = dead ? 0 : 1;
        p.s.render[3 * i + 0] = dead ? 0.0f : (float)p.s.x[i];
        p.s.render[3 * i + 1] = dead ? 0.0f : (float)p.s.y[i];
        p.s.render[3 * i + 2] = dead ? 0.0f : (float)p.s.z[i];
    }
    const unsigned dm = __ballot_sync(0xffffffffu, dead);
    if ((threadIdx.x & 31) == 0 && dm) atomicAdd(&p.s.ctr->n_dead, (unsigned long long)__popc(dm));
}

// ---------------------------------------------------------------- peer flag protocol
// One thread per peer: publish `step_id` in slot (slot_base + my rank) of every peer's flag block.
// Stream order guarantees the preceding kernel (K3 / K4) has completed, so its peer stores are done.
__global__ void k_peer_signal(const __grid_constant__ StepParams p, int slot_base)
{
    const int q = threadIdx.x;
    if (q < p.nranks && q != p.rank) {
        __threadfence_system();
        volatile unsigned long long *f = p.peers->sync[q] + slot_base + p.rank;
        *f = p.step_id;
        __threadfence_system();
    }
}

// One thread per peer: wait until every peer has published >= step_id in my flag block.  Bounded
// (60 s of globaltimer) so that a lost peer turns into an error flag, never into a hung GPU.
__global__ void k_peer_wait(const __grid_constant__ StepParams p, int slot_base)
{
    const int q = threadIdx.x;
    if (q < p.nranks && q != p.rank) {
        volatile unsigned long long *f = p.peers->sync[p.rank] + slot_base + q;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f < p.step_id) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 60000000000ull) {
                p.s.ctr->peer_timeout = 1;
                break;
            }
        }
        __threadfence_system();
    }
}

int launch_peer_signal(const StepParams &p, int slot_base, cudaStream_t st)
{
    k_peer_signal<<<1, MAX_RANKS, 0, st>>>(p, slot_base);
    return 1;
}
int launch_peer_wait(const StepParams &p, int slot_base, cudaStream_t st)
{
    k_peer_wait<<<1, MAX_RANKS, 0, st>>>(p, slot_base);
    return 1;
}

// Sharded upload: the rank's slice [i0,i1) of the freshly copied arrays goes straight into every
// peer's replica (coalesced NVLink peer stores), so that each host link carries n/P bodies, not n.
__global__ void __launch_bounds__(INT_THREADS) k_push_shard(const __grid_constant__ StepParams p, unsigned mask)
{
    const long long i = p.i0 + (long long)blockIdx.x * INT_THREADS + threadIdx.x;
    if (i >= p.i1) return;
    const DevState &s = p.s;
    const PeerTable &pt = *p.peers;
    double *const src[11] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, s.radius, s.rest, s.ff, s.fs};
    double *const *const dst[11] = {pt.x, pt.y, pt.z, pt.vx, pt.vy, pt.vz, pt.mass, pt.radius, pt.rest, pt.ff, pt.fs};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        if (!(mask & (1u << k))) continue;
        const double v = src[k][i];
        for (int q = 0; q < p.nranks; ++q)
            if (q != p.rank) dst[k][q][i] = v;
    }
    if (mask & PUSH_BEHAVIOR) {
        const uint8_t v = s.behavior[i];
        for (int q = 0; q < p.nranks; ++q)
            if (q != p.rank) pt.behavior[q][i] = v;
    }
    if (mask & PUSH_FLAGS) {
        const uint8_t v = s.flags[i];
        for (int q = 0; q < p.nranks; ++q)
            if (q != p.rank) pt.flags[q][i] = v;
    }
}

int launch_push_shard(const StepParams &p, unsigned mask, cudaStream_t st)
{
    const long long n_local = p.i1 - p.i0;
    if (n_local <= 0 || !mask) return 0;
    k_push_shard<<<(unsigned)((n_local + INT_THREADS - 1) / INT_THREADS), INT_THREADS, 0, st>>>(p, mask);
    return 1;
}

int launch_count_dead(const StepParams &p, cudaStream_t st)
{
    if (p.n <= 0) return 0;
    k_count_dead<<<(unsigned)((p.n + INT_THREADS - 1) / INT_THREADS), INT_THREADS, 0, st>>>(p);
    return 1;
}

// ---------------------------------------------------------------- K5: compaction
constexpr int CP_THREADS = 1024;

// pass 1: live bodies per block
__global__ void __launch_bounds__(CP_THREADS) k_compact_count(const uint8_t *flags, long long n, unsigned *block_sums)
{
    __shared__ unsigned wsum[CP_THREADS / 32];
    const long long i = (long long)blockIdx.x * CP_THREADS + threadIdx.x;
    const int live = (i < n && (flags[i] & NB_F_EXISTS)) ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < CP_THREADS / 32; ++w) t += wsum[w];
        block_sums[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of block sums (single CTA, sequential chunks), total → new_n
__global__ void __launch_bounds__(CP_THREADS) k_compact_scan(unsigned *block_sums, int n_blocks, long long *new_n)
{
    __shared__ unsigned warp_tot[CP_THREADS / 32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += CP_THREADS) {
        const int k = base + threadIdx.x;
        const unsigned v = k < n_blocks ? block_sums[k] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
        const unsigned c = carry;
        if (k < n_blocks) block_sums[k] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == CP_THREADS - 1) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *new_n = (long long)carry;
}

// pass 3: map[new] = old (stable)
__global__ void __launch_bounds__(CP_THREADS) k_compact_map(const uint8_t *flags, long long n, const unsigned *block_offs,
                                                            long long *map)
{
    __shared__ unsigned wsum[CP_THREADS / 32];
    const long long i = (long long)blockIdx.x * CP_THREADS + threadIdx.x;
    const int live = (i < n && (flags[i] & NB_F_EXISTS)) ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    unsigned off = block_offs[blockIdx.x];
    for (int k = 0; k < w; ++k) off += wsum[k];
    off += __popc(m & ((1u << lane) - 1u));
    if (live) map[off] = i;
}

int launch_compact_map(const DevState &s, long long n, long long *d_map, unsigned *d_block_sums, long long *d_new_n,
                       cudaStream_t st)
{
    if (n <= 0) {
        cudaMemsetAsync(d_new_n, 0, sizeof(long long), st);
        return 0;
    }
    const int nb = (int)((n + CP_THREADS - 1) / CP_THREADS);
    k_compact_count<<<nb, CP_THREADS, 0, st>>>(s.flags, n, d_block_sums);
    k_compact_scan<<<1, CP_THREADS, 0, st>>>(d_block_sums, nb, d_new_n);
    k_compact_map<<<nb, CP_THREADS, 0, st>>>(s.flags, n, d_block_sums, d_map);
    return 3;
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather(const T *src, T *dst, const long long *map, const long long *new_n)
{
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k < *new_n) dst[k] = src[map[k]];
}

int launch_gather_f64(const double *src, double *dst, const long long *d_map, const long long *d_new_n,
                      long long n_old, cudaStream_t st)
{
    if (n_old <= 0) return 0;
    k_gather<double><<<(unsigned)((n_old + 255) / 256), 256, 0, st>>>(src, dst, d_map, d_new_n);
    return 1;
}
int launch_gather_u8(const uint8_t *src, uint8_t *dst, const long long *d_map, const long long *d_new_n,
                     long long n_old, cudaStream_t st)
{
    if (n_old <= 0) return 0;
    k_gather<uint8_t><<<(unsigned)((n_old + 255) / 256), 256, 0, st>>>(src, dst, d_map, d_new_n);
    return 1;
}

template <typename T>
__global__ void __launch_bounds__(256) k_fill(T *dst, T v, long long n)
{
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k < n) dst[k] = v;
}
int launch_fill_f64(double *dst, double v, long long n, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_fill<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
    return 1;
}
int launch_fill_u8(uint8_t *dst, uint8_t v, long long n, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_fill<uint8_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
    return 1;
}

}  // namespace nb
