// nb_api.cu — host side of libnbody_b200.so: the C ABI of include/nbody_b200.h.
//
// One nb_sim owns the device image of BodyCollection.arr (cmd/body/body_collection.go:16)
// on one GPU, a stream, and (optionally) an NCCL communicator.  nb_step enqueues
//   K0 prep → K1 force+detect → [exchange pairs] → K3 resolve → K4 integrate → [exchange state]
// on the handle's stream — the block cmd/runner/computation-runner.go:285-320.
// There is no CPU fallback anywhere in this file: without a device every entry
// point fails with NB_ERR_NO_DEVICE / NB_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <unistd.h>
#include <string>
#include <vector>

#include "nb_internal.cuh"

using namespace nb;

// ---------------------------------------------------------------- NCCL (dlopen)
namespace {
struct NcclUid { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUid *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
constexpr int NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8;

bool load_nccl(std::string &err)
{
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
#define NB_SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                      \
    if (!g_nccl.field) { err = std::string("missing NCCL symbol ") + name; return false; }
    NB_SYM(GetUniqueId, "ncclGetUniqueId")
    NB_SYM(CommInitRank, "ncclCommInitRank")
    NB_SYM(CommDestroy, "ncclCommDestroy")
    NB_SYM(AllGather, "ncclAllGather")
    NB_SYM(GroupStart, "ncclGroupStart")
    NB_SYM(GroupEnd, "ncclGroupEnd")
    NB_SYM(GetErrorString, "ncclGetErrorString")
#undef NB_SYM
    g_nccl.ok = true;
    return true;
}
std::string g_create_err;
}  // namespace

// ---------------------------------------------------------------- handle
struct nb_sim {
    int device = 0;
    long long cap = 0, cap_pad = 0, n = 0;
    long long seg_cap = 0, hev_cap = 0;
    DevState d{};
    // partial sums (grown on demand)
    long long part_elems = 0;
    // compaction scratch
    double *scratch_f64 = nullptr;
    uint8_t *scratch_u8 = nullptr;
    long long *d_map = nullptr, *d_new_n = nullptr;
    unsigned *d_block_sums = nullptr;
    unsigned long long *d_pair_counts = nullptr;
    Counters *h_ctr = nullptr;                 // pinned
    unsigned long long *h_counts = nullptr;    // pinned [MAX_RANKS]
    long long *h_new_n = nullptr;              // pinned
    cudaStream_t st = nullptr;
    cudaEvent_t ev[7] = {};
    int rank = 0, nranks = 1;
    NcclComm comm = nullptr;
    std::string err;
    nb_step_result last{};
    bool pending = false;
    bool stepped = false;
    unsigned last_opts = 0;
    long long launches = 0;
    int force_R = 0;
    bool uniform_tiles = true;  // NB_UNIFORM_TILES=0 keeps every tile on the general (per-body mass) pass
    int res_cluster = 8;        // CTAs of the resolve cluster (portable maximum); NB_RES_CLUSTER=1: one CTA as in round 1
    int res_fast = 1;           // short event lists are scheduled in shared memory; NB_RES_FAST=0: always the general path
    StepParams last_params{};
    // fused peer-memory exchange (K4 pushes the shard state into every peer's replica)
    bool peer_push = false;
    PeerTable *d_peers = nullptr;
    unsigned long long *d_sync = nullptr;  // my flag block: 2*MAX_RANKS slots
    unsigned long long step_id = 0;
    unsigned long long upload_id = 0;  // sharded uploads so far (flag value of the PEER_SLOT_UP_* rounds)
    int2 *pairs_all_base = nullptr;
    std::vector<void *> ipc_opened;
    // library-owned pinned host buffers that every step fills with the Renderable snapshot
    float *h_render = nullptr;
    uint8_t *h_render_exists = nullptr;
    // CUDA graph of one cycle (single GPU): the ~17 stream calls of a cycle cost more CPU issue time
    // than the kernels of a small collection run (n = 1 k: ~30 us of 40), so a cycle whose
    // parameters repeat is captured once and replayed with one cudaGraphLaunch.
    bool graphs_enabled = true;
    cudaGraphExec_t gexec = nullptr;
    StepParams gkey{};        // parameters the cached graph was captured with (step_id zeroed)
    float *gkey_render = nullptr;
    long long g_launches = 0;  // kernel launches inside the cached graph
    StepParams prev_key{};    // parameters of the previous step (a graph is built on the first repeat)
    float *prev_render = nullptr;
    bool have_prev = false;
    unsigned long long graph_replays = 0, graph_captures = 0;
};

#define NB_CUDA(h, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return NB_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define NB_NCCL(h, call)                                                                              \
    do {                                                                                              \
        int r_ = (call);                                                                              \
        if (r_ != 0) {                                                                                \
            (h)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                         \
            return NB_ERR_COMM;                                                                       \
        }                                                                                             \
    } while (0)

static int fail(nb_handle h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

static long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

static double **f64_fields(DevState &d, int k)
{
    double **f[] = {&d.x, &d.y, &d.z, &d.vx, &d.vy, &d.vz, &d.mass, &d.radius, &d.rest, &d.ff, &d.fs};
    return f[k];
}
constexpr int N_F64 = 11;

extern "C" int nb_abi_version(void) { return NB_ABI_VERSION; }

extern "C" const char *nb_last_error(nb_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

static void free_all(nb_handle h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
    if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
    cudaFree(h->d_peers);
    cudaFree(h->d_sync);
    for (int k = 0; k < N_F64; ++k) cudaFree(*f64_fields(h->d, k));
    cudaFree(h->d.jx); cudaFree(h->d.jy); cudaFree(h->d.jz);
    cudaFree(h->d.jm); cudaFree(h->d.m0); cudaFree(h->d.computes0); cudaFree(h->d.fx); cudaFree(h->d.fy); cudaFree(h->d.fz);
    cudaFree(h->d.behavior); cudaFree(h->d.flags); cudaFree(h->d.tile_rmax); cudaFree(h->d.tile_muni); cudaFree(h->d.tile_dead);
    cudaFree(h->d.px); cudaFree(h->d.py); cudaFree(h->d.pz);
    cudaFree(h->d.render); cudaFree(h->d.render_exists);
    if (h->d.pairs_all != h->d.pairs) cudaFree(h->d.pairs_all);
    cudaFree(h->d.rs_ev); cudaFree(h->d.rs_list); cudaFree(h->d.rs_state);
    cudaFree(h->d.rs_queue); cudaFree(h->d.rs_cand); cudaFree(h->d.rs_active);
    cudaFree(h->d.rs_lkey); cudaFree(h->d.rs_pos); cudaFree(h->d.rs_candkey); cudaFree(h->d.rs_geo);
    cudaFree(h->d.adj_off); cudaFree(h->d.adj_cnt); cudaFree(h->d.rs_ctl);
    cudaFree(h->d.pairs); cudaFree(h->d.hev); cudaFree(h->d.head); cudaFree(h->d.ctr); cudaFree(h->d.zeros);
    cudaFree(h->scratch_f64); cudaFree(h->scratch_u8); cudaFree(h->d_map); cudaFree(h->d_new_n);
    cudaFree(h->d_block_sums); cudaFree(h->d_pair_counts);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->h_new_n) cudaFreeHost(h->h_new_n);
    if (h->h_render) cudaFreeHost(h->h_render);
    if (h->h_render_exists) cudaFreeHost(h->h_render_exists);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

// Scratch of the resolve kernel, sized for the events of all ranks.
static cudaError_t alloc_resolve_scratch(nb_handle h, int nranks)
{
    cudaFree(h->d.rs_ev); cudaFree(h->d.rs_list); cudaFree(h->d.rs_state);
    cudaFree(h->d.rs_queue); cudaFree(h->d.rs_cand); cudaFree(h->d.rs_active);
    cudaFree(h->d.rs_lkey); cudaFree(h->d.rs_pos); cudaFree(h->d.rs_candkey); cudaFree(h->d.rs_geo);
    h->d.rs_lkey = h->d.rs_candkey = nullptr; h->d.rs_pos = nullptr; h->d.rs_geo = nullptr;
    h->d.rs_ev = nullptr; h->d.rs_list = h->d.rs_state = h->d.rs_queue = h->d.rs_cand = h->d.rs_active = nullptr;
    const size_t E = (size_t)h->seg_cap * (size_t)nranks;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&h->d.rs_ev, E * sizeof(int2))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_list, 2 * E * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_state, E * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_queue, E * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_cand, 2 * E * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_active, 2 * E * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_lkey, 2 * E * sizeof(unsigned long long))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_pos, E * sizeof(int2))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_candkey, 2 * E * sizeof(unsigned long long))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&h->d.rs_geo, E * sizeof(ElasticGeo))) != cudaSuccess) return e;
    return cudaSuccess;
}

extern "C" int nb_create(int device, int64_t capacity, int64_t pair_capacity, nb_handle *out)
{
    if (!out || capacity < 0) return fail(nullptr, NB_ERR_INVALID, "nb_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, NB_ERR_NO_DEVICE,
                    std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, NB_ERR_INVALID, "nb_create: device out of range");
    nb_sim *h = new nb_sim();
    h->device = device;
    h->cap = capacity;
    h->cap_pad = round_up(capacity + MAX_RANKS, TJ);
    h->seg_cap = pair_capacity > 0 ? pair_capacity : 4 * capacity + 65536;
    h->hev_cap = h->seg_cap;
    if (const char *fr = getenv("NB_FORCE_R")) h->force_R = atoi(fr);
    if (const char *g = getenv("NB_GRAPH")) h->graphs_enabled = atoi(g) != 0;
    if (const char *u = getenv("NB_UNIFORM_TILES")) h->uniform_tiles = atoi(u) != 0;
    if (const char *c = getenv("NB_RES_CLUSTER")) h->res_cluster = std::max(1, std::min(8, atoi(c)));
    if (const char *c = getenv("NB_RES_FAST")) h->res_fast = atoi(c) != 0;
    auto bail = [&](const char *what, cudaError_t ce) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
        free_all(h);
        return NB_ERR_CUDA;
    };
#define NB_TRY(call)                                   \
    do {                                               \
        cudaError_t ce_ = (call);                      \
        if (ce_ != cudaSuccess) return bail(#call, ce_); \
    } while (0)
    NB_TRY(cudaSetDevice(device));
    NB_TRY(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    for (auto &evn : h->ev) NB_TRY(cudaEventCreate(&evn));
    const size_t fb = (size_t)h->cap_pad * sizeof(double);
    for (int k = 0; k < N_F64; ++k) {
        NB_TRY(cudaMalloc((void **)f64_fields(h->d, k), fb));
        NB_TRY(cudaMemsetAsync(*f64_fields(h->d, k), 0, fb, h->st));
    }
    NB_TRY(cudaMalloc((void **)&h->d.jx, fb));
    NB_TRY(cudaMalloc((void **)&h->d.jy, fb));
    NB_TRY(cudaMalloc((void **)&h->d.jz, fb));
    NB_TRY(cudaMalloc((void **)&h->d.jm, fb));
    NB_TRY(cudaMalloc((void **)&h->d.m0, fb));
    NB_TRY(cudaMemsetAsync(h->d.m0, 0, fb, h->st));
    NB_TRY(cudaMalloc((void **)&h->d.computes0, (size_t)h->cap_pad));
    NB_TRY(cudaMemsetAsync(h->d.computes0, 0, (size_t)h->cap_pad, h->st));
    NB_TRY(cudaMalloc((void **)&h->d.fx, fb));
    NB_TRY(cudaMalloc((void **)&h->d.fy, fb));
    NB_TRY(cudaMalloc((void **)&h->d.fz, fb));
    NB_TRY(cudaMemsetAsync(h->d.fx, 0, fb, h->st));
    NB_TRY(cudaMemsetAsync(h->d.fy, 0, fb, h->st));
    NB_TRY(cudaMemsetAsync(h->d.fz, 0, fb, h->st));
    NB_TRY(cudaMalloc((void **)&h->d.behavior, (size_t)h->cap_pad));
    NB_TRY(cudaMalloc((void **)&h->d.flags, (size_t)h->cap_pad));
    NB_TRY(cudaMemsetAsync(h->d.behavior, 0, (size_t)h->cap_pad, h->st));
    NB_TRY(cudaMemsetAsync(h->d.flags, 0, (size_t)h->cap_pad, h->st));
    NB_TRY(cudaMalloc((void **)&h->d.tile_rmax, (size_t)(h->cap_pad / TJ_SMALL) * sizeof(double)));
    NB_TRY(cudaMalloc((void **)&h->d.tile_muni, (size_t)(h->cap_pad / TJ_SMALL) * sizeof(double)));
    NB_TRY(cudaMemsetAsync(h->d.tile_muni, 0, (size_t)(h->cap_pad / TJ_SMALL) * sizeof(double), h->st));
    NB_TRY(cudaMalloc((void **)&h->d.tile_dead, (size_t)(h->cap_pad / TJ_SMALL)));
    NB_TRY(cudaMemsetAsync(h->d.tile_dead, 0, (size_t)(h->cap_pad / TJ_SMALL), h->st));
    NB_TRY(cudaMalloc((void **)&h->d.render, (size_t)h->cap_pad * 3 * sizeof(float)));
    NB_TRY(cudaMalloc((void **)&h->d.render_exists, (size_t)h->cap_pad));
    NB_TRY(cudaMemsetAsync(h->d.render, 0, (size_t)h->cap_pad * 3 * sizeof(float), h->st));
    NB_TRY(cudaMemsetAsync(h->d.render_exists, 0, (size_t)h->cap_pad, h->st));
    NB_TRY(cudaMalloc((void **)&h->d.pairs, (size_t)h->seg_cap * sizeof(int2)));
    NB_TRY(cudaMalloc((void **)&h->d.hev, (size_t)h->hev_cap * sizeof(nb_event)));
    NB_TRY(cudaMalloc((void **)&h->d.head, (size_t)h->cap_pad * sizeof(unsigned long long)));
    NB_TRY(cudaMemsetAsync(h->d.head, 0, (size_t)h->cap_pad * sizeof(unsigned long long), h->st));
    NB_TRY(cudaMalloc((void **)&h->d.adj_off, (size_t)h->cap_pad * sizeof(int)));
    NB_TRY(cudaMalloc((void **)&h->d.adj_cnt, (size_t)h->cap_pad * sizeof(int)));
    NB_TRY(cudaMemsetAsync(h->d.adj_off, 0, (size_t)h->cap_pad * sizeof(int), h->st));
    NB_TRY(cudaMemsetAsync(h->d.adj_cnt, 0, (size_t)h->cap_pad * sizeof(int), h->st));
    NB_TRY(alloc_resolve_scratch(h, 1));
    NB_TRY(cudaMalloc((void **)&h->d.rs_ctl, 8 * sizeof(int)));
    NB_TRY(cudaMemsetAsync(h->d.rs_ctl, 0, 8 * sizeof(int), h->st));
    NB_TRY(cudaMalloc((void **)&h->d.zeros, 1024 * sizeof(unsigned)));
    NB_TRY(cudaMemsetAsync(h->d.zeros, 0, 1024 * sizeof(unsigned), h->st));
    NB_TRY(cudaMalloc((void **)&h->d.ctr, sizeof(Counters)));
    NB_TRY(cudaMemsetAsync(h->d.ctr, 0, sizeof(Counters), h->st));
    NB_TRY(cudaMalloc((void **)&h->scratch_f64, fb));
    NB_TRY(cudaMalloc((void **)&h->scratch_u8, (size_t)h->cap_pad));
    NB_TRY(cudaMalloc((void **)&h->d_map, (size_t)h->cap_pad * sizeof(long long)));
    NB_TRY(cudaMalloc((void **)&h->d_new_n, sizeof(long long)));
    NB_TRY(cudaMalloc((void **)&h->d_block_sums, (size_t)(h->cap_pad / 1024 + 2) * sizeof(unsigned)));
    NB_TRY(cudaMalloc((void **)&h->d_pair_counts, MAX_RANKS * sizeof(unsigned long long)));
    NB_TRY(cudaMemsetAsync(h->d_pair_counts, 0, MAX_RANKS * sizeof(unsigned long long), h->st));
    h->d.pair_counts = h->d_pair_counts;
    h->d.pairs_all = h->d.pairs;
    NB_TRY(cudaMallocHost((void **)&h->h_ctr, sizeof(Counters)));
    NB_TRY(cudaMallocHost((void **)&h->h_counts, MAX_RANKS * sizeof(unsigned long long)));
    NB_TRY(cudaMallocHost((void **)&h->h_new_n, sizeof(long long)));
    NB_TRY(cudaStreamSynchronize(h->st));
#undef NB_TRY
    *out = h;
    return NB_OK;
}

extern "C" int nb_destroy(nb_handle h)
{
    if (!h) return NB_ERR_INVALID;
    free_all(h);
    return NB_OK;
}

// ---------------------------------------------------------------- state sync
static int finish_step(nb_handle h, nb_step_result *out);

static int copy_in_f64(nb_handle h, double *dst, const double *src, long long first, long long count, double dflt,
                       bool use_default)
{
    if (src) {
        NB_CUDA(h, cudaMemcpyAsync(dst + first, src, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->st));
    } else if (use_default) {
        h->launches += launch_fill_f64(dst + first, dflt, count, h->st);
    }
    return NB_OK;
}
static int copy_in_u8(nb_handle h, uint8_t *dst, const uint8_t *src, long long first, long long count, uint8_t dflt,
                      bool use_default)
{
    if (src) {
        NB_CUDA(h, cudaMemcpyAsync(dst + first, src, (size_t)count, cudaMemcpyHostToDevice, h->st));
    } else if (use_default) {
        h->launches += launch_fill_u8(dst + first, dflt, count, h->st);
    }
    return NB_OK;
}

static int write_range(nb_handle h, long long first, long long count, bool defaults, double rest_default,
                       const double *x, const double *y, const double *z, const double *vx, const double *vy,
                       const double *vz, const double *mass, const double *radius, const double *rest,
                       const double *ff, const double *fs, const uint8_t *behavior, const uint8_t *flags)
{
    if (count == 0) return NB_OK;
    NB_CUDA(h, cudaSetDevice(h->device));
    const double *src[N_F64] = {x, y, z, vx, vy, vz, mass, radius, rest, ff, fs};
    const double dfl[N_F64] = {0, 0, 0, 0, 0, 0, 0, 0, rest_default, 0, 0};
    for (int k = 0; k < N_F64; ++k) {
        int rc = copy_in_f64(h, *f64_fields(h->d, k), src[k], first, count, dfl[k], defaults);
        if (rc) return rc;
    }
    int rc = copy_in_u8(h, h->d.behavior, behavior, first, count, NB_ELASTIC, defaults);
    if (rc) return rc;
    rc = copy_in_u8(h, h->d.flags, flags, first, count, NB_F_EXISTS, defaults);
    if (rc) return rc;
    if (defaults) {  // a new body starts with fx = fy = fz = 0 (NewBody, body.go:73-75)
        for (double *f : {h->d.fx, h->d.fy, h->d.fz})
            NB_CUDA(h, cudaMemsetAsync(f + first, 0, (size_t)count * sizeof(double), h->st));
    }
    // host buffers may be reused by the caller as soon as we return
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_upload(nb_handle h, int64_t n, const double *x, const double *y, const double *z, const double *vx,
                         const double *vy, const double *vz, const double *mass, const double *radius,
                         const double *restitution, const double *frag_factor, const double *frag_step,
                         const uint8_t *behavior, const uint8_t *flags)
{
    if (!h || n < 0) return fail(h, NB_ERR_INVALID, "nb_upload: bad arguments");
    if (n > h->cap) return fail(h, NB_ERR_CAPACITY, "nb_upload: n exceeds capacity");
    if (n > 0 && (!x || !y || !z || !vx || !vy || !vz || !mass || !radius))
        return fail(h, NB_ERR_INVALID, "nb_upload: x,y,z,vx,vy,vz,mass,radius are required");
    h->n = n;
    h->stepped = false;
    return write_range(h, 0, n, true, 1.0, x, y, z, vx, vy, vz, mass, radius, restitution, frag_factor, frag_step,
                       behavior, flags);
}

// Sharded variant of nb_upload for several GPUs: every rank passes only ITS slice of the array (the
// i-range nb_plan gives for n) and the slices reach the other ranks' replicas over NVLink instead of
// n bodies crossing every rank's host link.  Collective: every rank calls it with the same n.
extern "C" int nb_upload_shard(nb_handle h, int64_t n, int64_t first, int64_t count, const double *x, const double *y,
                               const double *z, const double *vx, const double *vy, const double *vz,
                               const double *mass, const double *radius, const double *restitution,
                               const double *frag_factor, const double *frag_step, const uint8_t *behavior,
                               const uint8_t *flags)
{
    if (!h || n < 0) return fail(h, NB_ERR_INVALID, "nb_upload_shard: bad arguments");
    if (n > h->cap) return fail(h, NB_ERR_CAPACITY, "nb_upload_shard: n exceeds capacity");
    const long long shard = (n + h->nranks - 1) / h->nranks;
    const long long i0 = std::min<long long>(n, (long long)h->rank * shard), i1 = std::min<long long>(n, i0 + shard);
    if (first != i0 || count != i1 - i0)
        return fail(h, NB_ERR_INVALID, "nb_upload_shard: [first, first+count) must be this handle's i-range (nb_plan)");
    if (count > 0 && (!x || !y || !z || !vx || !vy || !vz || !mass || !radius))
        return fail(h, NB_ERR_INVALID, "nb_upload_shard: x,y,z,vx,vy,vz,mass,radius are required");
    if (h->pending) {
        int rc = finish_step(h, nullptr);
        if (rc) return rc;
    }
    NB_CUDA(h, cudaSetDevice(h->device));
    h->n = n;
    h->stepped = false;
    if (h->nranks == 1)
        return write_range(h, 0, n, true, 1.0, x, y, z, vx, vy, vz, mass, radius, restitution, frag_factor, frag_step,
                           behavior, flags);
    StepParams p{};
    p.s = h->d;
    p.peers = h->d_peers;
    p.n = n; p.i0 = i0; p.i1 = i1;
    p.rank = h->rank; p.nranks = h->nranks;
    p.step_id = ++h->upload_id;
    if (h->peer_push) {
        // nobody may write into my replica while a kernel of my last cycle still reads it (the render
        // snapshot of k_count_dead), and vice versa: one flag round before the slices move
        h->launches += launch_peer_signal(p, PEER_SLOT_UP_READY, h->st);
        h->launches += launch_peer_wait(p, PEER_SLOT_UP_READY, h->st);
    }
    const double *src[N_F64] = {x, y, z, vx, vy, vz, mass, radius, restitution, frag_factor, frag_step};
    const double dfl[N_F64] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 0};
    unsigned mask = 0;
    for (int k = 0; k < N_F64; ++k) {
        double *dst = *f64_fields(h->d, k);
        if (src[k]) {
            if (count > 0)
                NB_CUDA(h, cudaMemcpyAsync(dst + i0, src[k], (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->st));
            mask |= 1u << k;
        } else {
            h->launches += launch_fill_f64(dst, dfl[k], n, h->st);  // defaults need no exchange
        }
    }
    if (behavior) {
        if (count > 0) NB_CUDA(h, cudaMemcpyAsync(h->d.behavior + i0, behavior, (size_t)count, cudaMemcpyHostToDevice, h->st));
        mask |= PUSH_BEHAVIOR;
    } else {
        h->launches += launch_fill_u8(h->d.behavior, NB_ELASTIC, n, h->st);
    }
    if (flags) {
        if (count > 0) NB_CUDA(h, cudaMemcpyAsync(h->d.flags + i0, flags, (size_t)count, cudaMemcpyHostToDevice, h->st));
        mask |= PUSH_FLAGS;
    } else {
        h->launches += launch_fill_u8(h->d.flags, NB_F_EXISTS, n, h->st);
    }
    for (double *f : {h->d.fx, h->d.fy, h->d.fz})  // NewBody, body.go:73-75
        NB_CUDA(h, cudaMemsetAsync(f, 0, (size_t)n * sizeof(double), h->st));
    if (h->peer_push) {
        h->launches += launch_push_shard(p, mask, h->st);
        h->launches += launch_peer_signal(p, PEER_SLOT_UP_ARRIVED, h->st);
        h->launches += launch_peer_wait(p, PEER_SLOT_UP_ARRIVED, h->st);
    } else {
        NB_NCCL(h, g_nccl.GroupStart());
        for (int k = 0; k < N_F64; ++k) {
            if (!(mask & (1u << k))) continue;
            double *a = *f64_fields(h->d, k);
            NB_NCCL(h, g_nccl.AllGather(a + (long long)h->rank * shard, a, (size_t)shard, NCCL_FLOAT64, h->comm, h->st));
        }
        if (mask & PUSH_BEHAVIOR)
            NB_NCCL(h, g_nccl.AllGather(h->d.behavior + (long long)h->rank * shard, h->d.behavior, (size_t)shard,
                                        NCCL_UINT8, h->comm, h->st));
        if (mask & PUSH_FLAGS)
            NB_NCCL(h, g_nccl.AllGather(h->d.flags + (long long)h->rank * shard, h->d.flags, (size_t)shard, NCCL_UINT8,
                                        h->comm, h->st));
        NB_NCCL(h, g_nccl.GroupEnd());
    }
    // host buffers may be reused by the caller as soon as we return; the peers' slices have landed
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    NB_CUDA(h, cudaMemcpy(h->h_ctr, h->d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost));
    if (h->h_ctr->peer_timeout) return fail(h, NB_ERR_COMM, "sharded upload timed out: a rank did not join it");
    return NB_OK;
}

extern "C" int nb_patch(nb_handle h, int64_t first, int64_t count, const double *x, const double *y, const double *z,
                        const double *vx, const double *vy, const double *vz, const double *mass,
                        const double *radius, const double *restitution, const double *frag_factor,
                        const double *frag_step, const uint8_t *behavior, const uint8_t *flags)
{
    if (!h || first < 0 || count < 0 || first + count > h->n) return fail(h, NB_ERR_INVALID, "nb_patch: bad range");
    return write_range(h, first, count, false, 0.0, x, y, z, vx, vy, vz, mass, radius, restitution, frag_factor,
                       frag_step, behavior, flags);
}

extern "C" int nb_append(nb_handle h, int64_t count, double R, const double *x, const double *y, const double *z,
                         const double *vx, const double *vy, const double *vz, const double *mass,
                         const double *radius, const double *frag_factor, const double *frag_step,
                         const uint8_t *behavior, const uint8_t *flags)
{
    if (!h || count < 0) return fail(h, NB_ERR_INVALID, "nb_append: bad arguments");
    if (h->n + count > h->cap) return fail(h, NB_ERR_CAPACITY, "nb_append: capacity exceeded");
    if (count > 0 && (!x || !y || !z || !vx || !vy || !vz || !mass || !radius))
        return fail(h, NB_ERR_INVALID, "nb_append: x,y,z,vx,vy,vz,mass,radius are required");
    // arr[j].r = R for every added body (body_collection.go:276,288)
    int rc = write_range(h, h->n, count, true, R, x, y, z, vx, vy, vz, mass, radius, nullptr, frag_factor, frag_step,
                         behavior, flags);
    if (rc) return rc;
    h->n += count;
    return NB_OK;
}

extern "C" int nb_count(nb_handle h, int64_t *n)
{
    if (!h || !n) return NB_ERR_INVALID;
    *n = h->n;
    return NB_OK;
}

extern "C" int nb_compact(nb_handle h, int64_t *n_out, int64_t *old_index, int64_t old_index_cap)
{
    if (!h) return NB_ERR_INVALID;
    NB_CUDA(h, cudaSetDevice(h->device));
    const long long n_old = h->n;
    h->launches += launch_compact_map(h->d, n_old, h->d_map, h->d_block_sums, h->d_new_n, h->st);
    // gather through the scratch buffer; peers hold mapped pointers to these arrays, so with the
    // peer exchange active the result is copied back instead of swapping the pointers
    auto settle_f64 = [&](double **f) -> int {
        h->launches += launch_gather_f64(*f, h->scratch_f64, h->d_map, h->d_new_n, n_old, h->st);
        if (h->peer_push) {
            NB_CUDA(h, cudaMemcpyAsync(*f, h->scratch_f64, (size_t)n_old * sizeof(double), cudaMemcpyDeviceToDevice,
                                       h->st));
        } else {
            std::swap(*f, h->scratch_f64);
        }
        return NB_OK;
    };
    auto settle_u8 = [&](uint8_t **f) -> int {
        h->launches += launch_gather_u8(*f, h->scratch_u8, h->d_map, h->d_new_n, n_old, h->st);
        if (h->peer_push) {
            NB_CUDA(h, cudaMemcpyAsync(*f, h->scratch_u8, (size_t)n_old, cudaMemcpyDeviceToDevice, h->st));
        } else {
            std::swap(*f, h->scratch_u8);
        }
        return NB_OK;
    };
    for (int k = 0; k < N_F64; ++k)
        if (int rc = settle_f64(f64_fields(h->d, k))) return rc;
    double **outs[] = {&h->d.fx, &h->d.fy, &h->d.fz};
    for (auto f : outs)
        if (int rc = settle_f64(f)) return rc;
    if (int rc = settle_u8(&h->d.behavior)) return rc;
    if (int rc = settle_u8(&h->d.flags)) return rc;
    NB_CUDA(h, cudaMemcpyAsync(h->h_new_n, h->d_new_n, sizeof(long long), cudaMemcpyDeviceToHost, h->st));
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    const long long n_new = n_old > 0 ? *h->h_new_n : 0;
    if (old_index) {
        const long long m = std::min<long long>(n_new, old_index_cap);
        static_assert(sizeof(long long) == sizeof(int64_t), "int64");
        NB_CUDA(h, cudaMemcpy(old_index, h->d_map, (size_t)m * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    h->n = n_new;
    h->stepped = false;  // pair / event indices of the last step no longer match the array
    if (n_out) *n_out = n_new;
    return NB_OK;
}

// ---------------------------------------------------------------- step
static void chunking(long long n, int &tj, int &n_tiles, int &n_chunks, int &tiles_per_chunk)
{
    // a function of n only: the per-body summation order never depends on the grid or the rank count
    long long small_below = TJ_SMALL_BELOW;
    if (const char *e = getenv("NB_TJ_SMALL_BELOW")) small_below = atoll(e);  // development override
    tj = n < small_below ? TJ_SMALL : (n >= LARGE_N_BELOW ? TJ_HUGE : TJ_LARGE);
    n_tiles = (int)((n + tj - 1) / tj);
    // enough chunks that the CTA grid has >= ~40 rounds per SM (tail < ~1 %), within [32, MAX_CHUNKS]
    long long want = n > 0 ? (CHUNK_TARGET_CTAS * 512 + n - 1) / n : 1;
    if (const char *e = getenv("NB_CHUNKS")) want = atoll(e);  // development override (tools/kbench.py)
    // Large collections are the ones that get sharded over several GPUs: an eighth of 1 M bodies is 245
    // i-blocks, and 245 x 32 CTAs are 26.5 waves over 296 resident slots — the half-empty last wave costs
    // 2 % of K1 (measured at 8 GPUs).  Twice the slots halve the CTAs: at most one wave in 53 is partial, for
    // 0.13 ms more K4 traffic at 1 M on one GPU.  (A function of n only, like the rest.)
    const long long min_chunks = n >= LARGE_N_BELOW ? MAX_CHUNKS : MIN_CHUNKS;
    int cap = (int)std::max<long long>(min_chunks, std::min<long long>(want, MAX_CHUNKS));
    int s = std::min(n_tiles, cap);
    if (s < 1) s = 1;
    tiles_per_chunk = (n_tiles + s - 1) / s;
    if (tiles_per_chunk < 1) tiles_per_chunk = 1;
    n_chunks = n_tiles > 0 ? (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk : 1;
}

static int ensure_partials(nb_handle h, long long elems)
{
    if (elems <= h->part_elems) return NB_OK;
    cudaFree(h->d.px); cudaFree(h->d.py); cudaFree(h->d.pz);
    h->d.px = h->d.py = h->d.pz = nullptr;
    h->part_elems = 0;
    NB_CUDA(h, cudaMalloc((void **)&h->d.px, (size_t)elems * sizeof(double)));
    NB_CUDA(h, cudaMalloc((void **)&h->d.py, (size_t)elems * sizeof(double)));
    NB_CUDA(h, cudaMalloc((void **)&h->d.pz, (size_t)elems * sizeof(double)));
    h->part_elems = elems;
    return NB_OK;
}

static int finish_step(nb_handle h, nb_step_result *out)
{
    NB_CUDA(h, cudaSetDevice(h->device));
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    if (h->pending) {
        const Counters &c = *h->h_ctr;
        nb_step_result r{};
        r.n_bodies = h->n;
        // the event list holds collisions and subsumes; n_pairs counts the collisions
        r.n_pairs = (h->last_opts & NB_STEP_COLLISIONS) ? c.total_pairs - (long long)c.n_sub_events : 0;
        r.n_subsumed = (int64_t)c.n_subsumed;
        r.n_host_events = (int64_t)std::min<unsigned long long>(c.n_hev, (unsigned long long)h->hev_cap);
        r.n_resolved = (int64_t)c.n_resolved;
        r.n_culled = (int64_t)c.n_culled;
        r.n_dead = (int64_t)c.n_dead;
        r.resolve_rounds = c.rounds;
        r.pair_overflow = c.overflow ? 1 : (c.n_hev > (unsigned long long)h->hev_cap ? 2 : 0);
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]); r.ms_total = ms;
        if (h->last_opts & NB_STEP_PHASE_TIMINGS) {
            cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); r.ms_prep = ms;
            cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); r.ms_force = ms;
            cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); r.ms_exchange = ms;
            cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); r.ms_resolve = ms;
            cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); r.ms_integrate = ms;
            float ms2 = 0;
            cudaEventElapsedTime(&ms2, h->ev[5], h->ev[6]); r.ms_exchange += ms2;
        }
        h->last = r;
        h->pending = false;
    }
    if (out) *out = h->last;
    if (h->h_ctr->peer_timeout)
        return fail(h, NB_ERR_COMM, "peer exchange timed out: a rank did not reach this cycle");
    if (h->last.pair_overflow == 1)
        return fail(h, NB_ERR_PAIR_OVERFLOW, "collision pair capacity exceeded; step not applied");
    return NB_OK;
}

extern "C" int nb_sync(nb_handle h, nb_step_result *out)
{
    if (!h) return NB_ERR_INVALID;
    return finish_step(h, out);
}

static int exchange_pairs(nb_handle h, StepParams &p)
{
    // counts first (8 bytes per rank), then the first `mx` entries of every rank's list
    NB_NCCL(h, g_nccl.AllGather(&h->d.ctr->n_pairs, h->d_pair_counts, 1, NCCL_UINT64, h->comm, h->st));
    NB_CUDA(h, cudaMemcpyAsync(h->h_counts, h->d_pair_counts, h->nranks * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, h->st));
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    unsigned long long mx = 0;
    for (int r = 0; r < h->nranks; ++r) mx = std::max(mx, h->h_counts[r]);
    mx = std::min<unsigned long long>(mx, (unsigned long long)h->seg_cap);
    p.seg_stride = (long long)std::max<unsigned long long>(mx, 1ull);
    if (mx > 0)
        NB_NCCL(h, g_nccl.AllGather(h->d.pairs, h->d.pairs_all, (size_t)mx * 2, NCCL_UINT32, h->comm, h->st));
    return NB_OK;
}

// Enqueues one cycle on the handle's stream: K0 -> K1 -> [pair exchange] -> K3 -> K4 -> [state
// exchange] -> snapshot / counter copies.  With `capturing` the same calls are recorded into a CUDA
// graph instead (timing events become event-record nodes).  *launches = kernels issued.
static cudaError_t record_event(nb_handle h, int k, bool capturing)
{
    return capturing ? cudaEventRecordWithFlags(h->ev[k], h->st, cudaEventRecordExternal)
                     : cudaEventRecord(h->ev[k], h->st);
}

static int enqueue_cycle(nb_handle h, StepParams &p, uint32_t opts, bool capturing, long long *launches)
{
    long long nl = 0;
    const long long shard = (h->n + h->nranks - 1) / h->nranks;
    const bool phases = (opts & NB_STEP_PHASE_TIMINGS) != 0;
    NB_CUDA(h, record_event(h, 0, capturing));
    // K0 resets the step counters itself (one node less); an empty collection launches no K0
    if (p.n_tiles <= 0) NB_CUDA(h, cudaMemsetAsync(h->d.ctr, 0, sizeof(Counters), h->st));
    nl += launch_prep(p, h->st);
    if (phases) NB_CUDA(h, record_event(h, 1, capturing));
    nl += launch_force(p, h->st, h->force_R);
    if (phases) NB_CUDA(h, record_event(h, 2, capturing));
    if (h->nranks > 1 && (opts & NB_STEP_COLLISIONS)) {
        if (h->peer_push) {
            nl += launch_push_pairs(p, h->st);
            nl += launch_peer_signal(p, PEER_SLOT_PAIRS, h->st);
            nl += launch_peer_wait(p, PEER_SLOT_PAIRS, h->st);
        } else {
            int rc = exchange_pairs(h, p);
            if (rc) return rc;
        }
    }
    if (phases) NB_CUDA(h, record_event(h, 3, capturing));
    if (opts & NB_STEP_COLLISIONS) nl += launch_resolve(p, h->st);
    if (phases) NB_CUDA(h, record_event(h, 4, capturing));
    const bool advance = !(opts & NB_STEP_NO_INTEGRATE);
    if (h->peer_push && advance) {
        // nobody may overwrite my replica before I have finished reading this cycle's inputs (K1, K3),
        // and I may not overwrite a peer's before it has: publish "done reading", wait for everyone's
        nl += launch_peer_signal(p, PEER_SLOT_DONE, h->st);
        nl += launch_peer_wait(p, PEER_SLOT_DONE, h->st);
    }
    nl += launch_integrate(p, h->st);
    if (phases) NB_CUDA(h, record_event(h, 5, capturing));
    if (h->nranks > 1) {
        if (advance) {
            if (h->peer_push) {
                // K4 already stored the shard into every peer; wait until every peer's shard has landed here
                nl += launch_peer_signal(p, PEER_SLOT_ARRIVED, h->st);
                nl += launch_peer_wait(p, PEER_SLOT_ARRIVED, h->st);
            } else {
                NB_NCCL(h, g_nccl.GroupStart());
                double *arrs[] = {h->d.x, h->d.y, h->d.z, h->d.vx, h->d.vy, h->d.vz, h->d.rest};
                for (double *a : arrs)
                    NB_NCCL(h, g_nccl.AllGather(a + (long long)h->rank * shard, a, (size_t)shard, NCCL_FLOAT64,
                                                h->comm, h->st));
                // Body.fx,fy,fz: read back only for bodies that do not compute (fragmenting), whose owner may
                // change when Cycle moves the shard boundaries
                for (double *a : {h->d.fx, h->d.fy, h->d.fz})
                    NB_NCCL(h, g_nccl.AllGather(a + (long long)h->rank * shard, a, (size_t)shard, NCCL_FLOAT64,
                                                h->comm, h->st));
                NB_NCCL(h, g_nccl.AllGather(h->d.flags + (long long)h->rank * shard, h->d.flags, (size_t)shard,
                                            NCCL_UINT8, h->comm, h->st));
                NB_NCCL(h, g_nccl.GroupEnd());
            }
        }
        nl += launch_count_dead(p, h->st);
    }
    NB_CUDA(h, record_event(h, 6, capturing));
    if (h->h_render && h->n > 0) {  // snapshot rides the same stream: in host memory when the step is synced
        NB_CUDA(h, cudaMemcpyAsync(h->h_render, h->d.render, (size_t)h->n * 3 * sizeof(float), cudaMemcpyDeviceToHost,
                                   h->st));
        NB_CUDA(h, cudaMemcpyAsync(h->h_render_exists, h->d.render_exists, (size_t)h->n, cudaMemcpyDeviceToHost,
                                   h->st));
    }
    NB_CUDA(h, cudaMemcpyAsync(h->h_ctr, h->d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->st));
    NB_CUDA(h, cudaGetLastError());
    *launches = nl;
    return NB_OK;
}

static bool same_params(const StepParams &a, const StepParams &b) { return memcmp(&a, &b, sizeof(StepParams)) == 0; }

extern "C" int nb_step(nb_handle h, double time_scaling, double R, uint32_t opts, nb_step_result *out)
{
    if (!h) return NB_ERR_INVALID;
    if (h->pending) {
        int rc = finish_step(h, nullptr);
        if (rc) return rc;
    }
    NB_CUDA(h, cudaSetDevice(h->device));
    StepParams p{};
    p.n = h->n;
    chunking(h->n, p.tj, p.n_tiles, p.n_chunks, p.tiles_per_chunk);
    const long long shard = (h->n + h->nranks - 1) / h->nranks;
    p.i0 = std::min<long long>(h->n, (long long)h->rank * shard);
    p.i1 = std::min<long long>(h->n, p.i0 + shard);
    p.n_pad_local = round_up(std::max<long long>(p.i1 - p.i0, 1), 32);
    p.rank = h->rank;
    p.nranks = h->nranks;
    p.seg_cap = h->seg_cap;
    p.hev_cap = h->hev_cap;
    p.opts = opts;
    p.uniform_tiles = h->uniform_tiles ? 1 : 0;
    p.res_cluster = h->res_cluster;
    p.res_fast = h->res_fast;
    p.ts = time_scaling;
    p.R = R;
    int rc = ensure_partials(h, (long long)p.n_chunks * p.n_pad_local);
    if (rc) return rc;
    h->d.pair_counts = h->d_pair_counts;
    p.seg_stride = h->seg_cap;
    p.s = h->d;
    p.step_id = ++h->step_id;
    p.peers = h->peer_push ? h->d_peers : nullptr;
    if (h->peer_push) {  // this cycle's parity of the gathered pair buffers
        const long long par = (long long)(p.step_id & 1ull);
        p.s.pairs_all = h->pairs_all_base + par * h->nranks * h->seg_cap;
        p.s.pair_counts = h->d_pair_counts + par * MAX_RANKS;
    }

    // ---- one cycle: replay the cached graph, or enqueue (and, on the first repeat, capture) -------
    StepParams key = p;
    key.step_id = 0;  // only the peer-exchange kernels read it, and those never run inside a graph
    const bool graphable = h->graphs_enabled && h->nranks == 1 && h->n > 0;
    if (graphable && h->gexec && same_params(key, h->gkey) && h->gkey_render == h->h_render) {
        NB_CUDA(h, cudaGraphLaunch(h->gexec, h->st));
        h->launches += h->g_launches;
        h->graph_replays++;
    } else {
        const bool capture = graphable && h->have_prev && same_params(key, h->prev_key) && h->prev_render == h->h_render;
        long long nl = 0;
        if (capture) {
            if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
            cudaGraph_t graph = nullptr;
            NB_CUDA(h, cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_cycle(h, p, opts, true, &nl);
            const cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
            cudaError_t ci = cudaSuccess;
            if (rc == NB_OK && ce == cudaSuccess && graph) ci = cudaGraphInstantiate(&h->gexec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != NB_OK || ce != cudaSuccess || ci != cudaSuccess || !h->gexec) {
                // never fatal: this handle goes on with plain stream launches
                cudaGetLastError();
                h->gexec = nullptr;
                h->graphs_enabled = false;
                rc = enqueue_cycle(h, p, opts, false, &nl);
                if (rc) return rc;
                h->launches += nl;
            } else {
                h->gkey = key;
                h->gkey_render = h->h_render;
                h->g_launches = nl;
                h->graph_captures++;
                NB_CUDA(h, cudaGraphLaunch(h->gexec, h->st));
                h->launches += nl;
            }
        } else {
            rc = enqueue_cycle(h, p, opts, false, &nl);
            if (rc) return rc;
            h->launches += nl;
        }
    }
    h->prev_key = key;
    h->prev_render = h->h_render;
    h->have_prev = true;
    NB_CUDA(h, cudaGetLastError());
    h->pending = true;
    h->stepped = true;
    h->last_opts = opts;
    h->last_params = p;
    if (opts & NB_STEP_ASYNC) return NB_OK;
    return finish_step(h, out);
}

// ---------------------------------------------------------------- results
static int copy_out(nb_handle h, void *dst, const void *src, size_t bytes)
{
    if (!dst || bytes == 0) return NB_OK;
    NB_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st));
    return NB_OK;
}

extern "C" int nb_download_state(nb_handle h, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                                 double *mass, double *radius, double *restitution, uint8_t *behavior,
                                 uint8_t *flags)
{
    if (!h) return NB_ERR_INVALID;
    NB_CUDA(h, cudaSetDevice(h->device));
    const size_t fb = (size_t)h->n * sizeof(double);
    double *dst[] = {x, y, z, vx, vy, vz, mass, radius, restitution};
    for (int k = 0; k < 9; ++k) {
        int rc = copy_out(h, dst[k], *f64_fields(h->d, k), fb);
        if (rc) return rc;
    }
    int rc = copy_out(h, behavior, h->d.behavior, (size_t)h->n);
    if (rc) return rc;
    rc = copy_out(h, flags, h->d.flags, (size_t)h->n);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

// Range variant: bodies [first, first+count) only.  On several GPUs every handle holds the whole state
// after a cycle, so a host that wants each GPU to return its own slice (n/P bodies per host link)
// passes the handle's i-range here.
extern "C" int nb_download_state_range(nb_handle h, int64_t first, int64_t count, double *x, double *y, double *z,
                                       double *vx, double *vy, double *vz, double *mass, double *radius,
                                       double *restitution, uint8_t *behavior, uint8_t *flags)
{
    if (!h || first < 0 || count < 0 || first + count > h->n) return fail(h, NB_ERR_INVALID, "nb_download_state_range: bad range");
    NB_CUDA(h, cudaSetDevice(h->device));
    const size_t fb = (size_t)count * sizeof(double);
    double *dst[] = {x, y, z, vx, vy, vz, mass, radius, restitution};
    for (int k = 0; k < 9; ++k) {
        int rc = copy_out(h, dst[k], *f64_fields(h->d, k) + first, fb);
        if (rc) return rc;
    }
    int rc = copy_out(h, behavior, h->d.behavior + first, (size_t)count);
    if (rc) return rc;
    rc = copy_out(h, flags, h->d.flags + first, (size_t)count);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_download_render_range(nb_handle h, int64_t first, int64_t count, float *xyz, uint8_t *exists)
{
    if (!h || first < 0 || count < 0 || first + count > h->n) return fail(h, NB_ERR_INVALID, "nb_download_render_range: bad range");
    NB_CUDA(h, cudaSetDevice(h->device));
    int rc = copy_out(h, xyz, h->d.render + 3 * first, (size_t)count * 3 * sizeof(float));
    if (rc) return rc;
    rc = copy_out(h, exists, h->d.render_exists + first, (size_t)count);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_download_render(nb_handle h, float *xyz, uint8_t *exists)
{
    if (!h) return NB_ERR_INVALID;
    NB_CUDA(h, cudaSetDevice(h->device));
    int rc = copy_out(h, xyz, h->d.render, (size_t)h->n * 3 * sizeof(float));
    if (rc) return rc;
    rc = copy_out(h, exists, h->d.render_exists, (size_t)h->n);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_render_buffers(nb_handle h, float **xyz, uint8_t **exists)
{
    if (!h || !xyz || !exists) return NB_ERR_INVALID;
    NB_CUDA(h, cudaSetDevice(h->device));
    if (!h->h_render) {
        NB_CUDA(h, cudaMallocHost((void **)&h->h_render, (size_t)h->cap_pad * 3 * sizeof(float)));
        NB_CUDA(h, cudaMallocHost((void **)&h->h_render_exists, (size_t)h->cap_pad));
        memset(h->h_render, 0, (size_t)h->cap_pad * 3 * sizeof(float));
        memset(h->h_render_exists, 0, (size_t)h->cap_pad);
    }
    *xyz = h->h_render;
    *exists = h->h_render_exists;
    return NB_OK;
}

extern "C" int nb_get_forces(nb_handle h, double *fx, double *fy, double *fz)
{
    if (!h) return NB_ERR_INVALID;
    NB_CUDA(h, cudaSetDevice(h->device));
    const size_t fb = (size_t)h->n * sizeof(double);
    int rc = copy_out(h, fx, h->d.fx, fb);
    if (!rc) rc = copy_out(h, fy, h->d.fy, fb);
    if (!rc) rc = copy_out(h, fz, h->d.fz, fb);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_set_forces(nb_handle h, int64_t first, int64_t count, const double *fx, const double *fy,
                             const double *fz)
{
    if (!h || first < 0 || count < 0 || first + count > h->n) return fail(h, NB_ERR_INVALID, "nb_set_forces: bad range");
    NB_CUDA(h, cudaSetDevice(h->device));
    const double *src[3] = {fx, fy, fz};
    double *dst[3] = {h->d.fx, h->d.fy, h->d.fz};
    for (int k = 0; k < 3; ++k)
        if (src[k] && count > 0)
            NB_CUDA(h, cudaMemcpyAsync(dst[k] + first, src[k], (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->st));
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_get_pairs(nb_handle h, int32_t *i, int32_t *j, int64_t cap, int64_t *n)
{
    if (!h || !n) return NB_ERR_INVALID;
    int rc = finish_step(h, nullptr);
    if (rc && rc != NB_ERR_PAIR_OVERFLOW) return rc;
    *n = 0;
    if (!h->stepped || !(h->last_opts & NB_STEP_COLLISIONS)) return NB_OK;
    NB_CUDA(h, cudaSetDevice(h->device));
    std::vector<int2> all;
    if (h->peer_push)  // counts of the last cycle's parity live on the device
        NB_CUDA(h, cudaMemcpy(h->h_counts, h->last_params.s.pair_counts, h->nranks * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost));
    for (int r = 0; r < h->nranks; ++r) {
        unsigned long long c = h->nranks == 1 ? h->h_ctr->n_pairs : h->h_counts[r];
        c = std::min<unsigned long long>(c, (unsigned long long)h->seg_cap);
        if (!c) continue;
        const size_t off = all.size();
        all.resize(off + c);
        NB_CUDA(h, cudaMemcpy(all.data() + off, h->last_params.s.pairs_all + (long long)r * h->last_params.seg_stride,
                              c * sizeof(int2), cudaMemcpyDeviceToHost));
    }
    // collisions only: subsume events share the device list but are reported by nb_get_host_events
    all.erase(std::remove_if(all.begin(), all.end(), [](const int2 &a) { return (a.y & EV_SUBSUME_BIT) != 0; }),
              all.end());
    std::sort(all.begin(), all.end(), [](const int2 &a, const int2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    *n = (int64_t)all.size();
    const int64_t m = std::min<int64_t>(cap, *n);
    for (int64_t k = 0; k < m; ++k) {
        if (i) i[k] = all[k].x;
        if (j) j[k] = all[k].y;
    }
    return NB_OK;
}

extern "C" int nb_get_cycle_top_positions(nb_handle h, int64_t first, int64_t count, double *x, double *y, double *z)
{
    if (!h || first < 0 || count < 0 || first + count > h->n) return fail(h, NB_ERR_INVALID, "nb_get_cycle_top_positions: bad range");
    if (!h->stepped) return fail(h, NB_ERR_INVALID, "nb_get_cycle_top_positions: no step since the array last changed");
    NB_CUDA(h, cudaSetDevice(h->device));
    // the j-stream K0 built for the last cycle: the positions Compute and ProcessMods saw
    const size_t fb = (size_t)count * sizeof(double);
    int rc = copy_out(h, x, h->d.jx + first, fb);
    if (!rc) rc = copy_out(h, y, h->d.jy + first, fb);
    if (!rc) rc = copy_out(h, z, h->d.jz + first, fb);
    if (rc) return rc;
    NB_CUDA(h, cudaStreamSynchronize(h->st));
    return NB_OK;
}

extern "C" int nb_get_host_events(nb_handle h, nb_event *ev, int64_t cap, int64_t *n)
{
    if (!h || !n) return NB_ERR_INVALID;
    int rc = finish_step(h, nullptr);
    if (rc && rc != NB_ERR_PAIR_OVERFLOW) return rc;
    *n = 0;
    if (!h->stepped) return NB_OK;
    NB_CUDA(h, cudaSetDevice(h->device));
    const unsigned long long c = std::min<unsigned long long>(h->h_ctr->n_hev, (unsigned long long)h->hev_cap);
    std::vector<nb_event> all(c);
    if (c) NB_CUDA(h, cudaMemcpy(all.data(), h->d.hev, c * sizeof(nb_event), cudaMemcpyDeviceToHost));
    // NB_EV_FRAG_INIT records come in the reference's handling order: descending key (i, j) of the event that
    // raised them (applied == 1: a is the event's b1 = i, 2: a is its b2 = j); within one event b1 before b2
    auto key_of = [](const nb_event &e) {
        const unsigned long long i = (unsigned)(e.applied == 2 ? e.b : e.a), j = (unsigned)(e.applied == 2 ? e.a : e.b);
        return (i << 32) | j;
    };
    std::sort(all.begin(), all.end(), [&](const nb_event &a, const nb_event &b) {
        if (a.kind != b.kind) return a.kind < b.kind;
        if (a.kind == NB_EV_FRAG_INIT) {
            const unsigned long long ka = key_of(a), kb = key_of(b);
            if (ka != kb) return ka > kb;
            return a.applied < b.applied;
        }
        if (a.a != b.a) return a.a < b.a;
        return a.b < b.b;
    });
    *n = (int64_t)c;
    const int64_t m = std::min<int64_t>(cap, *n);
    if (ev) std::copy(all.begin(), all.begin() + m, ev);
    return NB_OK;
}

// ---------------------------------------------------------------- multi-GPU
extern "C" int nb_comm_unique_id(void *id128)
{
    if (!id128) return NB_ERR_INVALID;
    std::string err;
    if (!load_nccl(err)) return fail(nullptr, NB_ERR_COMM, err);
    NcclUid uid;
    int r = g_nccl.GetUniqueId(&uid);
    if (r) return fail(nullptr, NB_ERR_COMM, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    memcpy(id128, &uid, sizeof uid);
    return NB_OK;
}

// Maps every peer's replica of the arrays K4 updates (CUDA IPC across processes, plain UVA pointers
// + peer access inside one process) so that K4 can store its shard straight into them.  All ranks
// agree on the outcome: if any rank cannot map a peer, everybody falls back to the NCCL all-gather.
namespace {
struct PeerInfo {
    cudaIpcMemHandle_t ipc[PEER_ARRAYS];
    unsigned long long raw[PEER_ARRAYS];
    long long pid;
    int device, ok;
};
}  // namespace

static int setup_peer_push(nb_handle h)
{
    const int P = h->nranks;
    NB_CUDA(h, cudaMalloc((void **)&h->d_sync, PEER_SYNC_SLOTS * sizeof(unsigned long long)));
    NB_CUDA(h, cudaMemset(h->d_sync, 0, PEER_SYNC_SLOTS * sizeof(unsigned long long)));
    // gathered pair buffers, double-buffered by cycle parity (a fast peer may already be pushing the
    // next cycle's pairs while this rank's host still reads the last cycle's list)
    if (h->d.pairs_all != h->d.pairs) cudaFree(h->d.pairs_all);
    h->d.pairs_all = nullptr;
    NB_CUDA(h, cudaMalloc((void **)&h->d.pairs_all, (size_t)2 * P * h->seg_cap * sizeof(int2)));
    cudaFree(h->d_pair_counts);
    h->d_pair_counts = nullptr;
    NB_CUDA(h, cudaMalloc((void **)&h->d_pair_counts, 2 * MAX_RANKS * sizeof(unsigned long long)));
    NB_CUDA(h, cudaMemset(h->d_pair_counts, 0, 2 * MAX_RANKS * sizeof(unsigned long long)));
    h->pairs_all_base = h->d.pairs_all;
    void *mine[PEER_ARRAYS] = {h->d.x, h->d.y, h->d.z, h->d.vx, h->d.vy, h->d.vz, h->d.rest, h->d.flags, h->d_sync,
                               h->d.pairs_all, h->d_pair_counts, h->d.fx, h->d.fy, h->d.fz, h->d.mass, h->d.radius,
                               h->d.ff, h->d.fs, h->d.behavior};
    std::vector<PeerInfo> info((size_t)P);
    PeerInfo &me = info[(size_t)h->rank];
    memset(&me, 0, sizeof me);
    me.ok = 1;
    me.pid = (long long)getpid();
    me.device = h->device;
    for (int k = 0; k < PEER_ARRAYS; ++k) {
        me.raw[k] = (unsigned long long)(uintptr_t)mine[k];
        if (cudaIpcGetMemHandle(&me.ipc[k], mine[k]) != cudaSuccess) { me.ok = 0; cudaGetLastError(); }
    }
    PeerInfo *d_info = nullptr;
    NB_CUDA(h, cudaMalloc((void **)&d_info, (size_t)P * sizeof(PeerInfo)));
    auto gather = [&]() -> int {
        NB_CUDA(h, cudaMemcpyAsync(d_info + h->rank, &me, sizeof(PeerInfo), cudaMemcpyHostToDevice, h->st));
        NB_NCCL(h, g_nccl.AllGather(d_info + h->rank, d_info, sizeof(PeerInfo), NCCL_UINT8, h->comm, h->st));
        NB_CUDA(h, cudaMemcpyAsync(info.data(), d_info, (size_t)P * sizeof(PeerInfo), cudaMemcpyDeviceToHost, h->st));
        NB_CUDA(h, cudaStreamSynchronize(h->st));
        return NB_OK;
    };
    const PeerInfo mine_copy = me;
    if (int rc = gather()) { cudaFree(d_info); return rc; }
    PeerTable t;
    memset(&t, 0, sizeof t);
    int ok = 1;
    for (int q = 0; q < P && ok; ++q) ok = info[(size_t)q].ok;
    void **slots[PEER_ARRAYS] = {(void **)t.x, (void **)t.y, (void **)t.z, (void **)t.vx, (void **)t.vy, (void **)t.vz,
                                 (void **)t.rest, (void **)t.flags, (void **)t.sync, (void **)t.pairs_all,
                                 (void **)t.pair_counts, (void **)t.fx, (void **)t.fy, (void **)t.fz, (void **)t.mass,
                                 (void **)t.radius, (void **)t.ff, (void **)t.fs, (void **)t.behavior};
    for (int q = 0; q < P && ok; ++q) {
        const PeerInfo &pi = info[(size_t)q];
        for (int k = 0; k < PEER_ARRAYS && ok; ++k) {
            void *ptr = nullptr;
            if (q == h->rank) {
                ptr = mine[k];
            } else if (pi.pid == mine_copy.pid) {
                // same process (one host driving several handles): UVA pointer + peer access
                cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                cudaGetLastError();
                ptr = (void *)(uintptr_t)pi.raw[k];
            } else {
                if (cudaIpcOpenMemHandle(&ptr, pi.ipc[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    ok = 0;
                    cudaGetLastError();
                } else {
                    h->ipc_opened.push_back(ptr);
                }
            }
            slots[k][q] = ptr;
        }
    }
    // second round: everybody must have succeeded, otherwise everybody uses NCCL
    me = mine_copy;
    me.ok = ok;
    if (int rc = gather()) { cudaFree(d_info); return rc; }
    for (int q = 0; q < P; ++q) ok = ok && info[(size_t)q].ok;
    cudaFree(d_info);
    if (!ok) {
        // not an error (the NCCL all-gathers give the same bits), but never silent: nb_comm_mode reports 2
        h->peer_push = false;
        h->err = "peer-memory exchange unavailable (IPC / peer access failed); using NCCL all-gather";
        std::fprintf(stderr, "[WARN] libnbody_b200 rank %d: %s\n", h->rank, h->err.c_str());
        return NB_OK;
    }
    NB_CUDA(h, cudaMalloc((void **)&h->d_peers, sizeof(PeerTable)));
    NB_CUDA(h, cudaMemcpy(h->d_peers, &t, sizeof(PeerTable), cudaMemcpyHostToDevice));
    h->peer_push = true;
    h->step_id = 0;
    return NB_OK;
}

extern "C" int nb_comm_init(nb_handle h, int rank, int nranks, const void *id128)
{
    if (!h || !id128 || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(h, NB_ERR_INVALID, "nb_comm_init: bad arguments");
    std::string err;
    if (!load_nccl(err)) return fail(h, NB_ERR_COMM, err);
    NB_CUDA(h, cudaSetDevice(h->device));
    NcclUid uid;
    memcpy(&uid, id128, sizeof uid);
    NB_NCCL(h, g_nccl.CommInitRank(&h->comm, nranks, uid, rank));
    h->rank = rank;
    h->nranks = nranks;
    // gathered pair list: one segment per rank
    if (nranks > 1) {
        NB_CUDA(h, alloc_resolve_scratch(h, nranks));
        const char *e = getenv("NB_PEER_PUSH");
        if (!e || atoi(e) != 0) {
            int rc = setup_peer_push(h);
            if (rc != NB_OK) return rc;
        }
        if (!h->peer_push) {
            if (h->d.pairs_all != h->d.pairs) cudaFree(h->d.pairs_all);
            h->d.pairs_all = nullptr;
            NB_CUDA(h, cudaMalloc((void **)&h->d.pairs_all, (size_t)h->seg_cap * nranks * sizeof(int2)));
            h->pairs_all_base = h->d.pairs_all;
        }
    }
    return NB_OK;
}

extern "C" int nb_comm_mode(nb_handle h, int *mode)
{
    if (!h || !mode) return NB_ERR_INVALID;
    *mode = h->nranks <= 1 ? NB_COMM_SINGLE : (h->peer_push ? NB_COMM_PEER_PUSH : NB_COMM_NCCL);
    return NB_OK;
}

extern "C" int nb_shard_range(nb_handle h, int64_t *i0, int64_t *i1)
{
    if (!h) return NB_ERR_INVALID;
    const long long shard = (h->n + h->nranks - 1) / h->nranks;
    const long long a = std::min<long long>(h->n, (long long)h->rank * shard);
    if (i0) *i0 = a;
    if (i1) *i1 = std::min<long long>(h->n, a + shard);
    return NB_OK;
}

// Pure planning function (no device needed): the i-range of `rank` and the j-chunking for n bodies.
extern "C" int nb_plan(int64_t n, int rank, int nranks, int64_t *i0, int64_t *i1, int32_t *n_chunks,
                       int32_t *tiles_per_chunk)
{
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks) return NB_ERR_INVALID;
    const long long shard = (n + nranks - 1) / nranks;
    const long long a = std::min<long long>(n, (long long)rank * shard);
    if (i0) *i0 = a;
    if (i1) *i1 = std::min<long long>(n, a + shard);
    int tj, nt, nc, tpc;
    chunking(n, tj, nt, nc, tpc);
    if (n_chunks) *n_chunks = nc;
    if (tiles_per_chunk) *tiles_per_chunk = tpc;
    return NB_OK;
}

// ---------------------------------------------------------------- diagnostics
extern "C" int nb_measure_fp64_peak(int device, int iters, double *tflops, float *ms_out)
{
    if (!tflops || iters <= 0) return NB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, NB_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NB_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8;
    double *d_out = nullptr;
    if (cudaMalloc((void **)&d_out, sizeof(double)) != cudaSuccess) return NB_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch_fp64_peak(iters / 8 + 1, blocks, d_out, 0);  // warm-up
    cudaEventRecord(a, 0);
    launch_fp64_peak(iters, blocks, d_out, 0);
    cudaEventRecord(b, 0);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
    const double fmas = (double)blocks * 256.0 * (double)iters * 16.0 * 8.0;
    *tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return NB_OK;
}

// Issue-model probe (development diagnostic, see k_fp64_mix): DFMA TFLOP/s with `kind` selecting
// the number of ALU / MUFU instructions interleaved per 8 DFMA.
extern "C" int nb_probe_fp64_mix(int device, int kind, int iters, double *tflops)
{
    if (!tflops || iters <= 0) return NB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return NB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NB_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8;
    double *d_out = nullptr;
    if (cudaMalloc((void **)&d_out, sizeof(double)) != cudaSuccess) return NB_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    if (!launch_fp64_mix(kind, iters / 8 + 1, blocks, d_out, 0)) { cudaFree(d_out); return NB_ERR_INVALID; }
    cudaEventRecord(a, 0);
    launch_fp64_mix(kind, iters, blocks, d_out, 0);
    cudaEventRecord(b, 0);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    if (e != cudaSuccess) return NB_ERR_CUDA;
    *tflops = 2.0 * (double)blocks * 256.0 * (double)iters * 16.0 * 8.0 / (ms * 1e-3) / 1e12;
    return NB_OK;
}

extern "C" int nb_graph_stats(nb_handle h, int64_t *captures, int64_t *replays)
{
    if (!h) return NB_ERR_INVALID;
    if (captures) *captures = (int64_t)h->graph_captures;
    if (replays) *replays = (int64_t)h->graph_replays;
    return NB_OK;
}

extern "C" int nb_launch_count(nb_handle h, int64_t *launches)
{
    if (!h || !launches) return NB_ERR_INVALID;
    *launches = h->launches;
    return NB_OK;
}
