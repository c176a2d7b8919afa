// libtritd: C ABI (include/tritd.h), contexts, device state and the iteration driver of
// the B200-native TriTD-ADMM path.  Reference call being replaced:
//   [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)
//   fast_robust_triple_tensor/triple_decomp_ADMM.m:1-70
#include "../../include/tritd.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels_admm.cuh"
#include "kernels_aux.cuh"
#include "kernels_contract.cuh"
#include "kernels_fused.cuh"
#include "kernels_update.cuh"
#include <climits>

using namespace tritd;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// progress lines ("Iter %d, errL=%.2e, errO=%.2e", :60-62) go through a replaceable sink so that a MEX gateway can
// route them to mexPrintf (the MATLAB desktop does not show the process's stdout)
static void default_print(const char* line, void*) { fputs(line, stdout); fflush(stdout); }
static tritd_print_fn g_print = default_print;
static void* g_print_user = nullptr;
extern "C" void tritd_set_print(tritd_print_fn fn, void* user) { g_print = fn ? fn : default_print; g_print_user = fn ? user : nullptr; }
static void emit(const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_print(buf, g_print_user);
}

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(TRITD_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define ST_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != TRITD_OK) return _s; \
    } while (0)

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (single-rank use needs no NCCL at all; in a process that
// already loaded torch this resolves to torch's bundled libnccl.so.2)
// ---------------------------------------------------------------------------
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.h) return TRITD_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(TRITD_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define NCCL_SYM(field, name)                                                     \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                    \
    if (!g_nccl.field) return fail(TRITD_ERR_NCCL, "libnccl: missing symbol %s", name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.h = h;
    return TRITD_OK;
}

#define NCCL_TRY(expr)                                                                                       \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess)                                                                               \
            return fail(TRITD_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tritd_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    ncclComm_t comm = nullptr;
    int num_sms = 0;
    int64_t launches = 0;
    PFN_encodeTiled encode = nullptr;
    tritd_problem* cached = nullptr;     // device state of the last tritd_admm_f64 call, reused when the shape repeats
    bool cache_poisoned = false;         // the last call failed half-way: the cached state must not be reused
    tritd_problem* hcached = nullptr;    // factor-level state of the last triple_product / evaluate call (same reuse rule)
    // single-process multi-GPU (tritd_create_devices): the context the caller holds is a GROUP of one member
    // context per device (rank g of `nranks` = number of devices, no NCCL communicator, peer access instead of IPC)
    bool inproc = false;                 // member of a group
    std::vector<tritd_ctx*> sub;         // group: the members
    std::vector<tritd_problem*> gcached; // group: cached member problems of the last tritd_admm_f64 call
    unsigned xepoch = 1;                 // peer exchange: first unused epoch (advances identically on every rank; 0 = the zeroed mailbox)
};

struct RankCfg { int NT, KS; };
static RankCfg rank_cfg(int r) {
    const int R = r * r;
    return RankCfg{((R + 7) / 8), (R + 3) / 4};
}

static int ctx_common_init(tritd_ctx* c, int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(TRITD_ERR_CUDA, "no CUDA device available (%s); libtritd has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(TRITD_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    c->device = device;
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TRITD_ERR_CUDA, "device %d is sm_%d%d; libtritd is built for sm_100a only", device, prop.major,
                    prop.minor);
    c->num_sms = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(TRITD_ERR_CUDA, "cuTensorMapEncodeTiled not found");
    c->encode = (PFN_encodeTiled)fn;
    return TRITD_OK;
}

extern "C" int tritd_create(int device, tritd_ctx** out) {
    if (!out) return fail(TRITD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    tritd_ctx* c = new tritd_ctx();
    int s = ctx_common_init(c, device);
    if (s != TRITD_OK) { delete c; return s; }
    *out = c;
    return TRITD_OK;
}

extern "C" int tritd_nccl_unique_id(void* id_out) {
    if (!id_out) return fail(TRITD_ERR_INVALID, "id_out is NULL");
    ST_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == TRITD_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return TRITD_OK;
}

extern "C" int tritd_create_rank(int device, int rank, int nranks, const void* nccl_id, tritd_ctx** out) {
    if (!out) return fail(TRITD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(TRITD_ERR_INVALID, "bad rank %d / nranks %d", rank, nranks);
    if (nranks > 1 && !nccl_id) return fail(TRITD_ERR_INVALID, "nccl_id is NULL");
    tritd_ctx* c = new tritd_ctx();
    int s = ctx_common_init(c, device);
    if (s != TRITD_OK) { delete c; return s; }
    c->rank = rank;
    c->nranks = nranks;
    if (nranks > 1) {
        s = nccl_load();
        if (s != TRITD_OK) { tritd_destroy(c); return s; }
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            tritd_destroy(c);
            return fail(TRITD_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        }
    }
    *out = c;
    return TRITD_OK;
}

// Single-process multi-GPU context (what a MEX gateway needs: one MATLAB process, several GPUs).  The returned
// context takes and returns FULL tensors through tritd_admm_f64 / tritd_admm_ex_f64; inside, device g owns the mode-3
// slab tritd_slab_bounds(n3, ndev, g) and the per-iteration partials travel through the same NVLink peer mailboxes
// as in the one-process-per-GPU mode, mapped by plain peer access (no IPC, no NCCL).
extern "C" int tritd_create_devices(const int* devices, int ndev, tritd_ctx** out) {
    if (!out) return fail(TRITD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1) return fail(TRITD_ERR_INVALID, "no devices given");
    if (ndev == 1) return tritd_create(devices[0], out);
    if (ndev > 8) return fail(TRITD_ERR_UNSUPPORTED, "at most 8 devices per context (got %d)", ndev);
    for (int a = 0; a < ndev; ++a) for (int b = a + 1; b < ndev; ++b)
        if (devices[a] == devices[b]) return fail(TRITD_ERR_INVALID, "device %d listed twice", devices[a]);
    tritd_ctx* g = new tritd_ctx();
    int s = ctx_common_init(g, devices[0]);
    for (int q = 0; s == TRITD_OK && q < ndev; ++q) {
        tritd_ctx* m = new tritd_ctx();
        g->sub.push_back(m);
        s = ctx_common_init(m, devices[q]);
        m->rank = q; m->nranks = ndev; m->inproc = true;
    }
    for (int a = 0; s == TRITD_OK && a < ndev; ++a) {
        cudaSetDevice(devices[a]);
        for (int b = 0; b < ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[b], 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { s = fail(TRITD_ERR_UNSUPPORTED, "no peer access from device %d to device %d (%s)", devices[a], devices[b], cudaGetErrorString(e)); break; }
        }
    }
    if (s != TRITD_OK) { tritd_destroy(g); return s; }
    *out = g;
    return TRITD_OK;
}

extern "C" void tritd_problem_destroy(tritd_problem* p);
extern "C" int tritd_problem_get_E(tritd_problem* p, double* E_host);

extern "C" int tritd_trim(tritd_ctx* c) {
    if (!c) return fail(TRITD_ERR_INVALID, "ctx is NULL");
    if (c->cached) { tritd_problem_destroy(c->cached); c->cached = nullptr; }
    if (c->hcached) { tritd_problem_destroy(c->hcached); c->hcached = nullptr; }
    for (tritd_problem* q : c->gcached) tritd_problem_destroy(q);
    c->gcached.clear();
    return TRITD_OK;
}

extern "C" void tritd_destroy(tritd_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    tritd_trim(c);
    for (tritd_ctx* m : c->sub) tritd_destroy(m);
    c->sub.clear();
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" const char* tritd_last_error(void) { return g_err; }
extern "C" const char* tritd_version(void) { return "tritd-b200 0.1 (sm_100a)"; }
extern "C" int64_t tritd_launch_count(const tritd_ctx* c) {
    if (!c) return 0;
    int64_t n = c->launches;
    for (const tritd_ctx* m : c->sub) n += m->launches;
    return n;
}

extern "C" int tritd_set_stream(tritd_ctx* c, void* cuda_stream) {
    if (!c) return fail(TRITD_ERR_INVALID, "ctx is NULL");
    if (!c->sub.empty() && cuda_stream) return fail(TRITD_ERR_UNSUPPORTED, "a multi-device context runs on its own per-device streams");
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return TRITD_OK;
}

extern "C" int tritd_slab_bounds(int64_t n3, int nranks, int rank, int64_t* t0, int64_t* t1) {
    if (n3 < 0 || nranks < 1 || rank < 0 || rank >= nranks || !t0 || !t1) return fail(TRITD_ERR_INVALID, "bad slab query");
    const int64_t base = n3 / nranks, extra = n3 % nranks;
    *t0 = rank * base + (rank < extra ? rank : extra);
    *t1 = *t0 + base + (rank < extra ? 1 : 0);
    return TRITD_OK;
}

static int allreduce_sum(tritd_ctx* c, double* buf, size_t n) {
    if (c->nranks == 1) return TRITD_OK;
    NCCL_TRY(g_nccl.AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, c->stream));
    return TRITD_OK;
}

// ---------------------------------------------------------------------------
// problem (device-resident state of one solve)
// ---------------------------------------------------------------------------
struct tritd_problem {
    tritd_ctx* ctx = nullptr;
    int n1 = 0, n2 = 0, n3 = 0, r = 0, R = 0, RS = 0, NT = 0, KS = 0;
    int ld1 = 0, ldt = 0;
    size_t Np = 0;                       // padded elements of one N-array: ld1 * n2 * n3
    double *D = nullptr, *Z = nullptr, *YL = nullptr, *T = nullptr, *O = nullptr;   // Z: the sparse pair (E, Y_O) as one array (k_admm)
    double *A1 = nullptr, *B2 = nullptr, *C3 = nullptr, *A1T = nullptr;
    double *SA = nullptr, *SB = nullptr;
    double *bufA = nullptr;              // [rhsA (n1*RS) ; SC (RS*RS)] -- one all-reduce
    double *rhsB = nullptr, *rhsC = nullptr, *P = nullptr, *partM = nullptr;
    double *norm_part = nullptr, *norms = nullptr;
    double* gpart = nullptr;             // [3][kGramSlices][RS][RS] row-slice partial Grams of k_upd
    // N>1 peer exchange (kernels_xchg.cuh): local mailbox, the peers' mailboxes mapped with CUDA IPC
    bool xchg = false;
    double* box = nullptr;               // [A: nranks x slotA | B: nranks x slotB | N: nranks x 8 | S: 2 x nranks x RS^2 | flags (u32)]
    size_t slotA = 0, slotB = 0, offB = 0, offN = 0, offS = 0, offF = 0, box_doubles = 0;
    std::vector<void*> peer_map;         // cudaIpcOpenMemHandle mappings (nullptr for the own rank)
    double** peers = nullptr;            // device array [nranks] of mailbox bases
    unsigned xbase = 0;
    int upd_wave = 0;                    // k_upd CTAs resident at once (one wave)
    double* ones = nullptr;              // [64] vector of ones: the weights of a plain-sum RHS source
    double* Minv = nullptr;              // [3][RS][RS] inverses of the three ridge systems (written by k_upd's block 0)
    long long* dbg = nullptr;            // optional globaltimer stamps of k_upd (TRITD_DEBUG_STAMPS=1)
    long long* dbgA = nullptr;           // ... and of k_admm: [gridA][2] CTA start / end
    unsigned* flags = nullptr;           // [3][4] k_upd hand-shake words (A, B, C) + [12] the k_admm completion ticket
    int *tile0 = nullptr, *tile1 = nullptr;   // first / last i-tile of each k_mttkrp1 CTA
    IterState* st = nullptr;
    double *errHist = nullptr, *errL = nullptr, *errO = nullptr;
    IterState* st_host = nullptr;        // pinned mirror
    IterState* poll = nullptr;           // pinned [2]: asynchronous copies of the iteration scalars (tritd_problem_iterate)
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    CUtensorMap mapT, mapA1T;
    alignas(64) AdmmMaps maps;           // boxes [strips of a full i-tile][8*JG j][16 i] of D, Y_L, Z, T, O for k_admm
    alignas(64) AdmmMaps mapsLast;       // the same with the depth of the last i-tile (no box reaches past the tensor in i)
    double* partF = nullptr;             // [gridA][128][RS] fused mode-1 partials (next iteration's X1*F')
    int* ctaTab = nullptr;               // per k_admm CTA: (i-tile, index among the tile's CTAs, CTAs of that tile)
    int* tileCnt = nullptr;              // CTAs (= partials of X1*F') per i-tile
    int partSlots = 0;                   // the largest of them: partF is [i-tile][partSlots][128][RS]
    bool pre_inv = false;                // the ridge inverses of updates A / B are computed by the previous k_admm / k_ppass (see fill_ridge_job)
    int jgp = 2;                         // k_admm: column groups per stage asked for (AdmmCfg::JGP)
    int gridA = 0, tileH = 128, nitA = 1;  // k_admm: rows per i-tile (16 x consumer warps used) and number of i-tiles
    bool rhsA_ready = false;             // partF holds X1*F' of the current T
    int n_it = 0, n_jc = 0, gridM = 0, gridP = 0, gridF = 0, gi = 0;
    long unitsM = 0, unitsP = 0;
    size_t smemM = 0, smemP = 0;
    tritd_opts opts{};
    bool has_D = false, initialized = false;
    bool masked = false;                 // completion variant: unobserved entries of D hold NaN (tritd_problem_set_mask_*)
    int level = 0;                       // what is allocated: kLevelSolver / kLevelContract / kLevelFactors
    int printed_k = 0;
    int hist_cap = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;    // TRITD_NPHASE + 1 events per profiled iteration
    cudaGraphExec_t graph = nullptr;     // one steady-state iteration, captured once and replayed
    cudaGraphExec_t graphN = nullptr;    // kGraphBatch iterations in one graph
    int graph_launches = 0;              // kernels inside the graph
    bool graph_off = false;              // capture failed or TRITD_NO_GRAPH set: plain launches
    std::vector<void*> allocs;
};

// What a tritd_problem holds.  The solver needs everything; the standalone helpers only what they touch, so that
// e.g. triple_product(A,B,C) after a solve costs one N-sized buffer, not six plus the exchange set-up.
enum { kLevelSolver = 0,      // 6 N-arrays, tensor maps of k_admm, partials, (N>1) the peer exchange
       kLevelContract = 1,    // T + its tensor map, A1T, P, the MTTKRP partials: tritd_mttkrp_f64
       kLevelFactors = 2 };   // factors and the small state only: triple_product / evaluate (they allocate their N-arrays)

template <typename Tp>
static int dalloc(tritd_problem* p, Tp** ptr, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(Tp) + 256);
    if (e != cudaSuccess) return fail(TRITD_ERR_CUDA, "cudaMalloc(%zu bytes): %s", count * sizeof(Tp), cudaGetErrorString(e));
    p->allocs.push_back(q);
    *ptr = (Tp*)q;
    return TRITD_OK;
}

static int make_map(tritd_ctx* c, CUtensorMap* map, void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = getenv("TRITD_L2PROMO")) {
        const int v = atoi(e);
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
    CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TRITD_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return TRITD_OK;
}

// ---- kernel dispatch on the triple rank r (1..8) ---------------------------
#define TRITD_DISPATCH_R(r, CALL)                       \
    switch (r) {                                        \
        case 1: { CALL(1, 1); } break;                  \
        case 2: { CALL(1, 1); } break;                  \
        case 3: { CALL(2, 3); } break;                  \
        case 4: { CALL(2, 4); } break;                  \
        case 5: { CALL(4, 7); } break;                  \
        case 6: { CALL(5, 9); } break;                  \
        case 7: { CALL(7, 13); } break;                 \
        case 8: { CALL(8, 16); } break;                 \
        default: return fail(TRITD_ERR_UNSUPPORTED, "r=%d unsupported (1..%d)", r, TRITD_MAX_R); \
    }

// Launch of an iteration kernel.  TRITD_PDL=1: programmatic dependent launch (common.cuh, pdl_wait) -- the kernel may
// become resident while the previous kernel of the stream drains; inside a captured graph this becomes a programmatic
// edge.  Measured on B200 it LOSES 2.5-3.5 us per iteration against plain graph edges (cfg3 350.1 vs 341.6 us,
// 240 x 320 x 38: 92.1 vs 88.7 us; profiles/r02_pdl.log): the gaps between graph nodes are already ~1 us.  Off by default.
static bool g_pdl = getenv("TRITD_PDL") != nullptr;
template <typename... KA, typename... A>
static cudaError_t launch_k(void (*kern)(KA...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

static size_t smem_mttkrp1(int NT) { return (size_t)kStages * kCW * kBoxBytes + (size_t)NT * 8 * kPJ * 8 + 2 * kStages * 8; }
static size_t smem_ppass(int NT) { return (size_t)ppass_stages(NT) * (kCW * kBoxBytes + NT * 8 * 128) + 2 * ppass_stages(NT) * 8; }

static int launch_mttkrp1(tritd_problem* p, const CUtensorMap& map, const double* B2, const double* C3, double* rhs_out) {
    tritd_ctx* c = p->ctx;
    Mttkrp1Args a;
    a.B2 = B2; a.C3 = C3; a.part = p->partM; a.stop = &p->st->stop;
    a.n1 = p->n1; a.n2 = p->n2; a.n3 = p->n3; a.RS = p->RS; a.n_it = p->n_it; a.n_jc = p->n_jc; a.units = p->unitsM;
#define CALL(NT_, KS_) \
    k_mttkrp1<NT_><<<p->gridM, (kCW + 1) * 32, p->smemM, c->stream>>>(map, a);
    TRITD_DISPATCH_R(p->r, CALL)
#undef CALL
    CU_TRY(cudaGetLastError());
    const long tot = (long)p->n1 * p->RS;
    k_mttkrp1_reduce<<<(unsigned)((tot + 31) / 32), 256, 0, c->stream>>>(p->partM, (size_t)2 * 128 * p->RS, 128, rhs_out, p->n1,
                                                                       p->RS, p->tile0, p->tile1, p->gridM, &p->st->stop);
    CU_TRY(cudaGetLastError());
    c->launches += 2;
    return TRITD_OK;
}

static void fill_ridge_job(tritd_problem* p, int which, RidgeJob& j);

static int launch_ppass(tritd_problem* p, const CUtensorMap& mapT, bool with_inv_B = false) {
    tritd_ctx* c = p->ctx;
    PpassArgs a;
    memset(&a.inv, 0, sizeof(a.inv));
    if (with_inv_B) fill_ridge_job(p, 1, a.inv);
    a.st = p->st; a.R = p->R;
    // the scalar Gauss-Jordan inverse takes ~0.27 us per column; a row block of this pass ~0.05 us per (16-row chunk x n-tile)
    a.inv_rb = (int)std::lround(0.27 * p->R / (0.05 * ((p->n1 + 15) / 16) * p->NT));
    a.P = p->P; a.stop = &p->st->stop;
    a.n1 = p->n1; a.n2 = p->n2; a.n3 = p->n3; a.RS = p->RS;
    a.n_jb = p->n_jc; a.n_rb = (long)p->n3 * p->n_jc; a.units = p->unitsP;
#define CALL(NT_, KS_) \
    CU_TRY(launch_k(k_ppass<NT_, ppass_scalar_col(NT_, KS_)>, p->gridP, (kPW + 1) * 32, p->smemP, c->stream, mapT, p->mapA1T, a));
    TRITD_DISPATCH_R(p->r, CALL)
#undef CALL
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

static int launch_fused(tritd_problem* p, int mode, double* Lout) {
    tritd_ctx* c = p->ctx;
    FusedArgs a;
    a.D = p->D; a.O = Lout;
    a.A1 = p->A1; a.B2 = p->B2; a.C3 = p->C3; a.st = p->st; a.norm_part = p->norm_part;
    a.n1 = p->n1; a.n2 = p->n2; a.n3 = p->n3; a.ld1 = p->ld1; a.RS = p->RS;
    a.n_it = p->n_it; a.n_jc = p->n_jc; a.gi = p->gi;
#define CALL(NT_, KS_)                                                              \
    if (mode == 2) k_fused<KS_, 2><<<p->gridF, 256, 0, c->stream>>>(a);             \
    else k_fused<KS_, 1><<<p->gridF, 256, 0, c->stream>>>(a);
    TRITD_DISPATCH_R(p->r, CALL)
#undef CALL
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

// The fused iteration kernel (TMA in / DMMA / TMA out); also leaves the next X1*F' partials in partF.
static int launch_admm(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    AdmmArgs a;
    a.A1 = p->A1; a.B2 = p->B2; a.C3 = p->C3; a.st = p->st; a.norm_part = p->norm_part; a.partM = p->partF;
    a.norms = p->norms; a.errHist = p->errHist; a.errL = p->errL; a.errO = p->errO;
    a.ticket = p->flags + 12; a.finalize = c->nranks == 1 ? 1 : 0;
    a.peers = p->xchg ? p->peers : nullptr; a.norm_off = (long)(p->offN + 8 * (size_t)c->rank);
    a.nslots = p->xchg ? p->box + p->offN : nullptr;
    a.rank = c->rank; a.nranks = c->nranks; a.xbase = p->xbase;
    a.n1s = p->n1;
    memset(&a.inv, 0, sizeof(a.inv));
    if (p->pre_inv) fill_ridge_job(p, 0, a.inv);
    a.R = p->R;
    a.inv_stages = (int)std::ceil(0.27 * p->R / (p->jgp == 1 ? 1.3 : 2.5));     // ~0.27 us per column vs ~2.5 us per two-group stage
    a.cta_tab = p->ctaTab; a.part_slots = p->partSlots; a.dbg = p->dbgA;
    a.n1 = p->n1; a.n2 = p->n2; a.n3 = p->n3; a.RS = p->RS; a.n_jc = p->n_jc; a.tile_h = p->tileH;
#define CALL(NT_, KS_)                                                                                                       \
    if (p->masked && p->jgp == 1) launch_k(k_admm<KS_, NT_, true, 1>, p->gridA, kAdmmThreads, AdmmCfg<KS_, NT_, 1>::kSmem, c->stream, p->maps, p->mapsLast, a);  \
    else if (p->masked) launch_k(k_admm<KS_, NT_, true, 2>, p->gridA, kAdmmThreads, AdmmCfg<KS_, NT_, 2>::kSmem, c->stream, p->maps, p->mapsLast, a);  \
    else if (p->jgp == 1) launch_k(k_admm<KS_, NT_, false, 1>, p->gridA, kAdmmThreads, AdmmCfg<KS_, NT_, 1>::kSmem, c->stream, p->maps, p->mapsLast, a); \
    else launch_k(k_admm<KS_, NT_, false, 2>, p->gridA, kAdmmThreads, AdmmCfg<KS_, NT_, 2>::kSmem, c->stream, p->maps, p->mapsLast, a);
    TRITD_DISPATCH_R(p->r, CALL)
#undef CALL
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

// One factor update (k_upd): RHS rows from `src`, X = RHS * inv(S1 o S2 + alpha I), S_out = X'X, optionally X
// transposed.  which = 0/1/2 (A/B/C) selects the scratch inverse and the hand-shake flags.  apply == false: only
// reduce the RHS rows into rhs_out (the all-reduce comes next).
constexpr int kGramSlices = 8;     // row slices of k_upd's Gram phase
constexpr size_t kFlagsSC = 8, kFlagsFixed = 24;   // u32 indices inside a mailbox's flag area: [norms: 8][C3'C3: 2 x 8][exchanges ...]
static int gram_slices(int n) { return std::max(1, std::min(kGramSlices, (n + 63) / 64)); }

enum UpdSrc { kSrcDirect = 0, kSrcPartF = 1, kSrcPB = 2, kSrcPC = 3 };

// The ridge system of update `which` (0 = A: S_B o S_C + lambda2 I, 1 = B: S_A o S_C + lambda2 I), as a job for the first
// k_admm / k_ppass CTA to finish (the inverse is then ready before the update kernel starts; its block 0 skips it).
static void fill_ridge_job(tritd_problem* p, int which, RidgeJob& j) {
    tritd_ctx* c = p->ctx;
    memset(&j, 0, sizeof(j));
    j.enable = 1;
    j.S1 = which == 0 ? p->SB : p->SA;
    j.S2 = p->bufA + (size_t)p->n1 * p->RS; j.ns2 = 1; j.s2stride = 0;
    if (p->xchg) {
        j.S2 = p->box + p->offS; j.ns2 = c->nranks; j.s2stride = (long)p->RS * p->RS;
        j.sc_wait = 1; j.s2par_stride = (long)c->nranks * p->RS * p->RS;
        j.sc_flags = reinterpret_cast<const unsigned*>(p->box + p->offF) + kFlagsSC;
        j.nranks = c->nranks; j.xbase = p->xbase;
    }
    j.alpha = p->opts.lambda2;
    j.Minv = p->Minv + (size_t)which * p->RS * p->RS;
    j.done_flag = p->flags + 4 * which;
}

static int launch_upd(tritd_problem* p, int which, int src, bool apply, const double* rhs_direct, double* rhs_out,
                      const double* S1, const double* S2, double alpha, double* X, double* XT, int n, double* S_out, bool inv_here = true) {
    tritd_ctx* c = p->ctx;
    UpdArgs a;
    memset(&a, 0, sizeof(a));
    const long RS = p->RS;
    a.w = p->ones; a.wstride = 0; a.tile_h = INT_MAX; a.tile_stride = 0; a.row_stride = RS; a.wpr = 1;
    switch (src) {
        case kSrcDirect: a.v = rhs_direct; a.stride = 0; a.count = 1; break;
        case kSrcPartF:
            a.v = p->partF; a.stride = 128 * RS; a.count = p->partSlots; a.tile_cnt = p->tileCnt;
            a.tile_h = p->tileH; a.tile_stride = (long)p->partSlots * 128 * RS; a.wpr = 8;
            break;
        case kSrcPB: a.v = p->P; a.stride = (long)p->n2 * RS; a.count = p->n3; a.w = p->C3; a.wstride = RS; a.wpr = 8; break;
        case kSrcPC: a.v = p->P; a.stride = RS; a.count = p->n2; a.row_stride = (long)p->n2 * RS; a.w = p->B2; a.wstride = RS; a.wpr = 8; break;
        default: return fail(TRITD_ERR_INVALID, "bad update source");
    }
    a.S1 = S1; a.S2 = S2; a.alpha = alpha; a.gr = gram_slices(n);
    a.ns2 = 1; a.s2stride = 0; a.inv_here = inv_here ? 1 : 0;
    if (p->xchg && apply) {
        a.peers = p->peers; a.rank = c->rank; a.nranks = c->nranks; a.xbase = p->xbase;
        a.flag_area_off = (long)p->offF; a.sc_flag_off = (long)kFlagsSC; a.sc_off = (long)p->offS;
        a.s2par_stride = (long)c->nranks * p->RS * p->RS;
        // C3'C3 lives in the exchange mailbox as per-rank partials (two parity buffers), pushed by update C
        if (S2 == p->bufA + (size_t)p->n1 * p->RS) {
            a.S2 = p->box + p->offS; a.ns2 = c->nranks; a.s2stride = (long)p->RS * p->RS;
            a.sc_wait = 1; a.sc_flags = reinterpret_cast<const unsigned*>(p->box + p->offF) + kFlagsSC;
        }
        if (which == 2) a.sc_push = 1;
        if (which < 2) {
            // updates A and B: the exchange of the RHS rows (and, for A, of the local C3'C3) happens inside the kernel
            const int ex = which;
            a.xmerge = 1;
            a.push_off = (long)((ex == 0 ? 0 : p->offB) + (ex == 0 ? p->slotA : p->slotB) * c->rank);
            a.xbox = p->box + (ex == 0 ? 0 : p->offB);
            a.xslot = (long)(ex == 0 ? p->slotA : p->slotB);
        }
    }
    // Row CTAs that wait -- for block 0's inverse (when it is computed in this launch) or, N>1, for their counterparts
    // on the other ranks -- hold their slot meanwhile: keep such a grid within ONE wave (more rows per CTA) so that no
    // second wave starts its reduction only after the first one has left.  Light reductions are latency bound and
    // want one wave too (1024 rows, r = 8: 58 vs 82 us for updates B + C); only a reduction that streams a lot of P with
    // nothing to wait for keeps the finest grid, for the loads in flight (512^3, r = 6: 110 vs 130 us).
    const bool heavy = (double)a.count * n * p->RS * 8.0 > 64e6;
    if (apply && (a.xmerge || a.inv_here || !heavy))
        while (a.wpr > 1 && (n + 8 / a.wpr - 1) / (8 / a.wpr) + 1 > p->upd_wave) a.wpr >>= 1;
    a.gram_part = p->gpart + (size_t)which * kGramSlices * p->RS * p->RS; a.gram_cnt = p->flags + 16 + 64 * which;
    a.Minv = p->Minv + (size_t)which * p->RS * p->RS;
    a.rhs_out = rhs_out; a.X = X; a.XT = XT; a.gram_out = S_out;
    a.st = p->st; a.flags = p->flags + 4 * which; a.apply = apply ? 1 : 0;
    a.n = n; a.R = p->R; a.RS = p->RS; a.ldt = p->ldt;
    a.dbg = p->dbg ? p->dbg + 16 * which : nullptr;
    const size_t sm = upd_smem_bytes(p->RS);
    const int rows = 8 / a.wpr;
    const unsigned grid = (unsigned)((n + rows - 1) / rows + 1);
    a.gram_cap = (int)grid <= p->upd_wave ? (int)grid : kUpdMaxGramCtas;
    switch ((p->R + 15) / 16) {
        case 1: launch_k(k_upd<1>, grid, kUpdThreads, sm, c->stream, a); break;
        case 2: launch_k(k_upd<2>, grid, kUpdThreads, sm, c->stream, a); break;
        case 3: launch_k(k_upd<3>, grid, kUpdThreads, sm, c->stream, a); break;
        default: launch_k(k_upd<4>, grid, kUpdThreads, sm, c->stream, a); break;
    }
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

// ---- peer exchange set-up ------------------------------------------------------------------------------------
static void exchange_layout(tritd_problem* p) {
    const int nr = p->ctx->nranks;
    const size_t RS = p->RS;
    p->slotA = (size_t)2 * p->n1 * RS;                // 16-byte self-validating words (kernels_xchg.cuh): 2 doubles per element
    p->slotB = (size_t)2 * p->n2 * RS;
    p->offB = p->slotA * nr;
    p->offN = p->offB + p->slotB * nr;
    p->offS = p->offN + (size_t)8 * nr;               // C3'C3 partials: [2 parities][nr][RS*RS]
    p->offF = p->offS + (size_t)2 * nr * RS * RS;     // flags (u32): [(unused): 8] [C3'C3: 2 x 8]
    p->box_doubles = p->offF + kFlagsFixed / 2 + 8;
}

// the mailbox bases of all ranks are known: upload them, size the one-wave limit of the exchanging grids
static int exchange_finish(tritd_problem* p, const std::vector<double*>& bases) {
    tritd_ctx* c = p->ctx;
    const int nr = c->nranks;
    int s;
    if ((s = dalloc(p, &p->peers, (size_t)nr)) != TRITD_OK) return s;
    CU_TRY(cudaMemcpy(p->peers, bases.data(), sizeof(double*) * nr, cudaMemcpyHostToDevice));
    p->xchg = true;
    return TRITD_OK;
}

// One process per GPU: map every rank's mailbox (CUDA IPC handles exchanged through NCCL once per problem).  Every
// rank takes part in both collectives whatever happened locally: a local failure (allocation, IPC disabled in the
// container, no peer access) only lowers the vote, and then ALL ranks use the NCCL all-reduces.
static int setup_exchange(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    const int nr = c->nranks;
    p->xchg = false;
    if (nr < 2 || nr > 8 || getenv("TRITD_XCHG_NCCL")) return TRITD_OK;       // NCCL all-reduces instead
    exchange_layout(p);
    int ok = 1;
    if (dalloc(p, &p->box, p->box_doubles) != TRITD_OK) { ok = 0; p->box = nullptr; }
    if (ok && cudaMemsetAsync(p->box, 0, p->box_doubles * 8, c->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, p->box) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    unsigned char* hdev = nullptr;
    int s;
    if ((s = dalloc(p, &hdev, (size_t)64 * (nr + 1))) != TRITD_OK) return s;       // (without this scratch no collective is possible)
    CU_TRY(cudaMemcpyAsync(hdev + 64 * nr, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllGather(hdev + 64 * nr, hdev, 64, ncclChar, c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(nr);
    CU_TRY(cudaMemcpyAsync(all.data(), hdev, (size_t)64 * nr, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    std::vector<double*> bases(nr, nullptr);
    p->peer_map.assign(nr, nullptr);
    for (int r = 0; ok && r < nr; ++r) {
        if (r == c->rank) { bases[r] = p->box; continue; }
        void* q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        p->peer_map[r] = q;
        bases[r] = (double*)q;
    }
    {
        double* flag = reinterpret_cast<double*>(hdev);           // scratch, no longer needed
        const double mine_ok = ok;
        CU_TRY(cudaMemcpyAsync(flag, &mine_ok, 8, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclMin, c->comm, c->stream));
        double all_ok = 0.0;
        CU_TRY(cudaMemcpyAsync(&all_ok, flag, 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (all_ok < 0.5) {
            for (void*& q : p->peer_map) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
            if (c->rank == 0) fprintf(stderr, "libtritd: peer mailboxes unavailable (allocation / CUDA IPC / peer access failed on some rank); using NCCL all-reduces\n");
            return TRITD_OK;
        }
    }
    return exchange_finish(p, bases);
}

// Single process, one member problem per device: the mailboxes are plain device allocations reachable through peer
// access (enabled by tritd_create_devices).
static int setup_exchange_inproc(std::vector<tritd_problem*>& ps) {
    const int nr = (int)ps.size();
    std::vector<double*> bases(nr, nullptr);
    for (int r = 0; r < nr; ++r) {
        tritd_problem* p = ps[r];
        CU_TRY(cudaSetDevice(p->ctx->device));
        exchange_layout(p);
        ST_TRY(dalloc(p, &p->box, p->box_doubles));
        CU_TRY(cudaMemset(p->box, 0, p->box_doubles * 8));
        bases[r] = p->box;
    }
    for (int r = 0; r < nr; ++r) {
        CU_TRY(cudaSetDevice(ps[r]->ctx->device));
        ST_TRY(exchange_finish(ps[r], bases));
    }
    return TRITD_OK;
}

static int launch_small_gram(tritd_problem* p, const double* X, int n, double* S) {
    tritd_ctx* c = p->ctx;
    k_small_gram<<<(p->RS * p->RS + 255) / 256, 256, 0, c->stream>>>(X, n, p->RS, S, &p->st->stop);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

extern "C" void tritd_problem_destroy(tritd_problem* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (void* q : p->peer_map) if (q) cudaIpcCloseMemHandle(q);
    for (void* q : p->allocs) cudaFree(q);
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    if (p->graph) cudaGraphExecDestroy(p->graph);
    if (p->graphN) cudaGraphExecDestroy(p->graphN);
    if (p->st_host) cudaFreeHost(p->st_host);
    if (p->poll) { cudaFreeHost(p->poll); cudaEventDestroy(p->poll_ev[0]); cudaEventDestroy(p->poll_ev[1]); }
    delete p;
}

static int problem_create(tritd_ctx* c, int64_t n1, int64_t n2, int64_t n3, int r, int level, tritd_problem** out);
extern "C" int tritd_problem_create(tritd_ctx* c, int64_t n1, int64_t n2, int64_t n3, int r, tritd_problem** out) {
    return problem_create(c, n1, n2, n3, r, kLevelSolver, out);
}

static int problem_create(tritd_ctx* c, int64_t n1, int64_t n2, int64_t n3, int r, int level, tritd_problem** out) {
    if (!c || !out) return fail(TRITD_ERR_INVALID, "ctx/out is NULL");
    *out = nullptr;
    if (!c->sub.empty() && level == kLevelSolver)
        return fail(TRITD_ERR_UNSUPPORTED, "the staged solver API needs a single-device context; a multi-device context solves through tritd_admm_f64");
    if (n1 < 1 || n2 < 1 || n3 < 1) return fail(TRITD_ERR_INVALID, "tensor size %lld x %lld x %lld", (long long)n1, (long long)n2, (long long)n3);
    if (r < 1) return fail(TRITD_ERR_INVALID, "triple rank r=%d", r);
    if (r > TRITD_MAX_R) return fail(TRITD_ERR_UNSUPPORTED, "r=%d unsupported (1..%d)", r, TRITD_MAX_R);
    if (n1 > (1 << 24) || n2 > (1 << 24) || n3 > (1 << 24)) return fail(TRITD_ERR_INVALID, "mode size above 2^24");
    CU_TRY(cudaSetDevice(c->device));
    tritd_problem* p = new tritd_problem();
    p->ctx = c;
    p->level = level;
    const bool solver = level == kLevelSolver, contract = level <= kLevelContract;
    p->n1 = (int)n1; p->n2 = (int)n2; p->n3 = (int)n3; p->r = r; p->R = r * r;
    p->RS = (p->R + 7) / 8 * 8;
    const RankCfg rc = rank_cfg(r);
    p->NT = rc.NT; p->KS = rc.KS;
    p->n_it = (p->n1 + 127) / 128;
    p->n_jc = (p->n2 + kBoxRows - 1) / kBoxRows;
    const int nwr_ = (p->n1 + 15) / 16;                      // 16-row strips
    // k_admm stage size: two 8-column groups per stage, except on small slabs (few stages per CTA), where the finer
    // one-group stages balance the CTAs better and keep more loads in flight during the short kernel
    {
        const int nit_ = (nwr_ + 7) / 8;
        const long stages2 = (long)p->n_jc * p->n3 * 2 / std::max(1, c->num_sms / nit_);
        p->jgp = stages2 < 24 ? 1 : 2;
    }
    if (const char* e = getenv("TRITD_ADMM_JG")) p->jgp = atoi(e) == 1 ? 1 : 2;
    // Leading dimension: a multiple of 16 rows (padded rows exist and stay zero: TMA views the rows as (16, ld1/16)),
    // and DRAM-friendly: measured on B200 (profiles/r02_tile_experiments.md), the streaming kernels run 7-10 % faster when
    // every column starts on a 2 KB boundary than on a 128-byte one, with 256-byte alignment (multiples of 32 rows) in
    // between -- the 1 KB column chunks of an i-tile then never straddle a DRAM interleave block.  So: round up to 128
    // rows when that costs <= 1/8 extra memory, else to 32 rows.
    p->ld1 = 16 * nwr_;
    {
        const int l128 = (p->n1 + 127) & ~127, l32 = (p->n1 + 31) & ~31;
        if ((l128 - p->n1) * 8 <= p->n1) p->ld1 = l128;
        else if ((l32 - p->n1) * 8 <= p->n1) p->ld1 = l32;
    }
    if (const char* e = getenv("TRITD_LD1")) p->ld1 = std::max(16 * nwr_, atoi(e) & ~15);
    p->ldt = p->ld1;
    p->Np = (size_t)p->ld1 * p->n2 * p->n3;

    int s = TRITD_OK;
    auto bail = [&](int code) { tritd_problem_destroy(p); return code; };
#define PALLOC(ptr, count) if ((s = dalloc(p, &p->ptr, (count))) != TRITD_OK) return bail(s)
    if (solver) { PALLOC(D, p->Np); PALLOC(Z, p->Np); PALLOC(YL, p->Np); PALLOC(O, p->Np); }
    if (contract) PALLOC(T, p->Np);
    PALLOC(A1, (size_t)p->n1 * p->RS); PALLOC(B2, (size_t)p->n2 * p->RS); PALLOC(C3, (size_t)p->n3 * p->RS);
    PALLOC(A1T, (size_t)p->RS * p->ldt);
    PALLOC(SA, (size_t)p->RS * p->RS); PALLOC(SB, (size_t)p->RS * p->RS);
    PALLOC(gpart, (size_t)3 * kGramSlices * p->RS * p->RS);
    PALLOC(bufA, (size_t)p->n1 * p->RS + (size_t)p->RS * p->RS);
    PALLOC(rhsB, (size_t)p->n2 * p->RS); PALLOC(rhsC, (size_t)p->n3 * p->RS);
    if (contract) PALLOC(P, (size_t)p->n3 * p->n2 * p->RS);

    // grids: one contraction CTA per SM (its pipeline fills shared memory); the fused kernel by occupancy
    p->unitsM = (long)p->n_it * p->n_jc * p->n3;
    p->gridM = (int)std::min<long>(p->unitsM, std::max(2 * p->n_it, c->num_sms));
    p->unitsP = ((long)p->n3 * p->n_jc + kCW - 1) / kCW;
    p->gridP = (int)std::min<long>((long)p->n3 * p->n_jc, c->num_sms);       // row blocks are the unit of balance
    p->smemM = smem_mttkrp1(p->NT);
    p->smemP = smem_ppass(p->NT);
    int occ = 1;
#define CALL(NT_, KS_) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused<KS_, 1>, 256, 0));
    {
        auto q = [&]() -> int { TRITD_DISPATCH_R(r, CALL) return TRITD_OK; };
        if ((s = q()) != TRITD_OK) return bail(s);
    }
#undef CALL
    if (occ < 1) occ = 1;
    p->gi = std::max(1, c->num_sms * occ / p->n_it);
    p->gridF = p->gi * p->n_it;

    if (contract) PALLOC(partM, (size_t)p->gridM * 2 * 128 * p->RS);
    {
        PALLOC(Minv, (size_t)3 * p->RS * p->RS);
        PALLOC(flags, 16 + 3 * 64);
        PALLOC(ones, 64);
        { double h1[64]; for (double& x : h1) x = 1.0; cudaMemcpy(p->ones, h1, sizeof(h1), cudaMemcpyHostToDevice); }
        if (getenv("TRITD_DEBUG_STAMPS")) { PALLOC(dbg, 48); cudaMemset(p->dbg, 0, 48 * 8); }
        if (contract) {
            PALLOC(tile0, p->gridM); PALLOC(tile1, p->gridM);
            std::vector<int> t0(p->gridM), t1(p->gridM);
            const long per_it = (long)p->n_jc * p->n3;
            for (int q = 0; q < p->gridM; ++q) {
                const long u0 = p->unitsM * q / p->gridM, u1 = p->unitsM * (q + 1) / p->gridM;
                t0[q] = (int)(u0 / per_it);
                t1[q] = u1 > u0 ? (int)((u1 - 1) / per_it) : t0[q] - 1;
            }
            cudaMemcpy(p->tile0, t0.data(), sizeof(int) * p->gridM, cudaMemcpyHostToDevice);
            cudaMemcpy(p->tile1, t1.data(), sizeof(int) * p->gridM, cudaMemcpyHostToDevice);
        }
        cudaMemset(p->flags, 0, (16 + 3 * 64) * 4);
        // k_admm: one CTA per SM.  The i-tiles are as even as 16-row warp strips allow (240 rows -> 128 + 112,
        // 130 rows -> 80 + 50) and every tile gets the same number of CTAs.  (Handing a shallower last tile fewer
        // CTAs in proportion to its strips was tried twice and lost both times: 284 vs 267 us on 240 x 320 x 300 with four
        // state arrays, 212.9 vs 201.3 us with three -- the time of a stage does not follow its depth;
        // profiles/r02_tile_experiments.md, r02_zstate_experiments.md.)
        const int nwr = (p->n1 + 15) / 16;
        p->nitA = (nwr + 7) / 8;
        p->tileH = 16 * ((nwr + p->nitA - 1) / p->nitA);
        std::vector<int> per(p->nitA, std::max(c->num_sms / p->nitA, 1));
        p->gridA = 0; p->partSlots = 0;
        for (int x : per) { p->gridA += x; p->partSlots = std::max(p->partSlots, x); }
        if (solver) {
            PALLOC(partF, (size_t)p->nitA * p->partSlots * 128 * p->RS);
            PALLOC(ctaTab, 3 * p->gridA);
            PALLOC(tileCnt, p->nitA);
            std::vector<int> tab(3 * p->gridA);
            // interleave the tiles so neighbouring CTAs (launched together) work on the same columns
            std::vector<int> used(p->nitA, 0);
            int q = 0;
            for (int cta = 0; cta < p->gridA;) {
                if (used[q] < per[q]) { tab[3 * cta] = q; tab[3 * cta + 1] = used[q]++; tab[3 * cta + 2] = per[q]; ++cta; }
                q = (q + 1) % p->nitA;
            }
            cudaMemcpy(p->ctaTab, tab.data(), sizeof(int) * 3 * p->gridA, cudaMemcpyHostToDevice);
            cudaMemcpy(p->tileCnt, per.data(), sizeof(int) * p->nitA, cudaMemcpyHostToDevice);
            if (getenv("TRITD_DEBUG_STAMPS")) { PALLOC(dbgA, 2 * p->gridA); cudaMemset(p->dbgA, 0, 16 * p->gridA); }
        }
    }
    PALLOC(norm_part, (size_t)2 * std::max(std::max(p->gridF, c->num_sms), 1024));
    PALLOC(norms, 8);
    PALLOC(st, 1);
#undef PALLOC
    if (cudaMallocHost((void**)&p->st_host, sizeof(IterState)) != cudaSuccess) return bail(fail(TRITD_ERR_CUDA, "cudaMallocHost failed"));

    // opt in to the large dynamic shared-memory carve-outs once
    {
        auto q = [&]() -> int {
#define CALL(NT_, KS_)                                                                                            \
    CU_TRY(cudaFuncSetAttribute(k_mttkrp1<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smemM));    \
    CU_TRY(cudaFuncSetAttribute(k_ppass<NT_, ppass_scalar_col(NT_, KS_)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smemP));      \
    CU_TRY(cudaFuncSetAttribute(k_admm<KS_, NT_, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                (int)AdmmCfg<KS_, NT_, 2>::kSmem));                                               \
    CU_TRY(cudaFuncSetAttribute(k_admm<KS_, NT_, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                (int)AdmmCfg<KS_, NT_, 1>::kSmem));                                               \
    CU_TRY(cudaFuncSetAttribute(k_admm<KS_, NT_, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                (int)AdmmCfg<KS_, NT_, 2>::kSmem));                                               \
    CU_TRY(cudaFuncSetAttribute(k_admm<KS_, NT_, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                (int)AdmmCfg<KS_, NT_, 1>::kSmem));
            TRITD_DISPATCH_R(r, CALL)
#undef CALL
            const int usm = (int)upd_smem_bytes(p->RS);
            CU_TRY(cudaFuncSetAttribute(k_upd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
            CU_TRY(cudaFuncSetAttribute(k_upd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
            CU_TRY(cudaFuncSetAttribute(k_upd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
            CU_TRY(cudaFuncSetAttribute(k_upd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
            int occ = 0;
            switch ((p->R + 15) / 16) {
                case 1: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_upd<1>, kUpdThreads, (size_t)usm)); break;
                case 2: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_upd<2>, kUpdThreads, (size_t)usm)); break;
                case 3: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_upd<3>, kUpdThreads, (size_t)usm)); break;
                default: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_upd<4>, kUpdThreads, (size_t)usm)); break;
            }
            p->upd_wave = std::max(1, occ * c->num_sms);      // k_upd CTAs resident at once
            return TRITD_OK;
        };
        if ((s = q()) != TRITD_OK) return bail(s);
    }

    // zero everything once: pad rows / pad columns must be exact zeros forever
    cudaStream_t st = c->stream;
    for (double* q : {p->D, p->Z, p->YL, p->T, p->O})
        if (q && cudaMemsetAsync(q, 0, p->Np * sizeof(double), st) != cudaSuccess) return bail(fail(TRITD_ERR_CUDA, "memset failed"));
    cudaMemsetAsync(p->A1T, 0, (size_t)p->RS * p->ldt * sizeof(double), st);
    cudaMemsetAsync(p->st, 0, sizeof(IterState), st);

    // tensor maps: T as (i, j, t) with 128B-swizzled boxes [32 j][16 i]; A1T as (i, k) with boxes [RS k][16 i]
    {
        cuuint64_t dims[3] = {(cuuint64_t)p->n1, (cuuint64_t)p->n2, (cuuint64_t)p->n3};
        cuuint64_t str[2] = {(cuuint64_t)p->ld1 * 8, (cuuint64_t)p->ld1 * p->n2 * 8};
        cuuint32_t box[3] = {16, (cuuint32_t)kBoxRows, 1};
        if (contract && (s = make_map(c, &p->mapT, p->T, 3, dims, str, box)) != TRITD_OK) return bail(s);
        // k_admm views each N-array as (i_lo = 16, j, i_hi = ld1/16, t): one box = [8 i_hi][8 j][16 i_lo]
        cuuint64_t dims4[4] = {16, (cuuint64_t)p->n2, (cuuint64_t)(p->ld1 / 16), (cuuint64_t)p->n3};
        cuuint64_t str4[3] = {(cuuint64_t)p->ld1 * 8, 128, (cuuint64_t)p->ld1 * p->n2 * 8};
        int jgroups = 1;
        {
            auto q = [&]() -> int {
#define CALL(NT_, KS_) jgroups = p->jgp == 1 ? AdmmCfg<KS_, NT_, 1>::JG : AdmmCfg<KS_, NT_, 2>::JG;
                TRITD_DISPATCH_R(r, CALL)
#undef CALL
                return TRITD_OK;
            };
            if ((s = q()) != TRITD_OK) return bail(s);
        }
        cuuint32_t box4[4] = {16, (cuuint32_t)(8 * jgroups), (cuuint32_t)(p->tileH / 16), 1};
        const int strips = (p->n1 + 15) / 16, last_depth = strips - (p->nitA - 1) * (p->tileH / 16);
        cuuint32_t box4l[4] = {16, (cuuint32_t)(8 * jgroups), (cuuint32_t)last_depth, 1};
        struct { CUtensorMap* m; CUtensorMap* ml; double* base; } mm[5] = {
            {&p->maps.D, &p->mapsLast.D, p->D}, {&p->maps.YL, &p->mapsLast.YL, p->YL}, {&p->maps.Z, &p->mapsLast.Z, p->Z},
            {&p->maps.T, &p->mapsLast.T, p->T}, {&p->maps.O, &p->mapsLast.O, p->O}};
        for (auto& q : mm) {
            if (solver && (s = make_map(c, q.m, q.base, 4, dims4, str4, box4)) != TRITD_OK) return bail(s);
            if (solver && (s = make_map(c, q.ml, q.base, 4, dims4, str4, box4l)) != TRITD_OK) return bail(s);
        }
        cuuint64_t dims2[2] = {(cuuint64_t)p->n1, (cuuint64_t)p->RS};
        cuuint64_t str2[1] = {(cuuint64_t)p->ldt * 8};
        cuuint32_t box2[2] = {16, (cuuint32_t)p->RS};
        if ((s = make_map(c, &p->mapA1T, p->A1T, 2, dims2, str2, box2)) != TRITD_OK) return bail(s);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return bail(fail(TRITD_ERR_CUDA, "sync failed"));
    if (solver && !c->inproc && (s = setup_exchange(p)) != TRITD_OK) return bail(s);      // (groups: setup_exchange_inproc)
    *out = p;
    return TRITD_OK;
}

// dense column-major (ld = n1) <-> padded (ld = ld1) copies; kind = cudaMemcpyDefault works for host and device
static int copy_in(tritd_problem* p, double* dst_padded, const double* src_dense) {
    cudaStream_t st = p->ctx->stream;
    if (p->ld1 == p->n1) CU_TRY(cudaMemcpyAsync(dst_padded, src_dense, p->Np * sizeof(double), cudaMemcpyDefault, st));
    else CU_TRY(cudaMemcpy2DAsync(dst_padded, (size_t)p->ld1 * 8, src_dense, (size_t)p->n1 * 8, (size_t)p->n1 * 8,
                                  (size_t)p->n2 * p->n3, cudaMemcpyDefault, st));
    return TRITD_OK;
}
static int copy_out(tritd_problem* p, double* dst_dense, const double* src_padded) {
    cudaStream_t st = p->ctx->stream;
    if (p->ld1 == p->n1) CU_TRY(cudaMemcpyAsync(dst_dense, src_padded, p->Np * sizeof(double), cudaMemcpyDefault, st));
    else CU_TRY(cudaMemcpy2DAsync(dst_dense, (size_t)p->n1 * 8, src_padded, (size_t)p->ld1 * 8, (size_t)p->n1 * 8,
                                  (size_t)p->n2 * p->n3, cudaMemcpyDefault, st));
    return TRITD_OK;
}

extern "C" int tritd_problem_set_D_host(tritd_problem* p, const double* D_host) {
    if (!p || !D_host) return fail(TRITD_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(copy_in(p, p->D, D_host));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
    p->has_D = true; p->initialized = false; p->masked = false;
    return TRITD_OK;
}
extern "C" int tritd_problem_set_D_dev(tritd_problem* p, const double* D_dev) {
    if (!p || !D_dev) return fail(TRITD_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(copy_in(p, p->D, D_dev));
    p->has_D = true; p->initialized = false; p->masked = false;
    return TRITD_OK;
}

// Completion variant (opt-in): mark the entries with mask == 0 as unobserved.  The mask is folded into
// D itself (NaN), so the iteration moves not one byte more than the unmasked solver.
static int apply_mask(tritd_problem* p, const unsigned char* mask, bool on_host) {
    if (!p || !mask) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (!p->has_D) return fail(TRITD_ERR_INVALID, "tritd_problem_set_D_* must be called first");
    tritd_ctx* c = p->ctx;
    CU_TRY(cudaSetDevice(c->device));
    const size_t ncols = (size_t)p->n2 * p->n3, nm = (size_t)p->n1 * ncols;
    unsigned char* md = nullptr;
    if (on_host) {
        CU_TRY(cudaMalloc((void**)&md, nm));
        if (cudaMemcpyAsync(md, mask, nm, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { cudaFree(md); return fail(TRITD_ERR_CUDA, "mask copy failed"); }
    }
    const unsigned grid = (unsigned)std::min<size_t>((nm + 255) / 256, (size_t)c->num_sms * 16);
    k_apply_mask<<<grid, 256, 0, c->stream>>>(p->D, on_host ? md : mask, p->n1, p->ld1, ncols);
    const cudaError_t e = cudaGetLastError();
    c->launches += 1;
    cudaStreamSynchronize(c->stream);
    if (md) cudaFree(md);
    if (e != cudaSuccess) return fail(TRITD_ERR_CUDA, "k_apply_mask: %s", cudaGetErrorString(e));
    p->masked = true; p->initialized = false;
    return TRITD_OK;
}
extern "C" int tritd_problem_set_mask_host(tritd_problem* p, const unsigned char* mask_host) { return apply_mask(p, mask_host, true); }
extern "C" int tritd_problem_set_mask_dev(tritd_problem* p, const unsigned char* mask_dev) { return apply_mask(p, mask_dev, false); }

// MATLAB 3-D factor shapes <-> row-major n x RS "unfolded" layout (reshape_*_from_*, :111-130, run backwards)
static void pack_A(const double* A, int n1, int R, int RS, std::vector<double>& out) {
    out.assign((size_t)n1 * RS, 0.0);
    for (int k = 0; k < R; ++k) for (int i = 0; i < n1; ++i) out[(size_t)i * RS + k] = A[(size_t)k * n1 + i];
}
static void pack_B(const double* B, int n2, int r, int RS, std::vector<double>& out) {
    out.assign((size_t)n2 * RS, 0.0);   // B(p,j,s) at p + r*(j + n2*s)  ->  B2[j][p + r*s]
    for (int s = 0; s < r; ++s) for (int j = 0; j < n2; ++j) for (int q = 0; q < r; ++q)
        out[(size_t)j * RS + q + r * s] = B[(size_t)q + (size_t)r * (j + (size_t)n2 * s)];
}
static void pack_C(const double* C, int n3, int R, int RS, std::vector<double>& out) {
    out.assign((size_t)n3 * RS, 0.0);   // C(p,s,t) at (p + r*s) + R*t  ->  C3[t][p + r*s]
    for (int t = 0; t < n3; ++t) for (int k = 0; k < R; ++k) out[(size_t)t * RS + k] = C[(size_t)t * R + k];
}
static void unpack_A(const std::vector<double>& in, int n1, int R, int RS, double* A) {
    for (int k = 0; k < R; ++k) for (int i = 0; i < n1; ++i) A[(size_t)k * n1 + i] = in[(size_t)i * RS + k];
}
static void unpack_B(const std::vector<double>& in, int n2, int r, int RS, double* B) {
    for (int s = 0; s < r; ++s) for (int j = 0; j < n2; ++j) for (int q = 0; q < r; ++q)
        B[(size_t)q + (size_t)r * (j + (size_t)n2 * s)] = in[(size_t)j * RS + q + r * s];
}
static void unpack_C(const std::vector<double>& in, int n3, int R, int RS, double* C) {
    for (int t = 0; t < n3; ++t) for (int k = 0; k < R; ++k) C[(size_t)t * R + k] = in[(size_t)t * RS + k];
}

static int upload_factors(tritd_problem* p, const double* A0, const double* B0, const double* C0) {
    cudaStream_t st = p->ctx->stream;
    std::vector<double> ha, hb, hc;
    pack_A(A0, p->n1, p->R, p->RS, ha);
    pack_B(B0, p->n2, p->r, p->RS, hb);
    pack_C(C0, p->n3, p->R, p->RS, hc);
    CU_TRY(cudaMemcpyAsync(p->A1, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(p->B2, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(p->C3, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));           // (the staging vectors die here)
    return TRITD_OK;
}

// tritd_problem_init in three steps so that a group of member problems (one per device) can run the steps in
// lock-step with a host-side sum in between: (1) local state + partial ||D||^2, (2) the sum over the ranks,
// (3) normD, the small Grams, epochs.
static int problem_init_local(tritd_problem* p, const tritd_opts* o, const double* A0, const double* B0, const double* C0) {
    if (!p || !o || !A0 || !B0 || !C0) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (p->level != kLevelSolver) return fail(TRITD_ERR_INVALID, "not a solver problem");
    if (!p->has_D) return fail(TRITD_ERR_INVALID, "tritd_problem_set_D_* must be called first");
    if (o->maxIter < 1) return fail(TRITD_ERR_INVALID, "opts.maxIter = %d", o->maxIter);
    if (!(o->mu > 0.0)) return fail(TRITD_ERR_INVALID, "opts.mu must be positive");
    tritd_ctx* c = p->ctx;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    p->opts = *o;
    ST_TRY(upload_factors(p, A0, B0, C0));

    // O = E = Y_L = Y_O = 0 (:24-26); the first target T = D - O + (1/muL)*Y_L is D itself (:33)
    for (double* q : {p->Z, p->YL, p->O}) CU_TRY(cudaMemsetAsync(q, 0, p->Np * sizeof(double), st));
    if (p->masked) {      // completion variant: the first target is the zero-filled data (the callers' own convention, traffic_triple_comparison.m:34-35)
        k_fill_unobserved<<<(unsigned)std::min<size_t>((p->Np + 255) / 256, (size_t)c->num_sms * 16), 256, 0, st>>>(p->D, p->T, p->Np);
        CU_TRY(cudaGetLastError());
        c->launches += 1;
    } else {
        CU_TRY(cudaMemcpyAsync(p->T, p->D, p->Np * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }

    IterState h;
    memset(&h, 0, sizeof(h));
    h.muL = o->mu; h.muO = o->mu; h.muL_max = o->mu * 1e6; h.muO_max = o->mu * 1e6;
    h.rhoL = o->rho; h.rhoO = o->rho; h.lambda = o->lambda_; h.tol = o->tol; h.normD = 0.0;
    h.k = 0; h.stop = 0; h.status = 0; h.maxIter = o->maxIter; h.masked = p->masked ? 1 : 0;
    iter_state_derive(h);
    h.muO_prev = h.muO; h.thr_prev = h.thr;      // (Z = 0: E = Y_O = 0 whatever these are)
    *p->st_host = h;
    CU_TRY(cudaMemcpyAsync(p->st, p->st_host, sizeof(IterState), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(p->flags, 0, (16 + 3 * 64) * 4, st));

    // history buffers sized by maxIter
    if (o->maxIter > p->hist_cap) {
        double* q = nullptr;
        ST_TRY(dalloc(p, &q, (size_t)3 * o->maxIter));
        p->errHist = q;
        p->hist_cap = o->maxIter;
    }
    p->errL = p->errHist + p->hist_cap; p->errO = p->errHist + 2 * (size_t)p->hist_cap;
    CU_TRY(cudaMemsetAsync(p->errHist, 0, (size_t)3 * p->hist_cap * 8, st));

    // partial of normD^2 = sum(D(:).^2)  (:28) over this rank's slab
    k_sumsq_part<<<1024, 256, 0, st>>>(p->D, p->Np, p->norm_part, p->masked ? 1 : 0);
    k_sum_pairs<<<1, 256, 0, st>>>(p->norm_part, 1024, p->norms, nullptr);
    CU_TRY(cudaGetLastError());
    c->launches += 2;
    return TRITD_OK;
}

static int problem_init_finish(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    k_set_normD<<<1, 32, 0, st>>>(p->st, p->norms);
    c->launches += 1;
    // small Grams of the initial factors: SB = B2'B2, SC (partial over local rows) = C3'C3
    ST_TRY(launch_small_gram(p, p->B2, p->n2, p->SB));
    ST_TRY(launch_small_gram(p, p->C3, p->n3, p->bufA + (size_t)p->n1 * p->RS));
    p->xbase = c->xepoch;                 // epochs of this solve: xbase (initial C3'C3), xbase + 1 .. xbase + maxIter (same on every rank)
    c->xepoch += (unsigned)p->opts.maxIter + 1u;
    if (p->xchg) {
        // the local partial of C3'C3 goes to every mailbox like update C's later ones (parity 1, epoch xbase)
        k_push_sc<<<1, 256, 0, st>>>(p->bufA + (size_t)p->n1 * p->RS, p->peers, (long)(p->offS + ((size_t)c->nranks + c->rank) * p->RS * p->RS),
                                    (long)p->offF, (long)(kFlagsSC + 8 + c->rank), c->nranks, p->RS * p->RS, p->xbase);
        CU_TRY(cudaGetLastError());
        c->launches += 1;
    }
    CU_TRY(cudaStreamSynchronize(st));
    p->initialized = true;
    p->rhsA_ready = false;
    p->pre_inv = !(c->nranks > 1 && !p->xchg) && !getenv("TRITD_NO_PREINV");       // (NCCL path: C3'C3 arrives only with update A's all-reduce)
    p->graph_off = getenv("TRITD_NO_GRAPH") != nullptr;
    if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }     // opts (lambda2) are baked into the graphs
    if (p->graphN) { cudaGraphExecDestroy(p->graphN); p->graphN = nullptr; }
    p->printed_k = 0;
    return TRITD_OK;
}

extern "C" int tritd_problem_init(tritd_problem* p, const tritd_opts* o, const double* A0, const double* B0,
                                  const double* C0) {
    if (p && p->ctx->inproc) return fail(TRITD_ERR_UNSUPPORTED, "member problems of a multi-device context are driven by tritd_admm_f64");
    ST_TRY(problem_init_local(p, o, A0, B0, C0));
    ST_TRY(allreduce_sum(p->ctx, p->norms, 2));         // ||D||^2 over the slabs (one process per GPU: NCCL)
    return problem_init_finish(p);
}

// One ADMM iteration, enqueued on the context's stream (triple_decomp_ADMM.m:31-66).
static int enqueue_iteration(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    cudaStream_t st = c->stream;
    double* rhsA = p->bufA;
    double* SC = p->bufA + (size_t)p->n1 * p->RS;
    auto mark = [&]() -> int {           // phase boundary (only when profiling)
        if (!p->profiling) return TRITD_OK;
        cudaEvent_t e;
        CU_TRY(cudaEventCreate(&e));
        CU_TRY(cudaEventRecord(e, st));
        p->prof_ev.push_back(e);
        return TRITD_OK;
    };
    ST_TRY(mark());

    const bool multi = c->nranks > 1;
    // update_A (:73-81): RHS = X1*F', Gram = (B2'B2) o (C3'C3) + lambda2*I.  From the second iteration on the
    // previous k_admm left per-CTA partials of X1*F' (accumulated from registers); single rank: k_upd sums them,
    // applies the inverse and forms A1'A1 in one launch.
    const bool xc = p->xchg;             // N>1: partials travel through the peer mailboxes (else NCCL all-reduces)
    const size_t nA = (size_t)p->n1 * p->RS + (size_t)p->RS * p->RS;
    const bool direct_A = !p->rhsA_ready || (multi && !xc);
    if (!p->rhsA_ready) ST_TRY(launch_mttkrp1(p, p->mapT, p->B2, p->C3, rhsA));
    else if (multi && !xc) ST_TRY(launch_upd(p, 0, kSrcPartF, false, nullptr, rhsA, nullptr, nullptr, 0.0, nullptr, nullptr, p->n1, nullptr));
    ST_TRY(mark());
    if (multi && !xc) ST_TRY(allreduce_sum(c, p->bufA, nA));
    // (from the second iteration on the previous k_admm has already inverted update A's ridge system)
    ST_TRY(launch_upd(p, 0, direct_A ? kSrcDirect : kSrcPartF, true, rhsA, nullptr, p->SB, SC, p->opts.lambda2, p->A1,
                      p->A1T, p->n1, p->SA, !(p->pre_inv && p->rhsA_ready)));
    ST_TRY(mark());

    // update_B (:83-88) with the new A: RHS = X2*G' = sum_t C3(t,:) .* P(t,j,:), Gram = (A1'A1) o (C3'C3) + lambda2*I
    ST_TRY(launch_ppass(p, p->mapT, p->pre_inv));         // (its first CTA to finish inverts update B's ridge system)
    ST_TRY(mark());
    if (multi && !xc) {
        ST_TRY(launch_upd(p, 1, kSrcPB, false, nullptr, p->rhsB, nullptr, nullptr, 0.0, nullptr, nullptr, p->n2, nullptr));
        ST_TRY(allreduce_sum(c, p->rhsB, (size_t)p->n2 * p->RS));
    }
    ST_TRY(launch_upd(p, 1, multi && !xc ? kSrcDirect : kSrcPB, true, p->rhsB, nullptr, p->SA, SC, p->opts.lambda2, p->B2,
                      nullptr, p->n2, p->SB, !p->pre_inv));

    // update_C (:90-95) with the new A, B: slice-local; ridge fixed at 1e-9.  Leaves SC = C3'C3 over the local
    // slices in bufA, where the next exchange sums it over the ranks.
    ST_TRY(launch_upd(p, 2, kSrcPC, true, nullptr, nullptr, p->SA, p->SB, 1e-9, p->C3, nullptr, p->n3, SC));

    ST_TRY(mark());
    // L, O, E, duals, next T, residual norms (:38-59, :33); single rank: its last CTA also does :56-65
    ST_TRY(launch_admm(p));
    p->rhsA_ready = true;
    ST_TRY(mark());
    if (xc) {
        // nothing: k_admm's last CTA exchanged the residual sums and finalised the iteration
    } else if (multi) {
        ST_TRY(allreduce_sum(c, p->norms, 2));
        k_finalize<<<1, 256, 0, st>>>(p->st, p->norm_part, p->gridA, p->norms, 1, p->errHist, p->errL, p->errO);
        CU_TRY(cudaGetLastError());
        c->launches += 1;
    }
    ST_TRY(mark());
    return TRITD_OK;
}

// diagnostics: the globaltimer stamps k_upd leaves when TRITD_DEBUG_STAMPS is set (not part of the public header)
extern "C" int tritd_debug_stamps(tritd_problem* p, long long* out48) {
    if (!p || !p->dbg) return fail(TRITD_ERR_INVALID, "no debug stamps");
    CU_TRY(cudaMemcpy(out48, p->dbg, 48 * 8, cudaMemcpyDeviceToHost));
    return TRITD_OK;
}

// diagnostics: start / end globaltimer stamps of every k_admm CTA of the last launch + its (tile, index, CTAs of tile)
extern "C" int tritd_debug_admm_stamps(tritd_problem* p, long long* out, int* tab, int cap) {
    if (!p || !p->dbgA || cap < p->gridA) return fail(TRITD_ERR_INVALID, "no k_admm stamps (TRITD_DEBUG_STAMPS) or buffer too small");
    CU_TRY(cudaMemcpy(out, p->dbgA, 16 * (size_t)p->gridA, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(tab, p->ctaTab, 12 * (size_t)p->gridA, cudaMemcpyDeviceToHost));
    return p->gridA;
}

constexpr int kGraphBatch = 10;      // iterations per replayed graph (the cadence of the host's look at the stopping rule)
static int capture_graph(tritd_problem* p, int iters, cudaGraphExec_t* out, bool* failed) {
    tritd_ctx* c = p->ctx;
    const int64_t l0 = c->launches;
    cudaGraph_t g = nullptr;
    *failed = true;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return TRITD_OK; }
    int s = TRITD_OK;
    for (int i = 0; i < iters && s == TRITD_OK; ++i) s = enqueue_iteration(p);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    p->graph_launches = (int)((c->launches - l0) / iters);
    c->launches = l0;
    if (s != TRITD_OK || e != cudaSuccess || !g || cudaGraphInstantiate(out, g, 0) != cudaSuccess) {
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        *out = nullptr;
        return TRITD_OK;
    }
    cudaGraphDestroy(g);
    if (cudaGraphUpload(*out, c->stream) != cudaSuccess) cudaGetLastError();     // (so that the first replay is as cheap as the later ones)
    *failed = false;
    return TRITD_OK;
}

// `n` iterations: the first one (and profiled ones) as plain launches, every later one as a replay of a CUDA graph
// captured from the very same enqueue_iteration() -- all iteration state lives in device memory, so the graphs
// need no parameter updates.  Whole batches of kGraphBatch iterations replay ONE graph, inside which every kernel is
// a programmatic dependent of its predecessor (also across the iteration boundary).
static int run_iterations(tritd_problem* p, int n, bool eager = false) {
    tritd_ctx* c = p->ctx;
    if (c->inproc) CU_TRY(cudaSetDevice(c->device));      // members of a group are driven by one host thread
    while (n > 0) {
        if (p->profiling || !p->rhsA_ready || p->graph_off) { ST_TRY(enqueue_iteration(p)); --n; continue; }
        const bool batch = n >= kGraphBatch;
        cudaGraphExec_t& ge = batch ? p->graphN : p->graph;
        if (!ge) {
            bool failed = false;
            ST_TRY(capture_graph(p, batch ? kGraphBatch : 1, &ge, &failed));
            if (failed) { p->graph_off = true; continue; }
        }
        if (eager && !p->graphN) {
            // staged API (tritd_problem_enqueue): build the batch graph together with the first one, so that a caller who
            // times a later enqueue (bench.py after its warm-up) never has the capture inside its timed region
            bool failed = false;
            ST_TRY(capture_graph(p, kGraphBatch, &p->graphN, &failed));
            if (failed) { p->graph_off = true; continue; }
        }
        CU_TRY(cudaGraphLaunch(ge, c->stream));
        c->launches += (int64_t)p->graph_launches * (batch ? kGraphBatch : 1);
        n -= batch ? kGraphBatch : 1;
    }
    return TRITD_OK;
}

extern "C" int tritd_problem_set_profiling(tritd_problem* p, int enable) {
    if (!p) return fail(TRITD_ERR_INVALID, "NULL problem");
    p->profiling = enable != 0;
    return TRITD_OK;
}

extern "C" int tritd_problem_phase_ms(tritd_problem* p, double* ms_out, int32_t* iters_out) {
    if (!p || !ms_out) return fail(TRITD_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(p->ctx->device));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
    const size_t per = TRITD_NPHASE + 1, n = p->prof_ev.size() / per;
    for (int q = 0; q < TRITD_NPHASE; ++q) ms_out[q] = 0.0;
    for (size_t it = 0; it < n; ++it)
        for (int q = 0; q < TRITD_NPHASE; ++q) {
            float ms = 0.f;
            CU_TRY(cudaEventElapsedTime(&ms, p->prof_ev[it * per + q], p->prof_ev[it * per + q + 1]));
            ms_out[q] += ms;
        }
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    p->prof_ev.clear();
    if (iters_out) *iters_out = (int32_t)n;
    return TRITD_OK;
}

static int fetch_state(tritd_problem* p) {
    CU_TRY(cudaMemcpyAsync(p->st_host, p->st, sizeof(IterState), cudaMemcpyDeviceToHost, p->ctx->stream));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
    if (p->st_host->status == kStatusNumeric)
        return fail(TRITD_ERR_NUMERIC, "ridge system contains NaN / Inf at iteration %d (pinv: input must not contain NaN or Inf)",
                    p->st_host->k + 1);
    if (p->st_host->status == kStatusHang)
        return fail(TRITD_ERR_TIMEOUT, "a device-side wait (inter-CTA hand-shake or peer exchange) timed out at iteration %d: a peer rank died or the exchange is broken",
                    p->st_host->k + 1);
    return TRITD_OK;
}

extern "C" int tritd_problem_enqueue(tritd_problem* p, int32_t n) {
    if (!p || !p->initialized) return fail(TRITD_ERR_INVALID, "problem not initialised");
    CU_TRY(cudaSetDevice(p->ctx->device));
    return run_iterations(p, n, true);
}

extern "C" int tritd_problem_sync(tritd_problem* p) {
    if (!p) return fail(TRITD_ERR_INVALID, "NULL problem");
    CU_TRY(cudaSetDevice(p->ctx->device));
    return fetch_state(p);
}

static int print_progress(tritd_problem* p) {
    // "Iter %d, errL=%.2e, errO=%.2e" every 10th iteration (:60-62)
    const int k = p->st_host->k;
    if (!p->opts.disp || k / 10 == p->printed_k / 10) { p->printed_k = k; return TRITD_OK; }
    std::vector<double> eL(k), eO(k);
    CU_TRY(cudaMemcpy(eL.data(), p->errL, (size_t)k * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(eO.data(), p->errO, (size_t)k * 8, cudaMemcpyDeviceToHost));
    for (int q = p->printed_k / 10 * 10 + 10; q <= k; q += 10) emit("Iter %d, errL=%.2e, errO=%.2e\n", q, eL[q - 1], eO[q - 1]);
    p->printed_k = k;
    return TRITD_OK;
}

// Drive one problem -- or the member problems of a group (one per device, rank order), which iterate in lock-step:
// their kernels wait for one another through the peer mailboxes, so every iteration is enqueued on all devices
// before the host ever blocks.
// The stopping rule lives on the device (later launches of a stopped solve are no-ops).  The host follows it
// WITHOUT stalling the GPUs: after every batch of <= 10 iterations (the cadence of the reference's progress line)
// the iteration scalars of rank 0 are copied to a pinned slot, and the host looks at the slot of the batch BEFORE
// the one it has just enqueued -- the streams never run dry, and with one process per GPU no rank waits for its
// host while its peers spin in an exchange.  At most one batch of no-op launches follows a stop.  (All ranks take
// bitwise identical stopping decisions, so rank 0's scalars stand for all.)
static int iterate_many(const std::vector<tritd_problem*>& ps, int32_t max_more, int32_t* iters_total) {
    tritd_problem* p = ps[0];
    tritd_ctx* c = p->ctx;
    for (tritd_problem* q : ps) {
        if (!q || !q->initialized) return fail(TRITD_ERR_INVALID, "problem not initialised");
        CU_TRY(cudaSetDevice(q->ctx->device));
        ST_TRY(fetch_state(q));
    }
    CU_TRY(cudaSetDevice(c->device));
    int remaining = std::min<int>(max_more, p->opts.maxIter - p->st_host->k);
    if (!p->poll) {
        CU_TRY(cudaMallocHost((void**)&p->poll, 2 * sizeof(IterState)));
        CU_TRY(cudaEventCreateWithFlags(&p->poll_ev[0], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&p->poll_ev[1], cudaEventDisableTiming));
    }
    bool pending[2] = {false, false};
    int k_enq = p->st_host->k, slot = 0;
    auto look = [&](int sl) -> int {
        CU_TRY(cudaEventSynchronize(p->poll_ev[sl]));
        pending[sl] = false;
        *p->st_host = p->poll[sl];
        if (p->st_host->status == kStatusNumeric)
            return fail(TRITD_ERR_NUMERIC, "ridge system contains NaN / Inf at iteration %d (pinv: input must not contain NaN or Inf)",
                        p->st_host->k + 1);
        if (p->st_host->status == kStatusHang)
            return fail(TRITD_ERR_TIMEOUT, "a device-side wait timed out at iteration %d", p->st_host->k + 1);
        return print_progress(p);
    };
    while (remaining > 0 && !p->st_host->stop) {
        const int batch = std::min(remaining, 10 - k_enq % 10);
        if (ps.size() == 1) ST_TRY(run_iterations(p, batch));
        else
            for (int i = 0; i < batch; ++i)               // groups: iteration by iteration across the devices
                for (tritd_problem* q : ps) ST_TRY(run_iterations(q, 1));
        k_enq += batch; remaining -= batch;
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(cudaMemcpyAsync(&p->poll[slot], p->st, sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaEventRecord(p->poll_ev[slot], c->stream));
        pending[slot] = true;
        if (pending[slot ^ 1]) ST_TRY(look(slot ^ 1));
        slot ^= 1;
    }
    for (int q = 0; q < 2; ++q) { const int sl = (slot + q) & 1; if (pending[sl]) ST_TRY(look(sl)); }   // oldest first
    for (size_t g = ps.size(); g-- > 0;) {                 // rank 0 last: its device is current afterwards
        CU_TRY(cudaSetDevice(ps[g]->ctx->device));
        ST_TRY(fetch_state(ps[g]));
    }
    ST_TRY(print_progress(p));
    if (iters_total) *iters_total = p->st_host->k;
    return TRITD_OK;
}

extern "C" int tritd_problem_iterate(tritd_problem* p, int32_t max_more, int32_t* iters_total) {
    if (!p || !p->initialized) return fail(TRITD_ERR_INVALID, "problem not initialised");
    if (p->ctx->inproc) return fail(TRITD_ERR_UNSUPPORTED, "member problems of a multi-device context are driven by tritd_admm_f64");
    return iterate_many(std::vector<tritd_problem*>{p}, max_more, iters_total);
}

// O is not written inside the loop (see k_recover_O): materialise it into p->O on demand.
static int recover_O(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    const unsigned grid = (unsigned)std::min<size_t>((p->Np + 255) / 256, (size_t)c->num_sms * 16);
    k_recover_O<<<grid, 256, 0, c->stream>>>(p->D, p->T, p->YL, p->st, p->O, p->Np);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

static int reconstruct_to(tritd_problem* p, double* dst, int ld);
static int reconstruct_L(tritd_problem* p, double** Lbuf) {
    double* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, p->Np * sizeof(double));
    if (e != cudaSuccess) return fail(TRITD_ERR_CUDA, "cudaMalloc(L): %s", cudaGetErrorString(e));
    int s = reconstruct_to(p, q, p->ld1);          // writes every entry of the padded array (pad rows: zeros)
    if (s != TRITD_OK) { cudaFree(q); return s; }
    *Lbuf = q;
    return TRITD_OK;
}

extern "C" int tritd_problem_get(tritd_problem* p, double* A, double* B, double* C, double* O, double* L,
                                 double* errHist, double* errL, double* errO, int32_t* iters) {
    if (!p || !p->initialized) return fail(TRITD_ERR_INVALID, "problem not initialised");
    tritd_ctx* c = p->ctx;
    CU_TRY(cudaSetDevice(c->device));
    ST_TRY(fetch_state(p));
    const int k = p->st_host->k;
    std::vector<double> h;
    if (A) { h.resize((size_t)p->n1 * p->RS); CU_TRY(cudaMemcpy(h.data(), p->A1, h.size() * 8, cudaMemcpyDeviceToHost)); unpack_A(h, p->n1, p->R, p->RS, A); }
    if (B) { h.resize((size_t)p->n2 * p->RS); CU_TRY(cudaMemcpy(h.data(), p->B2, h.size() * 8, cudaMemcpyDeviceToHost)); unpack_B(h, p->n2, p->r, p->RS, B); }
    if (C) { h.resize((size_t)p->n3 * p->RS); CU_TRY(cudaMemcpy(h.data(), p->C3, h.size() * 8, cudaMemcpyDeviceToHost)); unpack_C(h, p->n3, p->R, p->RS, C); }
    if (O) { ST_TRY(recover_O(p)); ST_TRY(copy_out(p, O, p->O)); }
    if (L) {
        double* Lbuf = nullptr;
        ST_TRY(reconstruct_L(p, &Lbuf));
        int s = copy_out(p, L, Lbuf);
        cudaStreamSynchronize(c->stream);
        cudaFree(Lbuf);
        ST_TRY(s);
    }
    if (errHist && k) CU_TRY(cudaMemcpyAsync(errHist, p->errHist, (size_t)k * 8, cudaMemcpyDeviceToHost, c->stream));
    if (errL && k) CU_TRY(cudaMemcpyAsync(errL, p->errL, (size_t)k * 8, cudaMemcpyDeviceToHost, c->stream));
    if (errO && k) CU_TRY(cudaMemcpyAsync(errO, p->errO, (size_t)k * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (iters) *iters = k;
    return TRITD_OK;
}

extern "C" int tritd_problem_get_O_dev(tritd_problem* p, double* O_dev) {
    if (!p || !O_dev) return fail(TRITD_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(recover_O(p));
    return copy_out(p, O_dev, p->O);
}
extern "C" int tritd_problem_get_L_dev(tritd_problem* p, double* L_dev) {
    if (!p || !L_dev || !p->initialized) return fail(TRITD_ERR_INVALID, "NULL argument / not initialised");
    CU_TRY(cudaSetDevice(p->ctx->device));
    double* Lbuf = nullptr;
    ST_TRY(reconstruct_L(p, &Lbuf));
    int s = copy_out(p, L_dev, Lbuf);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(Lbuf);
    return s;
}

// E of the last finished iteration (the reference's header documents "O,E : sparse components (clone E)", :12).
// The loop keeps Z = R3 instead of the pair (E, Y_O) (k_admm): E = soft_threshold(Z, lambda/muO) is materialised into
// the p->O scratch on demand (stream order keeps an earlier copy of O out of it intact).
static int materialise_E(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    ST_TRY(fetch_state(p));              // (the scalars of the last finished iteration are final)
    const unsigned grid = (unsigned)std::min<size_t>((p->Np + 255) / 256, (size_t)c->num_sms * 16);
    k_E_from_Z<<<grid, 256, 0, c->stream>>>(p->Z, p->st, p->O, p->Np);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}
extern "C" int tritd_problem_get_E(tritd_problem* p, double* E_host) {
    if (!p || !E_host || !p->initialized) return fail(TRITD_ERR_INVALID, "NULL argument / not initialised");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(materialise_E(p));
    ST_TRY(copy_out(p, E_host, p->O));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
    return TRITD_OK;
}
extern "C" int tritd_problem_get_E_dev(tritd_problem* p, double* E_dev) {
    if (!p || !E_dev || !p->initialized) return fail(TRITD_ERR_INVALID, "NULL argument / not initialised");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(materialise_E(p));
    return copy_out(p, E_dev, p->O);
}
// How often the ridge solves of this solve took the truncating pseudo-inverse path and how many singular values
// that zeroed in total (what MATLAB's pinv does silently at :78/:86/:93).
extern "C" int tritd_problem_pinv_stats(tritd_problem* p, int32_t* fallbacks, int32_t* truncated) {
    if (!p) return fail(TRITD_ERR_INVALID, "NULL problem");
    CU_TRY(cudaSetDevice(p->ctx->device));
    ST_TRY(fetch_state(p));
    if (fallbacks) *fallbacks = p->st_host->pinv_fallbacks;
    if (truncated) *truncated = p->st_host->pinv_truncated;
    return TRITD_OK;
}

// ---------------------------------------------------------------------------
// one-call solver
// ---------------------------------------------------------------------------
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Multi-rank: may the cached problem be reused?  The decision must be the same on every rank -- creating a problem
// is collective (NCCL all-gather of the IPC handles), and a rank that frees its mailbox while a peer still has it
// mapped would be written into after the free.  So the ranks agree (NCCL min) on "every rank has a usable cache hit".
static int agree_on_cache_hit(tritd_ctx* c, bool local_hit, bool* all_hit) {
    *all_hit = local_hit;
    if (c->nranks == 1) return TRITD_OK;
    double* flag = nullptr;
    CU_TRY(cudaMalloc((void**)&flag, 8));
    const double mine = local_hit ? 1.0 : 0.0;
    double got = 0.0;
    int s = TRITD_OK;
    if (cudaMemcpyAsync(flag, &mine, 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) s = fail(TRITD_ERR_CUDA, "cache flag upload failed");
    if (s == TRITD_OK) {
        ncclResult_t r = g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclMin, c->comm, c->stream);
        if (r != ncclSuccess) s = fail(TRITD_ERR_NCCL, "ncclAllReduce(cache flag): %s", g_nccl.GetErrorString(r));
    }
    if (s == TRITD_OK && (cudaMemcpyAsync(&got, flag, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                          cudaStreamSynchronize(c->stream) != cudaSuccess))
        s = fail(TRITD_ERR_CUDA, "cache flag download failed");
    cudaFree(flag);
    *all_hit = got > 0.5;
    return s;
}

static int admm_group(tritd_ctx* G, const double* D_host, const unsigned char* mask_host, int64_t n1, int64_t n2, int64_t n3, int r,
                      const tritd_opts* o, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                      double* O, double* E, double* L, double* errHist, int32_t* iters_out, tritd_timing* tm);

static int admm_impl(tritd_ctx* c, const double* D_host, const unsigned char* mask_host, int64_t n1, int64_t n2, int64_t n3, int r,
                     const tritd_opts* o, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                     double* O, double* E, double* L, double* errHist, int32_t* iters_out, tritd_timing* tm) {
    if (!c || !D_host || !o || !A0 || !B0 || !C0 || !errHist) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (!c->sub.empty())
        return admm_group(c, D_host, mask_host, n1, n2, n3, r, o, A0, B0, C0, A, B, C, O, E, L, errHist, iters_out, tm);
    const double t_begin = now_ms();
    const int64_t launches0 = c->launches;
    // device state (6 N-sized arrays, tensor maps, the captured iteration graph) is kept in the context and
    // reused when the next call has the same shape and rank -- a MATLAB session typically calls the solver
    // repeatedly on equally sized data; tritd_trim() / tritd_destroy() release it
    tritd_problem* p = c->cached;
    const bool local_hit = p && !c->cache_poisoned && p->n1 == n1 && p->n2 == n2 && p->n3 == n3 && p->r == r;
    bool hit = local_hit;
    ST_TRY(agree_on_cache_hit(c, local_hit, &hit));
    if (!hit) { tritd_trim(c); p = nullptr; }
    if (!p) {
        ST_TRY(tritd_problem_create(c, n1, n2, n3, r, &p));
        c->cached = p;
        c->cache_poisoned = false;
    }
    // a failure below leaves the cached state in place but poisoned: the next call re-creates it on EVERY rank
    // (the decision above is collective), instead of this rank alone freeing memory its peers have mapped
    auto poison = [&](int code) { c->cache_poisoned = true; if (c->nranks == 1) tritd_trim(c); return code; };
    int s;
    double t0 = now_ms();
    if ((s = tritd_problem_set_D_host(p, D_host)) != TRITD_OK) return poison(s);
    if (mask_host && (s = tritd_problem_set_mask_host(p, mask_host)) != TRITD_OK) return poison(s);
    if ((s = tritd_problem_init(p, o, A0, B0, C0)) != TRITD_OK) return poison(s);
    const double t_h2d = now_ms() - t0;

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    int32_t k = 0;
    s = tritd_problem_iterate(p, o->maxIter, &k);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float it_ms = 0.f;
    cudaEventElapsedTime(&it_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (s != TRITD_OK) return poison(s);

    t0 = now_ms();
    s = tritd_problem_get(p, A, B, C, O, L, errHist, nullptr, nullptr, &k);
    if (s == TRITD_OK && E) s = tritd_problem_get_E(p, E);
    const double t_d2h = now_ms() - t0;
    if (s != TRITD_OK) return poison(s);
    if (iters_out) *iters_out = k;
    if (tm) {
        tm->h2d_ms = t_h2d; tm->iterate_ms = it_ms; tm->d2h_ms = t_d2h; tm->total_ms = now_ms() - t_begin;
        tm->iters = k; tm->launches = (int32_t)(c->launches - launches0);
    }
    return TRITD_OK;
}

// The one-call solver on a single-process multi-GPU context (tritd_create_devices): FULL tensors in and out; device g
// owns the mode-3 slab [t0_g, t1_g); A and B come back from device 0 (the replicas are bitwise equal), C / O / E / L
// slab by slab.
static int admm_group(tritd_ctx* G, const double* D_host, const unsigned char* mask_host, int64_t n1, int64_t n2, int64_t n3, int r,
                      const tritd_opts* o, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                      double* O, double* E, double* L, double* errHist, int32_t* iters_out, tritd_timing* tm) {
    if (!D_host || !o || !A0 || !B0 || !C0 || !errHist) return fail(TRITD_ERR_INVALID, "NULL argument");
    const int nd = (int)G->sub.size();
    if (n1 < 1 || n2 < 1 || n3 < nd) return fail(TRITD_ERR_INVALID, "tensor %lld x %lld x %lld cannot be split into %d mode-3 slabs", (long long)n1, (long long)n2, (long long)n3, nd);
    const double t_begin = now_ms();
    const int64_t launches0 = tritd_launch_count(G);
    std::vector<int64_t> t0(nd), t1(nd);
    for (int g = 0; g < nd; ++g) ST_TRY(tritd_slab_bounds(n3, nd, g, &t0[g], &t1[g]));
    std::vector<tritd_problem*>& ps = G->gcached;
    bool hit = (int)ps.size() == nd && !G->cache_poisoned;
    for (int g = 0; hit && g < nd; ++g) hit = ps[g]->n1 == n1 && ps[g]->n2 == n2 && ps[g]->n3 == t1[g] - t0[g] && ps[g]->r == r;
    auto fail_out = [&](int code) { tritd_trim(G); return code; };
    if (!hit) {
        tritd_trim(G);
        G->cache_poisoned = false;
        for (int g = 0; g < nd; ++g) {
            tritd_problem* q = nullptr;
            int s = problem_create(G->sub[g], n1, n2, t1[g] - t0[g], r, kLevelSolver, &q);
            if (s != TRITD_OK) return fail_out(s);
            ps.push_back(q);
        }
        int s = setup_exchange_inproc(ps);
        if (s != TRITD_OK) return fail_out(s);
    }
    const size_t slice = (size_t)n1 * n2, R = (size_t)r * r;
    int s;
    double tq = now_ms();
    for (int g = 0; g < nd; ++g) {
        if ((s = tritd_problem_set_D_host(ps[g], D_host + slice * t0[g])) != TRITD_OK) return fail_out(s);
        if (mask_host && (s = tritd_problem_set_mask_host(ps[g], mask_host + slice * t0[g])) != TRITD_OK) return fail_out(s);
    }
    for (int g = 0; g < nd; ++g)
        if ((s = problem_init_local(ps[g], o, A0, B0, C0 + R * t0[g])) != TRITD_OK) return fail_out(s);
    {   // ||D||^2: the slabs' partial sums added on the host in rank order, the total handed to every device
        double tot = 0.0;
        for (int g = 0; g < nd; ++g) {
            double part[2];
            cudaSetDevice(ps[g]->ctx->device);
            if (cudaMemcpyAsync(part, ps[g]->norms, 16, cudaMemcpyDeviceToHost, ps[g]->ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ps[g]->ctx->stream) != cudaSuccess)
                return fail_out(fail(TRITD_ERR_CUDA, "normD partial of device %d: %s", ps[g]->ctx->device, cudaGetErrorString(cudaGetLastError())));
            tot += part[0];
        }
        for (int g = 0; g < nd; ++g) {
            const double both[2] = {tot, 0.0};
            cudaSetDevice(ps[g]->ctx->device);
            if (cudaMemcpyAsync(ps[g]->norms, both, 16, cudaMemcpyHostToDevice, ps[g]->ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ps[g]->ctx->stream) != cudaSuccess)
                return fail_out(fail(TRITD_ERR_CUDA, "normD upload failed"));
        }
    }
    for (int g = 0; g < nd; ++g)
        if ((s = problem_init_finish(ps[g])) != TRITD_OK) return fail_out(s);
    const double t_h2d = now_ms() - tq;

    cudaSetDevice(ps[0]->ctx->device);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ps[0]->ctx->stream);
    int32_t k = 0;
    s = iterate_many(ps, o->maxIter, &k);
    cudaSetDevice(ps[0]->ctx->device);
    cudaEventRecord(e1, ps[0]->ctx->stream);
    cudaEventSynchronize(e1);
    float it_ms = 0.f;
    cudaEventElapsedTime(&it_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (s != TRITD_OK) return fail_out(s);
    if (getenv("TRITD_DEBUG_STAMPS")) {            // diagnostics: globaltimer stamps of the last iteration's three k_upd launches, per device
        for (int g = 0; g < nd; ++g) {
            long long h[48];
            if (!ps[g]->dbg) continue;
            cudaSetDevice(ps[g]->ctx->device);
            cudaMemcpy(h, ps[g]->dbg, sizeof(h), cudaMemcpyDeviceToHost);
            for (int w = 0; w < 3; ++w) {
                const long long* q = h + 16 * w; const long long b = q[0];
                fprintf(stderr, "dev %d upd %c: block0 inv_init=%lld inv_done=%lld gram_wait_done=%lld gram_computed=%lld end=%lld | block1 start=%lld reduced=%lld "
                        "inv_seen=%lld applied=%lld rows_done=%lld  (ns since block 0 start; start abs %lld)\n", g, "ABC"[w], q[13] - b, q[1] - b, q[2] - b,
                        q[12] - b, q[3] - b, q[4] - b, q[5] - b, q[6] - b, q[9] - b, q[7] - b, b);
            }
        }
    }

    tq = now_ms();
    for (int g = 0; g < nd && s == TRITD_OK; ++g) {
        s = tritd_problem_get(ps[g], g == 0 ? A : nullptr, g == 0 ? B : nullptr, C ? C + R * t0[g] : nullptr, O ? O + slice * t0[g] : nullptr,
                              L ? L + slice * t0[g] : nullptr, g == 0 ? errHist : nullptr, nullptr, nullptr, &k);
        if (s == TRITD_OK && E) s = tritd_problem_get_E(ps[g], E + slice * t0[g]);
    }
    const double t_d2h = now_ms() - tq;
    if (s != TRITD_OK) return fail_out(s);
    if (iters_out) *iters_out = k;
    if (tm) {
        tm->h2d_ms = t_h2d; tm->iterate_ms = it_ms; tm->d2h_ms = t_d2h; tm->total_ms = now_ms() - t_begin;
        tm->iters = k; tm->launches = (int32_t)(tritd_launch_count(G) - launches0);
    }
    return TRITD_OK;
}

extern "C" int tritd_admm_f64(tritd_ctx* c, const double* D_host, int64_t n1, int64_t n2, int64_t n3, int r,
                              const tritd_opts* o, const double* A0, const double* B0, const double* C0, double* A,
                              double* B, double* C, double* O, double* L, double* errHist, int32_t* iters_out,
                              tritd_timing* tm) {
    return admm_impl(c, D_host, nullptr, n1, n2, n3, r, o, A0, B0, C0, A, B, C, O, nullptr, L, errHist, iters_out, tm);
}

extern "C" int tritd_admm_ex_f64(tritd_ctx* c, const double* D_host, const unsigned char* mask_or_null, int64_t n1, int64_t n2,
                                 int64_t n3, int r, const tritd_opts* o, const double* A0, const double* B0, const double* C0,
                                 double* A, double* B, double* C, double* O, double* E_or_null, double* L_or_null,
                                 double* errHist, int32_t* iters_out, tritd_timing* tm) {
    return admm_impl(c, D_host, mask_or_null, n1, n2, n3, r, o, A0, B0, C0, A, B, C, O, E_or_null, L_or_null, errHist, iters_out, tm);
}

// ---------------------------------------------------------------------------
// standalone helpers
// ---------------------------------------------------------------------------
struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(double));
        return e == cudaSuccess ? TRITD_OK : fail(TRITD_ERR_CUDA, "cudaMalloc(%zu doubles): %s", n, cudaGetErrorString(e));
    }
};

static int check_r(int r) {
    if (r < 1) return fail(TRITD_ERR_INVALID, "triple rank r=%d", r);
    if (r > TRITD_MAX_R) return fail(TRITD_ERR_UNSUPPORTED, "r=%d unsupported (1..%d)", r, TRITD_MAX_R);
    return TRITD_OK;
}

// ---------------------------------------------------------------------------
// [A,B,C,errHist] = triple_decomp_ALS(X, r, opts)    (fast_robust_triple_tensor/triple_decomp_ALS.m:1-40)
// The same kernels as the ADMM sweep with the target fixed at X and the ridge 1e-9 in all three updates; the
// relative error is evaluated BEFORE the updates of an iteration (:15-16) and the stopping rule (:20-23)
// skips them.  Single rank.
// ---------------------------------------------------------------------------
static int als_iteration(tritd_problem* p) {
    tritd_ctx* c = p->ctx;
    double* rhsA = p->bufA;
    double* SC = p->bufA + (size_t)p->n1 * p->RS;
    ST_TRY(launch_fused(p, 2, nullptr));                                   // sum((X - Xhat).^2) partials
    k_finalize_als<<<1, 256, 0, c->stream>>>(p->st, p->norm_part, p->gridF, p->errHist);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    ST_TRY(launch_mttkrp1(p, p->mapT, p->B2, p->C3, rhsA));                // X1*F'
    ST_TRY(launch_upd(p, 0, kSrcDirect, true, rhsA, nullptr, p->SB, SC, 1e-9, p->A1, p->A1T, p->n1, p->SA));
    ST_TRY(launch_ppass(p, p->mapT));
    ST_TRY(launch_upd(p, 1, kSrcPB, true, nullptr, nullptr, p->SA, SC, 1e-9, p->B2, nullptr, p->n2, p->SB));
    ST_TRY(launch_upd(p, 2, kSrcPC, true, nullptr, nullptr, p->SA, p->SB, 1e-9, p->C3, nullptr, p->n3, SC));
    return TRITD_OK;
}

extern "C" int tritd_als_f64(tritd_ctx* c, const double* X_host, int64_t n1, int64_t n2, int64_t n3, int r, int32_t maxIter,
                             double tol, int32_t disp, const double* A0, const double* B0, const double* C0, double* A,
                             double* B, double* C, double* errHist, int32_t* iters_out) {
    if (!c || !X_host || !A0 || !B0 || !C0 || !errHist) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (c->nranks != 1) return fail(TRITD_ERR_UNSUPPORTED, "triple_decomp_ALS runs on a single-rank context");
    if (maxIter < 1) return fail(TRITD_ERR_INVALID, "opts.maxIter = %d", maxIter);
    tritd_problem* p = nullptr;
    ST_TRY(tritd_problem_create(c, n1, n2, n3, r, &p));
    auto bail = [&](int code) { tritd_problem_destroy(p); return code; };
    int s;
    if ((s = tritd_problem_set_D_host(p, X_host)) != TRITD_OK) return bail(s);
    // the ADMM state set-up also gives what ALS needs: T = X, ||X||, the small Grams of the initial factors
    tritd_opts o;
    memset(&o, 0, sizeof(o));
    o.mu = 1.0; o.rho = 1.0; o.lambda_ = 0.0; o.lambda2 = 1e-9; o.tol = tol; o.maxIter = maxIter; o.disp = 0;
    if ((s = tritd_problem_init(p, &o, A0, B0, C0)) != TRITD_OK) return bail(s);
    int done = 0, printed = 0;
    std::vector<double> eh(maxIter);
    while (done < maxIter && !p->st_host->stop) {
        const int batch = std::min(maxIter - done, 5);            // the reference prints every 5th iteration
        for (int i = 0; i < batch; ++i)
            if ((s = als_iteration(p)) != TRITD_OK) return bail(s);
        if ((s = fetch_state(p)) != TRITD_OK) return bail(s);
        done = p->st_host->k;
        if (disp) {
            if (cudaMemcpy(eh.data(), p->errHist, sizeof(double) * done, cudaMemcpyDeviceToHost) != cudaSuccess)
                return bail(fail(TRITD_ERR_CUDA, "errHist copy failed"));
            for (int k = printed + 1; k <= done; ++k)
                if (k % 5 == 0) emit("Iteration %d, relative error = %.4e\n", k, eh[k - 1]);
            printed = done;
        }
        if (p->st_host->stop) break;
    }
    if (p->st_host->status != 0) return bail(fail(TRITD_ERR_NUMERIC, "ridge system contains NaN / Inf (ALS iteration %d)", done));
    s = tritd_problem_get(p, A, B, C, nullptr, nullptr, errHist, nullptr, nullptr, nullptr);
    if (iters_out) *iters_out = done;
    tritd_problem_destroy(p);
    return s;
}

// Reconstruction L = triple_product(A,B,C) from the factors resident in `p`, written to `dst` (device) with leading
// dimension `ld` (even; n1 itself when n1 is even, so a dense caller buffer is written in place).
static int reconstruct_to(tritd_problem* p, double* dst, int ld) {
    tritd_ctx* c = p->ctx;
    FusedArgs a;
    a.D = nullptr; a.O = dst;
    a.A1 = p->A1; a.B2 = p->B2; a.C3 = p->C3; a.st = p->st; a.norm_part = p->norm_part;
    a.n1 = p->n1; a.n2 = p->n2; a.n3 = p->n3; a.ld1 = ld; a.RS = p->RS;
    a.n_it = p->n_it; a.n_jc = p->n_jc; a.gi = p->gi;
#define CALL(NT_, KS_) k_fused<KS_, 1><<<p->gridF, 256, 0, c->stream>>>(a);
    TRITD_DISPATCH_R(p->r, CALL)
#undef CALL
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

// dense (ld = n1) when the 16-byte vector stores of k_fused stay aligned, else padded to even
static int dense_ld(int n1) { return (n1 + 1) & ~1; }

// The factor-level state (factors, small scratch; no N-sized array) of the standalone helpers is kept in the context and
// reused while the shape repeats: the caller's post-solve triple_product(A,B,C) (traffic_triple_comparison.m:62) then costs
// the factor upload and one kernel, not a dozen allocations.
static int helper_problem(tritd_ctx* c, int64_t n1, int64_t n2, int64_t n3, int r, tritd_problem** out) {
    tritd_problem* p = c->hcached;
    if (p && !(p->n1 == n1 && p->n2 == n2 && p->n3 == n3 && p->r == r)) { tritd_problem_destroy(p); p = nullptr; c->hcached = nullptr; }
    if (!p) {
        ST_TRY(problem_create(c, n1, n2, n3, r, kLevelFactors, &p));
        c->hcached = p;
    }
    *out = p;
    return TRITD_OK;
}

static int triple_product_common(tritd_ctx* c, const double* A, const double* B, const double* C, int64_t n1, int64_t n2,
                                 int64_t n3, int r, double* Xhat, bool out_on_device) {
    if (!c || !A || !B || !C || !Xhat) return fail(TRITD_ERR_INVALID, "NULL argument");
    ST_TRY(check_r(r));
    tritd_problem* p = nullptr;
    ST_TRY(helper_problem(c, n1, n2, n3, r, &p));                         // factors + small state only
    auto done = [&](int code) { cudaStreamSynchronize(c->stream); return code; };
    int s = upload_factors(p, A, B, C);
    if (s != TRITD_OK) return done(s);
    const int ld = dense_ld(p->n1);
    const size_t ncols = (size_t)p->n2 * p->n3, N = (size_t)p->n1 * ncols;
    if (out_on_device && ld == p->n1) return done(reconstruct_to(p, Xhat, ld));      // straight into the caller's buffer
    DevBuf tmp;
    if ((s = tmp.alloc((size_t)ld * ncols)) != TRITD_OK) return done(s);
    if ((s = reconstruct_to(p, tmp.p, ld)) != TRITD_OK) return done(s);
    cudaError_t e;
    if (ld == p->n1) e = cudaMemcpyAsync(Xhat, tmp.p, N * 8, cudaMemcpyDefault, c->stream);
    else e = cudaMemcpy2DAsync(Xhat, (size_t)p->n1 * 8, tmp.p, (size_t)ld * 8, (size_t)p->n1 * 8, ncols, cudaMemcpyDefault, c->stream);
    if (e != cudaSuccess) return done(fail(TRITD_ERR_CUDA, "triple_product copy: %s", cudaGetErrorString(e)));
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return done(fail(TRITD_ERR_CUDA, "triple_product failed: %s", cudaGetErrorString(cudaGetLastError())));
    return done(TRITD_OK);
}

extern "C" int tritd_triple_product_f64(tritd_ctx* c, const double* A, const double* B, const double* C, int64_t n1,
                                        int64_t n2, int64_t n3, int r, double* Xhat) {
    return triple_product_common(c, A, B, C, n1, n2, n3, r, Xhat, false);
}
extern "C" int tritd_triple_product_dev_f64(tritd_ctx* c, const double* A, const double* B, const double* C, int64_t n1,
                                            int64_t n2, int64_t n3, int r, double* Xhat_dev) {
    return triple_product_common(c, A, B, C, n1, n2, n3, r, Xhat_dev, true);
}

// Design matrices / triple product of the original (Qi) triple decomposition -- origin_triple_tensor/buildF.m,
// buildG.m, buildH.m, triple_product.m.  which: 0 = F(B,C), 1 = G(A,C), 2 = H(A,B).
static int design_qi_dev(tritd_ctx* c, const double* U, size_t nU, const double* V, size_t nV, long na, long nb, int r, int which,
                         DevBuf& u, DevBuf& v, DevBuf& o) {
    const size_t total = (size_t)r * r * na * nb;
    ST_TRY(u.alloc(nU)); ST_TRY(v.alloc(nV)); ST_TRY(o.alloc(total));
    CU_TRY(cudaMemcpyAsync(u.p, U, nU * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(v.p, V, nV * 8, cudaMemcpyHostToDevice, c->stream));
    k_design_qi<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, c->stream>>>(u.p, v.p, o.p, na, nb, r, which);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

extern "C" int tritd_design_qi_f64(tritd_ctx* c, int which, const double* U, const double* V, int64_t na, int64_t nb, int r,
                                   double* out) {
    if (!c || !U || !V || !out) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (which < 0 || which > 2 || na < 1 || nb < 1) return fail(TRITD_ERR_INVALID, "bad design-matrix request");
    ST_TRY(check_r(r));
    CU_TRY(cudaSetDevice(c->device));
    const size_t R = (size_t)r * r;
    // which 0: U = B (r x n2 x r), V = C (r x r x n3); 1: U = A (n1 x r x r), V = C; 2: U = A, V = B (r x n2 x r)
    DevBuf u, v, o;
    ST_TRY(design_qi_dev(c, U, R * na, V, R * nb, (long)na, (long)nb, r, which, u, v, o));
    CU_TRY(cudaMemcpyAsync(out, o.p, R * na * nb * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return TRITD_OK;
}

extern "C" int tritd_triple_product_qi_f64(tritd_ctx* c, const double* A, const double* B, const double* C, int64_t n1,
                                           int64_t n2, int64_t n3, int r, double* Xhat) {
    if (!c || !A || !B || !C || !Xhat) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (n1 < 1 || n2 < 1 || n3 < 1) return fail(TRITD_ERR_INVALID, "bad size");
    ST_TRY(check_r(r));
    CU_TRY(cudaSetDevice(c->device));
    const size_t R = (size_t)r * r, N = (size_t)n1 * n2 * n3;
    DevBuf u, v, f, a, x;
    ST_TRY(design_qi_dev(c, B, R * n2, C, R * n3, (long)n2, (long)n3, r, 0, u, v, f));       // F (r^2 x n2 n3)
    ST_TRY(a.alloc(R * n1)); ST_TRY(x.alloc(N));
    CU_TRY(cudaMemcpyAsync(a.p, A, R * n1 * 8, cudaMemcpyHostToDevice, c->stream));
    // Xhat = reshape(A, [n1, r^2]) * F : column k = q + r*s of unfold(A,1) is A(:,q,s)
    k_unfold1_times<<<(unsigned)std::min<size_t>((N + 255) / 256, 148 * 16), 256, 0, c->stream>>>(a.p, f.p, x.p, (long)n1,
                                                                                                 (size_t)n2 * n3, (int)R);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    CU_TRY(cudaMemcpyAsync(Xhat, x.p, N * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return TRITD_OK;
}

// [rmse, nrmse] = evaluate(Xhat, gt, mask) with Xhat = triple_product(A,B,C) formed on the device
// (traffic_triple_comparison.m:194-202; the drivers' RRE).  gt is a dense n1 x n2 x n3 tensor (entries outside the
// mask are ignored), mask dense bytes or NULL (all true): the reconstruction never crosses PCIe.  `p` supplies the
// factors (and the reduction scratch): a kLevelFactors problem for the standalone call, the solver's own problem for
// tritd_problem_evaluate.  Device memory used: the reconstruction, gt and the mask -- nothing else.
static int evaluate_with(tritd_problem* p, const double* gt_host, const unsigned char* mask_host, double* rmse, double* nrmse) {
    tritd_ctx* c = p->ctx;
    const int ld = dense_ld(p->n1);
    const size_t ncols = (size_t)p->n2 * p->n3, N = (size_t)p->n1 * ncols;
    DevBuf L, gt;
    ST_TRY(L.alloc((size_t)ld * ncols)); ST_TRY(gt.alloc(N));
    unsigned char* mdev = nullptr;
    int s = reconstruct_to(p, L.p, ld);
    if (s == TRITD_OK && cudaMemcpyAsync(gt.p, gt_host, N * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        s = fail(TRITD_ERR_CUDA, "gt copy failed");
    if (s == TRITD_OK && mask_host) {
        if (cudaMalloc((void**)&mdev, N) != cudaSuccess) s = fail(TRITD_ERR_CUDA, "cudaMalloc(mask) failed");
        else if (cudaMemcpyAsync(mdev, mask_host, N, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) s = fail(TRITD_ERR_CUDA, "mask copy failed");
    }
    double sums[2] = {0.0, 0.0};
    if (s == TRITD_OK) {
        k_evaluate<<<1024, 256, 0, c->stream>>>(L.p, gt.p, mdev, p->n1, ld, p->n1, ncols, p->norm_part);
        k_sum_pairs<<<1, 256, 0, c->stream>>>(p->norm_part, 1024, p->norms, nullptr);
        c->launches += 2;
        if (cudaMemcpyAsync(sums, p->norms, 16, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess)
            s = fail(TRITD_ERR_CUDA, "evaluate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaStreamSynchronize(c->stream);
    if (mdev) cudaFree(mdev);
    if (s != TRITD_OK) return s;
    if (rmse) *rmse = sqrt(sums[0]);
    if (nrmse) *nrmse = sqrt(sums[0]) / sqrt(sums[1]);
    return TRITD_OK;
}

extern "C" int tritd_evaluate_f64(tritd_ctx* c, const double* A, const double* B, const double* C, int64_t n1, int64_t n2,
                                  int64_t n3, int r, const double* gt_host, const unsigned char* mask_host, double* rmse,
                                  double* nrmse) {
    if (!c || !A || !B || !C || !gt_host) return fail(TRITD_ERR_INVALID, "NULL argument");
    ST_TRY(check_r(r));
    tritd_problem* p = nullptr;
    ST_TRY(helper_problem(c, n1, n2, n3, r, &p));
    int s = upload_factors(p, A, B, C);
    if (s == TRITD_OK) s = evaluate_with(p, gt_host, mask_host, rmse, nrmse);
    return s;
}

// the same with the factors the solver holds on the device (after tritd_problem_iterate): nothing but gt crosses PCIe
extern "C" int tritd_problem_evaluate(tritd_problem* p, const double* gt_host, const unsigned char* mask_host, double* rmse,
                                      double* nrmse) {
    if (!p || !gt_host || !p->initialized) return fail(TRITD_ERR_INVALID, "NULL argument / not initialised");
    CU_TRY(cudaSetDevice(p->ctx->device));
    return evaluate_with(p, gt_host, mask_host, rmse, nrmse);
}

// ---- unfold / buildF / buildG / buildH / soft_threshold: device cores + host wrappers --------------------------
static int unfold_dev(tritd_ctx* c, const double* X, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn) {
    const size_t N = (size_t)n1 * n2 * n3;
    if (mode == 1) {                     // reshape only (unfold.m:6)
        if (Xn != X) CU_TRY(cudaMemcpyAsync(Xn, X, N * 8, cudaMemcpyDeviceToDevice, c->stream));
        return TRITD_OK;
    }
    long rows, cols, batch;
    if (mode == 2) { rows = n1; cols = n2; batch = n3; } else { rows = n1 * n2; cols = n3; batch = 1; }
    if (batch > 65535 || (cols + 31) / 32 > 65535) return fail(TRITD_ERR_INVALID, "unfold: dimension too large");
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)batch);
    const bool vec = rows % 2 == 0 && cols % 2 == 0 && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Xn % 16 == 0);
    if (vec) k_transpose_v2<<<grid, 256, 0, c->stream>>>(X, Xn, rows, cols);
    else k_transpose<<<grid, 256, 0, c->stream>>>(X, Xn, rows, cols);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

static int check_unfold_args(tritd_ctx* c, const double* X, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn) {
    if (!c || !X || !Xn) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (n1 < 1 || n2 < 1 || n3 < 1) return fail(TRITD_ERR_INVALID, "bad size");
    if (mode < 1 || mode > 3) return fail(TRITD_ERR_INVALID, "Mode must be 1, 2, or 3.");
    CU_TRY(cudaSetDevice(c->device));
    return TRITD_OK;
}

extern "C" int tritd_unfold_dev_f64(tritd_ctx* c, const double* X_dev, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn_dev) {
    ST_TRY(check_unfold_args(c, X_dev, n1, n2, n3, mode, Xn_dev));
    if (mode != 1 && X_dev == Xn_dev) return fail(TRITD_ERR_INVALID, "unfold: in-place transposition is not supported");
    return unfold_dev(c, X_dev, n1, n2, n3, mode, Xn_dev);
}

extern "C" int tritd_unfold_f64(tritd_ctx* c, const double* X, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn) {
    ST_TRY(check_unfold_args(c, X, n1, n2, n3, mode, Xn));
    const size_t N = (size_t)n1 * n2 * n3;
    if (mode == 1) { if (Xn != X) memcpy(Xn, X, N * 8); return TRITD_OK; }   // reshape only (unfold.m:6)
    DevBuf in, out;
    ST_TRY(in.alloc(N)); ST_TRY(out.alloc(N));
    CU_TRY(cudaMemcpyAsync(in.p, X, N * 8, cudaMemcpyHostToDevice, c->stream));
    ST_TRY(unfold_dev(c, in.p, n1, n2, n3, mode, out.p));
    CU_TRY(cudaMemcpyAsync(Xn, out.p, N * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return TRITD_OK;
}

// which: 0 = F(B,C), 1 = G(A,C), 2 = H(A,B); U, V are the MATLAB 3-D factor arrays on the device
static int design_dev(tritd_ctx* c, int which, const double* U, const double* V, int64_t na, int64_t nb, int r, double* out) {
    const int R = r * r;
    if ((size_t)R * (size_t)na >= ((size_t)1 << 31)) return fail(TRITD_ERR_INVALID, "design matrix: mode too large");
    dim3 grid((unsigned)(((size_t)R * na + 255) / 256), (unsigned)std::min<int64_t>(nb, 4096));
    k_build_design<<<grid, 256, 0, c->stream>>>(U, V, out, (int)na, (int)nb, r, which);
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

extern "C" int tritd_build_design_dev_f64(tritd_ctx* c, int which, const double* U_dev, const double* V_dev, int64_t na, int64_t nb,
                                          int r, double* out_dev) {
    if (!c || !U_dev || !V_dev || !out_dev || na < 1 || nb < 1 || which < 0 || which > 2) return fail(TRITD_ERR_INVALID, "bad argument");
    ST_TRY(check_r(r));
    CU_TRY(cudaSetDevice(c->device));
    return design_dev(c, which, U_dev, V_dev, na, nb, r, out_dev);
}

static int design_host(tritd_ctx* c, int which, const double* U, const double* V, int64_t na, int64_t nb, int r, double* out) {
    if (!c || !U || !V || !out || na < 1 || nb < 1) return fail(TRITD_ERR_INVALID, "bad argument");
    ST_TRY(check_r(r));
    CU_TRY(cudaSetDevice(c->device));
    const size_t R = (size_t)r * r, total = R * na * nb;
    DevBuf u, v, o;
    ST_TRY(u.alloc(R * na)); ST_TRY(v.alloc(R * nb)); ST_TRY(o.alloc(total));
    CU_TRY(cudaMemcpyAsync(u.p, U, R * na * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(v.p, V, R * nb * 8, cudaMemcpyHostToDevice, c->stream));
    ST_TRY(design_dev(c, which, u.p, v.p, na, nb, r, o.p));
    CU_TRY(cudaMemcpyAsync(out, o.p, total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return TRITD_OK;
}

extern "C" int tritd_buildF_f64(tritd_ctx* c, const double* B, const double* C, int64_t n2, int64_t n3, int r, double* F) {
    return design_host(c, 0, B, C, n2, n3, r, F);
}
extern "C" int tritd_buildG_f64(tritd_ctx* c, const double* A, const double* C, int64_t n1, int64_t n3, int r, double* G) {
    return design_host(c, 1, A, C, n1, n3, r, G);
}
extern "C" int tritd_buildH_f64(tritd_ctx* c, const double* A, const double* B, int64_t n1, int64_t n2, int r, double* H) {
    return design_host(c, 2, A, B, n1, n2, r, H);
}

static int soft_threshold_dev(tritd_ctx* c, const double* X, int64_t n, double lam, double* out) {
    const size_t n2 = (size_t)n / 2;
    const bool vec = ((uintptr_t)X % 16 == 0) && ((uintptr_t)out % 16 == 0) && n2 > 0;
    if (vec) {
        const unsigned grid = (unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)c->num_sms * 16);
        k_soft_threshold_v2<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const double2*>(X), reinterpret_cast<double2*>(out), n2, lam);
        if (n & 1) k_soft_threshold<<<1, 32, 0, c->stream>>>(X + 2 * n2, out + 2 * n2, 1, lam);
    } else {
        const unsigned grid = (unsigned)std::min<size_t>(((size_t)n + 255) / 256, (size_t)c->num_sms * 16);
        k_soft_threshold<<<grid, 256, 0, c->stream>>>(X, out, (size_t)n, lam);
    }
    CU_TRY(cudaGetLastError());
    c->launches += 1;
    return TRITD_OK;
}

extern "C" int tritd_soft_threshold_dev_f64(tritd_ctx* c, const double* X_dev, int64_t n, double lam, double* out_dev) {
    if (!c || !X_dev || !out_dev || n < 0) return fail(TRITD_ERR_INVALID, "bad argument");
    if (n == 0) return TRITD_OK;
    CU_TRY(cudaSetDevice(c->device));
    return soft_threshold_dev(c, X_dev, n, lam, out_dev);
}

extern "C" int tritd_soft_threshold_f64(tritd_ctx* c, const double* X, int64_t n, double lam, double* out) {
    if (!c || !X || !out || n < 0) return fail(TRITD_ERR_INVALID, "bad argument");
    if (n == 0) return TRITD_OK;
    CU_TRY(cudaSetDevice(c->device));
    DevBuf in, o;
    ST_TRY(in.alloc((size_t)n)); ST_TRY(o.alloc((size_t)n));
    CU_TRY(cudaMemcpyAsync(in.p, X, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    ST_TRY(soft_threshold_dev(c, in.p, n, lam, o.p));
    CU_TRY(cudaMemcpyAsync(out, o.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return TRITD_OK;
}

// Measured FP64 tensor-core (DMMA.8x8x4) peak of this GPU: the denominator of the "FP64-TC roofline %" north_star asks
// for (MEASURED_PEAKS.json holds only HBM and bf16 peaks).  One CTA of 16 warps per SM, 8 independent accumulator
// chains per warp, `ms_budget` milliseconds of back-to-back launches after a warm-up.
extern "C" int tritd_measure_dmma_peak(tritd_ctx* c, double ms_budget, double* tflops) {
    if (!c || !tflops) return fail(TRITD_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    DevBuf out;
    ST_TRY(out.alloc((size_t)c->num_sms * 512));
    const int iters = 4000;
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
    auto run = [&](int reps, float* ms) -> int {
        CU_TRY(cudaEventRecord(e0, c->stream));
        for (int q = 0; q < reps; ++q) k_dmma_peak<<<c->num_sms, 512, 0, c->stream>>>(out.p, iters);
        CU_TRY(cudaEventRecord(e1, c->stream));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(ms, e0, e1));
        c->launches += reps;
        return TRITD_OK;
    };
    float ms = 0.f;
    int s = run(3, &ms);                                          // warm-up + calibration
    if (s == TRITD_OK) {
        const int reps = std::max(3, (int)(ms_budget / std::max(ms / 3.0f, 1e-3f)));
        s = run(reps, &ms);
        if (s == TRITD_OK) *tflops = (double)reps * c->num_sms * 16.0 * iters * 8.0 * 512.0 / (ms * 1e-3) * 1e-12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return s;
}

// The three contractions of one sweep at fixed factors, through the SOLVER's own kernels: mode 1 = k_mttkrp1 (+ its
// fixed-order reduction), modes 2 / 3 = k_ppass followed by k_upd's strided-sum sources over t / over j (the very
// reductions update_B / update_C run; apply = 0 stops before the ridge solve).
extern "C" int tritd_mttkrp_f64(tritd_ctx* c, const double* X, const double* A, const double* B, const double* C,
                                int64_t n1, int64_t n2, int64_t n3, int r, int mode, double* rhs) {
    if (!c || !X || !A || !B || !C || !rhs) return fail(TRITD_ERR_INVALID, "NULL argument");
    if (mode < 1 || mode > 3) return fail(TRITD_ERR_INVALID, "Mode must be 1, 2, or 3.");
    ST_TRY(check_r(r));
    tritd_problem* p = nullptr;
    ST_TRY(problem_create(c, n1, n2, n3, r, kLevelContract, &p));
    cudaStream_t st = c->stream;
    auto body = [&]() -> int {
        ST_TRY(copy_in(p, p->T, X));
        ST_TRY(upload_factors(p, A, B, C));
        const int n = mode == 1 ? p->n1 : (mode == 2 ? p->n2 : p->n3);
        std::vector<double> h((size_t)n * p->RS);
        if (mode == 1) {
            ST_TRY(launch_mttkrp1(p, p->mapT, p->B2, p->C3, p->bufA));
            CU_TRY(cudaMemcpyAsync(h.data(), p->bufA, h.size() * 8, cudaMemcpyDeviceToHost, st));
        } else {
            // transposed copy of A1 for the TMA-fed pass
            std::vector<double> a1, a1t((size_t)p->RS * p->ldt, 0.0);
            pack_A(A, p->n1, p->R, p->RS, a1);
            for (int i = 0; i < p->n1; ++i) for (int k = 0; k < p->R; ++k) a1t[(size_t)k * p->ldt + i] = a1[(size_t)i * p->RS + k];
            CU_TRY(cudaMemcpyAsync(p->A1T, a1t.data(), a1t.size() * 8, cudaMemcpyHostToDevice, st));
            CU_TRY(cudaStreamSynchronize(st));
            ST_TRY(launch_ppass(p, p->mapT));
            double* out = mode == 2 ? p->rhsB : p->rhsC;
            ST_TRY(launch_upd(p, mode - 1, mode == 2 ? kSrcPB : kSrcPC, false, nullptr, out, nullptr, nullptr, 0.0, nullptr, nullptr, n, nullptr));
            CU_TRY(cudaMemcpyAsync(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost, st));
        }
        CU_TRY(cudaStreamSynchronize(st));
        for (int k = 0; k < p->R; ++k) for (int i = 0; i < n; ++i) rhs[(size_t)k * n + i] = h[(size_t)i * p->RS + k];
        return TRITD_OK;
    };
    int s = body();
    tritd_problem_destroy(p);
    return s;
}

// One factor update in isolation (update_A/B/C after the contraction, :77-78 / :86 / :93): X = RHS * pinv(S1 o S2 +
// alpha*I) through k_upd exactly as the solver launches it (direct RHS source): also returns the pseudo-inverse and
// X'X, plus info[0] = 1 when the truncating pinv path ran and info[1] = singular values it zeroed.
extern "C" int tritd_factor_update_f64(tritd_ctx* c, const double* rhs, int64_t n, int r, const double* S1, const double* S2,
                                       double alpha, double* X, double* Ginv_or_null, double* XtX_or_null, int32_t* info_or_null) {
    if (!c || !rhs || !S1 || !S2 || !X || n < 1) return fail(TRITD_ERR_INVALID, "bad argument");
    ST_TRY(check_r(r));
    tritd_problem* p = nullptr;
    // the update runs over "mode 1" of an n x 1 x 1 problem: only the factor-sized state is used
    ST_TRY(problem_create(c, n, 1, 1, r, kLevelFactors, &p));
    cudaStream_t st = c->stream;
    const int R = p->R, RS = p->RS;
    auto body = [&]() -> int {
        std::vector<double> h((size_t)n * RS, 0.0), g((size_t)RS * RS, 0.0);
        for (int k = 0; k < R; ++k) for (int64_t i = 0; i < n; ++i) h[(size_t)i * RS + k] = rhs[(size_t)k * n + i];
        CU_TRY(cudaMemcpyAsync(p->bufA, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        double* S2d = p->bufA + (size_t)p->n1 * RS;
        for (int a = 0; a < R; ++a) for (int b = 0; b < R; ++b) g[(size_t)a * RS + b] = S1[(size_t)b * R + a];
        CU_TRY(cudaMemcpyAsync(p->SB, g.data(), g.size() * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        for (int a = 0; a < R; ++a) for (int b = 0; b < R; ++b) g[(size_t)a * RS + b] = S2[(size_t)b * R + a];
        CU_TRY(cudaMemcpyAsync(S2d, g.data(), g.size() * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        ST_TRY(launch_upd(p, 0, kSrcDirect, true, p->bufA, nullptr, p->SB, S2d, alpha, p->A1, p->A1T, (int)n, p->SA));
        ST_TRY(fetch_state(p));
        CU_TRY(cudaMemcpy(h.data(), p->A1, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < R; ++k) for (int64_t i = 0; i < n; ++i) X[(size_t)k * n + i] = h[(size_t)i * RS + k];
        if (Ginv_or_null) {
            CU_TRY(cudaMemcpy(g.data(), p->Minv, g.size() * 8, cudaMemcpyDeviceToHost));
            for (int a = 0; a < R; ++a) for (int b = 0; b < R; ++b) Ginv_or_null[(size_t)b * R + a] = g[(size_t)a * RS + b];
        }
        if (XtX_or_null) {
            CU_TRY(cudaMemcpy(g.data(), p->SA, g.size() * 8, cudaMemcpyDeviceToHost));
            for (int a = 0; a < R; ++a) for (int b = 0; b < R; ++b) XtX_or_null[(size_t)b * R + a] = g[(size_t)a * RS + b];
        }
        if (info_or_null) { info_or_null[0] = p->st_host->pinv_fallbacks; info_or_null[1] = p->st_host->pinv_truncated; }
        return TRITD_OK;
    };
    int s = body();
    tritd_problem_destroy(p);
    return s;
}
