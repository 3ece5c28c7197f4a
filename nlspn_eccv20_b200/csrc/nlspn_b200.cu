// nlspn_b200.cu -- C ABI (include/nlspn_b200.h) over the sm_100a kernels.
// Host side only validates, picks a kernel variant and enqueues launches on the caller's
// stream.  No allocation, no synchronisation, no torch/ATen.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <type_traits>
#include <vector>

#include "../../include/nlspn_b200.h"
#include "kernels_fixed.cuh"
#include "kernels_persist.cuh"
#include "kernels_tiled.cuh"
#include "kernels_f64.cuh"
#include "kernels_gather.cuh"
#include "kernels_step.cuh"
#include "kernels_local.cuh"
#include "kernels_det.cuh"
#include "kernels_head.cuh"
#include "kernels_head2.cuh"
#include "kernels_head_wgrad.cuh"
#include "kernels_head_dgrad.cuh"

using namespace nlspn;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

// ---- tuning options.  Read ONCE from the environment when the library is loaded; afterwards only
// nlspn_set_option changes them (tests, tools).  Nothing on the call path touches getenv.
// -1 = "auto" where a default depends on the shape.
enum Opt { kOptTiled = 0, kOptPersist, kOptPdl, kOptFwdTH, kOptParamTH, kOptStateTma, kOptStateGather,
           kOptGatherCompact, kOptStateZero3, kOptStateMinB, kOptGroupImages, kOptStreamHint,
           kOptParamFactored, kOptDcnBlocked, kOptStateLocal, kOptLocalPrefetch, kOptLocalMinB, kOptSchedMinB, kOptHeadsTma, kOptHeadsRows, kOptHeadsRing, kOptHeadsPersist, kOptHeadsReuse, kOptHeadsKs, kOptHeadsWgradRoll, kOptCount };
struct OptDef { const char *name; const char *env; int def; };
const OptDef kOptDefs[kOptCount] = {
    {"tiled", "NLSPN_TILED", 1},
    {"persist", "NLSPN_PERSIST", -1},
    {"pdl", "NLSPN_PDL", 1},
    {"fwd_th", "NLSPN_FWD_TH", 4},
    {"param_th", "NLSPN_PARAM_TH", 4},
    {"state_tma", "NLSPN_STATE_TMA", -1},
    {"state_gather", "NLSPN_STATE_GATHER", -1},
    {"gather_compact", "NLSPN_GATHER_COMPACT", -1},
    {"state_zero3", "NLSPN_STATE_ZERO3", 1},
    {"state_minb", "NLSPN_STATE_MINB", 4},
    {"group_images", "NLSPN_GROUP_IMAGES", 0},
    {"stream_hint", "NLSPN_STREAM_HINT", -1},
    {"param_factored", "NLSPN_PARAM_FACTORED", 1},
    {"dcn_blocked", "NLSPN_DCN_BLOCKED", 1},
    {"state_local", "NLSPN_STATE_LOCAL", -1},
    {"local_prefetch", "NLSPN_LOCAL_PREFETCH", 0},
    {"local_minb", "NLSPN_LOCAL_MINB", 1},      // 0 / 1 / 2 = 4 / 5 / 6 CTAs per SM
    {"sched_minb", "NLSPN_SCHED_MINB", 5},      // CTAs per SM of the schedule build (4: 0.275 ms, 5: 0.249 ms per step)
    {"heads_tma", "NLSPN_HEADS_TMA", 1},         // head convolutions: activations by TMA ring (0: direct global loads)
    {"heads_rows", "NLSPN_HEADS_ROWS", 1},       // head convolutions: MN-major operand form (kernels_head2.cuh; K = 3, 5, W % 4 == 0)
    {"heads_ring", "NLSPN_HEADS_RING", 0},       // ring depth of the one-CTA-per-tile form for K = 3 (0 = deepest that keeps two CTAs per SM = 4; 3)
    {"heads_persist", "NLSPN_HEADS_PERSIST", 1}, // MN-major form as one persistent warp-specialised CTA per SM (0: one CTA per tile, two per SM)
    {"heads_reuse", "NLSPN_HEADS_REUSE", 1},     // persistent form: keep an input row's A tile in the tensor core's collector across its dy MMAs
    {"heads_ks", "NLSPN_HEADS_KS", 2},           // persistent form: channels per TMA stage / 8 (1 or 2: fewer, larger TMA operations)
    {"heads_wgrad_roll", "NLSPN_HEADS_WGRAD_ROLL", 1},   // head weight gradients: rolling column-strip form (0: one 32-pixel chunk per stage)
};
std::atomic<int> g_opt[kOptCount];
const bool g_opt_init = []() {
    for (int i = 0; i < kOptCount; ++i) {
        int v = kOptDefs[i].def;
        if (const char *e = getenv(kOptDefs[i].env)) v = atoi(e);
        g_opt[i].store(v, std::memory_order_relaxed);
    }
    return true;
}();
inline int opt(Opt o) { return g_opt[o].load(std::memory_order_relaxed); }

// ---- optional per-kernel-class timing (bench.py's roofline): CUDA events around every launch
enum ProfClass { kProfPrologue = 0, kProfIterFwd, kProfBwdState, kProfBwdParam, kProfFinalBwd,
                 kProfIterBwdV1, kProfDcnFwd, kProfDcnBwd, kProfBwdTable, kProfBwdGather, kProfHeads, kProfHeadsGrad, kProfClasses };
const char *const kProfNames[kProfClasses] = {"prologue_fwd_kernel", "iter_fwd_kernel", "bwd_state_kernel",
                                              "bwd_param_kernel", "final_bwd_kernel", "iter_bwd_kernel",
                                              "dcn_forward", "dcn_backward", "table_build_kernel",
                                              "bwd_gather_kernel", "head_fused_kernel", "head_wgrad_kernel"};
struct ProfRec { int cls; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::atomic<int> g_prof_on{0};
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;

cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    cudaStream_t st;
    int idx = -1;
    ProfScope(int cls, cudaStream_t s) : st(s)
    {
        if (!g_prof_on.load(std::memory_order_relaxed)) return;
        std::lock_guard<std::mutex> lock(g_prof_mu);
        ProfRec r{cls, prof_event(), prof_event()};
        cudaEventRecord(r.e0, st);
        g_prof_recs.push_back(r);
        idx = (int)g_prof_recs.size() - 1;
    }
    ~ProfScope()
    {
        if (idx < 0) return;
        std::lock_guard<std::mutex> lock(g_prof_mu);
        if (idx < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[idx].e1, st);
    }
};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

#define NLSPN_CHECK_LAUNCH(what)                              \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);  \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
    } while (0)

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_shape(int B, int H, int W, int K, int T)
{
    if (B <= 0 || H <= 0 || W <= 0 || T <= 0)
        return fail(NLSPN_ERR_SHAPE, "B, H, W, T must be positive (got %d, %d, %d, %d)", B, H, W, T);
    if (K != 3 && K != 5 && K != 7)
        return fail(NLSPN_ERR_KERNEL, "prop_kernel must be 3, 5 or 7 (got %d)", K);
    if (B > 65535) return fail(NLSPN_ERR_SHAPE, "B > 65535 is not supported (got %d)", B);
    // per-image plane index must fit int32; whole-tensor offsets are 64-bit
    if ((long)H * W > (1L << 30))
        return fail(NLSPN_ERR_SHAPE, "H*W too large (%ld)", (long)H * W);
    return 0;
}

dim3 grid_for(int P, int B) { return dim3((unsigned)((P + kBlock - 1) / kBlock), (unsigned)B, 1); }

#define DISPATCH_K(K_, ...)                                   \
    switch (K_) {                                             \
    case 3: { constexpr int KC = 3; __VA_ARGS__; } break;     \
    case 5: { constexpr int KC = 5; __VA_ARGS__; } break;     \
    case 7: { constexpr int KC = 7; __VA_ARGS__; } break;     \
    default: return fail(NLSPN_ERR_KERNEL, "prop_kernel must be 3, 5 or 7 (got %d)", K_); \
    }

} // namespace

extern "C" {

int nlspn_abi_version(void) { return NLSPN_ABI_VERSION; }

const char *nlspn_last_error(void) { return g_err; }

unsigned long long nlspn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nlspn_set_option(const char *name, int value)
{
    if (!name) return fail(NLSPN_ERR_NULL, "set_option: NULL name");
    for (int i = 0; i < kOptCount; ++i)
        if (strcmp(name, kOptDefs[i].name) == 0) {
            g_opt[i].store(value, std::memory_order_relaxed);
            return 0;
        }
    return fail(NLSPN_ERR_DOMAIN, "set_option: unknown option '%s'", name);
}

int nlspn_get_option(const char *name, int *value)
{
    if (!name || !value) return fail(NLSPN_ERR_NULL, "get_option: NULL argument");
    for (int i = 0; i < kOptCount; ++i)
        if (strcmp(name, kOptDefs[i].name) == 0) {
            *value = g_opt[i].load(std::memory_order_relaxed);
            return 0;
        }
    return fail(NLSPN_ERR_DOMAIN, "get_option: unknown option '%s'", name);
}

int nlspn_reset_options(void)
{
    for (int i = 0; i < kOptCount; ++i) g_opt[i].store(kOptDefs[i].def, std::memory_order_relaxed);
    return 0;
}

int nlspn_profile_enable(int on)
{
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto &r : g_prof_recs) {
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    g_prof_on.store(on ? 1 : 0);
    return 0;
}

int nlspn_profile_classes(void) { return kProfClasses; }

const char *nlspn_profile_class_name(int cls) { return cls >= 0 && cls < kProfClasses ? kProfNames[cls] : ""; }

int nlspn_profile_read(double *ms, long long *launches, int n)
{
    if (!ms || !launches || n < kProfClasses) return fail(NLSPN_ERR_NULL, "profile_read: need %d slots", kProfClasses);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
    for (auto &r : g_prof_recs) {
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(profile)");
        float t = 0.f;
        e = cudaEventElapsedTime(&t, r.e0, r.e1);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime(profile)");
        ms[r.cls] += t;
        launches[r.cls] += 1;
    }
    return 0;
}

int nlspn_device_info(int device, int *sm_count, int *l2_bytes)
{
    // one slot per device, filled once (DataParallel worker threads each drive their own device)
    constexpr int kMaxDev = 64;
    struct Slot { std::atomic<int> ready{0}; int sm = 0, l2 = 0, coop = 0; };
    static Slot slots[kMaxDev];
    static std::mutex mu;
    if (device < 0 || device >= kMaxDev) return fail(NLSPN_ERR_SHAPE, "device index %d out of range", device);
    Slot &s = slots[device];
    if (!s.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!s.ready.load(std::memory_order_relaxed)) {
            int sm = 0, l2 = 0;
            cudaError_t e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
            if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(SM count)");
            e = cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device);
            if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(L2 size)");
            s.sm = sm;
            s.l2 = l2;
            s.ready.store(1, std::memory_order_release);
        }
    }
    if (sm_count) *sm_count = s.sm;
    if (l2_bytes) *l2_bytes = s.l2;
    return 0;
}

// ---- image groups -----------------------------------------------------------------------
// Batch images propagate independently, so the host loop is GROUP-major: all T
// iterations of a group of G images are enqueued before the next group starts.  G is sized so
// that a group's offsets + affinities (+ the blocked scatter planes in the backward) stay
// resident in L2 across the iterations: HBM then sees each byte of them once per pass instead of
// once per iteration, and the backward's vector REDs hit L2 instead of missing to DRAM (ncu on
// the batch-major order: 958 us per bwd_state launch, long_scoreboard/lg_throttle stalls with
// every unit < 20 % busy; see profiles/).
static int group_images(int B, int H, int W, int K, bool backward)
{
    {
        const int g = opt(kOptGroupImages);
        if (g > 0) return g < B ? g : B;
    }
    // Measured on B200 (KITTI, B=8, K=3, T=18; profiles/): once every load of a thread is issued
    // before its first store the kernels no longer care whether a group is L2-resident
    // (G=1: 6.58, G=2: 6.82, G=4: 6.89, G=8: 7.09 Gpix*iter/s), and smaller groups cost launch
    // tails.  Default: the whole batch is one group.
    (void)H; (void)W; (void)K; (void)backward;
    return B;
}

// ---- TMA tensor maps -----------------------------------------------------------------------
// A [planes, H, W] fp32 array viewed as a 3-D tensor; box = one halo'd tile of one plane.
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so the
// library needs no link-time dependency on libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static bool tiled_enabled() { return opt(kOptTiled) != 0; }

// true when the tiled (TMA) kernels can serve this shape
static bool tiled_ok(const void *base, int W)
{
    return tiled_enabled() && (W % 4) == 0 && aligned16(base) && encode_tiled_fn() != nullptr;
}

// A tensor map only encodes (address, extents, box), so it can be cached by value: the torch caching
// allocator hands the same addresses back step after step, and cuTensorMapEncodeTiled (a driver call, 2-4 per
// API call) disappears from the steady-state launch path.  Small LRU, thread-safe.
struct MapKey {
    const void *base;
    long d2, d3;
    int H, W, box_w, box_h, box_c, rank;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && d2 == o.d2 && d3 == o.d3 && H == o.H && W == o.W && box_w == o.box_w &&
               box_h == o.box_h && box_c == o.box_c && rank == o.rank;
    }
};
struct MapSlot { MapKey key; CUtensorMap map; unsigned long long stamp; bool used; };
constexpr int kMapSlots = 32;
static MapSlot g_maps[kMapSlots];
static std::mutex g_maps_mu;
static unsigned long long g_maps_clock = 0;

static bool map_cache_get(const MapKey &k, CUtensorMap *out)
{
    std::lock_guard<std::mutex> lock(g_maps_mu);
    for (int i = 0; i < kMapSlots; ++i)
        if (g_maps[i].used && g_maps[i].key == k) {
            g_maps[i].stamp = ++g_maps_clock;
            *out = g_maps[i].map;
            return true;
        }
    return false;
}

static void map_cache_put(const MapKey &k, const CUtensorMap &m)
{
    std::lock_guard<std::mutex> lock(g_maps_mu);
    int victim = 0;
    for (int i = 0; i < kMapSlots; ++i) {
        if (!g_maps[i].used) { victim = i; break; }
        if (g_maps[i].stamp < g_maps[victim].stamp) victim = i;
    }
    g_maps[victim].key = k;
    g_maps[victim].map = m;
    g_maps[victim].stamp = ++g_maps_clock;
    g_maps[victim].used = true;
}

static int make_plane_map(CUtensorMap *map, const float *base, long planes, int H, int W, int box_w, int box_h)
{
    const MapKey key{base, planes, 0, H, W, box_w, box_h, 1, 3};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

// Launch with programmatic stream serialisation (PDL): the grid may become resident while the
// previous kernel in the stream is still draining; the kernel itself orders its dependent
// accesses with griddepcontrol.wait (tma::grid_dependency_wait).
static bool pdl_enabled() { return opt(kOptPdl) != 0; }

} // extern "C" (templates need C++ linkage)

// `chained` = the previous kernel in the stream is one of OUR PDL-aware kernels AND everything this kernel reads
// before its griddepcontrol.wait is older than that kernel.  The first launch after a plain producer (prologue,
// table build, a memset, the caller's own kernels) is launched without the attribute: the PDL contract only
// makes the producer's writes visible after the wait, and the early loads prefetch geometry.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_smem(bool chained, size_t smem, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                                   cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (chained && pdl_enabled()) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool chained, void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st,
                              Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (chained && pdl_enabled()) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" {

constexpr int kFwdTH = 4;     // default forward tile: 32 x 4 pixels (B200 sweep: TH 4: 45.0, 8: 44.2, 16: 36.5, 32: 26.1 Gpix*iter/s)

// forward tile height: 8, 16 or 32 rows (NLSPN_FWD_TH overrides the default; tuning knob)
static int fwd_tile_h()
{
    const int v = opt(kOptFwdTH);
    return (v == 4 || v == 8) ? v : kFwdTH;
}

// Tile heights: the B200 sweeps settled on 4 rows; 8 is kept as the one alternative.  (2, 16 and
// 32 were measured in round 1 and are no longer instantiated: they only cost compile time.)
#define DISPATCH_TH(TH_, ...)                                  \
    switch (TH_) {                                             \
    case 8: { constexpr int THC = 8; __VA_ARGS__; } break;     \
    default: { constexpr int THC = 4; __VA_ARGS__; } break;    \
    }
// [B, C, H, W] fp32 tensor viewed 4-D; box = {32 px, box_h rows, all C channels, 1 image}
static int make_geometry_map(CUtensorMap *map, const float *base, int B, int C, int H, int W, int box_h)
{
    const MapKey key{base, C, B, H, W, kTileW, box_h, C, 4};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)kTileW, (cuuint32_t)box_h, (cuuint32_t)C, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(4d) failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

// [B, C, H, W] fp32 tensor viewed 4-D with an arbitrary box {box_w px, box_h rows, box_c channels, 1 image}
static int make_nchw_map(CUtensorMap *map, const float *base, int B, int C, int H, int W, int box_w, int box_h, int box_c)
{
    const MapKey key{base, C, B, H, W, box_w, box_h, box_c, 4};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_c, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(4d) failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

// [B, 64, H, W] fp32 tensor with the dims ordered (x, channel, row, image) and a box {32 px, 8 channels, rows, 1}
// in the 128-byte swizzle with 32-byte atoms: the box is a column of MN-major tf32 MMA atoms (kernels_head2.cuh)
static int make_head_rows_map(CUtensorMap *map, const float *base, int B, int H, int W, int rows, int chans = 8)
{
    const MapKey key{base, kHeadCin, B, H, W, 32, rows, chans, 14};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)kHeadCin, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)kHeadCin * H * W * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)chans, (cuuint32_t)rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(head rows) failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

constexpr int kStateTH = 4;   // pass-A TMA tile: 32 x 4 pixels
constexpr int kParamTH = 4;   // default pass-B tile: 32 x 4 pixels (B200 sweep: TH 4: 1.02, 8: 1.10, 16: 1.27 ms)

static int param_tile_h()
{
    const int v = opt(kOptParamTH);
    return (v == 4 || v == 8) ? v : kParamTH;
}

static bool stream_hint_for(int B, int H, int W, int K)
{
    if (opt(kOptStreamHint) >= 0) return opt(kOptStreamHint) != 0;
    int dev = 0, l2 = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || nlspn_device_info(dev, nullptr, &l2) != 0) return true;
    return (double)B * H * W * 4.0 * 3.0 * K * K > 0.5 * (double)l2;   // geometry does not fit L2
}

struct FwdCall {
    const float *guidance, *confidence, *feat_init, *feat_fix;
    const float *gamma;
    int affinity;
    unsigned flags;
    int B, H, W, K, T;
    float *offset, *aff, *conf_fixed, *src;
    int S;
    float *list_feat;
    cudaStream_t st;
};

static int launch_prologue(const FwdCall &c, int b0, int nb)
{
    const int P = c.H * c.W, KK = c.K * c.K, N = KK - 1;
    const long o1 = (long)b0 * P;
    ProfScope prof__(kProfPrologue, c.st);
    const bool no_off = (c.flags & NLSPN_FLAG_NO_OFFSET) != 0;
    DISPATCH_K(c.K, (prologue_fwd_kernel<KC><<<grid_for(P, nb), kBlock, 0, c.st>>>(
                        c.guidance + o1 * (no_off ? 1 : 3) * N, c.confidence ? c.confidence + o1 : nullptr,
                        c.feat_init + o1, c.feat_fix ? c.feat_fix + o1 : nullptr, c.gamma, c.affinity,
                        c.flags, c.H, c.W, no_off ? nullptr : c.offset + o1 * 2 * KK, c.aff + o1 * KK,
                        c.conf_fixed ? c.conf_fixed + o1 : nullptr, c.src + o1)));
    NLSPN_CHECK_LAUNCH("prologue_fwd_kernel");
    return 0;
}

static int launch_iter_fwd(const FwdCall &c, int b0, int nb, int t, const CUtensorMap *src_map,
                           const CUtensorMap *list_map, bool stream)
{
    const int P = c.H * c.W, KK = c.K * c.K;
    const long BP = (long)c.B * P, o1 = (long)b0 * P;
    const float *src_prev;
    float *src_next = nullptr;
    const CUtensorMap *map = src_map;
    int plane_z;
    const bool use_src = c.conf_fixed || (c.flags & NLSPN_FLAG_BLEND_PRE);   // gather source != list_feat
    if (use_src) {
        src_prev = c.src + (long)((t - 1) % c.S) * BP + o1;
        plane_z = ((t - 1) % c.S) * c.B + b0;
        if (t < c.T) src_next = c.src + (long)(t % c.S) * BP + o1;
    } else if (t == 1) {
        src_prev = c.src + o1;
        plane_z = b0;
    } else {
        src_prev = c.list_feat + (long)(t - 2) * BP + o1;
        plane_z = (t - 2) * c.B + b0;
        map = list_map;
    }
    float *out = c.list_feat + (long)(t - 1) * BP + o1;
    const float *cf = c.conf_fixed ? c.conf_fixed + o1 : nullptr;
    const float *fx = c.feat_fix ? c.feat_fix + o1 : nullptr;
    ProfScope prof__(kProfIterFwd, c.st);
    if (c.flags & NLSPN_FLAG_NO_OFFSET) {
        cudaError_t e;
        if (stream)
            e = launch_pdl(t > 1, fixed_fwd_kernel<true>, grid_for(P, nb), dim3(kBlock), c.st, src_prev,
                           c.aff + o1 * KK, cf, fx, c.flags, c.H, c.W, out, src_next);
        else
            e = launch_pdl(t > 1, fixed_fwd_kernel<false>, grid_for(P, nb), dim3(kBlock), c.st, src_prev,
                           c.aff + o1 * KK, cf, fx, c.flags, c.H, c.W, out, src_next);
        if (e != cudaSuccess) return cuda_fail(e, "fixed_fwd_kernel");
        NLSPN_CHECK_LAUNCH("fixed_fwd_kernel");
        return 0;
    }
    if (map) {
        const int th = fwd_tile_h();
        dim3 grid((unsigned)((c.W + kTileW - 1) / kTileW), (unsigned)((c.H + th - 1) / th), (unsigned)nb);
        dim3 block(kTileW, th);
        if (stream) {
            DISPATCH_TH(th, DISPATCH_K(c.K, (launch_pdl(t > 1, iter_fwd_tiled_kernel<KC, THC, true>, grid, block, c.st,
                                *map, plane_z, src_prev, c.offset + o1 * 2 * KK, c.aff + o1 * KK, cf, fx,
                                c.flags, c.H, c.W, out, src_next))));
        } else {
            DISPATCH_TH(th, DISPATCH_K(c.K, (launch_pdl(t > 1, iter_fwd_tiled_kernel<KC, THC, false>, grid, block, c.st,
                                *map, plane_z, src_prev, c.offset + o1 * 2 * KK, c.aff + o1 * KK, cf, fx,
                                c.flags, c.H, c.W, out, src_next))));
        }
        NLSPN_CHECK_LAUNCH("iter_fwd_tiled_kernel");
        return 0;
    }
    DISPATCH_K(c.K, (iter_fwd_kernel<KC, true><<<grid_for(P, nb), kBlock, 0, c.st>>>(
                        src_prev, c.offset + o1 * 2 * KK, c.aff + o1 * KK, cf, fx,
                        nullptr, nullptr, c.flags, c.H, c.W, out, src_next)));
    NLSPN_CHECK_LAUNCH("iter_fwd_kernel");
    return 0;
}

// Persistent forward (kernels_persist.cuh): used when every pixel of the batch can own a
// co-resident thread.  *tried = 1 when the whole forward was enqueued here.
static int persist_capacity(int K, int *blocks_per_sm)
{
    (void)K;   // K == 3 only (checked by the caller)
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, persist_fwd_kernel<3>, kPersistBlock, 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    *blocks_per_sm = n;
    return 0;
}

static int try_persistent_forward(const FwdCall &c, bool prologue, int *tried)
{
    *tried = 0;
    if (c.K != 3) return 0;   // 24 / 48 taps of geometry do not fit the register file
    if (c.flags & NLSPN_FLAG_BLEND_PRE) return 0;   // upstream ordering: per-iteration kernels only
    if (opt(kOptPersist) == 0) return 0;
    int dev = 0, sms = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || nlspn_device_info(dev, &sms, nullptr) != 0) return 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return 0;
    int per_sm = 0;
    if (int rc = persist_capacity(c.K, &per_sm)) return rc;
    const dim3 pgrid((unsigned)((c.W + 31) / 32), (unsigned)((c.H + kPersistTH - 1) / kPersistTH), (unsigned)c.B);
    const long blocks_needed = (long)pgrid.x * pgrid.y * pgrid.z;
    if (per_sm <= 0 || blocks_needed > (long)sms * per_sm) return 0;   // not co-resident: per-iteration path
    if (prologue)
        if (int rc = launch_prologue(c, 0, c.B)) return rc;
    const float *offset = c.offset, *aff = c.aff, *conf = c.conf_fixed, *dep = c.feat_fix;
    unsigned flags = c.flags;
    int H = c.H, W = c.W, B = c.B, T = c.T, S = c.S;
    float *src = c.src, *list_feat = c.list_feat;
    void *args[] = {&offset, &aff, &conf, &dep, &flags, &H, &W, &B, &T, &src, &S, &list_feat};
    ProfScope prof__(kProfIterFwd, c.st);
    cudaError_t e = cudaSuccess;
    e = cudaLaunchCooperativeKernel((const void *)persist_fwd_kernel<3>, pgrid, dim3(32, kPersistTH), args, 0,
                                    c.st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(persist_fwd_kernel)");
    NLSPN_CHECK_LAUNCH("persist_fwd_kernel");
    *tried = 1;
    return 0;
}

// flag combinations the kernels implement
static int check_mode_flags(unsigned flags)
{
    const bool pre = flags & NLSPN_FLAG_BLEND_PRE, sampled = flags & NLSPN_FLAG_CONF_SAMPLED;
    if ((pre || sampled) && (flags & (NLSPN_FLAG_ALWAYS_CLIP | NLSPN_FLAG_NO_OFFSET | NLSPN_FLAG_BWD_PER_ITERATION)))
        return fail(NLSPN_ERR_DOMAIN, "BLEND_PRE / CONF_SAMPLED (upstream semantics) cannot be combined with "
                                      "ALWAYS_CLIP, NO_OFFSET or BWD_PER_ITERATION");
    if ((flags & NLSPN_FLAG_LEGACY) && !sampled)
        return fail(NLSPN_ERR_DOMAIN, "LEGACY only modifies CONF_SAMPLED");
    if ((flags & NLSPN_FLAG_DETERMINISTIC) &&
        (flags & (NLSPN_FLAG_NO_OFFSET | NLSPN_FLAG_BWD_PER_ITERATION | NLSPN_FLAG_CONF_SAMPLED)))
        return fail(NLSPN_ERR_DOMAIN, "DETERMINISTIC cannot be combined with NO_OFFSET, BWD_PER_ITERATION or CONF_SAMPLED "
                                      "(their scatters are atomic)");
    return 0;
}

static int check_fwd(const FwdCall &c, bool prologue, bool iters)
{
    if (int rc = check_shape(c.B, c.H, c.W, c.K, c.T)) return rc;
    if ((c.flags & NLSPN_FLAG_PRESERVE_INPUT) && !c.feat_fix)
        return fail(NLSPN_ERR_NULL, "PRESERVE_INPUT needs feat_fix");
    const bool no_off = (c.flags & NLSPN_FLAG_NO_OFFSET) != 0;
    if (no_off && c.K != 3)
        return fail(NLSPN_ERR_KERNEL, "fixed-local propagation (NO_OFFSET) is 3x3 only, as in the reference (got K=%d)", c.K);
    if (int rc = check_mode_flags(c.flags)) return rc;
    if (prologue && (c.flags & NLSPN_FLAG_CONF_SAMPLED) && (!c.confidence || c.conf_fixed))
        return fail(NLSPN_ERR_NULL, "CONF_SAMPLED needs `confidence` and a NULL `conf_fixed` (the loop does not pre-multiply)");
    if (prologue) {
        if (!c.guidance || !c.feat_init || (!c.offset && !no_off) || !c.aff || !c.src || !c.gamma)
            return fail(NLSPN_ERR_NULL, "prologue: guidance, feat_init, gamma, offset, aff, src are required");
        if (c.confidence && !c.conf_fixed && !(c.flags & NLSPN_FLAG_CONF_SAMPLED))
            return fail(NLSPN_ERR_NULL, "prologue: conf_fixed is required when confidence is given");
        if (c.affinity < NLSPN_AFF_AS || c.affinity > NLSPN_AFF_TGASS)
            return fail(NLSPN_ERR_AFFINITY, "unknown affinity mode %d", c.affinity);
    }
    if (iters) {
        if ((!c.offset && !no_off) || !c.aff || !c.src || !c.list_feat)
            return fail(NLSPN_ERR_NULL, "propagate: offset, aff, src, list_feat are required");
        if (c.S < 1 || ((c.conf_fixed || (c.flags & NLSPN_FLAG_BLEND_PRE)) && c.T > 1 && c.S < 2))
            return fail(NLSPN_ERR_SHAPE, "propagate: src needs S >= 2 planes with confidence / BLEND_PRE (got %d)", c.S);
    }
    return 0;
}

static int run_forward(const FwdCall &c, bool prologue, bool iters)
{
    if (int rc = check_fwd(c, prologue, iters)) return rc;
    const int G = group_images(c.B, c.H, c.W, c.K, false);
    CUtensorMap src_map, list_map;
    const CUtensorMap *psrc = nullptr, *plist = nullptr;
    if (iters && !(c.flags & NLSPN_FLAG_NO_OFFSET) && tiled_ok(c.src, c.W) && aligned16(c.list_feat)) {
        if (int rc = make_plane_map(&src_map, c.src, (long)c.S * c.B, c.H, c.W, kTileW + 2 * halo_for(c.K), fwd_tile_h() + 2 * halo_for(c.K))) return rc;
        if (int rc = make_plane_map(&list_map, c.list_feat, (long)c.T * c.B, c.H, c.W, kTileW + 2 * halo_for(c.K), fwd_tile_h() + 2 * halo_for(c.K))) return rc;
        psrc = &src_map;
        plist = &list_map;
    }
    const bool stream = stream_hint_for(G, c.H, c.W, c.K);
    if (iters && !(c.flags & NLSPN_FLAG_NO_OFFSET)) {
        int tried = 0;
        if (int rc = try_persistent_forward(c, prologue, &tried)) return rc;
        if (tried) return 0;
    }
    for (int b0 = 0; b0 < c.B; b0 += G) {
        const int nb = c.B - b0 < G ? c.B - b0 : G;
        if (prologue)
            if (int rc = launch_prologue(c, b0, nb)) return rc;
        if (iters)
            for (int t = 1; t <= c.T; ++t)
                if (int rc = launch_iter_fwd(c, b0, nb, t, psrc, plist, stream)) return rc;
    }
    return 0;
}

int nlspn_prologue_fwd(const float *guidance, const float *confidence, const float *feat_init,
                       const float *feat_fix, const float *gamma, int affinity, unsigned flags,
                       int B, int H, int W, int K,
                       float *offset, float *aff, float *conf_fixed, float *src0, void *stream)
{
    FwdCall c{guidance, confidence, feat_init, feat_fix, gamma, affinity, flags, B, H, W, K, 1,
              offset, aff, conf_fixed, src0, 1, nullptr, (cudaStream_t)stream};
    return run_forward(c, true, false);
}

int nlspn_propagate_fwd(const float *offset, const float *aff, const float *conf_fixed,
                        const float *feat_fix, unsigned flags, int B, int H, int W, int K, int T,
                        float *src, int S, float *list_feat, void *stream)
{
    FwdCall c{nullptr, nullptr, nullptr, feat_fix, nullptr, 0, flags, B, H, W, K, T,
              const_cast<float *>(offset), const_cast<float *>(aff), const_cast<float *>(conf_fixed), src, S,
              list_feat, (cudaStream_t)stream};
    return run_forward(c, false, true);
}

int nlspn_forward(const float *guidance, const float *confidence, const float *feat_init,
                  const float *feat_fix, const float *gamma, int affinity, unsigned flags,
                  int B, int H, int W, int K, int T,
                  float *offset, float *aff, float *conf_fixed, float *src, int S, float *list_feat,
                  void *stream)
{
    FwdCall c{guidance, confidence, feat_init, feat_fix, gamma, affinity, flags, B, H, W, K, T,
              offset, aff, conf_fixed, src, S, list_feat, (cudaStream_t)stream};
    return run_forward(c, true, true);
}

static size_t ws_bytes_v1(int B, int H, int W, int K)
{
    const size_t BP = (size_t)B * H * W;
    // scatter planes A, B + confidence-gradient accumulator + affinity-gradient accumulator
    return sizeof(float) * (3 * BP + (size_t)K * K * BP);
}

// taps per pass-B chunk: K = 3 one chunk of 9; K = 5 five chunks of 5 (B200, KITTI B=8, T=36: C = 5 / 7 / 9 / 13
// -> 5.01 / 5.69 / 5.30 / 6.86 ms); K = 7 six chunks of 9
constexpr int param_chunk(int K) { return K == 5 ? 5 : 9; }
constexpr int kGatherMinT = 8;
constexpr int kCompactMinT = 24;   // table compaction pays off from about two dozen iterations
constexpr int kGatherMinK = 5;   // pass A: RED scatter for K = 3, tabulated gather for K >= 5 (measured, kernels_gather.cuh)

// Which form of pass A a backward call with this shape takes (the workspace query must agree with the call).
static bool gather_form_selected(int H, int W, int K, int T)
{
    bool gather = K >= kGatherMinK && T >= kGatherMinT;   // the table build (once per call) must amortise
    if (opt(kOptStateGather) >= 0) gather = opt(kOptStateGather) != 0;
    if (H > 32767 || W > 65535) gather = false;           // table entries pack the source pixel as row << 16 | col (signed int)
    return gather;
}

// Pass A as a CTA-local transpose through shared memory (kernels_local.cuh): K = 3, tiled shapes.
static bool local_form_selected(int H, int W, int K, int T)
{
    (void)H;
    if (K != 3 || (W % 4) != 0 || !tiled_enabled() || gather_form_selected(H, W, K, T)) return false;
    return opt(kOptStateLocal) != 0;     // -1 (auto) and 1: on
}

static size_t ws_bytes_local(int B, int H, int W, int K, int T) { return ws_bytes_local_th<8>(B, H, W, K, T); }

static size_t ws_bytes_v2(int B, int H, int W, int K, int T, bool allow_local = true)
{
    if (allow_local && local_form_selected(H, W, K, T)) return ws_bytes_local(B, H, W, K, T);
    const size_t BP = (size_t)B * H * W;
    const ScatterGeo sg = scatter_geo(H, W);
    // three sets of four phase planes (read / scatter / being cleared) + confidence-gradient accumulator
    // + gy for every iteration + raw affinity-gradient accumulator (whole batch; a group uses a prefix)
    const size_t red_form = sizeof(float) * (3 * (size_t)B * sg.image + BP + (size_t)T * BP + (size_t)K * K * BP);
    // gather form (kernels_gather.cuh): entry table [cap][blocks] of 16 B + one set of planes + two overflow
    // planes + block counters + per-pixel overflow masks (8 B reserved) + the accumulators above
    const size_t NB = (size_t)B * sg.plane;
    const size_t gather_form = 16 * (size_t)gather_cap(K) * 32 * (size_t)B * (size_t)table_groups(sg.plane) + sizeof(float) * ((size_t)B * sg.image + 2 * BP) +
                               2 * sizeof(int) * 32 * (size_t)B * (size_t)table_groups(sg.plane) + 8 * BP +
                               sizeof(float) * (BP + (size_t)T * BP + (size_t)K * K * BP) + 64;
    return gather_form_selected(H, W, K, T) ? gather_form : red_form;
}

// deterministic form (kernels_det.cuh): per-block gamma partials, counts / row starts / scan tiles, the CSR entries
// (worst case: every corner of every neighbour tap in the image), one plain plane, the accumulators
static size_t ws_bytes_det(int B, int H, int W, int K, int T)
{
    const size_t P = (size_t)H * W, BP = (size_t)B * P, N = (size_t)K * K - 1;
    const size_t nblk = ((P + kBlock - 1) / kBlock) * B, ntiles = (BP + kDetScanTile - 1) / kDetScanTile;
    return sizeof(double) * nblk + sizeof(unsigned) * (2 * BP + 1 + ntiles) + 64 + sizeof(DetEntry) * 4 * N * BP +
           sizeof(float) * (2 * BP + (size_t)T * BP + (size_t)K * K * BP) + 64;
}

size_t nlspn_backward_workspace_bytes_ex(int B, int H, int W, int K, int T, unsigned flags)
{
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0 || T <= 0) return 0;
    if (flags & NLSPN_FLAG_DETERMINISTIC) return ws_bytes_det(B, H, W, K, T) + sizeof(double) * kGammaSlots;
    const size_t a = ws_bytes_v1(B, H, W, K), b = ws_bytes_v2(B, H, W, K, T);
    return (a > b ? a : b) + sizeof(double) * kGammaSlots;
}

size_t nlspn_backward_workspace_bytes(int B, int H, int W, int K, int T)
{
    return nlspn_backward_workspace_bytes_ex(B, H, W, K, T, 0u);
}

int nlspn_backward(const float *guidance, const float *feat_init, const float *feat_fix,
                   const float *confidence, const float *offset, const float *aff, const float *conf_fixed,
                   const float *src, int S, const float *list_feat, const float *const *g_list,
                   const float *g_offset_ext, const float *g_aff_ext, const float *gamma, int affinity,
                   unsigned flags, int B, int H, int W, int K, int T,
                   float *g_feat_init, float *g_guidance, float *g_confidence, double *g_gamma,
                   void *workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = check_shape(B, H, W, K, T)) return rc;
    const bool no_off = (flags & NLSPN_FLAG_NO_OFFSET) != 0;
    if (no_off && K != 3)
        return fail(NLSPN_ERR_KERNEL, "fixed-local propagation (NO_OFFSET) is 3x3 only (got K=%d)", K);
    if (!guidance || !feat_init || (!offset && !no_off) || !aff || !src || !list_feat || !g_list || !gamma ||
        !g_feat_init || !g_guidance || !g_gamma || !workspace)
        return fail(NLSPN_ERR_NULL, "backward: a required pointer is NULL");
    if (int rc = check_mode_flags(flags)) return rc;
    const bool sampled = (flags & NLSPN_FLAG_CONF_SAMPLED) != 0;
    const bool use_src = conf_fixed || (flags & NLSPN_FLAG_BLEND_PRE);
    if (sampled && (!confidence || conf_fixed))
        return fail(NLSPN_ERR_NULL, "backward: CONF_SAMPLED needs `confidence` and a NULL `conf_fixed`");
    if ((conf_fixed || sampled) && !g_confidence)
        return fail(NLSPN_ERR_NULL, "backward: g_confidence is required when confidence is used");
    if ((flags & NLSPN_FLAG_PRESERVE_INPUT) && !feat_fix)
        return fail(NLSPN_ERR_NULL, "backward: PRESERVE_INPUT needs feat_fix");
    if (affinity < NLSPN_AFF_AS || affinity > NLSPN_AFF_TGASS)
        return fail(NLSPN_ERR_AFFINITY, "unknown affinity mode %d", affinity);
    if (use_src && S < T)
        return fail(NLSPN_ERR_SHAPE, "backward: src must keep all T planes (S=%d, T=%d)", S, T);
    if (workspace_bytes < nlspn_backward_workspace_bytes_ex(B, H, W, K, T, flags) || !aligned16(workspace))
        return fail(NLSPN_ERR_WORKSPACE, "backward: workspace too small (%zu < %zu) or misaligned",
                    workspace_bytes, nlspn_backward_workspace_bytes_ex(B, H, W, K, T, flags));
    const int P = H * W, KK = K * K;
    const long BP = (long)B * P;
    const int N = KK - 1;
    cudaStream_t st = (cudaStream_t)stream;
    // workspace head: kGammaSlots fp64 partial sums of the gamma gradient
    double *gamma_slots = static_cast<double *>(workspace);
    float *ws = static_cast<float *>(workspace) + 2 * kGammaSlots;
    cudaError_t e = cudaMemsetAsync(gamma_slots, 0, sizeof(double) * kGammaSlots, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(gamma slots)");
    struct GammaReduce {   // runs on every return path after the final kernel was enqueued
        double *slots, *out;
        cudaStream_t st;
        bool armed = false;
        ~GammaReduce() { if (armed) gamma_reduce_kernel<<<1, 32, 0, st>>>(slots, out); }
    } gamma_reduce{gamma_slots, g_gamma, st};

    if (no_off) {
        // ---- fixed-local propagation (nlspnmodel.py:209-224): plain scatter planes, scalar REDs
        float *planeA = ws, *planeB = ws + BP, *g_conf_acc = ws + 2 * BP;
        float *gy_all = ws + 3 * BP, *g_aff_acc = gy_all + (long)T * BP;
        e = cudaMemsetAsync(ws, 0, sizeof(float) * 3 * BP, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
        const bool sh = stream_hint_for(B, H, W, K);
        for (int t = T; t >= 1; --t) {
            float *s_out = ((T - t) % 2 == 0) ? planeA : planeB;
            float *s_in = t == T ? nullptr : (((T - t) % 2 == 0) ? planeB : planeA);
            ProfScope prof__(kProfBwdState, st);
            if (sh)
                e = launch_pdl(t < T, fixed_state_kernel<true>, grid_for(P, B), dim3(kBlock), st, aff, conf_fixed,
                               feat_fix, list_feat + (long)(t - 1) * BP, g_list[t - 1], s_in, s_out,
                               gy_all + (long)(t - 1) * BP, g_conf_acc, flags, H, W);
            else
                e = launch_pdl(t < T, fixed_state_kernel<false>, grid_for(P, B), dim3(kBlock), st, aff, conf_fixed,
                               feat_fix, list_feat + (long)(t - 1) * BP, g_list[t - 1], s_in, s_out,
                               gy_all + (long)(t - 1) * BP, g_conf_acc, flags, H, W);
            if (e != cudaSuccess) return cuda_fail(e, "fixed_state_kernel");
            NLSPN_CHECK_LAUNCH("fixed_state_kernel");
        }
        const float *s_last = ((T - 1) % 2 == 0) ? planeA : planeB;
        {
            ProfScope prof__(kProfBwdParam, st);
            fixed_param_kernel<<<grid_for(P, B), kBlock, 0, st>>>(src, list_feat, gy_all, conf_fixed ? 1 : 0,
                                                                  H, W, T, BP, BP, g_aff_acc);
            NLSPN_CHECK_LAUNCH("fixed_param_kernel");
        }
        ProfScope prof__(kProfFinalBwd, st);
        final_bwd_kernel<3, false><<<grid_for(P, B), kBlock, 0, st>>>(
            guidance, feat_init, feat_fix, conf_fixed, s_last, g_aff_acc, g_conf_acc, nullptr, g_aff_ext, gamma,
            affinity, flags, H, W, g_feat_init, g_guidance, g_confidence, gamma_slots);
        NLSPN_CHECK_LAUNCH("final_bwd_kernel");
        gamma_reduce.armed = true;
        return 0;
    }

    if (flags & NLSPN_FLAG_BWD_PER_ITERATION) {
        // ---- v1: per-iteration accumulator RMW + scalar atomics (kept for cross-checking)
        float *planeA = ws, *planeB = ws + BP, *g_conf_acc = ws + 2 * BP, *g_aff_acc = ws + 3 * BP;
        e = cudaMemsetAsync(ws, 0, sizeof(float) * 3 * BP, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
        for (int t = T; t >= 1; --t) {
            const float *src_prev;
            if (conf_fixed) src_prev = src + (long)(t - 1) * BP;
            else src_prev = t == 1 ? src : list_feat + (long)(t - 2) * BP;
            float *s_out = ((T - t) % 2 == 0) ? planeA : planeB;
            float *s_in = t == T ? nullptr : (((T - t) % 2 == 0) ? planeB : planeA);
            ProfScope prof__(kProfIterBwdV1, st);
            DISPATCH_K(K, (iter_bwd_kernel<KC, true><<<grid_for(P, B), kBlock, 0, st>>>(
                              src_prev, offset, aff, conf_fixed, feat_fix, list_feat + (long)(t - 1) * BP,
                              g_list[t - 1], nullptr, s_in, s_out, g_guidance,
                              (long)3 * N * P, g_aff_acc, g_conf_acc, flags, t == T ? 1 : 0, H, W)));
            NLSPN_CHECK_LAUNCH("iter_bwd_kernel");
        }
        const float *s_last = ((T - 1) % 2 == 0) ? planeA : planeB;
        ProfScope prof__(kProfFinalBwd, st);
        DISPATCH_K(K, (final_bwd_kernel<KC, false><<<grid_for(P, B), kBlock, 0, st>>>(
                          guidance, feat_init, feat_fix, conf_fixed, s_last, g_aff_acc, g_conf_acc,
                          g_offset_ext, g_aff_ext, gamma, affinity, flags, H, W, g_feat_init, g_guidance,
                          g_confidence, gamma_slots)));
        NLSPN_CHECK_LAUNCH("final_bwd_kernel");
        gamma_reduce.armed = true;
        return 0;
    }

    // ---- v2: group-major; per group: pass A (state gradient, T launches, REDx4 scatter),
    //      pass B (parameter gradients, one launch), final (prologue backward)
    const ScatterGeo sg = scatter_geo(H, W);
    const int G = group_images(B, H, W, K, true);
    const bool stream_hint = stream_hint_for(G, H, W, K);
    const int state_minb = opt(kOptStateMinB);
    CUtensorMap src_map, list_map;
    const bool use_tiled = tiled_ok(src, W) && aligned16(list_feat);
    if (use_tiled) {
        if (int rc = make_plane_map(&src_map, src, (long)(use_src ? T : 1) * B, H, W, kTileW + 2 * halo_for(K), param_tile_h() + 2 * halo_for(K))) return rc;
        if (int rc = make_plane_map(&list_map, list_feat, (long)T * B, H, W, kTileW + 2 * halo_for(K), param_tile_h() + 2 * halo_for(K))) return rc;
    }
    // pass A with TMA-delivered geometry.  Measured on B200 (KITTI B=8): K=3 3.37 ms vs 3.28 ms with
    // plain coalesced loads (no gain: the RED sectors, not the load requests, bound the kernel);
    // K=5, T=36 16.5 ms vs 18.9 ms (24 taps of geometry no longer live in registers).  Default:
    // K=5 only; NLSPN_STATE_TMA=1/0 forces it on (K<=5; K=7 would need 75 KB of smem per CTA) / off.
    CUtensorMap off_map, aff_map;
    bool state_tma = use_tiled && K == 5 && aligned16(offset) && aligned16(aff);
    if (opt(kOptStateTma) >= 0)
        state_tma = use_tiled && K <= 5 && aligned16(offset) && aligned16(aff) && opt(kOptStateTma) != 0;
    if (state_tma) {
        if (int rc = make_geometry_map(&off_map, offset, B, 2 * KK, H, W, kStateTH)) return rc;
        if (int rc = make_geometry_map(&aff_map, aff, B, KK, H, W, kStateTH)) return rc;
    }
    // ---- deterministic form (NLSPN_FLAG_DETERMINISTIC, kernels_det.cuh): exact CSR table of the transposed
    // operator, rows sorted by source; one gather kernel per iteration, no atomics in any floating-point sum
    if (flags & NLSPN_FLAG_DETERMINISTIC) {
        if ((size_t)P > ((size_t)1 << 24) || (size_t)4 * N * (size_t)BP >= ((size_t)1 << 32))
            return fail(NLSPN_ERR_SHAPE, "DETERMINISTIC: H*W <= 2^24 and 4*(K*K-1)*B*H*W < 2^32 (split the batch)");
        const size_t nblk = ((size_t)(P + kBlock - 1) / kBlock) * B;
        const long ntiles = (BP + kDetScanTile - 1) / kDetScanTile;
        double *gamma_part = reinterpret_cast<double *>(ws);                 // [nblk]
        unsigned *count = reinterpret_cast<unsigned *>(gamma_part + nblk);   // [BP]   (then the fill cursor)
        unsigned *start = count + BP;                                        // [BP + 1]
        unsigned *tile_sum = start + BP + 1;                                 // [ntiles]
        uintptr_t ep = reinterpret_cast<uintptr_t>(tile_sum + ntiles);
        ep = (ep + 15) & ~(uintptr_t)15;
        DetEntry *entries = reinterpret_cast<DetEntry *>(ep);                // [<= 4 N BP]
        float *splane = reinterpret_cast<float *>(entries + (size_t)4 * N * BP);   // [BP]
        float *g_conf_acc = splane + BP;                                     // [BP]
        float *gy_all = g_conf_acc + BP;                                     // [T][BP]
        float *g_aff_acc = gy_all + (size_t)T * BP;                          // [B][KK][P]
        e = cudaMemsetAsync(gamma_part, 0, sizeof(double) * nblk + sizeof(unsigned) * (size_t)BP, st);   // + count
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(row counts)");
        e = cudaMemsetAsync(g_conf_acc, 0, sizeof(float) * (size_t)BP, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(confidence gradient)");
        {
            ProfScope prof__(kProfBwdTable, st);
            DISPATCH_K(K, (det_count_kernel<KC><<<grid_for(P, B), kBlock, 0, st>>>(offset, H, W, count)));
            NLSPN_CHECK_LAUNCH("det_count_kernel");
            det_scan_tiles_kernel<<<(unsigned)ntiles, 256, 0, st>>>(count, BP, tile_sum);
            NLSPN_CHECK_LAUNCH("det_scan_tiles_kernel");
            det_scan_sums_kernel<<<1, 1024, 0, st>>>(tile_sum, ntiles);
            NLSPN_CHECK_LAUNCH("det_scan_sums_kernel");
            det_scan_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(count, BP, tile_sum, start);
            NLSPN_CHECK_LAUNCH("det_scan_apply_kernel");
            DISPATCH_K(K, (det_fill_kernel<KC><<<grid_for(P, B), kBlock, 0, st>>>(offset, aff, H, W, start, count, entries)));
            NLSPN_CHECK_LAUNCH("det_fill_kernel");
            det_sort_kernel<<<(unsigned)((BP + kBlock / 32 - 1) / (kBlock / 32)), kBlock, 0, st>>>(start, BP, entries);
            NLSPN_CHECK_LAUNCH("det_sort_kernel");
        }
        for (int t = T; t >= 0; --t) {
            // t >= 1: gy_t from gy_{t+1};  t == 0: only the gathered plane (gradient reaching x_0)
            const float *gyn = t < T ? gy_all + (size_t)t * BP : nullptr;
            float *gyo = t >= 1 ? gy_all + (size_t)(t - 1) * BP : nullptr;
            const float *xt = t >= 1 ? list_feat + (long)(t - 1) * BP : nullptr;
            const float *ge = t >= 1 ? g_list[t - 1] : nullptr;
            ProfScope prof__(kProfBwdState, st);
            DISPATCH_K(K, (det_state_kernel<KC><<<grid_for(P, B), kBlock, 0, st>>>(
                              start, entries, aff, conf_fixed, feat_fix, xt, ge, gyn, gyo, splane, g_conf_acc, flags, H, W)));
            NLSPN_CHECK_LAUNCH("det_state_kernel");
        }
        {
            const int nch = (KK + param_chunk(K) - 1) / param_chunk(K);
            ProfScope prof__(kProfBwdParam, st);
            if (use_tiled) {
                const int pth = param_tile_h();
                dim3 tgrid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + pth - 1) / pth), (unsigned)(B * nch));
                dim3 tblock(kTileW, pth);
                DISPATCH_TH(pth, DISPATCH_K(K, (bwd_param_tiled_kernel<KC, param_chunk(KC), THC, 2><<<tgrid, tblock, 0, st>>>(
                                  src_map, list_map, B, 0, offset, aff, src, list_feat, gy_all, use_src ? 1 : 0, H, W, T,
                                  BP, BP, g_guidance, g_aff_acc))));
                NLSPN_CHECK_LAUNCH("bwd_param_tiled_kernel");
            } else {
                dim3 grid((unsigned)((P + kParamBlock - 1) / kParamBlock), (unsigned)B, (unsigned)nch);
                DISPATCH_K(K, (bwd_param_kernel<KC, param_chunk(KC)><<<grid, kParamBlock, 0, st>>>(
                                  offset, aff, src, list_feat, gy_all, use_src ? 1 : 0, H, W, T, BP, BP, g_guidance, g_aff_acc)));
                NLSPN_CHECK_LAUNCH("bwd_param_kernel");
            }
        }
        {
            ProfScope prof__(kProfFinalBwd, st);
            DISPATCH_K(K, (final_bwd_kernel<KC, 0, false><<<grid_for(P, B), kBlock, 0, st>>>(
                              guidance, feat_init, feat_fix, conf_fixed, splane, g_aff_acc, g_conf_acc, g_offset_ext,
                              g_aff_ext, gamma, affinity, flags, H, W, g_feat_init, g_guidance, g_confidence, gamma_part)));
            NLSPN_CHECK_LAUNCH("final_bwd_kernel");
            gamma_reduce_det_kernel<<<1, 256, 0, st>>>(gamma_part, (long)nblk, g_gamma);
            NLSPN_CHECK_LAUNCH("gamma_reduce_det_kernel");
        }
        return 0;
    }

    // ---- pass A in gather form (kernels_gather.cuh): table built once, then per iteration an elementwise gy
    // kernel and a gather-reduce kernel with plain stores.  Default for K >= 5; NLSPN_STATE_GATHER=0/1 overrides.
    const bool gather = gather_form_selected(H, W, K, T);
    if (gather) {
        const long NB = sg.plane;                       // blocks per image
        const int cap = gather_cap(K);
        float4 *entries = reinterpret_cast<float4 *>(ws);                     // [G][groups of 32 blocks][cap][32]
        float *planes = reinterpret_cast<float *>(entries + (size_t)G * table_groups(NB) * cap * 32);   // [G][sg.image]
        void *ovf = planes + (size_t)G * sg.image;                            // [G][P] masks, 8 B reserved each
        const long NBpad = table_groups(NB) * 32;
        const bool compact = opt(kOptGatherCompact) >= 0 ? opt(kOptGatherCompact) != 0 : T >= kCompactMinT;
        int *count = reinterpret_cast<int *>(static_cast<char *>(ovf) + 8 * (size_t)G * P);   // [G][NBpad]
        int *owner = count + (size_t)G * NBpad;                                               // [G][NBpad]
        float *f0 = reinterpret_cast<float *>(owner + (size_t)G * NBpad), *f1 = f0 + (size_t)G * P;
        float *g_conf_acc = f1 + (size_t)G * P;
        float *gy_all = g_conf_acc + (size_t)G * P;         // [T, G, P]
        float *g_aff_acc = gy_all + (size_t)T * G * P;      // [G, KK, P]
        for (int b0 = 0; b0 < B; b0 += G) {
            const int nb = B - b0 < G ? B - b0 : G;
            const long o1 = (long)b0 * P;
            const long GP = (long)nb * P;
            e = cudaMemsetAsync(f0, 0, sizeof(float) * 3 * (size_t)G * P, st);            // f0, f1, g_conf_acc
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
            e = cudaMemsetAsync(count, 0, sizeof(int) * (size_t)G * NBpad, st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(block counters)");
            const float *cf = conf_fixed ? conf_fixed + o1 : nullptr;
            const float *fx = feat_fix ? feat_fix + o1 : nullptr;
            const float *off_g = offset + o1 * 2 * KK, *aff_g = aff + o1 * KK;
            {
                ProfScope prof__(kProfBwdTable, st);
                DISPATCH_K(K, (table_build_kernel<KC><<<grid_for(P, nb), kBlock, 0, st>>>(
                                  off_g, aff_g, H, W, count, entries,
                                  static_cast<typename OvfMask<KC>::type *>(ovf))));
                NLSPN_CHECK_LAUNCH("table_build_kernel");
                // compaction (columns of a group sorted by entry count) costs about a third of the build and saves
                // ~11 % of every gather launch: worth it for long runs (K=5, T=36: 26.6 -> 25.9 ms; K=7, T=12: loses)
                if (compact) {
                    const dim3 cgrid((unsigned)((table_groups(NB) + kBlock / 32 - 1) / (kBlock / 32)), (unsigned)nb, 1);
                    DISPATCH_K(K, (table_compact_kernel<KC><<<cgrid, kBlock, 0, st>>>(H, W, count, entries, owner)));
                    NLSPN_CHECK_LAUNCH("table_compact_kernel");
                }
            }
            const dim3 ggrid((unsigned)((NBpad + kBlock - 1) / kBlock), (unsigned)nb, 1);
            for (int t = T; t >= 1; --t) {
                const float *xt = list_feat + (long)(t - 1) * BP + o1;
                const float *ge = g_list[t - 1] ? g_list[t - 1] + o1 : nullptr;
                float *gyo = gy_all + (long)(t - 1) * GP;
                const float *gyn = t < T ? gy_all + (long)t * GP : nullptr;
                float *f_out = ((T - t) % 2 == 0) ? f0 : f1;
                float *f_in = ((T - t) % 2 == 0) ? f1 : f0;
                {
                    ProfScope prof__(kProfBwdState, st);
                    DISPATCH_K(K, (e = launch_pdl(t < T, bwd_gy_kernel<KC>, grid_for(P, nb), dim3(kBlock), st, off_g, aff_g, cf,
                                                  fx, xt, ge, t < T ? (const float *)planes : (const float *)nullptr,
                                                  f_in, f_out, gyn,
                                                  static_cast<const typename OvfMask<KC>::type *>(ovf), gyo,
                                                  g_conf_acc, flags, H, W)));
                    if (e != cudaSuccess) return cuda_fail(e, "bwd_gy_kernel");
                    NLSPN_CHECK_LAUNCH("bwd_gy_kernel");
                }
                {
                    ProfScope prof__(kProfBwdGather, st);
                    DISPATCH_K(K, (e = launch_pdl(true, bwd_gather_kernel<KC>, ggrid, dim3(kBlock), st, (const int *)count,
                                                  compact ? (const int *)owner : (const int *)nullptr,
                                                  (const float4 *)entries, (const float *)gyo, H, W,
                                                  reinterpret_cast<float4 *>(planes))));
                    if (e != cudaSuccess) return cuda_fail(e, "bwd_gather_kernel");
                    NLSPN_CHECK_LAUNCH("bwd_gather_kernel");
                }
            }
            const float *f_last = ((T - 1) % 2 == 0) ? f0 : f1;
            {
                const int nch = (KK + param_chunk(K) - 1) / param_chunk(K);
                dim3 grid((unsigned)((P + kParamBlock - 1) / kParamBlock), (unsigned)nb, (unsigned)nch);
                ProfScope prof__(kProfBwdParam, st);
                if (use_tiled) {
                    const int pth = param_tile_h();
                    dim3 tgrid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + pth - 1) / pth),
                               (unsigned)(nb * nch));
                    dim3 tblock(kTileW, pth);
                    DISPATCH_TH(pth, DISPATCH_K(K, (bwd_param_tiled_kernel<KC, param_chunk(KC), THC, 2><<<tgrid, tblock, 0, st>>>(
                                      src_map, list_map, B, b0, off_g, aff_g, src + o1, list_feat + o1, gy_all,
                                      use_src ? 1 : 0, H, W, T, BP, GP, g_guidance + o1 * 3 * N, g_aff_acc))));
                    NLSPN_CHECK_LAUNCH("bwd_param_tiled_kernel");
                } else {
                    DISPATCH_K(K, (bwd_param_kernel<KC, param_chunk(KC)><<<grid, kParamBlock, 0, st>>>(
                                      off_g, aff_g, src + o1, list_feat + o1, gy_all, use_src ? 1 : 0, H, W, T, BP, GP,
                                      g_guidance + o1 * 3 * N, g_aff_acc)));
                    NLSPN_CHECK_LAUNCH("bwd_param_kernel");
                }
            }
            if (sampled) {
                e = cudaMemsetAsync(g_confidence + o1, 0, sizeof(float) * (size_t)nb * P, st);
                if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_confidence)");
            }
            ProfScope prof__(kProfFinalBwd, st);
            if (sampled) {
                DISPATCH_K(K, (final_bwd_kernel<KC, true, true><<<grid_for(P, nb), kBlock, 0, st>>>(
                                  guidance + o1 * 3 * N, feat_init + o1, fx, cf, planes, g_aff_acc, g_conf_acc,
                                  g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr,
                                  g_aff_ext ? g_aff_ext + o1 * KK : nullptr, gamma, affinity, flags, H, W,
                                  g_feat_init + o1, g_guidance + o1 * 3 * N, g_confidence + o1, gamma_slots,
                                  confidence + o1, gy_all, aff_g, f_last)));
            } else {
                DISPATCH_K(K, (final_bwd_kernel<KC, true, false><<<grid_for(P, nb), kBlock, 0, st>>>(
                                  guidance + o1 * 3 * N, feat_init + o1, fx, cf, planes, g_aff_acc, g_conf_acc,
                                  g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr,
                                  g_aff_ext ? g_aff_ext + o1 * KK : nullptr, gamma, affinity, flags, H, W,
                                  g_feat_init + o1, g_guidance + o1 * 3 * N, g_confidence ? g_confidence + o1 : nullptr,
                                  gamma_slots, nullptr, gy_all, aff_g, f_last)));
            }
            NLSPN_CHECK_LAUNCH("final_bwd_kernel");
        }
        gamma_reduce.armed = true;
        return 0;
    }

    // ---- pass A as a CTA-local transpose (kernels_local.cuh): K = 3 default.  One schedule build per call, then per
    // iteration ONE kernel whose only global scatter is a TMA tensor reduction of the tile's 48 x 24 region.
    const bool local_form = local_form_selected(H, W, K, T) && use_tiled;
    if (!local_form && local_form_selected(H, W, K, T) &&
        workspace_bytes < ws_bytes_v2(B, H, W, K, T, false) + sizeof(double) * kGammaSlots)
        return fail(NLSPN_ERR_WORKSPACE, "backward: tensors are not 16-byte aligned, so the scatter form is needed, "
                                         "and the workspace is too small for it (%zu < %zu)",
                    workspace_bytes, ws_bytes_v2(B, H, W, K, T, false) + sizeof(double) * kGammaSlots);
    auto run_local = [&](auto th_tag) -> int {
        constexpr int TH = decltype(th_tag)::value;
        constexpr int MB_LO = TH == 8 ? 4 : 1, MB_MID = TH == 8 ? 5 : 2, MB_HI = TH == 8 ? 6 : 3;   // CTAs per SM
        using L = LocalGeo<3, TH>;
        const PadGeo pg = pad_geo(H, W, L::R);
        const long tiles = local_tiles<TH>(H, W);
        float *planes = ws;                                              // [2][G][PH][PW]
        float *g_conf_acc = planes + 2 * (size_t)G * pg.plane;           // [G][P]
        float *gy_all = g_conf_acc + (size_t)G * P;                      // [T][G][P]
        float *g_aff_acc = gy_all + (size_t)T * G * P;                   // [G][KK][P]
        uintptr_t sp = reinterpret_cast<uintptr_t>(g_aff_acc + (size_t)G * KK * P);
        sp = (sp + 15) & ~(uintptr_t)15;
        uint4 *geo = reinterpret_cast<uint4 *>(sp);                      // [G][tiles][GEOV][256]
        unsigned short *sched_tab = reinterpret_cast<unsigned short *>(geo + (size_t)G * tiles * L::GEOV * L::NT);
        CUtensorMap plane_map;
        if (int rc = make_plane_map(&plane_map, planes, 2L * G, pg.PH, pg.PW, L::RW, L::RH)) return rc;
        cudaError_t smem_err = cudaSuccess;
        {
            auto set = [&smem_err](auto kernel) {
                const cudaError_t e1 = ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), (int)LocalSmem<3, TH>::bytes);
                if (e1 != cudaSuccess && smem_err == cudaSuccess) smem_err = e1;
            };
            set(bwd_state_local_kernel<3, TH, true, MB_LO>);
            set(bwd_state_local_kernel<3, TH, false, MB_LO>);
            set(bwd_state_local_kernel<3, TH, true, MB_MID>);
            set(bwd_state_local_kernel<3, TH, false, MB_MID>);
            set(bwd_state_local_kernel<3, TH, true, MB_HI>);
            set(bwd_state_local_kernel<3, TH, false, MB_HI>);
        }
        if (smem_err != cudaSuccess) return cuda_fail(smem_err, "cudaFuncSetAttribute(bwd_state_local_kernel)");
        const int local_minb = opt(kOptLocalMinB);
        // L2 prefetch distance (in tiles) of the packed records; -1 = one wave of resident CTAs.  Measured on B200
        // (KITTI B=8): no gain at 5 CTAs per SM (2.48 vs 2.46 ms of pass A per step), so the default is off.
        int pf_dist = opt(kOptLocalPrefetch);
        if (pf_dist < 0) {
            int dev = 0, sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess) nlspn_device_info(dev, &sms, nullptr);
            pf_dist = 5 * sms;
        }
        for (int b0 = 0; b0 < B; b0 += G) {
            const int nb = B - b0 < G ? B - b0 : G;
            const long o1 = (long)b0 * P;
            const long GP = (long)nb * P;
            e = cudaMemsetAsync(planes, 0, sizeof(float) * (2 * (size_t)G * pg.plane + (size_t)G * P), st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
            const float *cf = conf_fixed ? conf_fixed + o1 : nullptr;
            const float *fx = feat_fix ? feat_fix + o1 : nullptr;
            const float *off_g = offset + o1 * 2 * KK, *aff_g = aff + o1 * KK;
            const dim3 lgrid((unsigned)((W + L::TW - 1) / L::TW), (unsigned)((H + L::TH - 1) / L::TH), (unsigned)nb);
            const dim3 lblock(L::TW, L::TH);
            {
                ProfScope prof__(kProfBwdTable, st);
                if (opt(kOptSchedMinB) >= 5)
                    sched_build_kernel<3, TH, 5><<<lgrid, lblock, 0, st>>>(off_g, aff_g, cf, fx, flags, H, W, geo, sched_tab);
                else
                    sched_build_kernel<3, TH, 4><<<lgrid, lblock, 0, st>>>(off_g, aff_g, cf, fx, flags, H, W, geo, sched_tab);
                NLSPN_CHECK_LAUNCH("sched_build_kernel");
            }
            for (int t = T; t >= 1; --t) {
                const int io = (T - t) % 2;
                float *s_out = planes + (size_t)io * G * pg.plane;
                float *s_in = t == T ? nullptr : planes + (size_t)(1 - io) * G * pg.plane;
                const float *xt = list_feat + (long)(t - 1) * BP + o1;
                const float *ge = g_list[t - 1] ? g_list[t - 1] + o1 : nullptr;
                float *gyo = gy_all + (long)(t - 1) * GP;
                ProfScope prof__(kProfBwdState, st);
#define LOCAL_LAUNCH(SH_, MB_)                                                                                      \
    e = launch_pdl_smem(t < T, LocalSmem<3, TH>::bytes, bwd_state_local_kernel<3, TH, SH_, MB_>, lgrid, lblock, st, plane_map, \
                        io * G, (const uint4 *)geo, cf ? 1 : 0, xt, ge, s_in, s_out,                                \
                        (const unsigned short *)sched_tab, gyo, g_conf_acc, flags, H, W, pf_dist)
                if (stream_hint) {
                    if (local_minb == 0) LOCAL_LAUNCH(true, MB_LO);
                    else if (local_minb == 2) LOCAL_LAUNCH(true, MB_HI);
                    else LOCAL_LAUNCH(true, MB_MID);
                } else {
                    if (local_minb == 0) LOCAL_LAUNCH(false, MB_LO);
                    else if (local_minb == 2) LOCAL_LAUNCH(false, MB_HI);
                    else LOCAL_LAUNCH(false, MB_MID);
                }
#undef LOCAL_LAUNCH
                if (e != cudaSuccess) return cuda_fail(e, "bwd_state_local_kernel");
                NLSPN_CHECK_LAUNCH("bwd_state_local_kernel");
            }
            const float *s_last = planes + (size_t)((T - 1) % 2) * G * pg.plane;
            {
                ProfScope prof__(kProfBwdParam, st);
                const int pth = param_tile_h();
                dim3 tgrid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + pth - 1) / pth), (unsigned)nb);
                dim3 tblock(kTileW, pth);
                if (opt(kOptParamFactored) != 0) {
                    DISPATCH_TH(pth, (bwd_param_tiled_kernel<3, 9, THC, 2, true><<<tgrid, tblock, 0, st>>>(
                                         src_map, list_map, B, b0, off_g, aff_g, src + o1, list_feat + o1, gy_all,
                                         use_src ? 1 : 0, H, W, T, BP, GP, g_guidance + o1 * 3 * N, g_aff_acc)));
                } else {   // round-1 form (4 scalar LDS per tap, literal expressions): kept for A/B measurements
                    DISPATCH_TH(pth, (bwd_param_tiled_kernel<3, 9, THC, 2, false><<<tgrid, tblock, 0, st>>>(
                                         src_map, list_map, B, b0, off_g, aff_g, src + o1, list_feat + o1, gy_all,
                                         use_src ? 1 : 0, H, W, T, BP, GP, g_guidance + o1 * 3 * N, g_aff_acc)));
                }
                NLSPN_CHECK_LAUNCH("bwd_param_tiled_kernel");
            }
            if (sampled) {
                e = cudaMemsetAsync(g_confidence + o1, 0, sizeof(float) * (size_t)nb * P, st);
                if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_confidence)");
            }
            ProfScope prof__(kProfFinalBwd, st);
            if (sampled) {
                final_bwd_kernel<3, 2, true><<<grid_for(P, nb), kBlock, 0, st>>>(
                    guidance + o1 * 3 * N, feat_init + o1, fx, cf, s_last, g_aff_acc, g_conf_acc,
                    g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr, g_aff_ext ? g_aff_ext + o1 * KK : nullptr,
                    gamma, affinity, flags, H, W, g_feat_init + o1, g_guidance + o1 * 3 * N, g_confidence + o1,
                    gamma_slots, confidence + o1, nullptr, nullptr, nullptr, L::R, pg.PW, pg.plane);
            } else {
                final_bwd_kernel<3, 2, false><<<grid_for(P, nb), kBlock, 0, st>>>(
                    guidance + o1 * 3 * N, feat_init + o1, fx, cf, s_last, g_aff_acc, g_conf_acc,
                    g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr, g_aff_ext ? g_aff_ext + o1 * KK : nullptr,
                    gamma, affinity, flags, H, W, g_feat_init + o1, g_guidance + o1 * 3 * N,
                    g_confidence ? g_confidence + o1 : nullptr, gamma_slots, nullptr, nullptr, nullptr, nullptr,
                    L::R, pg.PW, pg.plane);
            }
            NLSPN_CHECK_LAUNCH("final_bwd_kernel");
        }
        gamma_reduce.armed = true;
        return 0;
    };
    // tile height 8 (256 threads).  32 x 16 tiles (3 instead of 4.5 region cells per pixel) were measured on B200 and
    // lose: pass A 2.88 / 3.28 / 3.49 ms at 1 / 2 / 3 CTAs per SM vs 2.46 ms (16-warp barriers)
    if (local_form) return run_local(std::integral_constant<int, 8>{});

    // scatter-plane sets rotate over three buffers so that clearing is a linear 16-byte-per-thread job
    // (NLSPN_STATE_ZERO3=0: two sets, every thread clears the four cells it read)
    const bool zero3 = opt(kOptStateZero3) != 0;
    float *sets[3] = {ws, ws + (long)G * sg.image, ws + 2 * (long)G * sg.image};
    float *g_conf_acc = sets[2] + (long)G * sg.image;
    float *gy_all = g_conf_acc + (long)G * P;          // [T, G, P]
    float *g_aff_acc = gy_all + (long)T * G * P;       // [G, KK, P]
    for (int b0 = 0; b0 < B; b0 += G) {
        const int nb = B - b0 < G ? B - b0 : G;
        const long o1 = (long)b0 * P;
        const long GP = (long)nb * P;                   // gy planes of this group are [T, nb, P]
        e = cudaMemsetAsync(sets[0], 0, sizeof(float) * (3 * (size_t)G * sg.image + (size_t)G * P), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
        const float *cf = conf_fixed ? conf_fixed + o1 : nullptr;
        const float *fx = feat_fix ? feat_fix + o1 : nullptr;
        const int nset = zero3 ? 3 : 2;
        for (int t = T; t >= 1; --t) {
            const int io = (T - t) % nset;
            float *s_out = sets[io];
            float *s_in = t == T ? nullptr : sets[(io + nset - 1) % nset];
            // the set iteration t+1 read becomes iteration t-1's target: this launch clears it
            float *s_zero = (zero3 && t < T && t > 1) ? sets[(io + 1) % 3] : nullptr;
            ProfScope prof__(kProfBwdState, st);
            const float *xt = list_feat + (long)(t - 1) * BP + o1;
            const float *ge = g_list[t - 1] ? g_list[t - 1] + o1 : nullptr;
            float *gyo = gy_all + (long)(t - 1) * GP;
#define STATE_LAUNCH(SH_, MB_)                                                                          \
    DISPATCH_K(K, (launch_pdl(t < T, bwd_state_kernel<KC, SH_, MB_>, grid_for(P, nb), dim3(kBlock), st,           \
                              offset + o1 * 2 * KK, aff + o1 * KK, cf, fx, xt, ge, s_in, s_out, gyo,       \
                              g_conf_acc, flags, H, W, s_zero)))
            if (state_tma) {
                dim3 sgrid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + kStateTH - 1) / kStateTH), (unsigned)nb);
                if (K == 3) {
                    e = launch_pdl(t < T, bwd_state_tma_kernel<3, kStateTH>, sgrid, dim3(kTileW, kStateTH), st, off_map,
                                   aff_map, b0, cf, fx, xt, ge, s_in, s_out, gyo, g_conf_acc, flags, H, W, s_zero);
                } else {
                    e = launch_pdl(t < T, bwd_state_tma_kernel<5, kStateTH>, sgrid, dim3(kTileW, kStateTH), st, off_map,
                                   aff_map, b0, cf, fx, xt, ge, s_in, s_out, gyo, g_conf_acc, flags, H, W, s_zero);
                }
                if (e != cudaSuccess) return cuda_fail(e, "bwd_state_tma_kernel");
            } else if (stream_hint) {
                switch (state_minb) {
                case 2: STATE_LAUNCH(true, 2); break;
                case 3: STATE_LAUNCH(true, 3); break;
                case 5: STATE_LAUNCH(true, 5); break;
                case 6: STATE_LAUNCH(true, 6); break;
                case 8: STATE_LAUNCH(true, 8); break;
                default: STATE_LAUNCH(true, 4); break;
                }
            } else {
                switch (state_minb) {
                case 2: STATE_LAUNCH(false, 2); break;
                case 3: STATE_LAUNCH(false, 3); break;
                case 5: STATE_LAUNCH(false, 5); break;
                case 6: STATE_LAUNCH(false, 6); break;
                case 8: STATE_LAUNCH(false, 8); break;
                default: STATE_LAUNCH(false, 4); break;
                }
            }
#undef STATE_LAUNCH
            NLSPN_CHECK_LAUNCH("bwd_state_kernel");
        }
        const float *s_last = sets[(T - 1) % nset];
        {
            const int nch = (KK + param_chunk(K) - 1) / param_chunk(K);
            dim3 grid((unsigned)((P + kParamBlock - 1) / kParamBlock), (unsigned)nb, (unsigned)nch);
            ProfScope prof__(kProfBwdParam, st);
            if (use_tiled) {
                const int pth = param_tile_h();
                dim3 tgrid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + pth - 1) / pth),
                           (unsigned)(nb * nch));
                dim3 tblock(kTileW, pth);
#define PARAM_ARGS                                                                                          \
    src_map, list_map, B, b0, offset + o1 * 2 * KK, aff + o1 * KK, src + o1, list_feat + o1, gy_all,          \
        use_src ? 1 : 0, H, W, T, BP, GP, g_guidance + o1 * 3 * N, g_aff_acc
                // (a deeper TMA pipeline -- 3 or 4 boxes in flight -- was measured in round 1: no gain)
                DISPATCH_TH(pth, DISPATCH_K(K, (bwd_param_tiled_kernel<KC, param_chunk(KC), THC, 2><<<tgrid, tblock, 0, st>>>(PARAM_ARGS))));
#undef PARAM_ARGS
                NLSPN_CHECK_LAUNCH("bwd_param_tiled_kernel");
            } else {
                DISPATCH_K(K, (bwd_param_kernel<KC, param_chunk(KC)><<<grid, kParamBlock, 0, st>>>(
                                  offset + o1 * 2 * KK, aff + o1 * KK, src + o1, list_feat + o1, gy_all,
                                  use_src ? 1 : 0, H, W, T, BP, GP, g_guidance + o1 * 3 * N, g_aff_acc)));
                NLSPN_CHECK_LAUNCH("bwd_param_kernel");
            }
        }
        if (sampled) {   // the confidence gradient is scattered by the final kernel
            e = cudaMemsetAsync(g_confidence + o1, 0, sizeof(float) * (size_t)nb * P, st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_confidence)");
        }
        ProfScope prof__(kProfFinalBwd, st);
        if (sampled) {
            DISPATCH_K(K, (final_bwd_kernel<KC, true, true><<<grid_for(P, nb), kBlock, 0, st>>>(
                              guidance + o1 * 3 * N, feat_init + o1, fx, cf, s_last, g_aff_acc, g_conf_acc,
                              g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr,
                              g_aff_ext ? g_aff_ext + o1 * KK : nullptr, gamma, affinity, flags, H, W,
                              g_feat_init + o1, g_guidance + o1 * 3 * N, g_confidence + o1, gamma_slots,
                              confidence + o1)));
        } else {
            DISPATCH_K(K, (final_bwd_kernel<KC, true, false><<<grid_for(P, nb), kBlock, 0, st>>>(
                              guidance + o1 * 3 * N, feat_init + o1, fx, cf, s_last, g_aff_acc, g_conf_acc,
                              g_offset_ext ? g_offset_ext + o1 * 2 * KK : nullptr,
                              g_aff_ext ? g_aff_ext + o1 * KK : nullptr, gamma, affinity, flags, H, W,
                              g_feat_init + o1, g_guidance + o1 * 3 * N, g_confidence ? g_confidence + o1 : nullptr,
                              gamma_slots, nullptr)));
        }
        NLSPN_CHECK_LAUNCH("final_bwd_kernel");
    }
    gamma_reduce.armed = true;
    return 0;
}

// ---- the three final head convolutions as one tcgen05 implicit GEMM (kernels_head.cuh, kernels_head2.cuh) ----
// `packed` = [K-major nine-tap blocks of kernels_head.cuh | per-stage blocks of kernels_head2.cuh (K = 3, 5)]
static long head_rows_packed_floats(int K) { return K == 3 ? HeadRows<3>::packed_floats : K == 5 ? HeadRows<5>::packed_floats : 0; }

size_t nlspn_heads_packed_floats(int K)
{
    if (K != 3 && K != 5 && K != 7) return 0;
    return (size_t)(head_packed_floats(K) + head_rows_packed_floats(K));
}

int nlspn_heads_pack(const float *w_id, const float *w_oa, const float *w_cf, int K, float *packed, void *stream)
{
    if (K != 3 && K != 5 && K != 7) return fail(NLSPN_ERR_KERNEL, "prop_kernel must be 3, 5 or 7 (got %d)", K);
    if (!w_id || !w_oa || !w_cf || !packed) return fail(NLSPN_ERR_NULL, "heads_pack: a required pointer is NULL");
    head_pack_weights_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w_id, w_oa, w_cf, 3 * (K * K - 1), head_np(K), packed);
    NLSPN_CHECK_LAUNCH("head_pack_weights_kernel");
    float *rows = packed + head_packed_floats(K);
    if (K == 3) head_rows_pack_kernel<3><<<64, 256, 0, (cudaStream_t)stream>>>(w_id, w_oa, w_cf, rows);
    else if (K == 5) head_rows_pack_kernel<5><<<64, 256, 0, (cudaStream_t)stream>>>(w_id, w_oa, w_cf, rows);
    NLSPN_CHECK_LAUNCH("head_rows_pack_kernel");
    return 0;
}

// the MN-major operand form applies: K = 3 or 5, rows that are 16-byte multiples, 16-byte aligned tensors
static bool head_rows_ok(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1, int W, int K)
{
    return opt(kOptHeadsRows) != 0 && (K == 3 || K == 5) && tiled_ok(id_fd1, W) && aligned16(oa_fd1) && aligned16(cf_fd1) &&
           aligned16(fe1);
}

static int launch_head_rows(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                            const float *packed, const float *bias, int B, int H, int W, int K, const HeadRowsOut &o,
                            cudaStream_t st)
{
    const int rows = (K == 3 ? HeadRows<3>::ROWS : HeadRows<5>::ROWS);
    CUtensorMap m_id, m_oa, m_cf, m_fe;
    const bool tiles_ks2 = opt(kOptHeadsPersist) != 0 && opt(kOptHeadsKs) == 2;
    const int chans = tiles_ks2 ? 16 : 8;
    if (int rc = make_head_rows_map(&m_id, id_fd1, B, H, W, rows, chans)) return rc;
    if (int rc = make_head_rows_map(&m_oa, oa_fd1, B, H, W, rows, chans)) return rc;
    if (int rc = make_head_rows_map(&m_cf, cf_fd1, B, H, W, rows, chans)) return rc;
    if (int rc = make_head_rows_map(&m_fe, fe1, B, H, W, rows, chans)) return rc;
    const float *rows_packed = packed + head_packed_floats(K);
    if (opt(kOptHeadsPersist) != 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || nlspn_device_info(dev, &sms, nullptr) != 0 || sms <= 0) sms = 148;
        const cudaError_t pe = tiles_ks2 && K == 5 ? head_persist_launch<5, 2>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, sms, opt(kOptHeadsReuse) != 0, o, st)
                             : tiles_ks2 ? head_persist_launch<3, 2>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, sms, opt(kOptHeadsReuse) != 0, o, st)
                             : K == 3 ? head_persist_launch<3>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, sms, opt(kOptHeadsReuse) != 0, o, st)
                                      : head_persist_launch<5>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, sms, opt(kOptHeadsReuse) != 0, o, st);
        if (pe != cudaSuccess) return cuda_fail(pe, "head_persist_kernel");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    const cudaError_t e = K == 3 ? (opt(kOptHeadsRing) == 3 ? head_rows_launch<3, 3>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, o, st)
                                                            : head_rows_launch<3>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, o, st))
                                 : head_rows_launch<5>(m_id, m_oa, m_cf, m_fe, rows_packed, bias, B, H, W, o, st);
    if (e != cudaSuccess) return cuda_fail(e, "head_rows_kernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int nlspn_heads_fwd(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                    const float *packed, const float *bias, int B, int H, int W, int K,
                    float *pred_init, float *guidance, float *confidence, void *stream)
{
    if (int rc = check_shape(B, H, W, K, 1)) return rc;
    if (!id_fd1 || !oa_fd1 || !cf_fd1 || !fe1 || !packed || !bias || !pred_init || !guidance || !confidence)
        return fail(NLSPN_ERR_NULL, "heads_fwd: a required pointer is NULL");
    if (H > 65535) return fail(NLSPN_ERR_SHAPE, "heads_fwd: H > 65535 is not supported (got %d)", H);
    if (!aligned16(packed)) return fail(NLSPN_ERR_ALIGN, "heads_fwd: the packed weights must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int N3 = 3 * (K * K - 1);
    ProfScope prof__(kProfHeads, st);
    if (head_rows_ok(id_fd1, oa_fd1, cf_fd1, fe1, W, K)) {
        HeadRowsOut o{pred_init, confidence, guidance, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0u};
        return launch_head_rows(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, B, H, W, K, o, st);
    }
    cudaError_t e;
    // activations by TMA when the rows are 16-byte multiples (option heads_tma = 0 keeps the direct-load kernel)
    const bool use_tma = opt(kOptHeadsTma) != 0 && tiled_ok(id_fd1, W) && aligned16(oa_fd1) && aligned16(cf_fd1) && aligned16(fe1);
    if (use_tma) {
        CUtensorMap m_id, m_oa, m_cf, m_fe;
        if (int rc = make_nchw_map(&m_id, id_fd1, B, kHeadCin, H, W, kHeadRawW, 3, kHeadChunk)) return rc;
        if (int rc = make_nchw_map(&m_oa, oa_fd1, B, kHeadCin, H, W, kHeadRawW, 3, kHeadChunk)) return rc;
        if (int rc = make_nchw_map(&m_cf, cf_fd1, B, kHeadCin, H, W, kHeadRawW, 3, kHeadChunk)) return rc;
        if (int rc = make_nchw_map(&m_fe, fe1, B, kHeadCin, H, W, kHeadRawW, 3, kHeadChunk)) return rc;
        switch (K) {
        case 3: e = head_launch_tma<head_np(3)>(m_id, m_oa, m_cf, m_fe, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        case 5: e = head_launch_tma<head_np(5)>(m_id, m_oa, m_cf, m_fe, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        default: e = head_launch_tma<head_np(7)>(m_id, m_oa, m_cf, m_fe, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        }
    } else {
        switch (K) {
        case 3: e = head_launch<head_np(3)>(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        case 5: e = head_launch<head_np(5)>(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        default: e = head_launch<head_np(7)>(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, B, H, W, N3, pred_init, guidance, confidence, st); break;
        }
    }
    if (e != cudaSuccess) return cuda_fail(e, "head_fused_kernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int nlspn_heads_prologue_supported(int W, int K)
{
    return opt(kOptHeadsRows) != 0 && (K == 3 || K == 5) && W % 4 == 0 && tiled_enabled() ? 1 : 0;
}

int nlspn_heads_prologue_fwd(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                             const float *packed, const float *bias, const float *feat_fix, const float *gamma,
                             int affinity, unsigned flags, int B, int H, int W, int K,
                             float *pred_init, float *confidence, float *guidance,
                             float *offset, float *aff, float *conf_fixed, float *src0, void *stream)
{
    if (int rc = check_shape(B, H, W, K, 1)) return rc;
    if (!id_fd1 || !oa_fd1 || !cf_fd1 || !fe1 || !packed || !bias || !gamma || !pred_init || !confidence || !offset || !aff || !src0)
        return fail(NLSPN_ERR_NULL, "heads_prologue_fwd: a required pointer is NULL");
    if (affinity < 0 || affinity > 3) return fail(NLSPN_ERR_AFFINITY, "heads_prologue_fwd: unknown affinity mode %d", affinity);
    if (flags & ~(unsigned)(NLSPN_FLAG_PRESERVE_INPUT | NLSPN_FLAG_ALWAYS_CLIP))
        return fail(NLSPN_ERR_DOMAIN, "heads_prologue_fwd: only PRESERVE_INPUT and ALWAYS_CLIP are implemented in the fused epilogue (flags 0x%x)", flags);
    if ((flags & NLSPN_FLAG_PRESERVE_INPUT) && !feat_fix)
        return fail(NLSPN_ERR_NULL, "heads_prologue_fwd: PRESERVE_INPUT needs feat_fix");
    if (H > 65535) return fail(NLSPN_ERR_SHAPE, "heads_prologue_fwd: H > 65535 is not supported (got %d)", H);
    if (!aligned16(packed)) return fail(NLSPN_ERR_ALIGN, "heads_prologue_fwd: the packed weights must be 16-byte aligned");
    if (!head_rows_ok(id_fd1, oa_fd1, cf_fd1, fe1, W, K))
        return fail(NLSPN_ERR_SHAPE, "heads_prologue_fwd: needs prop_kernel 3 or 5, W %% 4 == 0 and 16-byte aligned tensors "
                                     "(got K = %d, W = %d): call nlspn_heads_fwd + nlspn_prologue_fwd instead", K, W);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof__(kProfHeads, st);
    HeadRowsOut o{pred_init, confidence, guidance, feat_fix, gamma, offset, aff, conf_fixed, src0, affinity, flags};
    return launch_head_rows(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, B, H, W, K, o, st);
}

// ---- weight gradient of the head convolutions on tcgen05 (kernels_head_wgrad.cuh) ----
// [imgs, chans, H, W] fp32 tensor with the dims ordered (x, channel, row, image) and a box {32 px, box_c channels, box_r rows, 1}
// in the 128-byte swizzle: box row = one K-major tf32 operand row (32 pixels)
static int make_wgrad_map(CUtensorMap *map, const float *base, int imgs, int chans, int H, int W, int box_c, int box_r)
{
    const MapKey key{base, chans, imgs, H, W, 32, box_r, box_c, 15};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)chans, (cuuint64_t)H, (cuuint64_t)imgs};
    const cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)chans * H * W * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_c, (cuuint32_t)box_r, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(head wgrad) failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

int nlspn_heads_wgrad_supported(int W, int K)
{
    return (K == 3 || K == 5 || K == 7) && W > 0 && W % 4 == 0 && tiled_enabled() ? 1 : 0;
}

static int check_heads_grad(const char *who, int B, int H, int W, int K)
{
    if (int rc = check_shape(B, H, W, K, 1)) return rc;
    if (!nlspn_heads_wgrad_supported(W, K))
        return fail(NLSPN_ERR_SHAPE, "%s: needs W %% 4 == 0 and the TMA path (got W = %d): use the framework's own "
                                     "convolution gradients instead", who, W);
    if (H > 65535 || 3L * B > 0x7fffffffL) return fail(NLSPN_ERR_SHAPE, "%s: H > 65535 is not supported (got %d)", who, H);
    return 0;
}

// the tensor maps need the driver's cuTensorMapEncodeTiled: absent = no CUDA driver in this process
static int need_tma_driver(const char *who)
{
    return encode_tiled_fn() != nullptr ? 0 : fail(NLSPN_ERR_DOMAIN, "%s: cuTensorMapEncodeTiled is not available (no CUDA driver?)", who);
}

int nlspn_heads_grad_prep(const float *g_init, const float *pred_init, const float *g_guidance, const float *g_confidence,
                          const float *confidence, int B, int H, int W, int K, float *g_shift, float *g_bias, void *stream)
{
    if (int rc = check_heads_grad("heads_grad_prep", B, H, W, K)) return rc;
    if (!pred_init || !confidence || !g_shift) return fail(NLSPN_ERR_NULL, "heads_grad_prep: a required pointer is NULL");
    if (!aligned16(g_shift) || !aligned16(pred_init) || !aligned16(confidence) || !aligned16(g_init) || !aligned16(g_guidance) ||
        !aligned16(g_confidence))
        return fail(NLSPN_ERR_ALIGN, "heads_grad_prep: tensors must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int NT = 3 * (K * K - 1) + 2;
    if (g_bias) {
        const cudaError_t e = cudaMemsetAsync(g_bias, 0, sizeof(float) * NT, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_bias)");
    }
    const long quads = (long)H * (W / 4);
    dim3 grid((unsigned)((quads + kGradPrepQuads - 1) / kGradPrepQuads), (unsigned)NT, (unsigned)B);
    ProfScope prof__(kProfHeadsGrad, st);
    head_grad_prep_kernel<<<grid, kGradPrepThreads, 0, st>>>(g_init, pred_init, g_guidance, g_confidence, confidence, B, NT, H, W,
                                                             g_shift, g_bias);
    NLSPN_CHECK_LAUNCH("head_grad_prep_kernel");
    return 0;
}

int nlspn_heads_dgrad_one(const float *g_all, const float *w_id, const float *w_cf, int B, int H, int W, int K,
                          float *d_id_fd1, float *d_cf_fd1, void *stream)
{
    if (int rc = check_heads_grad("heads_dgrad_one", B, H, W, K)) return rc;
    if (!g_all || (d_id_fd1 && !w_id) || (d_cf_fd1 && !w_cf)) return fail(NLSPN_ERR_NULL, "heads_dgrad_one: a required pointer is NULL");
    if (!aligned16(g_all) || !aligned16(d_id_fd1) || !aligned16(d_cf_fd1))
        return fail(NLSPN_ERR_ALIGN, "heads_dgrad_one: tensors must be 16-byte aligned");
    if (!d_id_fd1 && !d_cf_fd1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long quads = (long)B * H * (W / 4);
    dim3 grid((unsigned)((quads + kDgradOneThreads - 1) / kDgradOneThreads), 2u);
    ProfScope prof__(kProfHeadsGrad, st);
    head_dgrad_one_kernel<<<grid, kDgradOneThreads, 0, st>>>(g_all, w_id, w_cf, B, 3 * (K * K - 1) + 2, H, W, d_id_fd1, d_cf_fd1);
    NLSPN_CHECK_LAUNCH("head_dgrad_one_kernel");
    return 0;
}

// ---- the two wide data gradients of the head convolutions as one tcgen05 GEMM (kernels_head_dgrad.cuh; prop_kernel 3) ----
size_t nlspn_heads_dgrad_packed_floats(int K) { return K == 3 ? (size_t)HeadDgrad::packed_floats : 0; }

int nlspn_heads_dgrad_supported(int W, int K) { return K == 3 && nlspn_heads_wgrad_supported(W, K) ? 1 : 0; }

int nlspn_heads_dgrad_pack(const float *w_id, const float *w_oa, const float *w_cf, int K, float *packed, void *stream)
{
    if (K != 3) return fail(NLSPN_ERR_KERNEL, "heads_dgrad_pack: implemented for prop_kernel 3 (got %d)", K);
    if (!w_id || !w_oa || !w_cf || !packed) return fail(NLSPN_ERR_NULL, "heads_dgrad_pack: a required pointer is NULL");
    head_dgrad_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w_id, w_oa, w_cf, 3 * (K * K - 1) + 2, packed);
    NLSPN_CHECK_LAUNCH("head_dgrad_pack_kernel");
    return 0;
}

int nlspn_heads_dgrad_wide(const float *g_shift, const float *packed, int B, int H, int W, int K, float *d_oa_fd1, float *d_fe1,
                           void *stream)
{
    if (int rc = check_heads_grad("heads_dgrad_wide", B, H, W, K)) return rc;
    if (K != 3) return fail(NLSPN_ERR_KERNEL, "heads_dgrad_wide: implemented for prop_kernel 3 (got %d): use the framework's "
                                              "convolution backward", K);
    if (!g_shift || !packed || !d_fe1) return fail(NLSPN_ERR_NULL, "heads_dgrad_wide: a required pointer is NULL");
    if (!aligned16(g_shift) || !aligned16(packed) || !aligned16(d_oa_fd1) || !aligned16(d_fe1))
        return fail(NLSPN_ERR_ALIGN, "heads_dgrad_wide: tensors must be 16-byte aligned");
    if (int rc = need_tma_driver("heads_dgrad_wide")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int NT = 3 * (K * K - 1) + 2;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || nlspn_device_info(dev, &sms, nullptr) != 0 || sms <= 0) sms = 148;
    // g_shift as [3 B images][NT channels][H][W]: dims (x, channel, row, image), box {32 px, 8 channels, R + 2 rows}, 32-byte-atom swizzle
    CUtensorMap mg;
    {
        const MapKey key{g_shift, NT, 3 * B, H, W, 32, HeadDgrad::ROWS, 8, 17};
        if (!map_cache_get(key, &mg)) {
            const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)NT, (cuuint64_t)H, (cuuint64_t)(3 * B)};
            const cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)NT * H * W * 4};
            const cuuint32_t box[4] = {32, 8, (cuuint32_t)HeadDgrad::ROWS, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            const CUresult r = encode_tiled_fn()(&mg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(g_shift), dims, strides,
                                                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(head dgrad) failed (CUresult %d)", (int)r);
            map_cache_put(key, mg);
        }
    }
    ProfScope prof__(kProfHeadsGrad, st);
    const cudaError_t e = head_dgrad_wide_launch(mg, packed, B, H, W, d_oa_fd1, d_fe1, sms, st);
    if (e != cudaSuccess) return cuda_fail(e, "head_dgrad_wide_kernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// gradient copies [3][B][NT][H][W] with the dims ordered (x, channel, copy, row, image), box {32, box_c, 3, 1, 1}, 128-byte swizzle
static int make_wgrad_g5_map(CUtensorMap *map, const float *base, int B, int NT, int H, int W, int box_c)
{
    const MapKey key{base, NT, B, H, W, 32, 3, box_c, 16};
    if (map_cache_get(key, map)) return 0;
    const cuuint64_t P4 = (cuuint64_t)H * W * 4;
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)NT, 3, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[4] = {P4, (cuuint64_t)B * NT * P4, (cuuint64_t)W * 4, (cuuint64_t)NT * P4};
    const cuuint32_t box[5] = {32, (cuuint32_t)box_c, 3, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NLSPN_ERR_SHAPE, "cuTensorMapEncodeTiled(head wgrad, 5d) failed (CUresult %d)", (int)r);
    map_cache_put(key, *map);
    return 0;
}

// one rolling launch: input tensors x0 (and x1), gradient channels n0 .. n0 + np - 1 (np = 32 or 8)
static int launch_wgrad_roll(int np, const float *x0, const float *x1, const CUtensorMap &mg, int B, int H, int W, int n0, HeadWgradX xs,
                             int ldw, float *dw_all, int sms, cudaStream_t st)
{
    CUtensorMap m0, m1;
    if (int rc = make_wgrad_map(&m0, x0, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
    cudaError_t e;
    if (x1) {
        if (int rc = make_wgrad_map(&m1, x1, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
        e = np == 32 ? head_wgrad_roll_launch<32, 2>(m0, m1, mg, B, H, W, n0, xs, ldw, dw_all, sms, st)
                     : head_wgrad_roll_launch<8, 2>(m0, m1, mg, B, H, W, n0, xs, ldw, dw_all, sms, st);
    } else {
        e = np == 32 ? head_wgrad_roll_launch<32, 1>(m0, m0, mg, B, H, W, n0, xs, ldw, dw_all, sms, st)
                     : head_wgrad_roll_launch<8, 1>(m0, m0, mg, B, H, W, n0, xs, ldw, dw_all, sms, st);
    }
    if (e != cudaSuccess) return cuda_fail(e, "head_wgrad_roll_kernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int nlspn_heads_wgrad(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1, const float *g_shift,
                      int B, int H, int W, int K, float *dw_all, void *stream)
{
    if (int rc = check_heads_grad("heads_wgrad", B, H, W, K)) return rc;
    if (!fe1 || !g_shift || !dw_all) return fail(NLSPN_ERR_NULL, "heads_wgrad: a required pointer is NULL");
    if (!aligned16(id_fd1) || !aligned16(oa_fd1) || !aligned16(cf_fd1) || !aligned16(fe1) || !aligned16(g_shift))
        return fail(NLSPN_ERR_ALIGN, "heads_wgrad: tensors must be 16-byte aligned");
    if (int rc = need_tma_driver("heads_wgrad")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int N3 = 3 * (K * K - 1), NT = N3 + 2, ldw = 2 * kHeadCin * 9;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || nlspn_device_info(dev, &sms, nullptr) != 0 || sms <= 0) sms = 148;
    cudaError_t e = cudaMemsetAsync(dw_all, 0, sizeof(float) * (size_t)NT * ldw, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(dw_all)");
    ProfScope prof__(kProfHeadsGrad, st);
    if (opt(kOptHeadsWgradRoll) != 0) {
        CUtensorMap mg32, mg8;
        if (int rc = make_wgrad_g5_map(&mg32, g_shift, B, NT, H, W, 32)) return rc;
        if (int rc = make_wgrad_g5_map(&mg8, g_shift, B, NT, H, W, 8)) return rc;
        // the shared fe1 (dW input channels 64..127, every gradient channel) and the guidance branch (input channels 0..63,
        // gradient channels 2..NT-1) side by side as one N = 128 operand, 32 gradient channels per launch
        for (int n0 = 0; n0 < NT; n0 += 32) {
            const HeadWgradX xs{{0, 2}, {NT, NT}, {kHeadCin, 0}};
            if (int rc = launch_wgrad_roll(32, fe1, oa_fd1, mg32, B, H, W, n0, xs, ldw, dw_all, sms, st)) return rc;
        }
        // the two one-channel heads: init (gradient channel 0) and confidence (channel 1), input channels 0..63 each
        if (id_fd1 && cf_fd1) {
            const HeadWgradX xs{{0, 1}, {1, 2}, {0, 0}};
            if (int rc = launch_wgrad_roll(8, id_fd1, cf_fd1, mg8, B, H, W, 0, xs, ldw, dw_all, sms, st)) return rc;
        } else if (id_fd1 || cf_fd1) {
            const int n = id_fd1 ? 0 : 1;
            const HeadWgradX xs{{n, 0}, {n + 1, 0}, {0, 0}};
            if (int rc = launch_wgrad_roll(8, id_fd1 ? id_fd1 : cf_fd1, nullptr, mg8, B, H, W, 0, xs, ldw, dw_all, sms, st)) return rc;
        }
        return 0;
    }
    // per-chunk form (option heads_wgrad_roll = 0): one launch per input tensor and block of gradient channels
    CUtensorMap mg32, mg8, mx;
    if (int rc = make_wgrad_map(&mg32, g_shift, 3 * B, NT, H, W, 32, 3)) return rc;
    if (int rc = make_wgrad_map(&mg8, g_shift, 3 * B, NT, H, W, 8, 3)) return rc;
    if (int rc = make_wgrad_map(&mx, fe1, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
    for (int n0 = 0; n0 < NT; n0 += 32) {
        e = head_wgrad_launch<32>(mx, mg32, B, H, W, n0, NT - n0 < 32 ? NT - n0 : 32, ldw, dw_all + kHeadCin * 9, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "head_wgrad_kernel<32>(fe1)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (oa_fd1) {
        if (int rc = make_wgrad_map(&mx, oa_fd1, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
        for (int n0 = 2; n0 < NT; n0 += 32) {
            e = head_wgrad_launch<32>(mx, mg32, B, H, W, n0, NT - n0 < 32 ? NT - n0 : 32, ldw, dw_all, sms, st);
            if (e != cudaSuccess) return cuda_fail(e, "head_wgrad_kernel<32>(off_aff)");
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    if (id_fd1) {
        if (int rc = make_wgrad_map(&mx, id_fd1, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
        e = head_wgrad_launch<8>(mx, mg8, B, H, W, 0, 1, ldw, dw_all, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "head_wgrad_kernel<8>(init)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (cf_fd1) {
        if (int rc = make_wgrad_map(&mx, cf_fd1, B, kHeadCin, H, W, kHeadCin, 1)) return rc;
        e = head_wgrad_launch<8>(mx, mg8, B, H, W, 1, 1, ldw, dw_all, sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "head_wgrad_kernel<8>(confidence)");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}

static int check_dcn_domain(int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                            int pad_w, int dilation_h, int dilation_w, int group,
                            int deformable_group, int C)
{
    if (C != 1 || group != 1 || deformable_group != 1)
        return fail(NLSPN_ERR_DOMAIN, "DCN: only C=1, group=1, deformable_group=1 (got %d, %d, %d)",
                    C, group, deformable_group);
    if (kernel_h != kernel_w) return fail(NLSPN_ERR_DOMAIN, "DCN: square kernels only");
    if (stride_h != 1 || stride_w != 1 || dilation_h != 1 || dilation_w != 1)
        return fail(NLSPN_ERR_DOMAIN, "DCN: stride 1 and dilation 1 only");
    if (pad_h != (kernel_h - 1) / 2 || pad_w != (kernel_w - 1) / 2)
        return fail(NLSPN_ERR_DOMAIN, "DCN: padding must be (K-1)/2");
    return 0;
}

int nlspn_dcn_forward(const float *input, const float *weight, const float *bias,
                      const float *offset, const float *mask,
                      int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                      int dilation_h, int dilation_w, int group, int deformable_group,
                      int im2col_step, int B, int C, int H, int W, float *output, void *stream)
{
    (void)im2col_step;
    if (int rc = check_dcn_domain(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                  dilation_w, group, deformable_group, C))
        return rc;
    if (int rc = check_shape(B, H, W, kernel_h, 1)) return rc;
    if (!input || !weight || !bias || !offset || !mask || !output)
        return fail(NLSPN_ERR_NULL, "dcn_forward: a required pointer is NULL");
    const int P = H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (tiled_ok(input, W)) {   // gather source delivered as TMA boxes (kernels_step.cuh)
        CUtensorMap in_map;
        if (int rc = make_plane_map(&in_map, input, B, H, W, kTileW + 2 * halo_for(kernel_h), kFwdTH + 2 * halo_for(kernel_h))) return rc;
        dim3 grid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + kFwdTH - 1) / kFwdTH), (unsigned)B);
        ProfScope prof__(kProfDcnFwd, st);
        DISPATCH_K(kernel_h, (dcn_fwd_tiled_kernel<KC, kFwdTH><<<grid, dim3(kTileW, kFwdTH), 0, st>>>(
                                 in_map, input, offset, mask, weight, bias, H, W, output)));
        NLSPN_CHECK_LAUNCH("dcn_fwd_tiled_kernel");
        return 0;
    }
    { ProfScope prof__(kProfDcnFwd, st);
    DISPATCH_K(kernel_h, (iter_fwd_kernel<KC, false><<<grid_for(P, B), kBlock, 0, st>>>(
                             input, offset, mask, nullptr, nullptr, weight, bias, 0u, H, W, output,
                             nullptr))); }
    NLSPN_CHECK_LAUNCH("iter_fwd_kernel<dcn>");
    return 0;
}

int nlspn_dcn_backward(const float *input, const float *weight, const float *bias,
                       const float *offset, const float *mask, const float *grad_output,
                       int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                       int dilation_h, int dilation_w, int group, int deformable_group,
                       int im2col_step, int B, int C, int H, int W,
                       float *grad_input, float *grad_offset, float *grad_mask,
                       float *grad_weight, float *grad_bias, void *stream)
{
    (void)im2col_step;
    (void)bias;
    if (int rc = check_dcn_domain(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                  dilation_w, group, deformable_group, C))
        return rc;
    if (int rc = check_shape(B, H, W, kernel_h, 1)) return rc;
    if (!input || !weight || !offset || !mask || !grad_output || !grad_input || !grad_offset ||
        !grad_mask || !grad_weight || !grad_bias)
        return fail(NLSPN_ERR_NULL, "dcn_backward: a required pointer is NULL");
    const int P = H * W;
    const long BP = (long)B * P;
    const int KK = kernel_h * kernel_h;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * BP, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_input)");
    e = cudaMemsetAsync(grad_weight, 0, sizeof(float) * KK, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_weight)");
    e = cudaMemsetAsync(grad_bias, 0, sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_bias)");
    { ProfScope prof__(kProfDcnBwd, st);
    DISPATCH_K(kernel_h, (iter_bwd_kernel<KC, false><<<grid_for(P, B), kBlock, 0, st>>>(
                             input, offset, mask, nullptr, nullptr, nullptr, grad_output, weight,
                             nullptr, grad_input, grad_offset, (long)2 * KK * P, grad_mask, nullptr,
                             0u, 1, H, W))); }
    NLSPN_CHECK_LAUNCH("iter_bwd_kernel<dcn>");
    DISPATCH_K(kernel_h, (dcn_wb_grad_kernel<KC><<<grid_for(P, B), kBlock, 0, st>>>(
                             input, offset, mask, grad_output, H, W, grad_weight, grad_bias)));
    NLSPN_CHECK_LAUNCH("dcn_wb_grad_kernel");
    return 0;
}

// ---- single fused iteration (kernels_step.cuh) --------------------------------------------------------
static int check_step(const float *src_prev, const float *offset, const float *aff, const float *feat_fix,
                      unsigned flags, int B, int H, int W, int K)
{
    if (int rc = check_shape(B, H, W, K, 1)) return rc;
    if (flags & ~(NLSPN_FLAG_PRESERVE_INPUT | NLSPN_FLAG_ALWAYS_CLIP | NLSPN_FLAG_NO_OFFSET))
        return fail(NLSPN_ERR_DOMAIN, "step: only PRESERVE_INPUT, ALWAYS_CLIP and NO_OFFSET apply to a single step");
    const bool no_off = (flags & NLSPN_FLAG_NO_OFFSET) != 0;
    if (no_off && K != 3) return fail(NLSPN_ERR_KERNEL, "fixed-local propagation (NO_OFFSET) is 3x3 only (got K=%d)", K);
    if (!src_prev || !aff || (!offset && !no_off)) return fail(NLSPN_ERR_NULL, "step: src_prev, aff (and offset) are required");
    if ((flags & NLSPN_FLAG_PRESERVE_INPUT) && !feat_fix) return fail(NLSPN_ERR_NULL, "step: PRESERVE_INPUT needs feat_fix");
    return 0;
}

int nlspn_step_fwd(const float *src_prev, const float *offset, const float *aff, const float *conf_fixed,
                   const float *feat_fix, unsigned flags, int B, int H, int W, int K,
                   float *out, float *src_next, void *stream)
{
    if (int rc = check_step(src_prev, offset, aff, feat_fix, flags, B, H, W, K)) return rc;
    if (!out) return fail(NLSPN_ERR_NULL, "step_fwd: out is required");
    if (conf_fixed && !src_next) return fail(NLSPN_ERR_NULL, "step_fwd: src_next is required with a confidence");
    const int P = H * W, KK = K * K;
    (void)KK;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof__(kProfIterFwd, st);
    if (flags & NLSPN_FLAG_NO_OFFSET) {
        cudaError_t e = launch_pdl(false, fixed_fwd_kernel<false>, grid_for(P, B), dim3(kBlock), st, src_prev, aff,
                                   conf_fixed, feat_fix, flags, H, W, out, src_next);
        if (e != cudaSuccess) return cuda_fail(e, "fixed_fwd_kernel");
        NLSPN_CHECK_LAUNCH("fixed_fwd_kernel");
        return 0;
    }
    if (tiled_ok(src_prev, W)) {
        CUtensorMap map;
        const int th = fwd_tile_h();
        if (int rc = make_plane_map(&map, src_prev, B, H, W, kTileW + 2 * halo_for(K), th + 2 * halo_for(K))) return rc;
        dim3 grid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + th - 1) / th), (unsigned)B);
        dim3 block(kTileW, th);
        DISPATCH_TH(th, DISPATCH_K(K, (launch_pdl(false, iter_fwd_tiled_kernel<KC, THC, false>, grid, block, st, map, 0,
                                                 src_prev, offset, aff, conf_fixed, feat_fix, flags, H, W, out,
                                                 src_next))));
        NLSPN_CHECK_LAUNCH("iter_fwd_tiled_kernel");
        return 0;
    }
    DISPATCH_K(K, (iter_fwd_kernel<KC, true><<<grid_for(P, B), kBlock, 0, st>>>(
                      src_prev, offset, aff, conf_fixed, feat_fix, nullptr, nullptr, flags, H, W, out, src_next)));
    NLSPN_CHECK_LAUNCH("iter_fwd_kernel");
    return 0;
}

size_t nlspn_step_bwd_workspace_bytes(int B, int H, int W, int K, unsigned flags)
{
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    if (flags & NLSPN_FLAG_NO_OFFSET) return 16;
    return sizeof(float) * (size_t)B * (size_t)scatter_geo(H, W).image + 16;   // four phase-shifted blocked planes
}

} // extern "C" (templates need C++ linkage)

// shared by nlspn_step_bwd (DCN = false) and nlspn_dcn_backward_ws (DCN = true)
template <bool DCN>
static int run_step_bwd(const float *src_prev, const float *offset, const float *aff, const float *conf_fixed,
                        const float *feat_fix, const float *out, const float *g_out, const float *g_src_next,
                        const float *weight, unsigned flags, int B, int H, int W, int K, float *g_src_prev,
                        float *g_offset, float *g_aff, float *g_conf, void *workspace, size_t workspace_bytes,
                        cudaStream_t st)
{
    const int P = H * W;
    const long BP = (long)B * P;
    const bool blocked = opt(kOptDcnBlocked) != 0 && tiled_ok(src_prev, W) && workspace && aligned16(workspace) &&
                         workspace_bytes >= nlspn_step_bwd_workspace_bytes(B, H, W, K, 0);
    ProfScope prof__(DCN ? kProfDcnBwd : kProfIterBwdV1, st);
    if (blocked) {
        const ScatterGeo sg = scatter_geo(H, W);
        float *planes = static_cast<float *>(workspace);
        cudaError_t e = cudaMemsetAsync(planes, 0, sizeof(float) * (size_t)B * sg.image, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(scatter planes)");
        CUtensorMap map;
        if (int rc = make_plane_map(&map, src_prev, B, H, W, kTileW + 2 * halo_for(K), kFwdTH + 2 * halo_for(K))) return rc;
        dim3 grid((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + kFwdTH - 1) / kFwdTH), (unsigned)B);
        DISPATCH_K(K, (step_bwd_tiled_kernel<KC, kFwdTH, DCN><<<grid, dim3(kTileW, kFwdTH), 0, st>>>(
                          map, src_prev, offset, aff, conf_fixed, feat_fix, out, g_out, g_src_next, weight, flags, H, W,
                          planes, g_offset, g_aff, g_conf)));
        NLSPN_CHECK_LAUNCH("step_bwd_tiled_kernel");
        scatter_collect_kernel<<<grid_for(P, B), kBlock, 0, st>>>(planes, H, W, g_src_prev);
        NLSPN_CHECK_LAUNCH("scatter_collect_kernel");
        return 0;
    }
    cudaError_t e = cudaMemsetAsync(g_src_prev, 0, sizeof(float) * BP, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_src_prev)");
    DISPATCH_K(K, (step_bwd_direct_kernel<KC, DCN><<<grid_for(P, B), kBlock, 0, st>>>(
                      src_prev, offset, aff, conf_fixed, feat_fix, out, g_out, g_src_next, weight, flags, H, W,
                      g_src_prev, g_offset, g_aff, g_conf)));
    NLSPN_CHECK_LAUNCH("step_bwd_direct_kernel");
    return 0;
}

extern "C" {

int nlspn_step_bwd(const float *src_prev, const float *offset, const float *aff, const float *conf_fixed,
                   const float *feat_fix, const float *out, const float *g_out, const float *g_src_next,
                   unsigned flags, int B, int H, int W, int K,
                   float *g_src_prev, float *g_offset, float *g_aff, float *g_conf,
                   void *workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = check_step(src_prev, offset, aff, feat_fix, flags, B, H, W, K)) return rc;
    const bool no_off = (flags & NLSPN_FLAG_NO_OFFSET) != 0;
    if (!g_src_prev || !g_aff || (!g_offset && !no_off)) return fail(NLSPN_ERR_NULL, "step_bwd: gradient outputs are required");
    if (!out && (g_src_next || (flags & NLSPN_FLAG_ALWAYS_CLIP)))
        return fail(NLSPN_ERR_NULL, "step_bwd: `out` (the forward's result) is required with g_src_next or ALWAYS_CLIP");
    if (conf_fixed && !g_conf) return fail(NLSPN_ERR_NULL, "step_bwd: g_conf is required with a confidence");
    cudaStream_t st = (cudaStream_t)stream;
    if (no_off) {
        const int P = H * W;
        cudaError_t e = cudaMemsetAsync(g_src_prev, 0, sizeof(float) * (size_t)B * P, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(g_src_prev)");
        ProfScope prof__(kProfIterBwdV1, st);
        fixed_step_bwd_kernel<<<grid_for(P, B), kBlock, 0, st>>>(src_prev, aff, conf_fixed, feat_fix, out, g_out,
                                                                 g_src_next, flags, H, W, g_src_prev, g_aff, g_conf);
        NLSPN_CHECK_LAUNCH("fixed_step_bwd_kernel");
        return 0;
    }
    return run_step_bwd<false>(src_prev, offset, aff, conf_fixed, feat_fix, out, g_out, g_src_next, nullptr, flags, B, H,
                               W, K, g_src_prev, g_offset, g_aff, g_conf, workspace, workspace_bytes, st);
}

size_t nlspn_dcn_backward_workspace_bytes(int B, int H, int W, int K)
{
    return nlspn_step_bwd_workspace_bytes(B, H, W, K, 0);
}

int nlspn_dcn_backward_ws(const float *input, const float *weight, const float *bias,
                          const float *offset, const float *mask, const float *grad_output,
                          int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                          int dilation_h, int dilation_w, int group, int deformable_group,
                          int im2col_step, int B, int C, int H, int W,
                          float *grad_input, float *grad_offset, float *grad_mask,
                          float *grad_weight, float *grad_bias, void *workspace, size_t workspace_bytes,
                          void *stream)
{
    (void)im2col_step;
    (void)bias;
    if (int rc = check_dcn_domain(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                  dilation_w, group, deformable_group, C))
        return rc;
    if (int rc = check_shape(B, H, W, kernel_h, 1)) return rc;
    if (!input || !weight || !offset || !mask || !grad_output || !grad_input || !grad_offset ||
        !grad_mask || !grad_weight || !grad_bias)
        return fail(NLSPN_ERR_NULL, "dcn_backward: a required pointer is NULL");
    const int P = H * W;
    const int KK = kernel_h * kernel_h;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_weight, 0, sizeof(float) * KK, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_weight)");
    e = cudaMemsetAsync(grad_bias, 0, sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_bias)");
    if (int rc = run_step_bwd<true>(input, offset, mask, nullptr, nullptr, nullptr, grad_output, nullptr, weight, 0u, B, H,
                                    W, kernel_h, grad_input, grad_offset, grad_mask, nullptr, workspace,
                                    workspace_bytes, st))
        return rc;
    DISPATCH_K(kernel_h, (dcn_wb_grad_kernel<KC><<<grid_for(P, B), kBlock, 0, st>>>(
                             input, offset, mask, grad_output, H, W, grad_weight, grad_bias)));
    NLSPN_CHECK_LAUNCH("dcn_wb_grad_kernel");
    return 0;
}

int nlspn_dcn_forward_f64(const double *input, const double *weight, const double *bias,
                          const double *offset, const double *mask,
                          int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                          int dilation_h, int dilation_w, int group, int deformable_group,
                          int im2col_step, int B, int C, int H, int W, double *output, void *stream)
{
    (void)im2col_step;
    if (int rc = check_dcn_domain(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                  dilation_w, group, deformable_group, C))
        return rc;
    if (int rc = check_shape(B, H, W, kernel_h, 1)) return rc;
    if (!input || !weight || !bias || !offset || !mask || !output)
        return fail(NLSPN_ERR_NULL, "dcn_forward_f64: a required pointer is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    { ProfScope prof__(kProfDcnFwd, st);
    DISPATCH_K(kernel_h, (dcn_fwd_f64_kernel<KC><<<grid_for(H * W, B), kBlock, 0, st>>>(
                             input, offset, mask, weight, bias, H, W, output))); }
    NLSPN_CHECK_LAUNCH("dcn_fwd_f64_kernel");
    return 0;
}

int nlspn_dcn_backward_f64(const double *input, const double *weight, const double *bias,
                           const double *offset, const double *mask, const double *grad_output,
                           int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                           int dilation_h, int dilation_w, int group, int deformable_group,
                           int im2col_step, int B, int C, int H, int W,
                           double *grad_input, double *grad_offset, double *grad_mask,
                           double *grad_weight, double *grad_bias, void *stream)
{
    (void)im2col_step;
    (void)bias;
    if (int rc = check_dcn_domain(kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                  dilation_w, group, deformable_group, C))
        return rc;
    if (int rc = check_shape(B, H, W, kernel_h, 1)) return rc;
    if (!input || !weight || !offset || !mask || !grad_output || !grad_input || !grad_offset ||
        !grad_mask || !grad_weight || !grad_bias)
        return fail(NLSPN_ERR_NULL, "dcn_backward_f64: a required pointer is NULL");
    const long BP = (long)B * H * W;
    const int KK = kernel_h * kernel_h;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(double) * BP, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_input)");
    e = cudaMemsetAsync(grad_weight, 0, sizeof(double) * KK, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_weight)");
    e = cudaMemsetAsync(grad_bias, 0, sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_bias)");
    { ProfScope prof__(kProfDcnBwd, st);
    DISPATCH_K(kernel_h, (dcn_bwd_f64_kernel<KC><<<grid_for(H * W, B), kBlock, 0, st>>>(
                             input, offset, mask, weight, grad_output, H, W, grad_input, grad_offset,
                             grad_mask, grad_weight, grad_bias))); }
    NLSPN_CHECK_LAUNCH("dcn_bwd_f64_kernel");
    return 0;
}

int nlspn_debug_indices(const float *offset, int B, int H, int W, int K, int32_t *idx, void *stream)
{
    if (int rc = check_shape(B, H, W, K, 1)) return rc;
    if (!offset || !idx) return fail(NLSPN_ERR_NULL, "debug_indices: NULL pointer");
    DISPATCH_K(K, (debug_indices_kernel<KC><<<grid_for(H * W, B), kBlock, 0, (cudaStream_t)stream>>>(
                      offset, H, W, idx)));
    NLSPN_CHECK_LAUNCH("debug_indices_kernel");
    return 0;
}

} // extern "C"
