// hem_b200.cu — host orchestration + C ABI (include/hem_b200.h) of the B200-native
// HEM Gaussian-mixture downsampler.  Kernels live in hem_kernels.cuh, bit-level
// arithmetic in hem_math.cuh.  Build: see csrc/Makefile (nvcc, sm_100a only).
//
// There is no CPU fallback anywhere in this file: every entry point either runs
// the CUDA path or returns an error.
#include "../../include/hem_b200.h"
#include "hem_kernels.cuh"
#include "hem_icp.cuh"

#include <cub/cub.cuh>
#include <nvtx3/nvToolsExt.h>          // header-only NVTX v3: ranges per stage for nsys / ncu --nvtx, no link dependency
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

using namespace hem;

namespace {

// NVTX range over a stage of a step (SURVEY.md §5 tracing hook); free when no profiler is attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
struct NvtxMark {      // consecutive stages inside one function: each mark closes the previous range
    bool open = false;
    void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
    ~NvtxMark() { if (open) nvtxRangePop(); }
};

struct NcclId { char internal[128]; };

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T> T* as() const { return (T*)p; }
};

struct Level {
    int64_t n = 0;
    DevBuf geo;              // float4[4*cap]
    DevBuf sh;               // float[cap*F], owned storage
    const float* shp = nullptr;   // where the SH rows are: sh.p, or the caller's device array (level 0, in place)
    DevBuf ex_xyz, ex_col, ex_op, ex_cov;   // de-interleaved copy for a bound host output (DMA source)
};

struct StepStats {
    int64_t n_in = 0, n_parents = 0, n_pairs = 0, n_orphans = 0, n_dropped = 0, n_out = 0, grid_cells = 0, n_certified = 0;
    int64_t n_parents_local = 0, n_pairs_local = 0;   // this rank's share of the gate / moment work (multi-GPU)
    int shards = 1;                                   // ranks the step was sharded over
    bool tight = false;                               // ellipsoid boxes were used
    int timed = 0;                                    // events of this step were recorded (ctx->timing at the time)
    bool timed_pairs = false, ms_read = true;         // ... and are still to be turned into ms[] (on the first hem_step_stats)
    double cell_size = 0, max_radius = 0;
    double ms_host = 0;       // host time spent enqueueing the step (up to its closing sync)
    double ms[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };   // total, prep, grid, gate, sumlw, moments, compact, k_gate, k_moments
};

struct StepPairs {
    std::vector<uint32_t> parent_idx, child;
    std::vector<int64_t> off;
    std::vector<float> wl, sumlw, radius;
};

enum { EV_BEGIN = 0, EV_PREP, EV_GRID, EV_GATE, EV_SUMLW, EV_MOM, EV_END, EV_GK0, EV_GK1, EV_MK0, EV_MK1, EV_COUNT };

} // namespace

struct hem_ctx {
    int device = 0;
    uint64_t seed = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    int timing = 0;                             // 0 off, 1 events around every stage, 2 only the level and the two hot kernels
    bool keep_pairs = false, smd_prefilter = true;
    bool tight_box = true;         // certified parents search the box of their Mahalanobis ellipsoid when no component of the level is uncertified
    int sumlw_mode = 0;            // 0 auto, 1 radix sort, 2 claim table
    // prep of the next level carried over from k_finish (no prep kernel, no host sync)
    bool prep_valid = false;
    int prep_level = -1;
    int64_t prep_n = 0;
    Ctl prep_stats;
    bool dense_hint = false;       // last step had more than 4 accepted pairs per component (or its claim table overflowed)
    std::vector<int64_t> claims_failed_n;      // per level: size of the level whose claim summation gave up in an earlier build
                                               // (a build of a same-sized level starts on the sort path instead of redoing the step)
    int redo_depth = 0;            // guards the redo recursion of run_step
    size_t pool_min = 0;           // pair pool size learned from an overflow
    // robust grid box of the current build (percentile range of the level-0 means, see k_boxhist)
    bool box_valid = false;
    float box_lo[3] = { 0, 0, 0 }, box_hi[3] = { 0, 0, 0 };
    double box_q = 5e-4;
    int64_t shard_min_n = 2000000; // multi-GPU: levels smaller than this are built redundantly on every rank (measured: a 1.1 M level loses to its exchanges at 2 ranks)
    uint32_t heavy_pairs = 16384;  // parents with more accepted pairs are accumulated by k_moments_heavy (0 = never)
    bool heavy_hint = false;       // a previous step met such a parent: launch k_moments_heavy from now on
    // the same for the gate, by rows of cells of the search box: a parent whose box has more rows than
    // max(heavy_rows, heavy_rows_rel x the box of a parent of geometric-mean radius) is scanned by the grid in
    // segments of GATE_SEG_ROWS rows instead of by one warp (heavy_rows = 0: never)
    int heavy_rows = 256;          // (1024 / 16 left parents of a few hundred rows to one warp: heavy-tailed stress cloud 31 -> 22 ms,
    double heavy_rows_rel = 8.0;   //  ordinary pyramids unchanged)
    bool heavy_gate_hint = false;
    double cell_scale = 0.55;
    double cell_occupancy = 1.3;   // lower bound of the cell edge in units of the mean component spacing
    bool rng_base_set = false;
    uint64_t rng_base = 0;
    int F = 0, Fp = 0;
    int sm_count = 148;
    uint64_t n_launches = 0, bytes_h2d = 0, bytes_d2h = 0;

    std::vector<Level> levels;
    std::vector<StepStats> stats;
    std::vector<StepPairs> pairs;

    // workspace (grow-only, reused across steps and builds)
    DevBuf in_xyz, in_col, in_op, in_cov, in_sh, in_w, in_mask;       // staged host inputs
    DevBuf out_a, out_b, out_c, out_d, out_e, out_f, out_g;           // staged outputs
    DevBuf radius, psafe, pflag, radius2, psafe2, pflag2;             // prep arrays, ping-pong by level parity
    DevBuf prank, key, key2, val, val2, sp0, sp1, sp2, sp3, cell_start;
    DevBuf ctlbuf, pair_off, pair_off_s, ptab, pinfo, csr_s, csr_pad, stage_geo, stage_sh, pool, chunks, csr_child, csr_wl, sk, sv, orank;
    DevBuf valid, vflag, vpos, cubtmp, tmp_geo, tmp_sh, claims, ovf, boxhist, heavy_part, heavy_done, seg_info, hv_hdr, seg_cnt, seg_off, seg_done, xstat;
    unsigned long long* h_counters = nullptr;   // pinned
    Ctl* h_ctl = nullptr;                       // pinned + mapped: kernels publish the control block into it
    Ctl* h_ctl_dev = nullptr;                   // device alias of h_ctl
    unsigned* h_seq = nullptr;                  // pinned + mapped: k_publish's "block is visible" word, polled by the host
    unsigned* h_seq_dev = nullptr;
    unsigned seq_next = 1;
    bool poll_sync = true;                      // end-of-level wait: poll h_seq instead of cudaStreamSynchronize
    // timing events, one set per level: recorded while the steps run, read once at the end of the build (a readout per
    // step cost an extra event synchronisation per level)
    std::vector<std::vector<cudaEvent_t>> evs;

    // independent stages of a step run beside the main chain on a second stream (fork / join by events):
    // control-block reset + parent-rank scan beside the grid build, orphan scan + copy beside the moment kernel
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // multi-GPU: the SH part of a sharded level's merged rows (180 of 244 bytes) is exchanged on its own stream while the
    // next step builds its grid and runs its gates (they need geometry only); first reader: the next moment / orphan kernels
    cudaStream_t xchg_stream = nullptr;
    cudaEvent_t ev_xfork = nullptr, ev_shx = nullptr;
    bool shx_pending = false;
    DevBuf ptab2;                                // parent table of odd levels (the SH exchange of level l reads ptab while level l+1 writes its own)
    bool overlap = true;
    DevBuf cubtmp2;                              // CUB temp storage of the second stream
    // copy/compute overlap: SH rows arrive and finished levels leave on a second stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_geo = nullptr, ev_sh = nullptr, ev_level = nullptr;
    bool sh_pending = false;                     // level-0 SH H2D still in flight on copy_stream
    // Geometry-first pipeline (host SH rows, one GPU): while the level-0 SH rows are still crossing PCIe, a step produces
    // the geometry rows of its output level only and stashes what the SH pass needs (pairs, offsets, parent table,
    // normalisers, orphan ranks); the SH passes run level by level once the rows have landed (sh_catch_up).
    struct ShStash { DevBuf csr, off, ptab, sumlw, orank; int64_t n = 0, np = 0, P = 0; };
    std::vector<ShStash> stash;
    DevBuf sh_ctl;                               // ticket counters of the SH passes
    int geo_first = 1;                           // option "geo_first": 0 off, 1 while the upload is in flight, 2 every level of a host-SH build (tests)
    int sh_level = 0;                            // levels 0..sh_level of this build have their SH rows
    bool bind_too_small = false;                 // a bound output buffer was smaller than its level (reported by the build)
    cudaStream_t d2h_stream = nullptr;           // bound outputs leave here (the SH upload owns copy_stream)
    int64_t sh_gather_chunk = 0;                 // > 0: only this rank's slice of the SH rows was uploaded; all-gather after the wait
    float* sh_gather_buf = nullptr;
    int sh_gather_row = 0;
    bool shard_upload = true;                    // multi-GPU: upload 1/world of a host input per rank + all-gather
    struct OutBind { float *xyz = nullptr, *color = nullptr, *opacity = nullptr, *cov6 = nullptr, *sh = nullptr; int64_t cap = 0; bool set = false; };
    std::vector<OutBind> binds;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;

    // ICP consumer (hem_icp.cuh)
    DevBuf icp_src, icp_tgt, icp_box, icp_key, icp_key2, icp_val, icp_val2, icp_rows, icp_start, icp_part, icp_state, icp_corr;
    hem::icp::State* h_icp = nullptr;           // pinned + mapped
    hem::icp::State* h_icp_dev = nullptr;
    cudaEvent_t ev_icp0 = nullptr, ev_icp1 = nullptr;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,       \
                     cudaGetErrorString(e_));                                                      \
            ctx->err = buf_;                                                                       \
            return HEM_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define CKR(expr)                                                                                  \
    do {                                                                                           \
        int r_ = (expr);                                                                           \
        if (r_ != HEM_OK) return r_;                                                               \
    } while (0)

int ensure(hem_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap && b.p) return HEM_OK;
    if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = std::max<size_t>(bytes + bytes / 8, 256);
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return HEM_OK;
}

void release(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// returns a device pointer holding `bytes` from `src` (staging through `stage` if src is host memory)
int to_device(hem_ctx* ctx, const void* src, size_t bytes, DevBuf& stage, const void** out)
{
    if (!src) { *out = nullptr; return HEM_OK; }
    if (is_device_ptr(src)) { *out = src; return HEM_OK; }
    CKR(ensure(ctx, stage, bytes));
    CK(cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytes_h2d += bytes;
    *out = stage.p;
    return HEM_OK;
}

constexpr uint32_t HEAVY_SEG_CAP = 1u << 20;    // row segments of heavy gate parents per step
int g_gate_tb = 0;    // HEM_GATE_TICKET in the environment forces the gate ticket size (A/B measurements)
bool g_pdl = true;   // HEM_PDL=0 in the environment turns programmatic dependent launch off (A/B measurements)

// Programmatic dependent launch: the kernel may become resident while its predecessor in the stream is
// still draining; every kernel launched this way calls pdl_enter() (griddepcontrol.wait) before it touches
// memory, so the data dependence is unchanged and only the launch latency between dependent kernels goes.
template <class... KArgs> struct PdlLaunch {
    void (*kernel)(KArgs...);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    template <class... Args> void operator()(Args&&... args)
    {
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through cudaGetLastError()
    }
};
template <class... KArgs> PdlLaunch<KArgs...> pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st)
{
    PdlLaunch<KArgs...> L;
    L.kernel = kernel;
    memset(&L.cfg, 0, sizeof L.cfg);
    L.cfg.gridDim = grid;
    L.cfg.blockDim = block;
    L.cfg.dynamicSmemBytes = smem;
    L.cfg.stream = st;
    memset(&L.attr, 0, sizeof L.attr);
    L.attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    L.attr.val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    return L;
}

int allgather_equal(hem_ctx* ctx, void* buf, size_t elem, int64_t chunk, cudaStream_t on = nullptr);

// Multi-GPU build from HOST arrays: every rank uploads only its 1/world slice of the rows over its own
// PCIe link and the ranks all-gather the slices over NVLink (the alternative — every rank uploading the
// whole cloud — multiplies the host-side traffic by the number of GPUs).  Every rank must pass the same
// kind of pointer (host or device) for the same argument, as it must pass the same values.
int to_device_sharded(hem_ctx* ctx, const void* src, int64_t rows, size_t row_bytes, DevBuf& stage, const void** out)
{
    if (!src) { *out = nullptr; return HEM_OK; }
    if (is_device_ptr(src)) { *out = src; return HEM_OK; }
    if (ctx->world <= 1 || !ctx->shard_upload)
        return to_device(ctx, src, (size_t)rows * row_bytes, stage, out);
    const int64_t chunk = (rows + ctx->world - 1) / ctx->world;
    CKR(ensure(ctx, stage, (size_t)chunk * ctx->world * row_bytes));
    const int64_t b = std::min<int64_t>(chunk * ctx->rank, rows), e = std::min<int64_t>(b + chunk, rows);
    if (e > b) {
        CK(cudaMemcpyAsync((char*)stage.p + (size_t)b * row_bytes, (const char*)src + (size_t)b * row_bytes,
                           (size_t)(e - b) * row_bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->bytes_h2d += (size_t)(e - b) * row_bytes;
    }
    CKR(allgather_equal(ctx, stage.p, row_bytes, chunk));
    *out = stage.p;
    return HEM_OK;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads); }

int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 32 && (max_value >> b)) ++b;
    return b;
}

template <class F> int cub_call(hem_ctx* ctx, DevBuf& tmp, F&& f);
template <class F> int cub_call(hem_ctx* ctx, F&& f) { return cub_call(ctx, ctx->cubtmp, std::forward<F>(f)); }
template <class F> int cub_call(hem_ctx* ctx, DevBuf& tmp, F&& f)
{
    // CUB's two-phase protocol costs a second host-side dispatch per call, which matters on the
    // small (launch-bound) levels: with a warm temp buffer the call goes straight through and the
    // size query only runs if CUB reports the buffer as too small.
    if (tmp.p && tmp.cap >= (size_t)8 << 20) {
        size_t have = tmp.cap;
        cudaError_t e = f(tmp.p, have);
        if (e == cudaSuccess) return HEM_OK;
        cudaGetLastError();
    }
    size_t bytes = 0;
    CK(f((void*)nullptr, bytes));
    CKR(ensure(ctx, tmp, std::max<size_t>(bytes, (size_t)8 << 20)));
    size_t have = tmp.cap;
    CK(f(tmp.p, have));
    return HEM_OK;
}

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (dlopen) so that the single-GPU library has no NCCL
// dependency and a process that already loaded torch's bundled NCCL shares it.
// ---------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib)
        for (const char* nm : names) {
            g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
    if (!g_nccl.lib) { g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { return dlsym(g_nccl.lib, n); };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclBroadcast");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Broadcast || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        g_nccl.err = "libnccl lacks a required symbol";
        g_nccl.lib = nullptr;
        return false;
    }
    return true;
}

#define NK(call)                                                                                   \
    do {                                                                                           \
        int e_ = (call);                                                                           \
        if (e_ != 0) {                                                                             \
            ctx->err = std::string(#call " failed: ") +                                            \
                       (g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "nccl error");         \
            return HEM_ERR_NCCL;                                                                   \
        }                                                                                          \
    } while (0)

// slab bounds: parents (in cell order) [bounds[r], bounds[r+1]) belong to rank r
// equal chunks of ceil(np/G) parents (the last ranks may own fewer): equal counts let the exchanges be single
// in-place ncclAllGather calls instead of G broadcasts
inline int64_t slab_chunk(int64_t np, int world) { return (np + world - 1) / world; }
inline int64_t slab_bound(int64_t np, int r, int world) { return std::min<int64_t>(slab_chunk(np, world) * r, np); }

// in-place all-gather: rank r owns elements [r*chunk, (r+1)*chunk) of `buf` (G*chunk elements long)
int allgather_equal(hem_ctx* ctx, void* buf, size_t elem, int64_t chunk, cudaStream_t on)
{
    if (ctx->world <= 1 || chunk <= 0) return HEM_OK;
    const size_t bytes = (size_t)chunk * elem;
    NK(g_nccl.AllGather((char*)buf + (size_t)ctx->rank * bytes, buf, bytes, /*ncclChar*/ 0, ctx->nccl_comm, on ? on : ctx->stream));
    return HEM_OK;
}

// the level-0 SH rows become visible to the main stream (and, after a sharded upload, complete on every rank)
int wait_sh(hem_ctx* ctx)
{
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_sh, 0));
    ctx->sh_pending = false;
    if (ctx->sh_gather_chunk > 0) {
        CKR(allgather_equal(ctx, ctx->sh_gather_buf, (size_t)ctx->sh_gather_row, ctx->sh_gather_chunk));
        ctx->sh_gather_chunk = 0;
    }
    return HEM_OK;
}

// percentile box of the means of a level (at most two histogram passes), see k_boxhist
int robust_box(hem_ctx* ctx, const float4* geo, int64_t n, const float bmin[3], const float bmax[3])
{
    cudaStream_t st = ctx->stream;
    CKR(ensure(ctx, ctx->boxhist, 3 * BOX_BINS * 4));
    std::vector<unsigned int> hh(3 * BOX_BINS);
    float lo[3] = { bmin[0], bmin[1], bmin[2] }, hi[3] = { bmax[0], bmax[1], bmax[2] };
    for (int pass = 0; pass < 2; ++pass) {
        float sc[3];
        bool any = false;
        for (int a = 0; a < 3; ++a) {
            const double w = (double)hi[a] - (double)lo[a];
            sc[a] = (w > 0 && std::isfinite(w)) ? (float)(BOX_BINS / w) : 0.0f;
            any |= sc[a] > 0;
        }
        if (!any) break;
        CK(cudaMemsetAsync(ctx->boxhist.p, 0, 3 * BOX_BINS * 4, st));
        k_boxhist<<<(unsigned)std::min<int64_t>(blocks_for(n, 256), (int64_t)ctx->sm_count * 4), 256, 0, st>>>(
            geo, n, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], ctx->boxhist.as<unsigned int>());
        CK(cudaGetLastError()); ctx->n_launches++;
        CK(cudaMemcpyAsync(hh.data(), ctx->boxhist.p, 3 * BOX_BINS * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        bool shrunk = false;
        for (int a = 0; a < 3; ++a) {
            if (!(sc[a] > 0)) continue;
            uint64_t tot = 0;
            for (int b = 0; b < BOX_BINS; ++b) tot += hh[a * BOX_BINS + b];
            const uint64_t cut = (uint64_t)(ctx->box_q * (double)tot);
            int b0 = 0, b1 = BOX_BINS - 1;
            uint64_t acc = 0;
            while (b0 < BOX_BINS - 1 && acc + hh[a * BOX_BINS + b0] <= cut) acc += hh[a * BOX_BINS + b0++];
            acc = 0;
            while (b1 > b0 && acc + hh[a * BOX_BINS + b1] <= cut) acc += hh[a * BOX_BINS + b1--];
            const double bw = ((double)hi[a] - (double)lo[a]) / BOX_BINS;
            const double nlo = (double)lo[a] + (b0 - 1) * bw, nhi = (double)lo[a] + (b1 + 2) * bw;   // one bin of margin
            const double oldw = (double)hi[a] - (double)lo[a];
            const float flo = (float)std::max(nlo, (double)lo[a]), fhi = (float)std::min(nhi, (double)hi[a]);
            if ((double)fhi - (double)flo < 0.125 * oldw) shrunk = true;
            lo[a] = flo; hi[a] = fhi;
        }
        if (!shrunk) break;       // the bulk fills the histogram range: resolution is fine
    }
    for (int a = 0; a < 3; ++a) { ctx->box_lo[a] = lo[a]; ctx->box_hi[a] = hi[a]; }
    ctx->box_valid = true;
    return HEM_OK;
}

// ---------------------------------------------------------------------------
// one clustering step: levels[l] -> levels[l+1]
// ---------------------------------------------------------------------------

// A bound output level leaves for the caller's buffers: the geometry arrays are de-interleaved by a kernel on the MAIN
// stream (the persistent gate / moment kernels of the next level would starve it elsewhere), the DMA runs on the
// download stream (the level-0 SH upload owns the copy stream; PCIe carries both directions at once).
// `do_geo` / `do_sh`: which half of the level to send (the geometry-first pipeline sends them at different times).
static int send_level(hem_ctx* ctx, size_t lv, bool do_geo, bool do_sh, bool* too_small)
{
    if (lv >= ctx->binds.size() || !ctx->binds[lv].set) return HEM_OK;
    cudaStream_t st = ctx->stream;
    const hem_ctx::OutBind& B = ctx->binds[lv];
    if (B.cap < ctx->levels[lv].n) { *too_small = true; return HEM_OK; }
    Level& LV = ctx->levels[lv];
    const int64_t m = LV.n;
    if (m <= 0) return HEM_OK;
    const bool hx = B.xyz && !is_device_ptr(B.xyz), hc = B.color && !is_device_ptr(B.color);
    const bool ho = B.opacity && !is_device_ptr(B.opacity), hv = B.cov6 && !is_device_ptr(B.cov6);
    float *dx = nullptr, *dc = nullptr, *dop = nullptr, *dcv = nullptr;
    if (do_geo) {
        if (hx) CKR(ensure(ctx, LV.ex_xyz, (size_t)m * 12));
        if (hc) CKR(ensure(ctx, LV.ex_col, (size_t)m * 12));
        if (ho) CKR(ensure(ctx, LV.ex_op, (size_t)m * 4));
        if (hv) CKR(ensure(ctx, LV.ex_cov, (size_t)m * 24));
        dx = B.xyz ? (hx ? LV.ex_xyz.as<float>() : B.xyz) : nullptr;
        dc = B.color ? (hc ? LV.ex_col.as<float>() : B.color) : nullptr;
        dop = B.opacity ? (ho ? LV.ex_op.as<float>() : B.opacity) : nullptr;
        dcv = B.cov6 ? (hv ? LV.ex_cov.as<float>() : B.cov6) : nullptr;
        if (dx || dc || dop || dcv) {
            k_export<<<blocks_for(m, 256), 256, 0, st>>>(LV.geo.as<float4>(), m, dx, dc, dop, dcv, nullptr, nullptr);
            CK(cudaGetLastError()); ctx->n_launches++;
        }
    }
    cudaStream_t cs = ctx->d2h_stream;
    CK(cudaEventRecord(ctx->ev_level, st));
    CK(cudaStreamWaitEvent(cs, ctx->ev_level, 0));
    if (do_sh && ctx->shx_pending) CK(cudaStreamWaitEvent(cs, ctx->ev_shx, 0));      // SH rows still in the exchange
    if (do_sh && B.sh && ctx->F > 0) {
        const bool to_dev = is_device_ptr(B.sh);
        CK(cudaMemcpyAsync(B.sh, LV.shp, (size_t)m * ctx->F * 4, to_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, cs));
        if (!to_dev) ctx->bytes_d2h += (size_t)m * ctx->F * 4;
    }
    if (do_geo) {
        if (hv) { CK(cudaMemcpyAsync(B.cov6, dcv, (size_t)m * 24, cudaMemcpyDeviceToHost, cs)); ctx->bytes_d2h += (size_t)m * 24; }
        if (hx) { CK(cudaMemcpyAsync(B.xyz, dx, (size_t)m * 12, cudaMemcpyDeviceToHost, cs)); ctx->bytes_d2h += (size_t)m * 12; }
        if (hc) { CK(cudaMemcpyAsync(B.color, dc, (size_t)m * 12, cudaMemcpyDeviceToHost, cs)); ctx->bytes_d2h += (size_t)m * 12; }
        if (ho) { CK(cudaMemcpyAsync(B.opacity, dop, (size_t)m * 4, cudaMemcpyDeviceToHost, cs)); ctx->bytes_d2h += (size_t)m * 4; }
    }
    return HEM_OK;
}

// SH rows of level l + 1 from the stash of step l (geometry-first pipeline): the moment kernel's SH half over the
// step's pairs, then the orphans' rows.  Level l's SH rows must be in place (wait_sh / the previous pass).
static int sh_pass(hem_ctx* ctx, int l)
{
    cudaStream_t st = ctx->stream;
    if ((size_t)l >= ctx->stash.size()) return HEM_OK;          // an empty level: nothing was stashed
    const hem_ctx::ShStash& H = ctx->stash[l];
    Level& in = ctx->levels[l];
    Level& out = ctx->levels[l + 1];
    const int Fp = ctx->Fp;
    if (ctx->sh_pending) CKR(wait_sh(ctx));
    if (H.n <= 0) return HEM_OK;
    NvtxRange nvtx("hem.sh_pass");
    CKR(ensure(ctx, ctx->sh_ctl, sizeof(Ctl)));
    CK(cudaMemsetAsync(ctx->sh_ctl.p, 0, sizeof(Ctl), st));
    Ctl* c = ctx->sh_ctl.as<Ctl>();
    if (H.np > 0 && H.P > 0) {
        MomArgs M;
        M.geo = in.geo.as<float4>(); M.sh = in.shp; M.Fp = Fp; M.ptab = H.ptab.as<uint2>(); M.pair_off_s = H.off.as<uint32_t>();
        M.csr_s = H.csr.as<uint2>(); M.sumlw = H.sumlw.as<float>();
        M.t_begin = 0; M.t_end = (int)H.np;
        M.counters = c->counters;
        M.pool_cap = ~0ull;
        M.hv_flag = nullptr; M.hv_done = nullptr; M.n_heavy = &c->n_heavy; M.heavy_thr = 0u;
        M.by_rank = 1; M.ogeo = out.geo.as<float4>(); M.osh = out.sh.as<float>();
        const unsigned mgrid = (unsigned)std::min<int64_t>(blocks_for((H.np + MOM_B - 1) / MOM_B, MOM_WARPS), (int64_t)ctx->sm_count * 4);
        if (Fp <= 64) k_moments<2, false, 2><<<mgrid, MOM_WARPS * 32, 0, st>>>(M);
        else if (Fp <= 128) k_moments<4, false, 2><<<mgrid, MOM_WARPS * 32, 0, st>>>(M);
        else k_moments<8, false, 2><<<mgrid, MOM_WARPS * 32, 0, st>>>(M);
        CK(cudaGetLastError()); ctx->n_launches++;
    }
    k_orphans<<<blocks_for(H.n, 256), 256, 0, st>>>(in.geo.as<float4>(), in.shp, Fp, H.n, H.sumlw.as<float>(), H.orank.as<uint32_t>(),
                                                    (uint32_t)H.np, out.geo.as<float4>(), out.sh.as<float>(), 0, 1);
    CK(cudaGetLastError()); ctx->n_launches++;
    return HEM_OK;
}

// run the SH passes of the stashed steps sh_level .. upto - 1 and send the bound SH outputs of their levels
static int sh_catch_up(hem_ctx* ctx, int upto, bool* too_small)
{
    for (int l = ctx->sh_level; l < upto; ++l) {
        CKR(sh_pass(ctx, l));
        ctx->sh_level = l + 1;
        CKR(send_level(ctx, (size_t)l + 1, false, true, too_small));
    }
    return HEM_OK;
}

// Host wait for the control block of a step: poll the mapped word k_publish writes last (returns a few microseconds
// after the kernel; a stream synchronisation costs the GPU ~6 us of idle more).  The stream is queried now and then so
// that a failed launch cannot leave the loop spinning.
static int wait_publish(hem_ctx* ctx, cudaStream_t st, unsigned seq_val)
{
    if (!ctx->poll_sync) { CK(cudaStreamSynchronize(st)); return HEM_OK; }
    volatile unsigned* seq = ctx->h_seq;
    for (unsigned spins = 0; *seq != seq_val; ++spins) {
        if ((spins & 0x3fffu) == 0x3fffu) {
            const cudaError_t q = cudaStreamQuery(st);
            if (q == cudaErrorNotReady) continue;
            CK(q);
            if (*seq != seq_val) CK(cudaStreamSynchronize(st));
            break;
        }
    }
    return HEM_OK;
}

int run_step(hem_ctx* ctx, int l, float hem_reduction, float distance_delta, float color_delta, float decay_rate,
             uint64_t rng_base)
{
    cudaStream_t st = ctx->stream;
    Level& in = ctx->levels[l];
    Level& out = ctx->levels[l + 1];
    StepStats& S = ctx->stats[l];
    const int64_t n = in.n;
    const int Fp = ctx->Fp;
    DevBuf& PTAB = (l & 1) ? ctx->ptab2 : ctx->ptab;
    const float pp = 1.0f / hem_reduction;
    S = StepStats();
    S.n_in = n;
    const auto host_t0 = std::chrono::steady_clock::now();
    cudaEvent_t* ev = nullptr;
    if (ctx->timing) {
        if (ctx->evs.size() <= (size_t)l) ctx->evs.resize((size_t)l + 1);
        if (ctx->evs[l].empty()) {
            ctx->evs[l].assign(EV_COUNT, nullptr);
            for (int i = 0; i < EV_COUNT; ++i) CK(cudaEventCreate(&ctx->evs[l][i]));
        }
        ev = ctx->evs[l].data();
    }
    char nvtx_name[48];
    snprintf(nvtx_name, sizeof nvtx_name, "hem.step level %d (n=%lld)", l, (long long)in.n);
    NvtxRange nvtx_step(nvtx_name);
    NvtxMark nvtx;
    nvtx.next("S1 prep");
    if (ctx->timing) CK(cudaEventRecord(ev[EV_BEGIN], st));

    if (n == 0) { out.n = 0; return HEM_OK; }
    if (n >= (int64_t)1 << 31) { ctx->err = "more than 2^31-1 components per level is not supported"; return HEM_ERR_INVALID; }

    const float4* geo = in.geo.as<float4>();
    const float* sh = in.shp;

    // Geometry-first (see hem_ctx::ShStash): as long as the level-0 SH rows are still crossing PCIe, this step makes the
    // geometry rows of its output level only.  Once they have landed, the stashed steps get their SH passes and the
    // build goes on with fused steps.
    bool defer_sh = false;
    {
        bool in_flight = false;
        if (ctx->sh_pending && ctx->sh_gather_chunk == 0) {
            const cudaError_t q = cudaEventQuery(ctx->ev_sh);
            if (q == cudaErrorNotReady || ctx->geo_first == 2) in_flight = true; else CK(q);
        }
        defer_sh = ctx->geo_first && in_flight && ctx->world == 1 && Fp > 0 && !ctx->keep_pairs && !ctx->heavy_hint;
        if (!defer_sh && ctx->sh_level < l) {
            bool small = false;
            CKR(sh_catch_up(ctx, l, &small));
            if (small) ctx->bind_too_small = true;
        }
    }

    // ---- S1 prep --------------------------------------------------------------------------
    // control block + the arrays that must start from zero live in one allocation: one memset per step
    //   [ Ctl | pair_cnt_s (n+1) | pair_cnt (n+1) | sumlw (n) | cnt_c (n) | heavy-parent list ]
    const size_t off_cnt_s = 256, off_cnt = off_cnt_s + ((size_t)(n + 1 + 64) * 4 + 255) / 256 * 256;   // +64: all-gather padding
    const size_t off_sumlw = off_cnt + ((size_t)(n + 1) * 4 + 255) / 256 * 256;
    const size_t off_cntc = off_sumlw + ((size_t)(n + 64) * 4 + 255) / 256 * 256;              // +64: all-gather padding
    const size_t off_heavy = off_cntc + ((size_t)n * 4 + 255) / 256 * 256;      // one byte per moment ticket
    const size_t ctl_bytes = (off_heavy + (size_t)(n / MOM_B + 2) + 15) / 16 * 16;
    CKR(ensure(ctx, ctx->ctlbuf, ctl_bytes));
    Ctl* d_ctl = ctx->ctlbuf.as<Ctl>();
    unsigned long long* d_counters = d_ctl->counters;
    uint32_t* d_pair_cnt_s = (uint32_t*)((char*)ctx->ctlbuf.p + off_cnt_s);
    uint32_t* d_pair_cnt = (uint32_t*)((char*)ctx->ctlbuf.p + off_cnt);
    float* d_sumlw = (float*)((char*)ctx->ctlbuf.p + off_sumlw);
    uint32_t* d_cnt_c = (uint32_t*)((char*)ctx->ctlbuf.p + off_cntc);
    uint8_t* d_heavy = (uint8_t*)ctx->ctlbuf.p + off_heavy;
    const int set = l & 1;
    DevBuf& RAD = set ? ctx->radius2 : ctx->radius;
    DevBuf& SAFE = set ? ctx->psafe2 : ctx->psafe;
    DevBuf& PFLAG = set ? ctx->pflag2 : ctx->pflag;
    DevBuf& RAD_N = set ? ctx->radius : ctx->radius2;
    DevBuf& SAFE_N = set ? ctx->psafe : ctx->psafe2;
    DevBuf& PFLAG_N = set ? ctx->pflag : ctx->pflag2;
    const bool carried = ctx->prep_valid && ctx->prep_level == l && ctx->prep_n == n;
    if (!carried) {
        CKR(ensure(ctx, RAD, n * 4));
        CKR(ensure(ctx, SAFE, n));
        CKR(ensure(ctx, PFLAG, n * 4));
    }
    CKR(ensure(ctx, ctx->prank, n * 4));
    // fork 1 (levels whose prep was carried over): the control-block reset and the parent-rank scan do not depend on
    // the grid build and run beside it on the second stream; the main chain joins before the first kernel that needs them
    const bool fork1 = carried && ctx->overlap && ctx->aux_stream;
    cudaStream_t s1 = fork1 ? ctx->aux_stream : st;
    bool join1_pending = false;
    if (fork1) {
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    }
    pdl(k_zero, (unsigned)std::min<size_t>((ctl_bytes / 16 + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, s1)(ctx->ctlbuf.as<uint4>(), ctl_bytes / 16);
    CK(cudaGetLastError()); ctx->n_launches++;
    Ctl ps;
    if (carried) {
        // radii, flags, certificates and statistics of this level were produced by the previous step's k_finish
        ps = ctx->prep_stats;
    } else {
        pdl(k_prep, (unsigned)std::min<int64_t>(blocks_for(n, 256), (int64_t)ctx->sm_count * 4), 256, 0, st)(
            geo, n, distance_delta, RAD.as<float>(), PFLAG.as<uint32_t>(), SAFE.as<uint8_t>(), d_ctl);
        CK(cudaGetLastError()); ctx->n_launches++;
        const unsigned seq_a = ctx->seq_next++;
        pdl(k_publish, 1, 32, 0, st)(d_ctl, ctx->h_ctl_dev, ctx->h_seq_dev, seq_a);
        CK(cudaGetLastError());
        CKR(wait_publish(ctx, st, seq_a));                              // sync A (level 0 / after an erasing filter only)
        ps = *ctx->h_ctl;
    }
    ctx->prep_valid = false;
    CKR(cub_call(ctx, fork1 ? ctx->cubtmp2 : ctx->cubtmp, [&](void* t, size_t& b) {
        return cub::DeviceScan::ExclusiveSum(t, b, PFLAG.as<uint32_t>(), ctx->prank.as<uint32_t>(), (int)n, s1);
    }));
    if (fork1) { CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream)); join1_pending = true; }
    auto join1 = [&]() -> int {
        if (join1_pending) { CK(cudaStreamWaitEvent(st, ctx->ev_join, 0)); join1_pending = false; }
        return HEM_OK;
    };
    const int64_t np = (int64_t)ps.n_parents;
    S.n_parents = np;
    S.n_certified = (int64_t)ps.n_certified;
    float max_radius = 0.0f;
    memcpy(&max_radius, &ps.max_radius_bits, 4);
    S.max_radius = max_radius;
    if (ctx->timing == 1) CK(cudaEventRecord(ev[EV_PREP], st));

    // next level store: upper bound n + np rows (a NaN-radius parent is both a merged row and an orphan)
    const int64_t cap_rows = n + np;
    CKR(ensure(ctx, out.geo, (size_t)cap_rows * 64));
    CKR(ensure(ctx, out.sh, (size_t)cap_rows * std::max(Fp, 1) * 4));
    out.shp = out.sh.as<float>();
    CKR(ensure(ctx, ctx->valid, cap_rows));
    CKR(ensure(ctx, RAD_N, cap_rows * 4));
    CKR(ensure(ctx, SAFE_N, cap_rows));
    CKR(ensure(ctx, PFLAG_N, cap_rows * 4));

    int64_t P = 0;
    size_t pool_cap_used = 0, chunk_cap_final = 0;
    const bool have_pairs = np > 0;   // parents without a usable radius simply find no child
    bool launched_heavy = false, launched_gate_heavy = false;
    // fork 2: the orphans (mixture.cpp:251-253) depend on sumLw only; their scan and copy run beside the moment kernel
    // on the second stream and are joined in front of k_finish
    bool fork2 = false;
    auto orphan_stage = [&]() -> int {
        CKR(ensure(ctx, ctx->orank, n * 4));
        if (ctx->shx_pending) {       // the SH rows of this level are still being exchanged: their first readers follow
            CK(cudaStreamWaitEvent(st, ctx->ev_shx, 0));
            if (ctx->aux_stream) CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_shx, 0));
            ctx->shx_pending = false;
        }
        fork2 = ctx->overlap && ctx->aux_stream && !(ctx->sh_pending && ctx->sh_gather_chunk > 0);
        cudaStream_t s2 = fork2 ? ctx->aux_stream : st;
        if (fork2) {
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
            if (ctx->sh_pending && !defer_sh) CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_sh, 0));    // level-0 SH rows still arriving
        } else if (ctx->sh_pending && !defer_sh) {
            CKR(wait_sh(ctx));
        }
        CKR(cub_call(ctx, fork2 ? ctx->cubtmp2 : ctx->cubtmp, [&](void* t, size_t& b) {
            cub::TransformInputIterator<uint32_t, IsOrphan, const float*> it(d_sumlw, IsOrphan());
            return cub::DeviceScan::ExclusiveSum(t, b, it, ctx->orank.as<uint32_t>(), (int)n, s2);
        }));
        pdl(k_orphans, blocks_for(n, 256), 256, 0, s2)(geo, sh, Fp, n, d_sumlw, ctx->orank.as<uint32_t>(),
                                                       (uint32_t)np, out.geo.as<float4>(), out.sh.as<float>(), 1, defer_sh ? 0 : 1);
        CK(cudaGetLastError()); ctx->n_launches++;
        if (fork2) CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
        return HEM_OK;
    };
    if (have_pairs) {
        // ---- S2 grid --------------------------------------------------------------------------
        nvtx.next("S2 grid");
        GridP g;
        {
            float bmin[3], bmax[3];
            for (int a = 0; a < 3; ++a) {
                bmin[a] = ps.bbmin_c[a] ? u2f(~ps.bbmin_c[a]) : 0.0f;
                bmax[a] = ps.bbmax_u[a] ? u2f(ps.bbmax_u[a]) : 0.0f;
                if (!std::isfinite(bmin[a])) bmin[a] = -FLT_MAX;
                if (!std::isfinite(bmax[a])) bmax[a] = FLT_MAX;
            }
            // the grid spans the robust (percentile) box of the build's level 0; later levels lie inside it
            if (!ctx->box_valid) CKR(robust_box(ctx, geo, n, bmin, bmax));
            for (int a = 0; a < 3; ++a) {
                bmin[a] = std::max(bmin[a], ctx->box_lo[a]);
                bmax[a] = std::max(bmin[a], std::min(bmax[a], ctx->box_hi[a]));
            }
            // cell edge = cell_scale x geometric mean of the parent radii
            double h = ps.n_radius ? ctx->cell_scale * std::exp2((double)ps.sum_radius_fx / 65536.0 / (double)ps.n_radius - 128.0) : 1.0;
            if (!(h > 0)) h = 1.0;
            const double max_cells = 1 << 24;
            double ext[3];
            for (int a = 0; a < 3; ++a) ext[a] = std::max(0.0, (double)bmax[a] - (double)bmin[a]);
            // ... but never much finer than the component spacing: on the coarse levels the merged Gaussians are
            // few and wide, and a radius-derived cell would leave most cells empty (the gate then spends its time
            // on cell-range lookups instead of candidates)
            if (ctx->cell_occupancy > 0 && n > 0) {
                const double vol = ext[0] * ext[1] * ext[2];
                if (vol > 0 && std::isfinite(vol)) h = std::max(h, ctx->cell_occupancy * std::cbrt(vol / (double)n));
            }
            for (int it = 0; it < 64; ++it) {
                double cells = 1;
                bool too_long = false;
                for (int a = 0; a < 3; ++a) {
                    double d = std::floor(ext[a] / h) + 1;
                    if (d > 2047) too_long = true;
                    cells *= d;
                }
                if (cells <= max_cells && !too_long) break;
                h *= 1.26;
            }
            g.ox = bmin[0]; g.oy = bmin[1]; g.oz = bmin[2];
            g.inv_h = (float)(1.0 / h);
            g.h = 1.0f / g.inv_h;
            g.gx = (int)std::min(2047.0, std::floor(ext[0] / h)) + 1;
            g.gy = (int)std::min(2047.0, std::floor(ext[1] / h)) + 1;
            g.gz = (int)std::min(2047.0, std::floor(ext[2] / h)) + 1;
            S.cell_size = h;
        }
        const uint32_t ncell = (uint32_t)g.gx * g.gy * g.gz;
        S.grid_cells = ncell;
        CKR(ensure(ctx, ctx->key, n * 4)); CKR(ensure(ctx, ctx->key2, n * 4));
        CKR(ensure(ctx, ctx->val, n * 4)); CKR(ensure(ctx, ctx->val2, n * 4));
        CKR(ensure(ctx, ctx->sp0, n * 16)); CKR(ensure(ctx, ctx->sp1, n * 16));
        CKR(ensure(ctx, ctx->sp2, n * 16)); CKR(ensure(ctx, ctx->sp3, n * 16));
        // keys: non-parents by cell, parents by cell behind them (cell id + ncell): one sort puts the parents in their own
        // cell-ordered block at slots [n - np, n), and `cell_start` of the first ncell cells indexes non-parents only
        CKR(ensure(ctx, ctx->cell_start, ((size_t)2 * ncell + 1) * 4));
        pdl(k_cellkey, blocks_for(n, 256), 256, 0, st)(geo, n, g, ncell, ctx->key.as<uint32_t>(), ctx->val.as<uint32_t>());
        CK(cudaGetLastError()); ctx->n_launches++;
        const int kbits = bits_for((uint64_t)2 * ncell - 1);
        CKR(cub_call(ctx, [&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, ctx->key.as<uint32_t>(), ctx->key2.as<uint32_t>(),
                                                   ctx->val.as<uint32_t>(), ctx->val2.as<uint32_t>(), (int)n, 0, kbits, st);
        }));
        // the thread that places a parent into the planes also writes its gate record (radius, rank, inverse covariance,
        // cell box): parents are the slots [n - np, n).  Ellipsoid boxes only while the certified pre-test is active and
        // every component of the level can be pre-tested.
        CKR(ensure(ctx, PTAB, np * 8));
        CKR(ensure(ctx, ctx->pinfo, (size_t)np * PINFO_F4 * 16));
        const float kthr_ = distance_delta * distance_delta * 0.5f;
        const bool pretest = ctx->smd_prefilter && kthr_ <= 1e4f;
        const float s_box = (ctx->tight_box && pretest && ps.n_uncert == 0) ? (2.0f * kthr_ + 0.5f) / 0.99f * 1.001f : 0.0f;
        S.tight = s_box > 0.0f;
        CKR(join1());       // parent ranks (second stream)
        pdl(k_sorted_gather, blocks_for(n, 256), 256, 0, st)(geo, ctx->key2.as<uint32_t>(), ctx->val2.as<uint32_t>(), n,
                                                             2u * ncell, ctx->sp0.as<float4>(), ctx->sp1.as<float4>(),
                                                             ctx->sp2.as<float4>(), ctx->sp3.as<float4>(),
                                                             ctx->cell_start.as<uint32_t>(), (uint32_t)(n - np),
                                                             ctx->prank.as<uint32_t>(), RAD.as<float>(), SAFE.as<uint8_t>(), g, s_box,
                                                             PTAB.as<uint2>(), ctx->pinfo.as<float4>());
        CK(cudaGetLastError()); ctx->n_launches++;
        if (ctx->timing == 1) CK(cudaEventRecord(ev[EV_GRID], st));

        // ---- S3 gate ----------------------------------------------------------------------------
        nvtx.next("S3 gate");
        // levels below shard_min_n are not worth an exchange: every rank builds them whole (identical results)
        const int G = (ctx->world > 1 && n >= ctx->shard_min_n) ? ctx->world : 1;
        const int R = G > 1 ? ctx->rank : 0;
        std::vector<int64_t> tb(G + 1);
        for (int q = 0; q <= G; ++q) tb[q] = slab_bound(np, q, G);
        const int64_t my_np = tb[R + 1] - tb[R];
        S.n_parents_local = my_np; S.shards = G;
        CKR(ensure(ctx, ctx->pair_off, (np + 1) * 4));
        CKR(ensure(ctx, ctx->pair_off_s, (np + 1) * 4));
        size_t pool_cap = std::max<size_t>(std::max<size_t>(ctx->pool.cap / 8, ctx->pool_min), std::max<size_t>((size_t)8 * n / G, 1 << 16));
        unsigned long long nchunks = 0;
        // Single GPU, claim-table regime: nothing between the gate and the end of the step needs a size on
        // the host (launches are sized by capacities, counts are read on the device), so the step runs with
        // ONE host sync at its end; a pool overflow is detected there and the step is redone with a larger pool.
        if (ctx->sumlw_mode == 0 && (size_t)l < ctx->claims_failed_n.size() && ctx->claims_failed_n[l] == n) ctx->dense_hint = true;
        const bool guess_claims = ctx->sumlw_mode == 2 || (ctx->sumlw_mode == 0 && !ctx->dense_hint);
        const bool fast = (G == 1) && !ctx->keep_pairs && guess_claims;
        size_t chunk_cap_used = 0;
        for (int attempt = 0; attempt < 8; ++attempt) {
            const size_t chunk_cap = 2 * (size_t)my_np + pool_cap / 32 + 16;      // one record per parent present in a batch of 32: <= parents + batches + tickets
            chunk_cap_used = chunk_cap;
            CKR(ensure(ctx, ctx->pool, pool_cap * 8));
            CKR(ensure(ctx, ctx->chunks, chunk_cap * 16));
            if (attempt > 0) {      // pool overflow retry: reopen the counters and the counts
                CK(cudaMemsetAsync(d_counters, 0, C_COUNT * sizeof(unsigned long long), st));
                CK(cudaMemsetAsync(d_pair_cnt_s, 0, (size_t)(n + 1) * 4, st));
                CK(cudaMemsetAsync(d_pair_cnt, 0, (size_t)(n + 1) * 4, st));
            }
            GateArgs A;
            A.sp0 = ctx->sp0.as<float4>(); A.sp1 = ctx->sp1.as<float4>(); A.sp2 = ctx->sp2.as<float4>();
            A.sp3 = ctx->sp3.as<float4>(); A.cell_start = ctx->cell_start.as<uint32_t>();
            A.pinfo = ctx->pinfo.as<float4>(); A.nc = (uint32_t)(n - np); A.g = g; A.t_begin = (int)tb[R]; A.t_end = (int)tb[R + 1];
            A.cthr = color_delta * color_delta * 0.5f;           // mixture.cpp:123
            A.kthr = distance_delta * distance_delta * 0.5f;     // mixture.cpp:128
            A.decay2 = decay_rate * decay_rate;                  // mixture.cpp:58,61
            A.smd_rej = (ctx->smd_prefilter && A.kthr <= 1e4f) ? 2.0f * A.kthr + 0.5f : INFINITY;
            A.all_cert = ps.n_uncert == 0 ? 1 : 0;
            A.pair_cnt_s = d_pair_cnt_s;
            A.pair_cnt = (G == 1) ? d_pair_cnt : nullptr;      // used only by the canonical-CSR (sort / white-box) path
            A.pool = ctx->pool.as<uint2>(); A.pool_cap = pool_cap;
            A.chunks = ctx->chunks.as<uint4>(); A.chunk_cap = chunk_cap;
            A.counters = d_counters;
            // heavy parents of the gate (single GPU): reported always, deferred once the hint is set
            constexpr uint32_t SEG_CAP = HEAVY_SEG_CAP;
            int row_thr = 0;
            if (ctx->heavy_rows > 0) {
                const double gm = ps.n_radius ? std::exp2((double)ps.sum_radius_fx / 65536.0 / (double)ps.n_radius - 128.0) : S.cell_size;
                const double side = 2.0 * gm / S.cell_size + 1.0;
                row_thr = (int)std::min(1e9, std::max((double)ctx->heavy_rows, ctx->heavy_rows_rel * side * side));
            }
            launched_gate_heavy = row_thr > 0 && ctx->heavy_gate_hint && G == 1 && my_np > 0;
            A.row_report = (G == 1) ? row_thr : 0;
            A.row_defer = launched_gate_heavy ? row_thr : 0;
            A.n_heavy_gate = &d_ctl->n_heavy_gate;
            A.seg_cnt = nullptr; A.seg_off = nullptr; A.seg_done = nullptr;
            A.seg_info = nullptr; A.seg_cap = SEG_CAP; A.hv_hdr = nullptr;
            if (launched_gate_heavy) {
                CKR(ensure(ctx, ctx->seg_info, (size_t)SEG_CAP * 16));
                CKR(ensure(ctx, ctx->hv_hdr, 16));
                CK(cudaMemsetAsync(ctx->hv_hdr.p, 0, 16, st));
                CKR(ensure(ctx, ctx->seg_cnt, (size_t)SEG_CAP * 4));
                CKR(ensure(ctx, ctx->seg_off, (size_t)SEG_CAP * 4));
                if (ctx->seg_done.cap < (size_t)np * 4) {
                    CKR(ensure(ctx, ctx->seg_done, (size_t)np * 4));
                    CK(cudaMemsetAsync(ctx->seg_done.p, 0, ctx->seg_done.cap, st));
                }
                A.seg_cnt = ctx->seg_cnt.as<uint32_t>();
                A.seg_off = ctx->seg_off.as<uint32_t>(); A.seg_done = ctx->seg_done.as<uint32_t>();
                A.seg_info = ctx->seg_info.as<uint4>();
                A.hv_hdr = ctx->hv_hdr.as<uint32_t>();
            }
            if (ctx->timing) CK(cudaEventRecord(ev[EV_GK0], st));
            if (my_np > 0) {
                // ticket size: GATE_TB parents amortise the ticket/pool atomics on big levels; on small levels
                // (fewer parents than a few per resident warp) one parent per ticket keeps every warp busy and
                // the per-warp chain of dependent loads short
                const int64_t resident_warps = (int64_t)ctx->sm_count * HEM_GATE_MINBLOCKS * GATE_WARPS;
                A.tb = (int)std::max<int64_t>(1, std::min<int64_t>(GATE_TB, my_np / (2 * resident_warps)));
                if (g_gate_tb > 0) A.tb = std::min(g_gate_tb, GATE_TB);
                pdl(k_gate<false>, (unsigned)std::min<int64_t>(blocks_for((my_np + A.tb - 1) / A.tb, GATE_WARPS), (int64_t)ctx->sm_count * HEM_GATE_MINBLOCKS), GATE_WARPS * 32, 0, st)(A);
                if (launched_gate_heavy) {
                    // the flagged parents, cut into segments of rows, shared by the whole grid
                    k_gate<true><<<(unsigned)ctx->sm_count * HEM_GATE_MINBLOCKS, GATE_WARPS * 32, 0, st>>>(A);
                    ctx->n_launches += 1;
                }
                CK(cudaGetLastError());
                ctx->n_launches++;
            }
            if (ctx->timing) CK(cudaEventRecord(ev[EV_GK1], st));
            if (G == 1) {       // single GPU: counts are complete, scan before the one host sync of this stage
                CKR(cub_call(ctx, [&](void* tt, size_t& bb) {
                    return cub::DeviceScan::ExclusiveSum(tt, bb, d_pair_cnt_s, ctx->pair_off_s.as<uint32_t>(),
                                                         (int)(np + 1), st);
                }));
            }
            if (fast) break;
            pdl(k_publish, 1, 32, 0, st)(d_ctl, ctx->h_ctl_dev, (unsigned*)nullptr, 0u);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));                              // sync B
            memcpy(ctx->h_counters, ctx->h_ctl->counters, C_COUNT * sizeof(unsigned long long));
            nchunks = ctx->h_counters[C_POOL] >> POOL_BITS;            // packed: chunks above bit 40, pairs below
            ctx->h_counters[C_POOL] &= (1ull << POOL_BITS) - 1ull;
            if (ctx->h_counters[C_POOL] <= pool_cap && nchunks <= chunk_cap) break;
            pool_cap = (size_t)ctx->h_counters[C_POOL] + (size_t)ctx->h_counters[C_POOL] / 4 + 1024;
            if (attempt == 7) { ctx->err = "pair pool overflow persisted"; return HEM_ERR_STATE; }
        }
        pool_cap_used = pool_cap; chunk_cap_final = chunk_cap_used;
        if (!fast) S.n_pairs_local = (int64_t)ctx->h_counters[C_POOL];
        std::vector<int64_t> pb(G + 1, 0);          // pair offsets at the slab bounds
        if (fast) {
            pb[1] = (int64_t)pool_cap;              // capacity; the true count is read on the device / at the final sync
        } else if (G == 1) {
            pb[1] = (int64_t)ctx->h_counters[C_POOL];
        } else {
            // exchange the per-parent counts, scan them everywhere, fetch the offsets at the slab bounds
            CKR(allgather_equal(ctx, d_pair_cnt_s, 4, slab_chunk(np, G)));
            CKR(cub_call(ctx, [&](void* tt, size_t& bb) {
                return cub::DeviceScan::ExclusiveSum(tt, bb, d_pair_cnt_s, ctx->pair_off_s.as<uint32_t>(),
                                                     (int)(np + 1), st);
            }));
            std::vector<uint32_t> hb(G + 1, 0);
            for (int q = 0; q <= G; ++q)
                CK(cudaMemcpyAsync(&hb[q], ctx->pair_off_s.as<uint32_t>() + tb[q], 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));                              // sync B2 (multi-GPU only)
            for (int q = 0; q <= G; ++q) pb[q] = hb[q];
        }
        P = pb[G];                                  // fast flow: an upper bound (capacity) until the final sync
        if (P >= (int64_t)1 << 32) { ctx->err = "more than 2^32-1 accepted pairs per level is not supported"; return HEM_ERR_INVALID; }
        S.n_pairs = P;
        CKR(ensure(ctx, ctx->csr_s, std::max<int64_t>(P, 1) * 8));
        CKR(ensure(ctx, ctx->csr_child, std::max<int64_t>(P, 1) * 4));
        CKR(ensure(ctx, ctx->csr_wl, std::max<int64_t>(P, 1) * 4));
        // claim-table regime decided here so that the single-GPU chunk copy can file the claims in the same pass
        // (dense_hint is also what a failed claim summation leaves behind: a few children claimed by thousands of
        // parents overflow the table even when P <= 4n — the redo must then take the sort path, not the table again)
        const bool use_claims = fast || ctx->sumlw_mode == 2 || (ctx->sumlw_mode == 0 && P <= 4 * n && !ctx->dense_hint);
        const bool fuse_claims = use_claims && G == 1 && P > 0;
        if (use_claims && P > 0) {
            CKR(ensure(ctx, ctx->claims, (size_t)n * (CLAIM_K1 + CLAIM_K) * 8));      // tier 1 (n x 4), then tier 2 (n x 16)
            CKR(ensure(ctx, ctx->ovf, (size_t)P * 16));
        }
        uint2* const d_claims = ctx->claims.as<uint2>();
        uint2* const d_claims2 = d_claims ? d_claims + (size_t)n * CLAIM_K1 : nullptr;
        if (G == 1) {
            if (fast || nchunks > 0) {
                const int64_t launch_pairs = fast ? (int64_t)pool_cap : P;          // one lane per pool entry
                pdl(k_chunk_copy, (unsigned)std::min<int64_t>(blocks_for(launch_pairs, 256), (int64_t)ctx->sm_count * 8), 256, 0, st)(
                    ctx->chunks.as<uint4>(), d_counters, (unsigned long long)chunk_cap_used, (unsigned long long)pool_cap, ctx->pool.as<uint2>(),
                    ctx->pair_off_s.as<uint32_t>(), 0ll, ctx->csr_s.as<uint2>(), ctx->sp0.as<float4>(), PTAB.as<uint2>(), d_cnt_c,
                    fuse_claims ? d_claims : nullptr, d_claims2, ctx->ovf.as<uint4>(), (unsigned long long)P,
                    launched_gate_heavy ? ctx->seg_off.as<uint32_t>() : nullptr);
                CK(cudaGetLastError()); ctx->n_launches++;
            }
        } else {
            // every rank gets every slab's pairs: segments padded to the longest one, ONE in-place all-gather,
            // then unpacked into the contiguous cell-order CSR
            int64_t M = 0;
            for (int q = 0; q < G; ++q) M = std::max(M, pb[q + 1] - pb[q]);
            CKR(ensure(ctx, ctx->csr_pad, (size_t)std::max<int64_t>(M, 1) * G * 8));
            if (nchunks > 0) {
                pdl(k_chunk_copy, (unsigned)std::min<int64_t>(blocks_for(pb[R + 1] - pb[R], 256), (int64_t)ctx->sm_count * 8), 256, 0, st)(
                    ctx->chunks.as<uint4>(), d_counters, (unsigned long long)chunk_cap_used, (unsigned long long)pool_cap, ctx->pool.as<uint2>(),
                    ctx->pair_off_s.as<uint32_t>(), (long long)((int64_t)R * M - pb[R]), ctx->csr_pad.as<uint2>(),
                    ctx->sp0.as<float4>(), nullptr, nullptr, nullptr, nullptr, nullptr, 0ull, nullptr);
                CK(cudaGetLastError()); ctx->n_launches++;
            }
            CKR(allgather_equal(ctx, ctx->csr_pad.p, 8, M));
            for (int q = 0; q < G; ++q)
                if (pb[q + 1] > pb[q])
                    CK(cudaMemcpyAsync(ctx->csr_s.as<uint2>() + pb[q], ctx->csr_pad.as<uint2>() + (size_t)q * M,
                                       (size_t)(pb[q + 1] - pb[q]) * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (ctx->timing == 1) CK(cudaEventRecord(ev[EV_GATE], st));

        // ---- per-child normaliser ------------------------------------------------------------------
        nvtx.next("S3 sumLw");
        // sparse regime (the benchmark's): claim table, 2 launches; dense regime: stable radix sort of the
        // canonical CSR by child.  Both add a child's claims in ascending parent rank (bit-identical).
        const bool need_canon = !use_claims || ctx->keep_pairs;
        if (need_canon) {
            if (G > 1) {    // sharded: counts arrive in cell order, put them in canonical order here
                k_cnt_canon<<<blocks_for(np, 256), 256, 0, st>>>(PTAB.as<uint2>(), d_pair_cnt_s, (int)np, d_pair_cnt);
                CK(cudaGetLastError()); ctx->n_launches++;
            }
            CKR(cub_call(ctx, [&](void* tt, size_t& bb) {
                return cub::DeviceScan::ExclusiveSum(tt, bb, d_pair_cnt, ctx->pair_off.as<uint32_t>(), (int)(np + 1), st);
            }));
            if (P > 0) {
                k_csr_canon<<<blocks_for(np * 32, 256), 256, 0, st>>>(PTAB.as<uint2>(), ctx->pair_off_s.as<uint32_t>(),
                                                                       ctx->pair_off.as<uint32_t>(), ctx->csr_s.as<uint2>(), (int)np,
                                                                       ctx->csr_child.as<uint32_t>(), ctx->csr_wl.as<float>());
                CK(cudaGetLastError()); ctx->n_launches++;
            }
        }
        if (P > 0 && use_claims) {
            // sharded step: every rank holds all pairs, but files and sums the claims of its own range of children only
            // (1/G of the atomics and of the table traffic); the G slices of sumLw are then all-gathered (4 bytes per child)
            const int64_t cchunk = (n + G - 1) / G;
            const int64_t c_lo = G > 1 ? std::min<int64_t>(cchunk * R, n) : 0, c_hi = G > 1 ? std::min<int64_t>(c_lo + cchunk, n) : n;
            if (!fuse_claims) {
                pdl(k_claims, blocks_for(np * 32, 256), 256, 0, st)(PTAB.as<uint2>(), ctx->pair_off_s.as<uint32_t>(),
                                                                    ctx->csr_s.as<uint2>(), (int)np, d_cnt_c, d_claims, d_claims2,
                                                                    ctx->ovf.as<uint4>(), (unsigned long long)P, fast ? (unsigned long long)pool_cap : ~0ull, d_counters,
                                                                    (uint32_t)c_lo, (uint32_t)c_hi);
                CK(cudaGetLastError()); ctx->n_launches++;
            }
            if (c_hi > c_lo) {
                pdl(k_sumlw_claims, blocks_for(c_hi - c_lo, 256), 256, 0, st)(d_cnt_c, d_claims, d_claims2, ctx->ovf.as<uint4>(), d_counters,
                                                                              ctx->sumlw_mode == 2 ? ~0ull : 4096ull, c_lo, c_hi,
                                                                              fuse_claims ? ctx->sp0.as<float4>() : nullptr, d_sumlw);
                CK(cudaGetLastError()); ctx->n_launches++;
            }
            if (G > 1) {
                CKR(allgather_equal(ctx, d_sumlw, 4, cchunk));
                CKR(ensure(ctx, ctx->xstat, 64 * 8));
                k_status_put<<<1, 1, 0, st>>>(d_counters, ctx->xstat.as<unsigned long long>(), R);
                CKR(allgather_equal(ctx, ctx->xstat.p, 8, 1));
                k_status_or<<<1, 1, 0, st>>>(d_counters, ctx->xstat.as<unsigned long long>(), G);
                CK(cudaGetLastError()); ctx->n_launches += 2;
            }
        } else if (P > 0) {
            CKR(ensure(ctx, ctx->sk, P * 4));
            CKR(ensure(ctx, ctx->sv, P * 4));
            const int cbits = bits_for((uint64_t)n - 1);
            CKR(cub_call(ctx, [&](void* tt, size_t& bb) {
                return cub::DeviceRadixSort::SortPairs(tt, bb, ctx->csr_child.as<uint32_t>(), ctx->sk.as<uint32_t>(),
                                                       ctx->csr_wl.as<float>(), ctx->sv.as<float>(), (int)P, 0, cbits, st);
            }));
            k_segsum<<<blocks_for(P, 256), 256, 0, st>>>(ctx->sk.as<uint32_t>(), ctx->sv.as<float>(), P, d_sumlw);
            CK(cudaGetLastError()); ctx->n_launches++;
        }
        if (ctx->timing == 1) CK(cudaEventRecord(ev[EV_SUMLW], st));
        CKR(orphan_stage());

        // ---- S4 moments -------------------------------------------------------------------------------
        nvtx.next("S4 moments");
        MomArgs M;
        M.geo = geo; M.sh = sh; M.Fp = Fp; M.ptab = PTAB.as<uint2>(); M.pair_off_s = ctx->pair_off_s.as<uint32_t>();
        M.csr_s = ctx->csr_s.as<uint2>(); M.sumlw = d_sumlw;
        M.t_begin = (int)tb[R]; M.t_end = (int)tb[R + 1];
        M.counters = d_counters;
        M.pool_cap = fast ? (unsigned long long)pool_cap : ~0ull;
        M.hv_flag = d_heavy;
        M.hv_done = nullptr;
        M.n_heavy = &d_ctl->n_heavy;
        M.heavy_thr = (G == 1) ? ctx->heavy_pairs : 0u;
        if (G == 1) {
            M.by_rank = 1; M.ogeo = out.geo.as<float4>(); M.osh = out.sh.as<float>();
        } else {
            CKR(ensure(ctx, ctx->stage_geo, (size_t)slab_chunk(np, G) * G * 64));
            CKR(ensure(ctx, ctx->stage_sh, (size_t)slab_chunk(np, G) * G * std::max(Fp, 1) * 4));
            M.by_rank = 0; M.ogeo = ctx->stage_geo.as<float4>(); M.osh = ctx->stage_sh.as<float>();
        }
        if (ctx->sh_pending && !defer_sh) CKR(wait_sh(ctx));
        if (ctx->timing) CK(cudaEventRecord(ev[EV_MK0], st));
        if (my_np > 0) {
            const unsigned mgrid = (unsigned)std::min<int64_t>(blocks_for((my_np + MOM_B - 1) / MOM_B, MOM_WARPS), (int64_t)ctx->sm_count * 4);
            // Parents above heavy_pairs pairs are split over the whole grid (k_moments_heavy).  An ordinary cloud has
            // none, so the deferring variant and the heavy kernel run only once such a parent has been reported (the
            // step that meets the first one is redone, like a pool overflow).  Sharded runs keep the ordinary kernel:
            // a redo decided by one rank alone would take the ranks out of step (not validated on several GPUs yet).
            launched_heavy = ctx->heavy_pairs && ctx->heavy_hint && G == 1;
            if (defer_sh) {          // geometry rows only (never together with the heavy path: heavy_hint ends the deferral)
                if (Fp <= 64) pdl(k_moments<2, false, 1>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else if (Fp <= 128) pdl(k_moments<4, false, 1>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else pdl(k_moments<8, false, 1>, mgrid, MOM_WARPS * 32, 0, st)(M);
            } else if (launched_heavy) {
                if (Fp <= 64) pdl(k_moments<2, true>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else if (Fp <= 128) pdl(k_moments<4, true>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else pdl(k_moments<8, true>, mgrid, MOM_WARPS * 32, 0, st)(M);
            } else {
                if (Fp <= 64) pdl(k_moments<2, false>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else if (Fp <= 128) pdl(k_moments<4, false>, mgrid, MOM_WARPS * 32, 0, st)(M);
                else pdl(k_moments<8, false>, mgrid, MOM_WARPS * 32, 0, st)(M);
            }
            if (launched_heavy) {
                const int nsh = Fp <= 64 ? 2 : (Fp <= 128 ? 4 : 8);
                const size_t items = (size_t)std::max<int64_t>(P, (int64_t)pool_cap) / HEAVY_CHUNK + (size_t)np + 8;
                CKR(ensure(ctx, ctx->heavy_part, items * (size_t)(16 + 32 * nsh) * 4));
                if (ctx->heavy_done.cap < (size_t)np * 4) {        // per-parent item counters: zero once, the kernel re-zeroes what it used
                    CKR(ensure(ctx, ctx->heavy_done, (size_t)np * 4));
                    CK(cudaMemsetAsync(ctx->heavy_done.p, 0, ctx->heavy_done.cap, st));
                }
                M.hv_done = ctx->heavy_done.as<uint32_t>();
                const unsigned hgrid = (unsigned)ctx->sm_count * 2;
                if (nsh == 2) pdl(k_moments_heavy<2>, hgrid, MOM_WARPS * 32, 0, st)(M, ctx->heavy_part.as<float>());
                else if (nsh == 4) pdl(k_moments_heavy<4>, hgrid, MOM_WARPS * 32, 0, st)(M, ctx->heavy_part.as<float>());
                else pdl(k_moments_heavy<8>, hgrid, MOM_WARPS * 32, 0, st)(M, ctx->heavy_part.as<float>());
                CK(cudaGetLastError());
                ctx->n_launches++;
            }
            CK(cudaGetLastError());
            ctx->n_launches++;
        }
        if (ctx->timing) CK(cudaEventRecord(ev[EV_MK1], st));
        if (G > 1) {
            // all-gather the merged rows (staged in cell order, one contiguous range per rank), then put
            // them at their canonical place: every rank ends the step with the complete next level
            // geometry rows first (what the next step's prep, grid and gates read) ...
            CKR(allgather_equal(ctx, ctx->stage_geo.p, 64, slab_chunk(np, G)));
            k_permute_rows<<<blocks_for(np * 32, 256), 256, 0, st>>>(PTAB.as<uint2>(), (int)np, ctx->stage_geo.as<float4>(),
                                                                      ctx->stage_sh.as<float>(), Fp, out.geo.as<float4>(),
                                                                      out.sh.as<float>(), 1, 0);
            CK(cudaGetLastError()); ctx->n_launches++;
            // ... the SH rows on the exchange stream, beside whatever follows on the main stream
            if (Fp > 0) {
                cudaStream_t xs = (ctx->overlap && ctx->xchg_stream) ? ctx->xchg_stream : st;
                if (xs != st) {
                    CK(cudaEventRecord(ctx->ev_xfork, st));
                    CK(cudaStreamWaitEvent(xs, ctx->ev_xfork, 0));
                }
                CKR(allgather_equal(ctx, ctx->stage_sh.p, (size_t)Fp * 4, slab_chunk(np, G), xs));
                k_permute_rows<<<blocks_for(np * 32, 256), 256, 0, xs>>>(PTAB.as<uint2>(), (int)np, ctx->stage_geo.as<float4>(),
                                                                          ctx->stage_sh.as<float>(), Fp, out.geo.as<float4>(),
                                                                          out.sh.as<float>(), 0, 1);
                CK(cudaGetLastError()); ctx->n_launches++;
                if (xs != st) { CK(cudaEventRecord(ctx->ev_shx, xs)); ctx->shx_pending = true; }
            }
        }
    } else {
        CKR(join1());
        if (ctx->timing == 1) {
            CK(cudaEventRecord(ev[EV_GRID], st));
            CK(cudaEventRecord(ev[EV_GATE], st));
            CK(cudaEventRecord(ev[EV_SUMLW], st));
        }
        CKR(orphan_stage());
    }
    if (ctx->timing == 1) CK(cudaEventRecord(ev[EV_MOM], st));

    // ---- S5 filter ---------------------------------------------------------------------------------------
    nvtx.next("S5 finish");
    if (fork2) CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));        // the orphan rows are in place
    // determinant, fresh flag and validity of every pre-filter row: launched over the row capacity,
    // the actual row count is read on the device, so one host sync closes the step
    pdl(k_finish, blocks_for(cap_rows, 256), 256, 0, st)(out.geo.as<float4>(), np, ctx->orank.as<uint32_t>(), d_sumlw, n,
                                                         ctx->seed, rng_base, pp, distance_delta, ctx->valid.as<uint8_t>(),
                                                         RAD_N.as<float>(), SAFE_N.as<uint8_t>(), PFLAG_N.as<uint32_t>(),
                                                         pool_cap_used ? (unsigned long long)pool_cap_used : ~0ull, d_ctl);
    CK(cudaGetLastError()); ctx->n_launches++;
    const unsigned seq_val = ctx->seq_next++;
    pdl(k_publish, 1, 32, 0, st)(d_ctl, ctx->h_ctl_dev, ctx->h_seq_dev, seq_val);
    CK(cudaGetLastError());
    if (ctx->timing) CK(cudaEventRecord(ev[EV_END], st));
    const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    CKR(wait_publish(ctx, st, seq_val));                                // sync C: the one host wait of a level
    memcpy(ctx->h_counters, ctx->h_ctl->counters, C_COUNT * sizeof(unsigned long long));
    if (have_pairs) {
        // fast flow: the pair / chunk counts arrive only now (in the other flows this re-reads known values)
        const unsigned long long packed = ctx->h_counters[C_POOL];
        const unsigned long long pairs = packed & ((1ull << POOL_BITS) - 1ull), nch = packed >> POOL_BITS;
        const bool overflow = pairs > pool_cap_used || nch > chunk_cap_final;
        const bool claims_failed = ctx->h_counters[C_FAIL] != 0;
        const bool heavy_missed = !overflow && ctx->world == 1 && ctx->heavy_pairs && !launched_heavy && ctx->h_ctl->n_heavy > 0;
        if (heavy_missed) ctx->heavy_hint = true;
        if (launched_gate_heavy) {
            // more deferred parents / segments than the index holds: give the heavy gate up for this context
            if (ctx->h_counters[C_CHUNK] > (unsigned long long)HEAVY_SEG_CAP) {      // segments listed (published by k_gate<true>)
                ctx->heavy_rows = 0;
                ctx->heavy_gate_hint = false;
                ctx->prep_valid = carried;
                ctx->prep_level = l; ctx->prep_n = n; ctx->prep_stats = ps;
                if (++ctx->redo_depth > 4) { ctx->redo_depth = 0; ctx->err = "clustering step did not settle after 4 redos"; return HEM_ERR_STATE; }
                const int rc = run_step(ctx, l, hem_reduction, distance_delta, color_delta, decay_rate, rng_base);
                ctx->redo_depth = 0;
                return rc;
            }
        }
        const bool gate_heavy_missed = !overflow && ctx->world == 1 && ctx->heavy_rows > 0 && !launched_gate_heavy &&
                                       !ctx->heavy_gate_hint && ctx->h_ctl->n_heavy_gate > 0;
        if (gate_heavy_missed) ctx->heavy_gate_hint = true;
        if (overflow || claims_failed || heavy_missed || gate_heavy_missed) {
            // nothing was corrupted (pool writes are guarded, the claim summation bailed out): redo the step
            // with a larger pool / through the sort path; the carried prep of THIS level is still intact
            if (overflow) ctx->pool_min = (size_t)pairs + (size_t)pairs / 4 + 1024;
            if (claims_failed) {
                ctx->dense_hint = true;
                if (ctx->claims_failed_n.size() <= (size_t)l) ctx->claims_failed_n.resize((size_t)l + 1, -1);
                ctx->claims_failed_n[l] = n;
            }
            ctx->prep_valid = carried;
            ctx->prep_level = l; ctx->prep_n = n; ctx->prep_stats = ps;
            if (++ctx->redo_depth > 4) { ctx->redo_depth = 0; ctx->err = "clustering step did not settle after 4 redos (pool / claim table)"; return HEM_ERR_STATE; }
            const int rc = run_step(ctx, l, hem_reduction, distance_delta, color_delta, decay_rate, rng_base);
            ctx->redo_depth = 0;
            return rc;
        }
        if (S.shards == 1) {        // sharded steps: P is the global count from the exchange, `pairs` is this rank's share
            P = (int64_t)pairs;
            S.n_pairs = P;
        }
        S.n_pairs_local = (int64_t)pairs;
    }
    ctx->dense_hint = P > 4 * n;
    const int64_t norph = (int64_t)ctx->h_counters[C_NORPH];
    const int64_t npre = np + norph;
    const int64_t ninvalid = (int64_t)ctx->h_counters[C_INVALID];
    S.n_orphans = norph;
    S.n_dropped = ninvalid;
    out.n = npre;
    if (defer_sh) {
        // what the SH pass of this step needs, out of the buffers the next step reuses
        if (ctx->stash.size() <= (size_t)l) ctx->stash.resize((size_t)l + 1);
        hem_ctx::ShStash& H = ctx->stash[l];
        H.n = n; H.np = have_pairs ? np : 0; H.P = have_pairs ? P : 0;
        CKR(ensure(ctx, H.sumlw, (size_t)n * 4));
        CKR(ensure(ctx, H.orank, (size_t)n * 4));
        CK(cudaMemcpyAsync(H.sumlw.p, d_sumlw, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(H.orank.p, ctx->orank.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        if (H.np > 0 && H.P > 0) {
            CKR(ensure(ctx, H.csr, (size_t)H.P * 8));
            CKR(ensure(ctx, H.off, (size_t)(H.np + 1) * 4));
            CKR(ensure(ctx, H.ptab, (size_t)H.np * 8));
            CK(cudaMemcpyAsync(H.csr.p, ctx->csr_s.p, (size_t)H.P * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(H.off.p, ctx->pair_off_s.p, (size_t)(H.np + 1) * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(H.ptab.p, PTAB.p, (size_t)H.np * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (ninvalid > 0) {          // the filter below moves whole rows: bring the SH rows up to this level first
            bool small = false;
            CKR(sh_catch_up(ctx, l, &small));
            if (small) ctx->bind_too_small = true;
            CKR(sh_pass(ctx, l));    // (its bound output leaves after the filter, with the geometry)
            ctx->sh_level = l + 1;
        }
    } else {
        ctx->sh_level = l + 1;
    }
    if (ninvalid > 0) {
        // rare path: erase invalid outputs keeping order (mixture.cpp:262-282)
        CKR(ensure(ctx, ctx->vflag, npre * 4));
        CKR(ensure(ctx, ctx->vpos, npre * 4));
        CKR(ensure(ctx, ctx->tmp_geo, (size_t)npre * 64));
        CKR(ensure(ctx, ctx->tmp_sh, (size_t)npre * std::max(Fp, 1) * 4));
        k_valid_flag<<<blocks_for(npre, 256), 256, 0, st>>>(ctx->valid.as<uint8_t>(), npre, ctx->vflag.as<uint32_t>());
        CK(cudaGetLastError()); ctx->n_launches++;
        CKR(cub_call(ctx, [&](void* t, size_t& b) {
            return cub::DeviceScan::ExclusiveSum(t, b, ctx->vflag.as<uint32_t>(), ctx->vpos.as<uint32_t>(), (int)npre, st);
        }));
        k_compact_rows<<<blocks_for(npre * 32, 256), 256, 0, st>>>(out.geo.as<float4>(), out.sh.as<float>(), Fp, npre,
                                                                    ctx->valid.as<uint8_t>(), ctx->vpos.as<uint32_t>(),
                                                                    ctx->tmp_geo.as<float4>(), ctx->tmp_sh.as<float>());
        CK(cudaGetLastError()); ctx->n_launches++;
        const int64_t nout = npre - ninvalid;
        CK(cudaMemcpyAsync(out.geo.p, ctx->tmp_geo.p, (size_t)nout * 64, cudaMemcpyDeviceToDevice, st));
        if (Fp > 0) CK(cudaMemcpyAsync(out.sh.p, ctx->tmp_sh.p, (size_t)nout * Fp * 4, cudaMemcpyDeviceToDevice, st));
        out.n = nout;
    }
    S.n_out = out.n;
    S.ms_host = host_ms;
    // hand the next level its prep (invalid as soon as the filter erased rows: positions shifted)
    ctx->prep_valid = (ninvalid == 0);
    ctx->prep_level = l + 1;
    ctx->prep_n = npre;
    ctx->prep_stats = Ctl();
    {
        const Ctl& hc = *ctx->h_ctl;
        Ctl& nx = ctx->prep_stats;
        for (int a = 0; a < 3; ++a) { nx.bbmax_u[a] = hc.nx_bbmax_u[a]; nx.bbmin_c[a] = hc.nx_bbmin_c[a]; }
        nx.max_radius_bits = hc.nx_max_radius_bits; nx.n_radius = hc.nx_n_radius; nx.n_parents = hc.nx_n_parents;
        nx.n_certified = hc.nx_n_certified; nx.sum_radius_fx = hc.nx_sum_radius_fx; nx.n_uncert = hc.nx_n_uncert;
    }
    S.timed = ctx->timing; S.timed_pairs = have_pairs; S.ms_read = !ctx->timing;

    if (ctx->keep_pairs) {
        StepPairs& K = ctx->pairs[l];
        K = StepPairs();
        K.sumlw.resize(n);
        CK(cudaMemcpyAsync(K.sumlw.data(), d_sumlw, n * 4, cudaMemcpyDeviceToHost, st));
        std::vector<uint32_t> pflag(n);
        std::vector<float> rad(n);
        CK(cudaMemcpyAsync(pflag.data(), PFLAG.p, n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(rad.data(), RAD.p, n * 4, cudaMemcpyDeviceToHost, st));
        std::vector<uint32_t> off32(np + 1, 0);
        K.child.resize(P);
        K.wl.resize(P);
        if (have_pairs) {
            CK(cudaMemcpyAsync(off32.data(), ctx->pair_off.p, (np + 1) * 4, cudaMemcpyDeviceToHost, st));
            if (P > 0) {
                CK(cudaMemcpyAsync(K.child.data(), ctx->csr_child.p, P * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(K.wl.data(), ctx->csr_wl.p, P * 4, cudaMemcpyDeviceToHost, st));
            }
        }
        CK(cudaStreamSynchronize(st));
        K.off.assign(off32.begin(), off32.end());
        for (int64_t i = 0; i < n; ++i)
            if (pflag[i]) { K.parent_idx.push_back((uint32_t)i); K.radius.push_back(rad[i]); }
    }
    return HEM_OK;
}

int from_device(hem_ctx* ctx, void* dst, size_t bytes, DevBuf& stage, void** dev_out, bool* staged)
{
    *staged = false;
    if (!dst) { *dev_out = nullptr; return HEM_OK; }
    if (is_device_ptr(dst)) { *dev_out = dst; return HEM_OK; }
    CKR(ensure(ctx, stage, bytes));
    *dev_out = stage.p;
    *staged = true;
    return HEM_OK;
}

// one level out in the reference's array layout, enqueued on `st` (no synchronisation here)
int export_level(hem_ctx* ctx, int level, float* xyz, float* color, float* opacity, float* cov6, float* sh,
                 float* weight, uint8_t* is_parent, cudaStream_t st)
{
    Level& L = ctx->levels[level];
    const int64_t n = L.n;
    if (n == 0) return HEM_OK;
    const int F = ctx->F;
    void *dx, *dc, *dop, *dcv, *dw, *df;
    bool sx, sc, sop, scv, sw, sf;
    CKR(from_device(ctx, xyz, (size_t)n * 12, ctx->out_a, &dx, &sx));
    CKR(from_device(ctx, color, (size_t)n * 12, ctx->out_b, &dc, &sc));
    CKR(from_device(ctx, opacity, (size_t)n * 4, ctx->out_c, &dop, &sop));
    CKR(from_device(ctx, cov6, (size_t)n * 24, ctx->out_d, &dcv, &scv));
    CKR(from_device(ctx, weight, (size_t)n * 4, ctx->out_f, &dw, &sw));
    CKR(from_device(ctx, is_parent, (size_t)n, ctx->out_g, &df, &sf));
    if (sh && F > 0) {       // rows are stored in the caller's layout: plain DMA (D2H or D2D), issued first (largest)
        const bool to_dev = is_device_ptr(sh);
        CK(cudaMemcpyAsync(sh, L.shp, (size_t)n * F * 4, to_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (!to_dev) ctx->bytes_d2h += (size_t)n * F * 4;
    }
    if (dx || dc || dop || dcv || dw || df) {
        k_export<<<blocks_for(n, 256), 256, 0, st>>>(L.geo.as<float4>(), n, (float*)dx, (float*)dc, (float*)dop, (float*)dcv,
                                                      (float*)dw, (uint8_t*)df);
        CK(cudaGetLastError()); ctx->n_launches++;
    }
    if (sx) CK(cudaMemcpyAsync(xyz, dx, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    if (sc) CK(cudaMemcpyAsync(color, dc, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    if (sop) CK(cudaMemcpyAsync(opacity, dop, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (scv) CK(cudaMemcpyAsync(cov6, dcv, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    if (sw) CK(cudaMemcpyAsync(weight, dw, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (sf) CK(cudaMemcpyAsync(is_parent, df, (size_t)n, cudaMemcpyDeviceToHost, st));
    ctx->bytes_d2h += (sx ? (size_t)n * 12 : 0) + (sc ? (size_t)n * 12 : 0) + (sop ? (size_t)n * 4 : 0) +
                      (scv ? (size_t)n * 24 : 0) + (sw ? (size_t)n * 4 : 0) + (sf ? (size_t)n : 0);
    return HEM_OK;
}

} // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int hem_version(void) { return 1000; }

const char* hem_last_error(const hem_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

hem_ctx* hem_create(int device, uint64_t seed)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= count) { g_create_error = "device index out of range"; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_create_error = "cudaGetDeviceProperties failed"; return nullptr; }
    if (prop.major != 10) {
        g_create_error = "libhem_b200 is built for sm_100a (Blackwell B200) only; device is sm_" +
                         std::to_string(prop.major) + std::to_string(prop.minor);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return nullptr; }
    hem_ctx* ctx = new hem_ctx();
    ctx->device = device;
    ctx->seed = seed;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_counters, 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_ctl, sizeof(Ctl), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->h_ctl_dev, ctx->h_ctl, 0) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_seq, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->h_seq_dev, ctx->h_seq, 0) != cudaSuccess) {
        g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return nullptr;
    }
    ctx->stream = ctx->own_stream;
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->xchg_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_xfork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_shx, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (const char* cs = getenv("HEM_OVERLAP")) ctx->overlap = atoi(cs) != 0;
    if (const char* cs = getenv("HEM_POLL_SYNC")) ctx->poll_sync = atoi(cs) != 0;
    if (const char* cs = getenv("HEM_GEO_FIRST")) ctx->geo_first = atoi(cs);
    *ctx->h_seq = 0u;
    if (const char* cs = getenv("HEM_HEAVY_ROWS")) ctx->heavy_rows = atoi(cs);
    if (const char* cs = getenv("HEM_HEAVY_REL")) ctx->heavy_rows_rel = atof(cs);
    cudaEventCreateWithFlags(&ctx->ev_geo, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_sh, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_level, cudaEventDisableTiming);
    if (const char* cs = getenv("HEM_PDL")) g_pdl = atoi(cs) != 0;
    if (const char* cs = getenv("HEM_GATE_TICKET")) g_gate_tb = atoi(cs);
    if (const char* cs = getenv("HEM_CELL_OCC")) { double v = atof(cs); if (v >= 0) ctx->cell_occupancy = v; }
    if (const char* cs = getenv("HEM_CELL_SCALE")) { double v = atof(cs); if (v > 0) ctx->cell_scale = v; }
    return ctx;
}

void hem_destroy(hem_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (Level& L : ctx->levels) { release(L.geo); release(L.sh); release(L.ex_xyz); release(L.ex_col); release(L.ex_op); release(L.ex_cov); }
    DevBuf* all[] = { &ctx->in_xyz, &ctx->in_col, &ctx->in_op, &ctx->in_cov, &ctx->in_sh, &ctx->in_w, &ctx->in_mask,
                      &ctx->out_a, &ctx->out_b, &ctx->out_c, &ctx->out_d, &ctx->out_e, &ctx->out_f, &ctx->out_g,
                      &ctx->radius, &ctx->psafe, &ctx->pflag, &ctx->radius2, &ctx->psafe2, &ctx->pflag2, &ctx->prank, &ctx->key, &ctx->key2, &ctx->val, &ctx->val2,
                      &ctx->sp0, &ctx->sp1, &ctx->sp2, &ctx->sp3, &ctx->cell_start,
                      &ctx->ctlbuf, &ctx->pair_off, &ctx->pair_off_s, &ctx->ptab, &ctx->ptab2, &ctx->pinfo, &ctx->csr_s, &ctx->csr_pad, &ctx->stage_geo, &ctx->stage_sh, &ctx->pool, &ctx->chunks, &ctx->csr_child,
                      &ctx->csr_wl, &ctx->sk, &ctx->sv, &ctx->orank, &ctx->valid, &ctx->vflag,
                      &ctx->vpos, &ctx->cubtmp, &ctx->tmp_geo, &ctx->tmp_sh, &ctx->claims, &ctx->ovf, &ctx->boxhist, &ctx->heavy_part, &ctx->heavy_done, &ctx->seg_info, &ctx->hv_hdr, &ctx->xstat,
                      &ctx->seg_cnt, &ctx->seg_off, &ctx->seg_done,
                      &ctx->icp_src, &ctx->icp_tgt, &ctx->icp_box, &ctx->icp_key, &ctx->icp_key2, &ctx->icp_val, &ctx->icp_val2,
                      &ctx->icp_rows, &ctx->icp_start, &ctx->icp_part, &ctx->icp_state, &ctx->icp_corr };
    for (DevBuf* b : all) release(*b);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->h_seq) cudaFreeHost(ctx->h_seq);
    if (ctx->h_icp) cudaFreeHost(ctx->h_icp);
    for (cudaEvent_t e : { ctx->ev_icp0, ctx->ev_icp1 }) if (e) cudaEventDestroy(e);
    for (auto& set : ctx->evs) for (cudaEvent_t e : set) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    for (auto& h : ctx->stash) for (DevBuf* b : { &h.csr, &h.off, &h.ptab, &h.sumlw, &h.orank }) if (b->p) cudaFree(b->p);
    if (ctx->sh_ctl.p) cudaFree(ctx->sh_ctl.p);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->xchg_stream) { cudaStreamSynchronize(ctx->xchg_stream); cudaStreamDestroy(ctx->xchg_stream); }
    for (cudaEvent_t e : { ctx->ev_xfork, ctx->ev_shx }) if (e) cudaEventDestroy(e);
    release(ctx->ptab2);
    for (cudaEvent_t e : { ctx->ev_fork, ctx->ev_join }) if (e) cudaEventDestroy(e);
    release(ctx->cubtmp2);
    for (cudaEvent_t e : { ctx->ev_geo, ctx->ev_sh, ctx->ev_level }) if (e) cudaEventDestroy(e);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int hem_set_stream(hem_ctx* ctx, void* stream)
{
    if (!ctx) return HEM_ERR_INVALID;
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return HEM_OK;
}

int hem_set_option(hem_ctx* ctx, const char* name, double value)
{
    if (!ctx || !name) return HEM_ERR_INVALID;
    std::string k(name);
    if (k == "timing") ctx->timing = value == 2 ? 2 : (value != 0 ? 1 : 0);
    else if (k == "keep_pairs") ctx->keep_pairs = value != 0;
    else if (k == "smd_prefilter") ctx->smd_prefilter = value != 0;
    else if (k == "tight_box") ctx->tight_box = value != 0;
    else if (k == "sumlw_mode") ctx->sumlw_mode = (int)value;
    else if (k == "overlap") ctx->overlap = value != 0;
    else if (k == "poll_sync") ctx->poll_sync = value != 0;
    else if (k == "geo_first") ctx->geo_first = value == 2 ? 2 : (value != 0 ? 1 : 0);
    else if (k == "shard_min_n") ctx->shard_min_n = (int64_t)value;
    else if (k == "box_quantile") { if (!(value >= 0 && value < 0.5)) { ctx->err = "box_quantile must be in [0, 0.5)"; return HEM_ERR_INVALID; } ctx->box_q = value; }
    else if (k == "shard_upload") ctx->shard_upload = value != 0;
    else if (k == "heavy_rows_rel") { if (!(value >= 0)) { ctx->err = "heavy_rows_rel must be >= 0"; return HEM_ERR_INVALID; } ctx->heavy_rows_rel = value; }
    else if (k == "heavy_rows") { if (!(value >= 0) || value > 1e9) { ctx->err = "heavy_rows out of range"; return HEM_ERR_INVALID; } ctx->heavy_rows = (int)value; }
    else if (k == "heavy_pairs") { if (!(value >= 0) || value > 4e9) { ctx->err = "heavy_pairs out of range"; return HEM_ERR_INVALID; } ctx->heavy_pairs = (uint32_t)value; }
    else if (k == "cell_occupancy") { if (!(value >= 0)) { ctx->err = "cell_occupancy must be >= 0"; return HEM_ERR_INVALID; } ctx->cell_occupancy = value; }
    else if (k == "cell_scale") { if (!(value > 0)) { ctx->err = "cell_scale must be > 0"; return HEM_ERR_INVALID; } ctx->cell_scale = value; }
    else if (k == "rng_base") { ctx->rng_base_set = value >= 0; ctx->rng_base = (uint64_t)value; }
    else { ctx->err = "unknown option: " + k; return HEM_ERR_INVALID; }
    return HEM_OK;
}

int hem_build_pyramid(hem_ctx* ctx, const float* xyz, const float* color, const float* opacity, const float* cov6,
                      const float* sh, const float* weight, int64_t n, int32_t sh_floats, uint32_t cluster_level,
                      float hem_reduction, float distance_delta, float color_delta, float decay_rate,
                      const uint8_t* level0_parent_mask)
{
    if (!ctx) return HEM_ERR_INVALID;
    NvtxRange nvtx_build("hem_build_pyramid");
    ctx->err.clear();
    if (n < 0 || sh_floats < 0 || sh_floats > 256) { ctx->err = "n must be >= 0 and 0 <= sh_floats <= 256"; return HEM_ERR_INVALID; }
    if (n > 0 && (!xyz || !color || !opacity || !cov6 || (sh_floats > 0 && !sh))) { ctx->err = "null input array"; return HEM_ERR_INVALID; }
    if (!(hem_reduction > 0)) { ctx->err = "hem_reduction must be > 0"; return HEM_ERR_INVALID; }
    if (cluster_level > 64) { ctx->err = "cluster_level > 64"; return HEM_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->F = sh_floats;
    ctx->Fp = sh_floats;          // rows are stored unpadded
    const int Fp = ctx->Fp;
    if (ctx->levels.size() < (size_t)cluster_level + 1) ctx->levels.resize((size_t)cluster_level + 1);
    for (Level& L : ctx->levels) L.n = 0;
    ctx->stats.assign(cluster_level, StepStats());
    ctx->pairs.assign(ctx->keep_pairs ? cluster_level : 0, StepPairs());

    ctx->prep_valid = false;
    ctx->box_valid = false;
    if (ctx->shx_pending) { cudaStreamSynchronize(ctx->xchg_stream); ctx->shx_pending = false; }
    ctx->sh_pending = false;       // a build that failed half-way must not leave a stale wait / all-gather behind
    ctx->sh_gather_chunk = 0;
    // level 0 = the input (Mixture::initMixture, mixture.cpp:287-333)
    Level& L0 = ctx->levels[0];
    L0.n = n;
    if (n > 0) {
        const void *dx, *dc, *dop, *dcv, *dw, *dm;
        CKR(to_device_sharded(ctx, xyz, n, 12, ctx->in_xyz, &dx));
        CKR(to_device_sharded(ctx, color, n, 12, ctx->in_col, &dc));
        CKR(to_device_sharded(ctx, opacity, n, 4, ctx->in_op, &dop));
        CKR(to_device_sharded(ctx, cov6, n, 24, ctx->in_cov, &dcv));
        CKR(to_device_sharded(ctx, weight, n, 4, ctx->in_w, &dw));
        CKR(to_device_sharded(ctx, level0_parent_mask, n, 1, ctx->in_mask, &dm));
        CKR(ensure(ctx, L0.geo, (size_t)n * 64));
        k_import<<<blocks_for(n, 256), 256, 0, st>>>((const float*)dx, (const float*)dc, (const float*)dop,
                                                      (const float*)dcv, (const float*)dw, (const uint8_t*)dm,
                                                      L0.geo.as<float4>(), n, ctx->seed, 1.0f / hem_reduction);
        CK(cudaGetLastError()); ctx->n_launches++;
        // SH rows keep the caller's [N,F] layout: a device array is used in place (no copy at all),
        // a host array is DMA'd straight into the level store
        if (Fp > 0 && is_device_ptr(sh)) {
            L0.shp = sh;
        } else {
            const bool shard = ctx->world > 1 && ctx->shard_upload && Fp > 0;
            const int64_t chunk = shard ? (n + ctx->world - 1) / ctx->world : n;
            CKR(ensure(ctx, L0.sh, (size_t)(shard ? chunk * ctx->world : n) * std::max(Fp, 1) * 4));
            L0.shp = L0.sh.as<float>();
            ctx->sh_gather_chunk = 0;
            if (shard) {
                // this rank's slice on the copy stream; the all-gather over NVLink follows the wait on the main stream
                const int64_t b = std::min<int64_t>(chunk * ctx->rank, n), e = std::min<int64_t>(b + chunk, n);
                CK(cudaEventRecord(ctx->ev_geo, st));
                CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_geo, 0));
                if (e > b) {
                    CK(cudaMemcpyAsync(L0.sh.as<float>() + (size_t)b * Fp, sh + (size_t)b * Fp, (size_t)(e - b) * Fp * 4,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
                    ctx->bytes_h2d += (size_t)(e - b) * Fp * 4;
                }
                CK(cudaEventRecord(ctx->ev_sh, ctx->copy_stream));
                ctx->sh_pending = true;
                ctx->sh_gather_chunk = chunk;
                ctx->sh_gather_buf = L0.sh.as<float>();
                ctx->sh_gather_row = Fp * 4;
            } else if (Fp > 0) {
                // the SH rows (3/4 of the bytes) are first read by the level-0 moment kernel: their H2D
                // runs on the copy stream, behind the geometry, under the prep / grid / gate kernels
                CK(cudaEventRecord(ctx->ev_geo, st));
                CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_geo, 0));
                CK(cudaMemcpyAsync(L0.sh.p, sh, (size_t)n * Fp * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
                CK(cudaEventRecord(ctx->ev_sh, ctx->copy_stream));
                ctx->sh_pending = true;
                ctx->bytes_h2d += (size_t)n * Fp * 4;
            }
        }
    }
    uint64_t base = ctx->rng_base_set ? ctx->rng_base : (uint64_t)n;
    ctx->bind_too_small = false;
    ctx->sh_level = 0;
    for (auto& h : ctx->stash) h.n = h.np = h.P = 0;
    for (uint32_t l = 0; l < cluster_level; ++l) {
        CKR(run_step(ctx, (int)l, hem_reduction, distance_delta, color_delta, decay_rate, base));
        base += (uint64_t)(ctx->stats[l].n_parents + ctx->stats[l].n_orphans);    // pre-filter size, mixture.cpp:257-259
        // a bound output leaves while the next level is being built (its SH rows later, if the step left them to the SH pass)
        CKR(send_level(ctx, (size_t)l + 1, true, ctx->sh_level >= (int)l + 1, &ctx->bind_too_small));
    }
    CKR(sh_catch_up(ctx, (int)cluster_level, &ctx->bind_too_small));
    const bool bind_too_small = ctx->bind_too_small;
    if (ctx->sh_pending) CKR(wait_sh(ctx));
    if (ctx->shx_pending) { CK(cudaStreamSynchronize(ctx->xchg_stream)); ctx->shx_pending = false; }
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->d2h_stream));
    if (bind_too_small) { ctx->err = "a bound output buffer is smaller than its level (use hem_level_size + hem_level_copy)"; return HEM_ERR_INVALID; }
    return HEM_OK;
}

int hem_num_levels(const hem_ctx* ctx) { return ctx ? (int)ctx->stats.size() + (ctx->levels.empty() ? 0 : 1) : 0; }

int64_t hem_level_size(const hem_ctx* ctx, int level)
{
    if (!ctx || level < 0 || level > (int)ctx->stats.size() || (size_t)level >= ctx->levels.size()) return -1;
    return ctx->levels[level].n;
}

int hem_level_copy(hem_ctx* ctx, int level, float* xyz, float* color, float* opacity, float* cov6, float* sh,
                   float* weight, uint8_t* is_parent)
{
    if (!ctx) return HEM_ERR_INVALID;
    if (level < 0 || level > (int)ctx->stats.size() || (size_t)level >= ctx->levels.size()) { ctx->err = "level not built"; return HEM_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CKR(export_level(ctx, level, xyz, color, opacity, cov6, sh, weight, is_parent, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HEM_OK;
}

int hem_level_registration_inputs(hem_ctx* ctx, int level, double* points, double* colors, double* cov33, double* normals)
{
    if (!ctx) return HEM_ERR_INVALID;
    if (level < 0 || level > (int)ctx->stats.size() || (size_t)level >= ctx->levels.size()) { ctx->err = "level not built"; return HEM_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Level& L = ctx->levels[level];
    const int64_t n = L.n;
    if (n == 0) return HEM_OK;
    double* outs[4] = { points, colors, cov33, normals };
    const size_t per[4] = { 24, 24, 72, 24 };
    DevBuf* stage[4] = { &ctx->out_a, &ctx->out_b, &ctx->out_d, &ctx->out_c };
    double* dev[4];
    bool staged[4];
    for (int k = 0; k < 4; ++k) {
        staged[k] = false;
        dev[k] = outs[k];
        if (outs[k] && !is_device_ptr(outs[k])) {
            CKR(ensure(ctx, *stage[k], (size_t)n * per[k]));
            dev[k] = stage[k]->as<double>();
            staged[k] = true;
        }
    }
    k_registration_inputs<<<blocks_for(n, 256), 256, 0, st>>>(L.geo.as<float4>(), n, dev[0], dev[1], dev[2], dev[3]);
    CK(cudaGetLastError()); ctx->n_launches++;
    for (int k = 0; k < 4; ++k)
        if (staged[k]) {
            CK(cudaMemcpyAsync(outs[k], dev[k], (size_t)n * per[k], cudaMemcpyDeviceToHost, st));
            ctx->bytes_d2h += (size_t)n * per[k];
        }
    CK(cudaStreamSynchronize(st));
    return HEM_OK;
}

int hem_level_from_mixture(hem_ctx* ctx, int level, float* scaling, float* rotation, float* eigenvalues, float* eigenvectors)
{
    if (!ctx) return HEM_ERR_INVALID;
    if (level < 0 || level > (int)ctx->stats.size() || (size_t)level >= ctx->levels.size()) { ctx->err = "level not built"; return HEM_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Level& L = ctx->levels[level];
    const int64_t n = L.n;
    if (n == 0) return HEM_OK;
    float* outs[4] = { scaling, rotation, eigenvalues, eigenvectors };
    const size_t per[4] = { 12, 16, 12, 36 };
    DevBuf* stage[4] = { &ctx->out_a, &ctx->out_b, &ctx->out_c, &ctx->out_d };
    float* dev[4];
    bool staged[4];
    for (int k = 0; k < 4; ++k) {
        staged[k] = false;
        dev[k] = outs[k];
        if (outs[k] && !is_device_ptr(outs[k])) {
            CKR(ensure(ctx, *stage[k], (size_t)n * per[k]));
            dev[k] = stage[k]->as<float>();
            staged[k] = true;
        }
    }
    k_from_mixture<<<blocks_for(n, 128), 128, 0, st>>>(L.geo.as<float4>(), n, dev[0], dev[1], dev[2], dev[3]);
    CK(cudaGetLastError()); ctx->n_launches++;
    for (int k = 0; k < 4; ++k)
        if (staged[k]) {
            CK(cudaMemcpyAsync(outs[k], dev[k], (size_t)n * per[k], cudaMemcpyDeviceToHost, st));
            ctx->bytes_d2h += (size_t)n * per[k];
        }
    CK(cudaStreamSynchronize(st));
    return HEM_OK;
}

int hem_bind_output(hem_ctx* ctx, int level, float* xyz, float* color, float* opacity, float* cov6, float* sh,
                    int64_t capacity_rows)
{
    if (!ctx || level < 1 || level > 64 || capacity_rows < 0) return HEM_ERR_INVALID;
    if (ctx->binds.size() <= (size_t)level) ctx->binds.resize((size_t)level + 1);
    hem_ctx::OutBind& B = ctx->binds[level];
    B.xyz = xyz; B.color = color; B.opacity = opacity; B.cov6 = cov6; B.sh = sh; B.cap = capacity_rows;
    B.set = xyz || color || opacity || cov6 || sh;
    return HEM_OK;
}

int hem_clear_outputs(hem_ctx* ctx)
{
    if (!ctx) return HEM_ERR_INVALID;
    ctx->binds.clear();
    return HEM_OK;
}

// Event times of a step, read on the first request after its build (a cudaEventElapsedTime costs ~3 us: reading all
// of them at the end of every build was 0.23 ms of an 8-level pyramid).  The build has synchronised its streams.
static void read_step_times(hem_ctx* ctx, int l)
{
    StepStats& S = ctx->stats[l];
    if (S.ms_read) return;
    S.ms_read = true;
    if (!S.timed || (size_t)l >= ctx->evs.size() || ctx->evs[l].empty() || S.n_in == 0) return;
    cudaSetDevice(ctx->device);
    cudaEvent_t* ev = ctx->evs[l].data();
    float ms = 0;
    const int pairs[7][2] = { { EV_BEGIN, EV_END }, { EV_BEGIN, EV_PREP }, { EV_PREP, EV_GRID }, { EV_GRID, EV_GATE },
                              { EV_GATE, EV_SUMLW }, { EV_SUMLW, EV_MOM }, { EV_MOM, EV_END } };
    for (int k = 0; k < (S.timed == 1 ? 7 : 1); ++k)
        if (cudaEventElapsedTime(&ms, ev[pairs[k][0]], ev[pairs[k][1]]) == cudaSuccess) S.ms[k] = ms; else cudaGetLastError();
    if (S.timed_pairs) {
        if (cudaEventElapsedTime(&ms, ev[EV_GK0], ev[EV_GK1]) == cudaSuccess) S.ms[7] = ms; else cudaGetLastError();
        if (cudaEventElapsedTime(&ms, ev[EV_MK0], ev[EV_MK1]) == cudaSuccess) S.ms[8] = ms; else cudaGetLastError();
    }
}

int hem_step_stats(const hem_ctx* ctx, int step, double* out)
{
    if (!ctx || !out || step < 0 || step >= (int)ctx->stats.size()) return HEM_ERR_INVALID;
    read_step_times(const_cast<hem_ctx*>(ctx), step);
    const StepStats& S = ctx->stats[step];
    out[0] = (double)S.n_in; out[1] = (double)S.n_parents; out[2] = (double)S.n_pairs; out[3] = (double)S.n_orphans;
    out[4] = (double)S.n_dropped; out[5] = (double)S.n_out; out[6] = (double)S.grid_cells; out[7] = S.cell_size;
    out[8] = S.max_radius;
    for (int i = 0; i < 9; ++i) out[9 + i] = S.ms[i];
    out[18] = (double)S.n_certified;
    out[19] = S.ms_host;
    out[20] = (double)S.n_parents_local; out[21] = (double)S.n_pairs_local; out[22] = (double)S.shards; out[23] = S.tight ? 1.0 : 0.0;
    return HEM_OK;
}

int hem_step_pairs(hem_ctx* ctx, int step, uint32_t* parent_idx, int64_t* off, uint32_t* child, float* wl,
                   float* sumlw, float* radius)
{
    if (!ctx) return HEM_ERR_INVALID;
    if (step < 0 || step >= (int)ctx->pairs.size()) { ctx->err = "pairs not kept (set option keep_pairs before building)"; return HEM_ERR_STATE; }
    const StepPairs& K = ctx->pairs[step];
    if (parent_idx) memcpy(parent_idx, K.parent_idx.data(), K.parent_idx.size() * 4);
    if (off) memcpy(off, K.off.data(), K.off.size() * 8);
    if (child) memcpy(child, K.child.data(), K.child.size() * 4);
    if (wl) memcpy(wl, K.wl.data(), K.wl.size() * 4);
    if (sumlw) memcpy(sumlw, K.sumlw.data(), K.sumlw.size() * 4);
    if (radius) memcpy(radius, K.radius.data(), K.radius.size() * 4);
    return HEM_OK;
}

int hem_get_info(const hem_ctx* ctx, const char* name, double* out)
{
    if (!ctx || !name || !out) return HEM_ERR_INVALID;
    std::string k(name);
    if (k == "launches") *out = (double)ctx->n_launches;
    else if (k == "bytes_h2d") *out = (double)ctx->bytes_h2d;
    else if (k == "bytes_d2h") *out = (double)ctx->bytes_d2h;
    else if (k == "sm_count") *out = (double)ctx->sm_count;
    else if (k == "world") *out = (double)ctx->world;
    else if (k == "rank") *out = (double)ctx->rank;
    else return HEM_ERR_INVALID;
    return HEM_OK;
}

int hem_nccl_unique_id(uint8_t* id_out)
{
    if (!id_out) return HEM_ERR_INVALID;
    if (!nccl_load()) { g_create_error = g_nccl.err; return HEM_ERR_NCCL; }
    NcclId id;
    memset(&id, 0, sizeof id);
    if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return HEM_ERR_NCCL; }
    memcpy(id_out, &id, HEM_NCCL_ID_BYTES);
    return HEM_OK;
}

int hem_comm_init(hem_ctx* ctx, const uint8_t* id, int rank, int world)
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return HEM_ERR_INVALID;
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return HEM_OK; }
    if (world > 64) { ctx->err = "more than 64 ranks are not supported (in-place exchange padding)"; return HEM_ERR_INVALID; }
    if (!nccl_load()) { ctx->err = g_nccl.err; return HEM_ERR_NCCL; }
    CK(cudaSetDevice(ctx->device));
    NcclId nid;
    memcpy(&nid, id, HEM_NCCL_ID_BYTES);
    void* comm = nullptr;
    NK(g_nccl.CommInitRank(&comm, world, nid, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return HEM_OK;
}

int hem_slab_range(int64_t n_parents, int rank, int world, int64_t* begin, int64_t* end)
{
    if (world < 1 || rank < 0 || rank >= world || n_parents < 0 || !begin || !end) return HEM_ERR_INVALID;
    *begin = slab_bound(n_parents, rank, world);
    *end = slab_bound(n_parents, rank + 1, world);
    return HEM_OK;
}

} // extern "C"

#include "hem_icp_host.inl"
