// hem_kernels.cuh — the CUDA kernels of one HEM clustering step (sm_100a).
//
// One step = reference Mixture::createClusterLevel (src/cpp_ext/src/mixture.cpp:66-285).
// Stage map (DESIGN.md has the data-flow picture and the byte accounting):
//   S1  k_prep            parent radius, bbox / radius statistics            mixture.cpp:70-95
//   S2  k_cellkey + radix sort + k_sorted_gather   uniform grid              pointindex.cpp:55-108
//   S3  k_gate            radius / colour / KLD / parent-exclusion gates,    mixture.cpp:106-164
//                         likelihood wL, pooled pair emission
//       k_chunk_copy      pool -> parent-major CSR (ascending parent index)
//       radix sort + k_segsum   per-child normaliser sumLw, ascending parent order  mixture.cpp:143-163
//   S4  k_moments         fused weight/mean/cov/colour/opacity/SH moments    mixture.cpp:168-247
//   S5  k_orphans         orphan copy, flags, validity                       mixture.cpp:251-282
//       k_compact_rows    validity filter compaction (only if something was dropped)
#pragma once
#include "hem_math.cuh"
#include <cfloat>

namespace hem {

// ---------------------------------------------------------------------------
// Layouts
// ---------------------------------------------------------------------------
// Canonical level store (index = canonical component index of the level):
//   geo : 4 x float4 per component, one aligned 64-byte row
//         q0 = (mean.x, mean.y, mean.z, weight)
//         q1 = (c00, c01, c02, c11)
//         q2 = (c12, c22, opacity, det(cov))
//         q3 = (col.r, col.g, col.b, is_parent as uint bits)
//   sh  : Fp floats per component, Fp = F rounded up to a multiple of 4 (16-byte rows,
//         192 B = six full 32-byte sectors at SH degree 3)
// Cell-sorted working set of a step (index = slot in cell order), SoA float4 planes:
//   sp0 = (mean.x, mean.y, mean.z, bits(canonical index | is_parent << 31))
//   sp1 = (col.r, col.g, col.b, opacity)
//   sp2 = (c00, c01, c02, c11)
//   sp3 = (c12, c22, det, 0)
struct GridP {
    float ox, oy, oz, inv_h;
    int gx, gy, gz;
};

// device counters (unsigned long long each)
enum { C_POOL = 0, C_CHUNK = 1, C_INVALID = 2, C_NORPH = 3, C_NPAR = 4, C_COUNT = 8 };

// statistics block written by k_prep
struct PrepStats {
    int bbmin[3];          // ordered-int encoded floats
    int bbmax[3];
    unsigned int max_radius_bits;
    unsigned int n_radius;        // parents with finite positive radius
    double sum_radius;
};

__device__ __forceinline__ int f2ord(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ord2f(int i)
{
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__device__ __forceinline__ int cell_axis(float x, float o, float inv_h, int g)
{
    // monotone non-decreasing in x: float sub, mul by a positive constant, floor, clamp
    int c = __float2int_rd(fm(fs(x, o), inv_h));     // NaN -> 0
    return min(max(c, 0), g - 1);
}

// ---------------------------------------------------------------------------
// import / export between the reference's array layout and the level store
// ---------------------------------------------------------------------------
__global__ void k_import(const float* __restrict__ xyz, const float* __restrict__ col,
                         const float* __restrict__ op, const float* __restrict__ cov6,
                         const float* __restrict__ weight, const uint8_t* __restrict__ mask,
                         float4* __restrict__ geo, int64_t n, uint64_t seed, float pp)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sym3 c = { cov6[6 * i], cov6[6 * i + 1], cov6[6 * i + 2], cov6[6 * i + 3], cov6[6 * i + 4], cov6[6 * i + 5] };
    float w = weight ? weight[i] : 1.0f;                                  // mixture.cpp:315
    bool flag = mask ? (mask[i] != 0) : parent_flag(seed, (uint64_t)i, pp);   // mixture.cpp:330
    geo[4 * i + 0] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], w);
    geo[4 * i + 1] = make_float4(c.e00, c.e01, c.e02, c.e11);
    geo[4 * i + 2] = make_float4(c.e12, c.e22, op[i], det_sym3(c));
    geo[4 * i + 3] = make_float4(col[3 * i], col[3 * i + 1], col[3 * i + 2], __uint_as_float(flag ? 1u : 0u));
}

__global__ void k_pad_sh(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int F, int Fp)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * Fp) return;
    int64_t i = e / Fp;
    int q = (int)(e - i * Fp);
    dst[e] = q < F ? src[i * F + q] : 0.0f;
}

__global__ void k_unpad_sh(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int F, int Fp)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * F) return;
    int64_t i = e / F;
    int q = (int)(e - i * F);
    dst[e] = src[i * Fp + q];
}

__global__ void k_export(const float4* __restrict__ geo, int64_t n, float* __restrict__ xyz,
                         float* __restrict__ col, float* __restrict__ op, float* __restrict__ cov6,
                         float* __restrict__ weight, uint8_t* __restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q0 = geo[4 * i], q1 = geo[4 * i + 1], q2 = geo[4 * i + 2], q3 = geo[4 * i + 3];
    if (xyz) { xyz[3 * i] = q0.x; xyz[3 * i + 1] = q0.y; xyz[3 * i + 2] = q0.z; }
    if (col) { col[3 * i] = q3.x; col[3 * i + 1] = q3.y; col[3 * i + 2] = q3.z; }
    if (op) op[i] = q2.z;
    if (cov6) {
        cov6[6 * i] = q1.x; cov6[6 * i + 1] = q1.y; cov6[6 * i + 2] = q1.z;
        cov6[6 * i + 3] = q1.w; cov6[6 * i + 4] = q2.x; cov6[6 * i + 5] = q2.y;
    }
    if (weight) weight[i] = q0.w;
    if (flag) flag[i] = (uint8_t)(__float_as_uint(q3.w) & 1u);
}

// ---------------------------------------------------------------------------
// S1: parent radii + statistics.  mixture.cpp:70-95
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prep(const float4* __restrict__ geo, int64_t n, float distance_delta, float* __restrict__ radius,
       uint32_t* __restrict__ pflag, PrepStats* __restrict__ stats)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bmin[3] = { INT_MAX, INT_MAX, INT_MAX }, bmax[3] = { INT_MIN, INT_MIN, INT_MIN };
    float rad = 0.0f;
    unsigned int rbits = 0, rcount = 0;
    double rsum = 0.0;
    if (i < n) {
        float4 q0 = geo[4 * i], q3 = geo[4 * i + 3];
        bool flag = (__float_as_uint(q3.w) & 1u) != 0;
        pflag[i] = flag ? 1u : 0u;
        if (q0.x == q0.x) { bmin[0] = bmax[0] = f2ord(q0.x); }
        if (q0.y == q0.y) { bmin[1] = bmax[1] = f2ord(q0.y); }
        if (q0.z == q0.z) { bmin[2] = bmax[2] = f2ord(q0.z); }
        if (flag) {
            float4 q1 = geo[4 * i + 1], q2 = geo[4 * i + 2];
            Sym3 c = { q1.x, q1.y, q1.z, q1.w, q2.x, q2.y };
            rad = fm(distance_delta, __fsqrt_rn(eig_max_sym3(c)));       // mixture.cpp:88
            if (rad > 0.0f && rad < FLT_MAX) { rbits = __float_as_uint(rad); rcount = 1; rsum = (double)rad; }
        }
        radius[i] = rad;
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        bmin[a] = __reduce_min_sync(full, bmin[a]);
        bmax[a] = __reduce_max_sync(full, bmax[a]);
    }
    rbits = __reduce_max_sync(full, rbits);
    rcount = __reduce_add_sync(full, rcount);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(full, rsum, o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (bmin[a] != INT_MAX) atomicMin(&stats->bbmin[a], bmin[a]);
            if (bmax[a] != INT_MIN) atomicMax(&stats->bbmax[a], bmax[a]);
        }
        if (rcount) {
            atomicMax(&stats->max_radius_bits, rbits);
            atomicAdd(&stats->n_radius, rcount);
            atomicAdd(&stats->sum_radius, rsum);
        }
    }
}

// ---------------------------------------------------------------------------
// S2: uniform grid.  Replaces PointIndex::create (pointindex.cpp:55-108): cell
// keys -> radix sort (CUB, in hem_b200.cu) -> cell-sorted SoA planes + offsets.
// ---------------------------------------------------------------------------
__global__ void k_cellkey(const float4* __restrict__ geo, int64_t n, GridP g, uint32_t* __restrict__ key,
                          uint32_t* __restrict__ val)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q0 = geo[4 * i];
    int cx = cell_axis(q0.x, g.ox, g.inv_h, g.gx);
    int cy = cell_axis(q0.y, g.oy, g.inv_h, g.gy);
    int cz = cell_axis(q0.z, g.oz, g.inv_h, g.gz);
    key[i] = (uint32_t)((cz * g.gy + cy) * g.gx + cx);
    val[i] = (uint32_t)i;
}

__global__ void k_sorted_gather(const float4* __restrict__ geo, const uint32_t* __restrict__ skey,
                                const uint32_t* __restrict__ sval, int64_t n, uint32_t ncell,
                                float4* __restrict__ sp0, float4* __restrict__ sp1, float4* __restrict__ sp2,
                                float4* __restrict__ sp3, uint8_t* __restrict__ sflag,
                                uint32_t* __restrict__ cell_start)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t i = sval[j], key = skey[j];
    float4 q0 = geo[4 * (size_t)i], q1 = geo[4 * (size_t)i + 1], q2 = geo[4 * (size_t)i + 2], q3 = geo[4 * (size_t)i + 3];
    uint32_t flag = __float_as_uint(q3.w) & 1u;
    sp0[j] = make_float4(q0.x, q0.y, q0.z, __uint_as_float(i | (flag << 31)));
    sp1[j] = make_float4(q3.x, q3.y, q3.z, q2.z);
    sp2[j] = q1;
    sp3[j] = make_float4(q2.x, q2.y, q2.w, 0.0f);
    sflag[j] = (uint8_t)flag;
    // segment offsets: cell_start[c] = first slot whose key >= c
    uint32_t lo = (j == 0) ? 0u : skey[j - 1] + 1u;
    if (j == 0 || key != skey[j - 1])
        for (uint32_t c = lo; c <= key; ++c) cell_start[c] = (uint32_t)j;
    if (j == n - 1)
        for (uint32_t c = key + 1; c <= ncell; ++c) cell_start[c] = (uint32_t)n;
}

// ---------------------------------------------------------------------------
// S3: gates + likelihood.  One warp per parent, parents taken in cell order so
// that neighbouring warps stream the same cell runs out of L2.
// ---------------------------------------------------------------------------
constexpr int GATE_WARPS = 8;
constexpr int GATE_OUTCAP = 96;

struct GateArgs {
    const float4* sp0; const float4* sp1; const float4* sp2; const float4* sp3;
    const uint32_t* cell_start;
    const uint32_t* psorted;     // [np] sorted slot of each parent, in cell order
    const uint32_t* prank;       // [n]  canonical parent rank of component i
    const float* radius;         // [n]
    const float4* geo;           // canonical rows (parent weight)
    GridP g;
    int np;
    float cthr, kthr, decay2;
    uint32_t* pair_cnt;          // [np+1] by canonical parent rank
    uint2* pool; unsigned long long pool_cap;
    uint4* chunks; unsigned long long chunk_cap;
    unsigned long long* counters;
};

__global__ void __launch_bounds__(GATE_WARPS * 32)
k_gate(GateArgs A)
{
    __shared__ uint32_t s_q[GATE_WARPS][64];
    __shared__ uint2 s_out[GATE_WARPS][GATE_OUTCAP];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * GATE_WARPS + warp;
    if (t >= A.np) return;

    const uint32_t pj = A.psorted[t];
    const float4 p0 = A.sp0[pj], p1 = A.sp1[pj], p2 = A.sp2[pj], p3 = A.sp3[pj];
    const uint32_t s = __float_as_uint(p0.w) & 0x7fffffffu;
    const uint32_t srank = A.prank[s];
    const float rad = A.radius[s];
    const float pw = A.geo[4 * (size_t)s].w;
    const float r2 = fm(rad, rad);
    const Sym3 pc = { p2.x, p2.y, p2.z, p2.w, p3.x, p3.y };
    const float pdet = p3.z;
    const Sym3 ip = inverse_sym3(pc, pdet);

    uint32_t* q = s_q[warp];
    uint2* ob = s_out[warp];
    int qn = 0, on = 0;
    uint32_t total = 0;

    auto flush = [&]() {
        if (on > 0) {
            unsigned long long base = 0, cid = 0;
            if (lane == 0) {
                base = atomicAdd(&A.counters[C_POOL], (unsigned long long)on);
                cid = atomicAdd(&A.counters[C_CHUNK], 1ull);
            }
            base = __shfl_sync(full, base, 0);
            if (base + (unsigned long long)on <= A.pool_cap)
                for (int k = lane; k < on; k += 32) A.pool[base + k] = ob[k];
            if (lane == 0 && cid < A.chunk_cap)
                A.chunks[cid] = make_uint4(srank, (uint32_t)base, (uint32_t)on, total);
            total += (uint32_t)on;
            on = 0;
            __syncwarp();
        }
    };

    // candidates that passed the radius + parent-exclusion test: colour gate, KLD gate, likelihood
    auto process = [&](int m) {
        bool acc = false;
        uint32_t cidx = 0;
        float wl = 0.0f;
        if (lane < m) {
            const uint32_t jj = q[lane];
            const float4 c1 = A.sp1[jj];
            // colour gate, mixture.cpp:121-124 / gaussian.hpp:111-114 (NaN passes)
            const float cd2 = sqdist3(c1.x, c1.y, c1.z, p1.x, p1.y, p1.z);
            const float cdist = __fsqrt_rn(cd2);
            if (!(cdist > A.cthr)) {
                const float4 c0 = A.sp0[jj], c2 = A.sp2[jj], c3 = A.sp3[jj];
                // KLD(child || parent), gaussian.hpp:106-109
                const float dx = fs(c0.x, p0.x), dy = fs(c0.y, p0.y), dz = fs(c0.z, p0.z);
                const float tx = fa(fa(fm(ip.e00, dx), fm(ip.e01, dy)), fm(ip.e02, dz));
                const float ty = fa(fa(fm(ip.e01, dx), fm(ip.e11, dy)), fm(ip.e12, dz));
                const float tz = fa(fa(fm(ip.e02, dx), fm(ip.e12, dy)), fm(ip.e22, dz));
                const float smd = fa(fa(fm(dx, tx), fm(dy, ty)), fm(dz, tz));
                const float t00 = fa(fa(fm(c2.x, ip.e00), fm(c2.y, ip.e01)), fm(c2.z, ip.e02));
                const float t11 = fa(fa(fm(c2.y, ip.e01), fm(c2.w, ip.e11)), fm(c3.x, ip.e12));
                const float t22 = fa(fa(fm(c2.z, ip.e02), fm(c3.x, ip.e12)), fm(c3.y, ip.e22));
                const float tr = fa(fa(t00, t11), t22);
                const float cdet = c3.z;
                const float kl = fm(0.5f, fs(fs(fa(smd, tr), 3.0f), logf_glibc(fd(cdet, pdet))));
                if (!(kl > A.kthr)) {
                    acc = true;
                    cidx = __float_as_uint(c0.w) & 0x7fffffffu;
                    // hemLikelihoodOpacity, mixture.cpp:54-64; clamp + parent weight, :155-156
                    const float dd = __fsqrt_rn(sqdist3(p0.x, p0.y, p0.z, c0.x, c0.y, c0.z));
                    const float distWeight = expf(fd(fm(-dd, dd), A.decay2));
                    const float colorInfluence = expf(fd(fm(-cdist, cdist), A.decay2));
                    const float L = fm(fm(fm(distWeight, c1.w), colorInfluence), __fsqrt_rn(cdet));
                    wl = fm(pw, clamp_ref(L, FLT_MIN, 1e8f));
                }
            }
        }
        const unsigned am = __ballot_sync(full, acc);
        if (acc) ob[on + __popc(am & ((1u << lane) - 1u))] = make_uint2(cidx, __float_as_uint(wl));
        on += __popc(am);
        __syncwarp();
        if (on > GATE_OUTCAP - 32) flush();
    };

    if (rad > 0.0f) {        // NaN / zero radius: no child can pass `d2 < r*r` (pointindex.cpp:138)
        // conservative cell range (see DESIGN.md "grid range"): the float test can only pass
        // for |dx| <= rad*(1+5*2^-24); pad for the rounding of the range ends themselves.
        const float rr = rad * 1.000001f;
        const float padx = (fabsf(p0.x) + rr) * 2.4e-7f, pady = (fabsf(p0.y) + rr) * 2.4e-7f,
                    padz = (fabsf(p0.z) + rr) * 2.4e-7f;
        const int x0 = cell_axis(p0.x - (rr + padx), A.g.ox, A.g.inv_h, A.g.gx);
        const int x1 = cell_axis(p0.x + (rr + padx), A.g.ox, A.g.inv_h, A.g.gx);
        const int y0 = cell_axis(p0.y - (rr + pady), A.g.oy, A.g.inv_h, A.g.gy);
        const int y1 = cell_axis(p0.y + (rr + pady), A.g.oy, A.g.inv_h, A.g.gy);
        const int z0 = cell_axis(p0.z - (rr + padz), A.g.oz, A.g.inv_h, A.g.gz);
        const int z1 = cell_axis(p0.z + (rr + padz), A.g.oz, A.g.inv_h, A.g.gz);
        const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
        const int nrows = ny * nz;

        for (int rb = 0; rb < nrows; rb += 32) {
            // each lane prefetches the [begin,end) run of one (y,z) row of cells
            uint32_t ra = 0, re = 0;
            const int rrow = rb + lane;
            if (rrow < nrows) {
                const int y = y0 + rrow % ny, z = z0 + rrow / ny;
                const uint32_t base = (uint32_t)((z * A.g.gy + y) * A.g.gx);
                ra = A.cell_start[base + x0];
                re = A.cell_start[base + x1 + 1];
            }
            const int cnt = min(32, nrows - rb);
            for (int k = 0; k < cnt; ++k) {
                const uint32_t a = __shfl_sync(full, ra, k), e = __shfl_sync(full, re, k);
                for (uint32_t b = a; b < e; b += 32) {
                    const uint32_t jj = b + lane;
                    bool in = false;
                    if (jj < e) {
                        const float4 c0 = A.sp0[jj];
                        const uint32_t w = __float_as_uint(c0.w);
                        // radius test, pointindex.cpp:138 (query - point); parent exclusion, mixture.cpp:131-133
                        const float d2 = sqdist3(p0.x, p0.y, p0.z, c0.x, c0.y, c0.z);
                        in = (d2 < r2) && !((w >> 31) && ((w & 0x7fffffffu) != s));
                    }
                    const unsigned m = __ballot_sync(full, in);
                    if (in) q[qn + __popc(m & ((1u << lane) - 1u))] = jj;
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) {
                        process(32);
                        const uint32_t mv = (lane + 32 < qn) ? q[lane + 32] : 0u;
                        __syncwarp();
                        q[lane] = mv;
                        qn -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        if (qn > 0) process(qn);
        flush();
    }
    if (lane == 0) A.pair_cnt[srank] = total;
}

// pool -> CSR in ascending canonical parent order
__global__ void k_chunk_copy(const uint4* __restrict__ chunks, unsigned long long nchunks,
                             const uint2* __restrict__ pool, const uint32_t* __restrict__ pair_off,
                             uint32_t* __restrict__ csr_child, float* __restrict__ csr_wl)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long c = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= nchunks) return;
    const uint4 ch = chunks[c];
    const size_t dst = (size_t)pair_off[ch.x] + ch.w;
    for (uint32_t k = lane; k < ch.z; k += 32) {
        const uint2 p = pool[(size_t)ch.y + k];
        csr_child[dst + k] = p.x;
        csr_wl[dst + k] = __uint_as_float(p.y);
    }
}

// per-child normaliser: pairs stably sorted by child (parents ascending inside a
// child), summed left to right like mixture.cpp:143-163
__global__ void k_segsum(const uint32_t* __restrict__ key, const float* __restrict__ val, int64_t P,
                         float* __restrict__ sumlw)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t c = key[k];
    if (k > 0 && key[k - 1] == c) return;
    float s = 0.0f;
    for (int64_t m = k; m < P && key[m] == c; ++m) s = fa(s, val[m]);
    sumlw[c] = s;
}

// ---------------------------------------------------------------------------
// S4: fused parent moments.  One warp per parent.  mixture.cpp:168-247
// ---------------------------------------------------------------------------
constexpr int MOM_WARPS = 8;

struct MomArgs {
    const float4* geo; const float* sh; int Fp;
    const uint32_t* psorted; const float4* sp0; const uint32_t* prank; int np;
    const uint32_t* pair_off; const uint32_t* csr_child; const float* csr_wl; const float* sumlw;
    float4* ogeo; float* osh; uint8_t* valid;
    uint64_t seed, rng_base; float pp;
    unsigned long long* counters;
};

__global__ void __launch_bounds__(MOM_WARPS * 32)
k_moments(MomArgs A)
{
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * MOM_WARPS + warp;
    if (t >= A.np) return;
    const uint32_t s = __float_as_uint(A.sp0[A.psorted[t]].w) & 0x7fffffffu;
    const uint32_t srank = A.prank[s];
    const uint32_t a = A.pair_off[srank], b = A.pair_off[srank + 1];
    const float4 pq0 = A.geo[4 * (size_t)s];
    const int Fp = A.Fp;

    float acc[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = 0.0f;
    float sh0 = 0.0f, sh1 = 0.0f;

    for (uint32_t k0 = a; k0 < b; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint32_t ch = 0;
        float w = 0.0f;
        float tm[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) tm[i] = 0.0f;
        if (k < b) {
            ch = A.csr_child[k];
            const float wl = A.csr_wl[k];
            const float sl = A.sumlw[ch];
            const float4 q0 = A.geo[4 * (size_t)ch], q1 = A.geo[4 * (size_t)ch + 1],
                         q2 = A.geo[4 * (size_t)ch + 2], q3 = A.geo[4 * (size_t)ch + 3];
            w = fm(fd(wl, sl), q0.w);                       // r_is * child.weight, mixture.cpp:196-197
            const float dx = q0.x - pq0.x, dy = q0.y - pq0.y, dz = q0.z - pq0.z;
            tm[0] = w;
            tm[1] = w * q0.x; tm[2] = w * q0.y; tm[3] = w * q0.z;
            tm[4] = w * q3.x; tm[5] = w * q3.y; tm[6] = w * q3.z;
            tm[7] = w * (q1.x + dx * dx); tm[8] = w * (q1.y + dx * dy); tm[9] = w * (q1.z + dx * dz);
            tm[10] = w * (q1.w + dy * dy); tm[11] = w * (q2.x + dy * dz); tm[12] = w * (q2.y + dz * dz);
            tm[13] = w * q2.z;
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            float v = tm[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
            acc[i] += v;
        }
        // SH rows: the whole warp reads one child's row (coalesced), children in list order
        const int cnt = (int)min(32u, b - k0);
        int c = 0;
        for (; c + 4 <= cnt; c += 4) {
            float wc[4], v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                wc[u] = __shfl_sync(full, w, c + u);
                const uint32_t cc = __shfl_sync(full, ch, c + u);
                const float* row = A.sh + (size_t)cc * Fp;
                v0[u] = lane < Fp ? row[lane] : 0.0f;
                v1[u] = lane + 32 < Fp ? row[lane + 32] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { sh0 = fmaf(v0[u], wc[u], sh0); sh1 = fmaf(v1[u], wc[u], sh1); }
        }
        for (; c < cnt; ++c) {
            const float wc = __shfl_sync(full, w, c);
            const uint32_t cc = __shfl_sync(full, ch, c);
            const float* row = A.sh + (size_t)cc * Fp;
            if (lane < Fp) sh0 = fmaf(row[lane], wc, sh0);
            if (lane + 32 < Fp) sh1 = fmaf(row[lane + 32], wc, sh1);
        }
    }

    // normalise, mixture.cpp:219-226
    const float inv_w = fd(1.0f, acc[0]);
    const float mx = inv_w * acc[1], my = inv_w * acc[2], mz = inv_w * acc[3];
    const float dmx = mx - pq0.x, dmy = my - pq0.y, dmz = mz - pq0.z;
    Sym3 cv;
    cv.e00 = fs(fm(inv_w, acc[7]), fm(dmx, dmx)); cv.e01 = fs(fm(inv_w, acc[8]), fm(dmx, dmy));
    cv.e02 = fs(fm(inv_w, acc[9]), fm(dmx, dmz)); cv.e11 = fs(fm(inv_w, acc[10]), fm(dmy, dmy));
    cv.e12 = fs(fm(inv_w, acc[11]), fm(dmy, dmz)); cv.e22 = fs(fm(inv_w, acc[12]), fm(dmz, dmz));
    const float dt = det_sym3(cv);
    // pre-filter position of a merged parent = its rank; flag drawn there, mixture.cpp:257-259
    const bool flag = parent_flag(A.seed, A.rng_base + (uint64_t)srank, A.pp);
    float4 o;
    if (lane == 0) o = make_float4(mx, my, mz, acc[0]);
    else if (lane == 1) o = make_float4(cv.e00, cv.e01, cv.e02, cv.e11);
    else if (lane == 2) o = make_float4(cv.e12, cv.e22, inv_w * acc[13], dt);
    else o = make_float4(inv_w * acc[4], inv_w * acc[5], inv_w * acc[6], __uint_as_float(flag ? 1u : 0u));
    if (lane < 4) A.ogeo[4 * (size_t)srank + lane] = o;
    float* orow = A.osh + (size_t)srank * Fp;
    if (lane < Fp) orow[lane] = inv_w * sh0;
    if (lane + 32 < Fp) orow[lane + 32] = inv_w * sh1;
    if (lane == 0) {
        // validity filter predicate, mixture.cpp:268-270
        const bool bad = (mx != mx || my != my || mz != mz) || (dt != dt) || dt <= 0.0f;
        A.valid[srank] = bad ? 0 : 1;
        if (bad) atomicAdd(&A.counters[C_INVALID], 1ull);
    }
}

// ---------------------------------------------------------------------------
// S5: orphans (mixture.cpp:251-253), their flags (:257-259) and validity (:262-282)
// ---------------------------------------------------------------------------
__global__ void k_orphan_flag(const float* __restrict__ sumlw, int64_t n, uint32_t* __restrict__ oflag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) oflag[i] = sumlw[i] == 0.0f ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
k_orphans(const float4* __restrict__ geo, const float* __restrict__ sh, int Fp, int64_t n,
          const uint32_t* __restrict__ oflag, const uint32_t* __restrict__ orank, uint32_t np,
          float4* __restrict__ ogeo, float* __restrict__ osh, uint8_t* __restrict__ valid,
          uint64_t seed, uint64_t rng_base, float pp, unsigned long long* counters)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t wbase = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (wbase >= n) return;
    const int64_t i = wbase + lane;
    const bool orphan = i < n && oflag[i] != 0;
    unsigned m = __ballot_sync(full, orphan);
    const uint32_t myrank = orphan ? orank[i] : 0u;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int64_t ci = wbase + src;
        const size_t dst = (size_t)np + __shfl_sync(full, myrank, src);
        float4 v;
        if (lane < 4) v = geo[4 * (size_t)ci + lane];
        // lanes 0..3 hold q0..q3; validity needs q0 (mean) and q2.w (det)
        const float mx = __shfl_sync(full, v.x, 0), my = __shfl_sync(full, v.y, 0), mz = __shfl_sync(full, v.z, 0);
        const float dt = __shfl_sync(full, v.w, 2);
        if (lane == 3) v.w = __uint_as_float(parent_flag(seed, rng_base + (uint64_t)dst, pp) ? 1u : 0u);
        if (lane < 4) ogeo[4 * dst + lane] = v;
        const float* srow = sh + (size_t)ci * Fp;
        float* drow = osh + dst * Fp;
        if (lane < Fp) drow[lane] = srow[lane];
        if (lane + 32 < Fp) drow[lane + 32] = srow[lane + 32];
        if (lane == 0) {
            const bool bad = (mx != mx || my != my || mz != mz) || (dt != dt) || dt <= 0.0f;
            valid[dst] = bad ? 0 : 1;
            if (bad) atomicAdd(&counters[C_INVALID], 1ull);
        }
    }
}

__global__ void k_valid_flag(const uint8_t* __restrict__ valid, int64_t n, uint32_t* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = valid[i];
}

// erase invalid outputs keeping order, mixture.cpp:262-282
__global__ void k_compact_rows(const float4* __restrict__ geo, const float* __restrict__ sh, int Fp, int64_t n,
                               const uint8_t* __restrict__ valid, const uint32_t* __restrict__ pos,
                               float4* __restrict__ ogeo, float* __restrict__ osh)
{
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n || !valid[j]) return;
    const size_t d = pos[j];
    if (lane < 4) ogeo[4 * d + lane] = geo[4 * (size_t)j + lane];
    for (int q = lane; q < Fp; q += 32) osh[d * Fp + q] = sh[(size_t)j * Fp + q];
}

} // namespace hem
