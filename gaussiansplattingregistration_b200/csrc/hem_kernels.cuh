// hem_kernels.cuh — the CUDA kernels of one HEM clustering step (sm_100a).
//
// One step = reference Mixture::createClusterLevel (src/cpp_ext/src/mixture.cpp:66-285).
// Stage map (DESIGN.md has the data-flow picture and the byte accounting):
//   S1  k_prep (level 0) / k_finish of the previous step (levels >= 1)
//                         parent radius, certificates, bbox / radius statistics     mixture.cpp:70-95
//       k_boxhist         percentile box of the level-0 means (grid extent only)
//   S2  k_cellkey + radix sort + k_sorted_gather (+ per-parent gate records)   uniform grid   pointindex.cpp:55-108
//   S3  k_gate            radius / colour / KLD / parent-exclusion gates,           mixture.cpp:106-164
//                         likelihood wL, pooled pair emission
//       k_chunk_copy      pool -> cell-order CSR (+ claim-table insertion on one GPU)
//       k_claims / k_sumlw_claims   per-child normaliser sumLw, ascending parent order  mixture.cpp:143-163
//       (k_cnt_canon / k_csr_canon / radix sort / k_segsum: the sort-based path of the same sum)
//   S4  k_moments         fused weight/mean/cov/colour/opacity/SH moments           mixture.cpp:168-247
//   S5  k_orphans         orphan copy                                               mixture.cpp:251-282
//       k_finish          det, Philox flag, validity, next step's S1
//       k_valid_flag / k_compact_rows   validity filter compaction (only if something was dropped)
//   I/O k_import / k_export / k_registration_inputs, k_zero / k_publish (control block)
#pragma once
#include "hem_math.cuh"
#include <cfloat>

namespace hem {

// Programmatic dependent launch (see pdl() in hem_b200.cu): let the next kernel of the stream become
// resident early, then wait until the previous kernel has completed and its writes are visible.  Must be
// the first statement of every kernel launched through pdl(); a no-op under an ordinary launch.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------
// Layouts
// ---------------------------------------------------------------------------
// Canonical level store (index = canonical component index of the level):
//   geo : 4 x float4 per component, one aligned 64-byte row
//         q0 = (mean.x, mean.y, mean.z, weight)
//         q1 = (c00, c01, c02, c11)
//         q2 = (c12, c22, opacity, det(cov))
//         q3 = (col.r, col.g, col.b, uint bits: bit0 = is_parent, bit1 = covariance certified for the smd pre-test)
//   sh  : F floats per component, unpadded (180-byte rows at SH degree 3), exactly the caller's
//         [N,F] layout: a device-resident input is used in place and levels are copied out by DMA
// Cell-sorted working set of a step, SoA float4 planes.  Slots [0, n - np) hold the NON-PARENT components in grid-cell
// order (what `cell_start` indexes and the gate scans), slots [n - np, n) the parents in grid-cell order (slot of the
// parent at position t = n - np + t): a parent can only claim non-parents and itself (mixture.cpp:131-133), so a third
// of the candidates never enters the scan and the parents need no separate selection pass.
//   sp0 = (mean.x, mean.y, mean.z, bits(canonical index | is_parent << 31))
//   sp1 = (col.r, col.g, col.b, opacity)
//   sp2 = (c00, c01, c02, c11)
//   sp3 = (c12, c22, det, bits: 1 = covariance certified for the smd pre-test)
struct GridP {
    float ox, oy, oz, inv_h, h;     // h = 1/inv_h evaluated in FP32 (cell edge used for row culling)
    int gx, gy, gz;
};

// Per-step control block.  Everything in it (and the per-parent / per-child arrays that follow
// it in the same allocation) starts from zero, so ONE memset opens a step.
enum { C_POOL = 0, C_CHUNK = 1, C_INVALID = 2, C_NORPH = 3, C_WORK = 4, C_WORK2 = 5, C_OVF = 6, C_FAIL = 7, C_COUNT = 8 };

struct Ctl {
    unsigned long long counters[C_COUNT];
    // bounding box of the means in a zero-initialisable encoding: u(x) = ord(x) ^ 0x80000000 is
    // monotone in x as an unsigned; bbmax_u = max u(x), bbmin_c = max ~u(x); 0 means "no finite mean"
    unsigned int bbmax_u[3];
    unsigned int bbmin_c[3];
    unsigned int max_radius_bits;   // positive floats order like their bit patterns
    unsigned int n_radius;          // parents with finite positive radius
    unsigned int n_parents;
    unsigned int n_certified;       // parents whose covariance is certified for the smd pre-test
    unsigned int n_heavy;           // moment tickets deferred to k_moments_heavy in this step
    unsigned int n_heavy_gate;      // parents whose search box has more rows of cells than the gate's row threshold
    unsigned long long sum_radius_fx;   // sum of (log2(radius) + 128) in 2^-16 units over the parents with a finite positive
                                        // radius: the grid cell is sized from the GEOMETRIC mean radius (robust against the
                                        // heavy tail of splat scales); integer adds are order-independent, so the cell size —
                                        // and with it every summation order — is deterministic
    // the same statistics for the NEXT level, produced by k_finish while it finalises the output rows
    // (so the next step starts without a prep kernel and without a host sync)
    unsigned int nx_bbmax_u[3];
    unsigned int nx_bbmin_c[3];
    unsigned int nx_max_radius_bits, nx_n_radius, nx_n_parents, nx_n_certified;
    unsigned long long nx_sum_radius_fx;
    // components whose covariance is NOT certified for the smd pre-test (hem_math.cuh), this level / next level:
    // with none of them, certified parents may search the box of their Mahalanobis ellipsoid instead of their sphere
    unsigned int n_uncert, nx_n_uncert;
};

// Control block -> mapped pinned host memory.  The host reads step sizes after a stream sync; a DMA
// readback would queue behind the bulk level downloads that share the D2H copy engine.
// `seq` (optional, mapped as well) receives `seq_val` once the block is visible to the host: the host may poll it
// instead of paying a stream synchronisation (one block only).
__global__ void k_publish(const Ctl* __restrict__ d, Ctl* __restrict__ h_mapped, unsigned* __restrict__ seq = nullptr,
                          unsigned seq_val = 0u)
{
    pdl_enter();
    const unsigned* s = reinterpret_cast<const unsigned*>(d);
    unsigned* o = reinterpret_cast<unsigned*>(h_mapped);
    for (unsigned i = threadIdx.x; i < sizeof(Ctl) / 4; i += blockDim.x) o[i] = s[i];
    if (seq) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(seq) = seq_val;
    }
}

// zero-fill by a kernel (a cudaMemsetAsync may be routed through a copy engine and queue behind bulk DMA)
__global__ void k_zero(uint4* __restrict__ p, size_t n16)
{
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ unsigned long long radius_fx(float r) { return (unsigned long long)((log2f(r) + 128.0f) * 65536.0f); }

__device__ __forceinline__ int f2ord(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ unsigned int f2u(float f) { return (unsigned int)f2ord(f) ^ 0x80000000u; }
__host__ __device__ __forceinline__ float u2f(unsigned int u)
{
    int i = (int)(u ^ 0x80000000u);
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
    const float dt = det_sym3(c);
    geo[4 * i + 2] = make_float4(c.e12, c.e22, op[i], dt);
    geo[4 * i + 3] = make_float4(col[3 * i], col[3 * i + 1], col[3 * i + 2],
                                 __uint_as_float((flag ? 1u : 0u) | (child_cov_certified(c, dt) ? 2u : 0u)));
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
// "Next" row: what the reference builds on the host right after the HEM step
// to feed Open3D (src/utils/point_cloud_converter.py:31-49): points as FP64, colours = sh2rgb(DC) =
// 0.28209479177387814 * c + 0.5 in FP64 (src/utils/graphics_utils.py:72-73), covariances unpacked to
// [N,3,3] FP64 (src/utils/general_utils.py:32-40), and — the reference's own TODO
// (point_cloud_converter.py:44-47) — a normal from the covariance: the eigenvector of the smallest
// eigenvalue (FP64 closed form + cross products), instead of a KD-tree estimate_normals pass.
// ---------------------------------------------------------------------------
__global__ void k_registration_inputs(const float4* __restrict__ geo, int64_t n, double* __restrict__ points,
                                      double* __restrict__ colors, double* __restrict__ cov33,
                                      double* __restrict__ normals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q0 = geo[4 * i], q1 = geo[4 * i + 1], q2 = geo[4 * i + 2], q3 = geo[4 * i + 3];
    if (points) { points[3 * i] = (double)q0.x; points[3 * i + 1] = (double)q0.y; points[3 * i + 2] = (double)q0.z; }
    if (colors) {
        colors[3 * i] = (double)q3.x * 0.28209479177387814 + 0.5;
        colors[3 * i + 1] = (double)q3.y * 0.28209479177387814 + 0.5;
        colors[3 * i + 2] = (double)q3.z * 0.28209479177387814 + 0.5;
    }
    const double a = q1.x, b = q1.y, c = q1.z, d = q1.w, e = q2.x, f = q2.y;   // [a b c; b d e; c e f]
    if (cov33) {
        double* o = cov33 + 9 * i;
        o[0] = a; o[1] = b; o[2] = c; o[3] = b; o[4] = d; o[5] = e; o[6] = c; o[7] = e; o[8] = f;
    }
    if (normals) {
        // smallest eigenvalue by the trigonometric closed form in FP64, eigenvector = the largest of the
        // three cross products of rows of (A - lambda I)
        const double m = (a + d + f) / 3.0;
        const double k00 = a - m, k11 = d - m, k22 = f - m;
        const double p2 = (k00 * k00 + k11 * k11 + k22 * k22 + 2.0 * (b * b + c * c + e * e)) / 6.0;
        double nx = 0.0, ny = 0.0, nz = 1.0;
        if (p2 > 0.0) {
            const double pp = sqrt(p2);
            const double det = (k00 * (k11 * k22 - e * e) - b * (b * k22 - e * c) + c * (b * e - k11 * c)) / (pp * pp * pp);
            const double r = fmin(1.0, fmax(-1.0, 0.5 * det));
            const double phi = acos(r) / 3.0;
            const double lmin = m + 2.0 * pp * cos(phi + 2.0943951023931953);      // + 2*pi/3: the smallest root
            const double r0x = a - lmin, r0y = b, r0z = c, r1x = b, r1y = d - lmin, r1z = e, r2x = c, r2y = e, r2z = f - lmin;
            double cx[3] = { r0y * r1z - r0z * r1y, r0y * r2z - r0z * r2y, r1y * r2z - r1z * r2y };
            double cy[3] = { r0z * r1x - r0x * r1z, r0z * r2x - r0x * r2z, r1z * r2x - r1x * r2z };
            double cz[3] = { r0x * r1y - r0y * r1x, r0x * r2y - r0y * r2x, r1x * r2y - r1y * r2x };
            int best = 0;
            double bn = -1.0;
            for (int k = 0; k < 3; ++k) {
                const double nn = cx[k] * cx[k] + cy[k] * cy[k] + cz[k] * cz[k];
                if (nn > bn) { bn = nn; best = k; }
            }
            if (bn > 0.0) {
                const double inv = 1.0 / sqrt(bn);
                nx = cx[best] * inv; ny = cy[best] * inv; nz = cz[best] * inv;
                // sign convention: the largest-magnitude component is positive (deterministic)
                const double ax = fabs(nx), ay = fabs(ny), az = fabs(nz);
                const double lead = (ax >= ay && ax >= az) ? nx : (ay >= az ? ny : nz);
                if (lead < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
            }
        }
        normals[3 * i] = nx; normals[3 * i + 1] = ny; normals[3 * i + 2] = nz;
    }
}

// ---------------------------------------------------------------------------
// "Next" row (second half): what GaussianModel.from_mixture derives from a level's covariances
// (src/models/gaussian_model.py:141-153 -> decompose_covariance_matrix :242-265 -> matrices_to_quaternions,
// src/utils/general_utils.py:94-100), one thread per component:
//   * eigen-decomposition of the symmetric 3x3 covariance (torch.linalg.eigh there; cyclic Jacobi in FP64 here),
//     eigenvalues ascending, eigenvectors as columns, sign fixed so that the largest-magnitude component of every
//     eigenvector is positive (eigh leaves the sign to the LAPACK / cuSOLVER build: the reference's own output is
//     defined only up to these signs);
//   * the reference's axis-correspondence reorder AS WRITTEN: corr[i] = argmax_j |V[j][i]| (the axis eigenvector i
//     points along, first maximum), sorted_eigenvalues[corr[i]] = eigenvalues[i], and ROW i of V goes to row corr[i]
//     (the reference scatters rows of the eigenvector matrix, not columns); a slot no eigenvector maps to stays 0,
//     of two eigenvectors mapping to the same axis the later one wins;
//   * `scaling` = the reordered eigenvalues (variances, not standard deviations — as in the reference);
//   * `rotation` = (w, x, y, z) with w = sqrt(1 + trace)/2, x = (R21 - R12)/(4w), ... evaluated in FP32.
// ---------------------------------------------------------------------------
__global__ void k_from_mixture(const float4* __restrict__ geo, int64_t n, float* __restrict__ scaling,
                               float* __restrict__ rotation, float* __restrict__ evals, float* __restrict__ evecs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q1 = geo[4 * i + 1], q2 = geo[4 * i + 2];
    double a[3][3] = { { q1.x, q1.y, q1.z }, { q1.y, q1.w, q2.x }, { q1.z, q2.x, q2.y } };
    double v[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    for (int sweep = 0; sweep < 8; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (!(off > 0.0)) break;                         // diagonal (or NaN): done
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {                // A <- A J
                const double akp = a[k][p], akq = a[k][q];
                a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {                // A <- J^T A
                const double apk = a[p][k], aqk = a[q][k];
                a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {                // V <- V J
                const double vkp = v[k][p], vkq = v[k][q];
                v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq;
            }
        }
    }
    // ascending eigenvalues, columns permuted alongside
    double lam[3] = { a[0][0], a[1][1], a[2][2] };
    int ord[3] = { 0, 1, 2 };
#define HEM_SW(x, y) { if (lam[ord[y]] < lam[ord[x]]) { const int t_ = ord[x]; ord[x] = ord[y]; ord[y] = t_; } }
    HEM_SW(0, 1) HEM_SW(1, 2) HEM_SW(0, 1)
#undef HEM_SW
    float V[3][3], L[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int o = ord[c];
        L[c] = (float)lam[o];
        double x = v[0][o], y = v[1][o], z = v[2][o];
        const double ax = fabs(x), ay = fabs(y), az = fabs(z);
        const double lead = (ax >= ay && ax >= az) ? x : (ay >= az ? y : z);
        if (lead < 0.0) { x = -x; y = -y; z = -z; }
        V[0][c] = (float)x; V[1][c] = (float)y; V[2][c] = (float)z;
    }
    if (evals) { evals[3 * i] = L[0]; evals[3 * i + 1] = L[1]; evals[3 * i + 2] = L[2]; }
    if (evecs)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) evecs[9 * i + 3 * r + c] = V[r][c];
    // the reference's reorder (gaussian_model.py:252-263)
    float SL[3] = { 0.0f, 0.0f, 0.0f }, R[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        int corr = 0;
        float best = fabsf(V[0][e]);
        if (fabsf(V[1][e]) > best) { best = fabsf(V[1][e]); corr = 1; }
        if (fabsf(V[2][e]) > best) { corr = 2; }
        SL[corr] = L[e];
        R[corr][0] = V[e][0]; R[corr][1] = V[e][1]; R[corr][2] = V[e][2];
    }
    if (scaling) { scaling[3 * i] = SL[0]; scaling[3 * i + 1] = SL[1]; scaling[3 * i + 2] = SL[2]; }
    if (rotation) {
        const float w = sqrtf(1.0f + (R[0][0] + R[1][1] + R[2][2])) / 2.0f;
        rotation[4 * i] = w;
        rotation[4 * i + 1] = (R[2][1] - R[1][2]) / (4.0f * w);
        rotation[4 * i + 2] = (R[0][2] - R[2][0]) / (4.0f * w);
        rotation[4 * i + 3] = (R[1][0] - R[0][1]) / (4.0f * w);
    }
}

// ---------------------------------------------------------------------------
// S1: parent radii + statistics.  mixture.cpp:70-95
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prep(const float4* __restrict__ geo, int64_t n, float distance_delta, float* __restrict__ radius,
       uint32_t* __restrict__ pflag, uint8_t* __restrict__ psafe, Ctl* __restrict__ ctl)
{
    pdl_enter();
    // * The FP64 closed-form eigenvalue is ~400 instructions and only one component in
    //   hemReduction is a parent: the parents of a 256-component tile are compacted through
    //   shared memory so that the eigen code runs on full warps.
    // * Grid-stride over tiles with the statistics kept in registers: one set of atomics per
    //   BLOCK at the end (per-warp atomics on nine shared addresses serialised the whole kernel).
    __shared__ float s_cov[256][6];
    __shared__ uint16_t s_src[256];
    __shared__ float s_rad[256];
    __shared__ uint8_t s_safe[256];
    __shared__ int s_woff[9];
    __shared__ int s_red[8][8];
    __shared__ unsigned long long s_redd[8];
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int bmax[3] = { 0u, 0u, 0u }, bminc[3] = { 0u, 0u, 0u };
    unsigned int rbits = 0, rcount = 0, pcount = 0, ccount = 0, ucount = 0;
    unsigned long long rsum = 0ull;
    const int64_t ntiles = (n + 255) / 256;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i = tile * 256 + tid;
        bool flag = false;
        float4 q1 = make_float4(0, 0, 0, 0), q2 = q1;
        if (i < n) {
            const float4 q0 = geo[4 * i], q3 = geo[4 * i + 3];
            flag = (__float_as_uint(q3.w) & 1u) != 0;
            ucount += (__float_as_uint(q3.w) & 2u) ? 0u : 1u;
            pflag[i] = flag ? 1u : 0u;
            if (q0.x == q0.x) { const unsigned o = f2u(q0.x); bmax[0] = max(bmax[0], o); bminc[0] = max(bminc[0], ~o); }
            if (q0.y == q0.y) { const unsigned o = f2u(q0.y); bmax[1] = max(bmax[1], o); bminc[1] = max(bminc[1], ~o); }
            if (q0.z == q0.z) { const unsigned o = f2u(q0.z); bmax[2] = max(bmax[2], o); bminc[2] = max(bminc[2], ~o); }
            if (flag) { q1 = geo[4 * i + 1]; q2 = geo[4 * i + 2]; }
        }
        s_rad[tid] = 0.0f;
        s_safe[tid] = 0;
        const unsigned bal = __ballot_sync(full, flag);
        if (lane == 0) s_woff[warp + 1] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            s_woff[0] = 0;
            for (int w = 0; w < 8; ++w) s_woff[w + 1] += s_woff[w];
        }
        __syncthreads();
        if (flag) {
            const int slot = s_woff[warp] + __popc(bal & ((1u << lane) - 1u));
            s_cov[slot][0] = q1.x; s_cov[slot][1] = q1.y; s_cov[slot][2] = q1.z;
            s_cov[slot][3] = q1.w; s_cov[slot][4] = q2.x; s_cov[slot][5] = q2.y;
            s_src[slot] = (uint16_t)tid;
        }
        __syncthreads();
        const int npar = s_woff[8];
        if (tid == 0) pcount += (unsigned)npar;
        if (tid < npar) {
            const Sym3 c = { s_cov[tid][0], s_cov[tid][1], s_cov[tid][2], s_cov[tid][3], s_cov[tid][4], s_cov[tid][5] };
            double emin, emax;
            const float rad = fm(distance_delta, __fsqrt_rn(eig_max_sym3(c, &emin, &emax)));       // mixture.cpp:88
            if (rad > 0.0f && rad < FLT_MAX) { rbits = max(rbits, __float_as_uint(rad)); rcount += 1; rsum += radius_fx(rad); }
            s_rad[s_src[tid]] = rad;
            const bool cert = parent_cov_certified(c, det_sym3(c), emin, emax);
            s_safe[s_src[tid]] = cert ? 1 : 0;
            ccount += cert ? 1u : 0u;
        }
        __syncthreads();
        if (i < n) { radius[i] = s_rad[tid]; psafe[i] = s_safe[tid]; }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        bmax[a] = __reduce_max_sync(full, bmax[a]);
        bminc[a] = __reduce_max_sync(full, bminc[a]);
    }
    rbits = __reduce_max_sync(full, rbits);
    rcount = __reduce_add_sync(full, rcount);
    ccount = __reduce_add_sync(full, ccount);
    if (lane == 0 && ccount) atomicAdd(&ctl->n_certified, ccount);
    ucount = __reduce_add_sync(full, ucount);
    if (lane == 0 && ucount) atomicAdd(&ctl->n_uncert, ucount);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(full, rsum, o);
    if (lane == 0) {
        for (int a = 0; a < 3; ++a) { s_red[warp][a] = (int)bminc[a]; s_red[warp][3 + a] = (int)bmax[a]; }
        s_red[warp][6] = (int)rbits; s_red[warp][7] = (int)rcount;
        s_redd[warp] = rsum;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) {
            for (int a = 0; a < 3; ++a) { bminc[a] = max(bminc[a], (unsigned)s_red[w][a]); bmax[a] = max(bmax[a], (unsigned)s_red[w][3 + a]); }
            rbits = max(rbits, (unsigned)s_red[w][6]); rcount += (unsigned)s_red[w][7];
            rsum += s_redd[w];
        }
        for (int a = 0; a < 3; ++a) {
            if (bmax[a]) atomicMax(&ctl->bbmax_u[a], bmax[a]);
            if (bminc[a]) atomicMax(&ctl->bbmin_c[a], bminc[a]);
        }
        if (pcount) atomicAdd(&ctl->n_parents, pcount);
        if (rcount) {
            atomicMax(&ctl->max_radius_bits, rbits);
            atomicAdd(&ctl->n_radius, rcount);
            atomicAdd(&ctl->sum_radius_fx, rsum);
        }
    }
}

// ---------------------------------------------------------------------------
// Robust grid box.  Real splat scenes carry a few far outliers (floaters); a grid over the full
// bounding box then degenerates to a handful of cells (measured: 20 outliers at 1000x the scene
// extent made the gate 600x slower).  The grid box is the [q, 1-q] percentile range of the means
// per axis, from 1024-bin histograms; everything outside is clamped into the border cells, which the
// row culling already treats as unbounded — the accepted sets do not depend on the box at all.
// ---------------------------------------------------------------------------
constexpr int BOX_BINS = 1024;

__global__ void __launch_bounds__(256)
k_boxhist(const float4* __restrict__ geo, int64_t n, float lox, float loy, float loz, float sx, float sy, float sz,
          unsigned int* __restrict__ hist)
{
    __shared__ unsigned int s_h[3 * BOX_BINS];
    for (int i = threadIdx.x; i < 3 * BOX_BINS; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 q0 = geo[4 * i];
        if (q0.x == q0.x) atomicAdd(&s_h[min(max(__float2int_rd((q0.x - lox) * sx), 0), BOX_BINS - 1)], 1u);
        if (q0.y == q0.y) atomicAdd(&s_h[BOX_BINS + min(max(__float2int_rd((q0.y - loy) * sy), 0), BOX_BINS - 1)], 1u);
        if (q0.z == q0.z) atomicAdd(&s_h[2 * BOX_BINS + min(max(__float2int_rd((q0.z - loz) * sz), 0), BOX_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * BOX_BINS; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// ---------------------------------------------------------------------------
// S2: uniform grid.  Replaces PointIndex::create (pointindex.cpp:55-108): cell
// keys -> radix sort (CUB, in hem_b200.cu) -> cell-sorted SoA planes + offsets.
// ---------------------------------------------------------------------------
__global__ void k_cellkey(const float4* __restrict__ geo, int64_t n, GridP g, uint32_t ncell, uint32_t* __restrict__ key,
                          uint32_t* __restrict__ val)
{
    pdl_enter();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q0 = geo[4 * i];
    const uint32_t flag = __float_as_uint(geo[4 * i + 3].w) & 1u;
    int cx = cell_axis(q0.x, g.ox, g.inv_h, g.gx);
    int cy = cell_axis(q0.y, g.oy, g.inv_h, g.gy);
    int cz = cell_axis(q0.z, g.oz, g.inv_h, g.gz);
    key[i] = (uint32_t)((cz * g.gy + cy) * g.gx + cx) + (flag ? ncell : 0u);      // parents behind all non-parents
    val[i] = (uint32_t)i;
}

// conservative cell box of a parent's search sphere along one axis (DESIGN.md "grid range")
__device__ __forceinline__ void cell_range(float q, float rr, float o, float inv_h, int g, int& lo, int& hi)
{
    const float pad = (fabsf(q) + rr) * 2.4e-7f;
    lo = cell_axis(q - (rr + pad), o, inv_h, g);
    hi = cell_axis(q + (rr + pad), o, inv_h, g);
}

// Per-parent record the gate starts from (7 float4 at cell-order position t), written by the thread that places the
// parent into the cell-sorted planes: the parent's four plane rows, radius, weight, rank, index, FP32 inverse
// covariance and its cell box - the sphere's, or the Mahalanobis ellipsoid's for certified parents (see `s_box`).
constexpr int PINFO_F4 = 8;
__device__ __forceinline__ void ptab_record(int t, uint32_t s, float4 p0, float4 p1, float4 p2, float4 p3, float weight,
                                            uint32_t srank, float rad, uint32_t safe, const GridP& g, float s_box,
                                            uint2* __restrict__ ptab, float4* __restrict__ pinfo)
{
    ptab[t] = make_uint2(s, srank);
    const Sym3 pc = { p2.x, p2.y, p2.z, p2.w, p3.x, p3.y };
    const Sym3 ip = inverse_sym3(pc, p3.z);
    int y0 = 0, y1 = 0, z0 = 0, z1 = 0;
    float hx = __uint_as_float(0x7f800000u);          // half-extent of the search box along x (+inf: the sphere's chord decides)
    if (rad > 0.0f) {
        const float rr = rad * 1.00001f;
        float ry = rr, rz = rr;
        // Ellipsoid box (s_box > 0: no component of the level is uncertified, and this parent is certified).  The
        // certified pre-test (hem_math.cuh) rejects every candidate whose squared Mahalanobis distance exceeds
        // 2*kthr + 0.5 in FP32; with d'P^-1 d >= d_i^2 / P_ii that is every candidate outside the axis-aligned box of
        // the ellipsoid d'P^-1 d <= s_box, s_box = (2*kthr + 0.5) / 0.99 * 1.001 (the certificate bounds the FP32 value
        // from below by 0.99 of the exact one).  Searching sphere-box INTERSECT ellipsoid-box leaves membership untouched.
        if (s_box > 0.0f && safe) {
            hx = sqrtf(s_box * p2.x) * 1.0001f;
            ry = fminf(rr, sqrtf(s_box * p2.w) * 1.0001f);
            rz = fminf(rr, sqrtf(s_box * p3.y) * 1.0001f);
        }
        cell_range(p0.y, ry, g.oy, g.inv_h, g.gy, y0, y1);
        cell_range(p0.z, rz, g.oz, g.inv_h, g.gz, z0, z1);
    }
    float4* o = pinfo + (size_t)t * PINFO_F4;
    o[0] = p0; o[1] = p1; o[2] = p2; o[3] = make_float4(p3.x, p3.y, p3.z, hx);
    o[4] = make_float4(rad, weight, __uint_as_float(srank | (safe << 31)), __uint_as_float(s));
    o[5] = make_float4(ip.e00, ip.e01, ip.e02, ip.e11);
    o[6] = make_float4(ip.e12, ip.e22, __uint_as_float((uint32_t)y0 | ((uint32_t)y1 << 16)),
                       __uint_as_float((uint32_t)z0 | ((uint32_t)z1 << 16)));
    // what every lane of the gate warp would otherwise recompute per parent: x range of the ellipsoid box (the whole
    // grid when the parent searches its sphere), 1 / rows-per-z-slice of its box, squared padded radius
    int xlo = 0, xhi = g.gx - 1;
    if (hx < 3.0e38f) cell_range(p0.x, hx, g.ox, g.inv_h, g.gx, xlo, xhi);
    const float rr = rad * 1.00001f;
    o[7] = make_float4(__int_as_float(xlo), __int_as_float(xhi), 1.0f / (float)(y1 - y0 + 1), rr * rr);
}

__global__ void k_sorted_gather(const float4* __restrict__ geo, const uint32_t* __restrict__ skey,
                                const uint32_t* __restrict__ sval, int64_t n, uint32_t ncell,
                                float4* __restrict__ sp0, float4* __restrict__ sp1, float4* __restrict__ sp2,
                                float4* __restrict__ sp3, uint32_t* __restrict__ cell_start,
                                // parents (slots [nc, n), whole warps of them but one): their gate records
                                uint32_t nc, const uint32_t* __restrict__ prank, const float* __restrict__ radius,
                                const uint8_t* __restrict__ psafe, GridP g, float s_box, uint2* __restrict__ ptab,
                                float4* __restrict__ pinfo)
{
    pdl_enter();
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t i = sval[j], key = skey[j];
    float4 q0 = geo[4 * (size_t)i], q1 = geo[4 * (size_t)i + 1], q2 = geo[4 * (size_t)i + 2], q3 = geo[4 * (size_t)i + 3];
    uint32_t flag = __float_as_uint(q3.w) & 1u;
    const float4 r0 = make_float4(q0.x, q0.y, q0.z, __uint_as_float(i | (flag << 31)));
    const float4 r1 = make_float4(q3.x, q3.y, q3.z, q2.z);
    const float4 r3 = make_float4(q2.x, q2.y, q2.w, __uint_as_float((__float_as_uint(q3.w) >> 1) & 1u));
    sp0[j] = r0;
    sp1[j] = r1;
    sp2[j] = q1;
    sp3[j] = r3;
    if (pinfo && j >= (int64_t)nc)
        ptab_record((int)(j - (int64_t)nc), i, r0, r1, q1, r3, q0.w, prank[i], radius[i], (uint32_t)psafe[i], g, s_box, ptab, pinfo);
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
#ifndef HEM_GATE_MINBLOCKS
#define HEM_GATE_MINBLOCKS 4
#endif
#ifndef HEM_GATE_TB
#define HEM_GATE_TB 8
#endif
#ifndef HEM_MOM_B
#define HEM_MOM_B 8
#endif
constexpr int GATE_WARPS = 8;
constexpr int GATE_OUTCAP = 64;         // accepted children waiting for the likelihood: flushed 32 at a time, so < 32 + 32
constexpr int GATE_TB = HEM_GATE_TB;        // parents per work ticket
// pool and chunk counters share one 64-bit word: pairs in the low 40 bits, chunks above
constexpr int POOL_BITS = 40;
// a step whose pair pool overflowed is redone by the host; until then nothing may follow the (unbounded) counts
__device__ __forceinline__ bool pool_overflowed(const unsigned long long* counters, unsigned long long pool_cap)
{
    return (counters[C_POOL] & ((1ull << POOL_BITS) - 1ull)) > pool_cap;
}

// per-parent record in cell order, built once per step (ptab_record, in k_sorted_gather) so that a gate warp starts
// from ONE contiguous 80-byte read instead of a chain of dependent gathers
//   [0..3] = sp0..sp3 of the parent's slot,
//   [4] = (radius, weight, bits(canonical rank | certified << 31), bits(canonical index))
//   [5] = (ip00, ip01, ip02, ip11)   the FP32 inverse covariance of the KLD gate, computed once per
//   [6] = (ip12, ip22, bits(y0 | y1 << 16), bits(z0 | z1 << 16))   parent by a THREAD instead of a warp
//   [7] = (bits(xlo), bits(xhi), 1 / (y1 - y0 + 1), (radius * 1.00001)^2)

struct GateArgs {
    const float4* sp0; const float4* sp1; const float4* sp2; const float4* sp3;
    const uint32_t* cell_start;
    const float4* pinfo;         // [np * PINFO_F4]
    uint32_t nc;                 // non-parents of the level = slot of the first parent
    GridP g;
    int t_begin, t_end;          // this rank's slab of parents (cell order)
    int tb;                      // parents per work ticket (1..GATE_TB)
    float cthr, kthr, decay2;
    float smd_rej;               // smd pre-test threshold 2*kthr + 0.5 (+inf disables it)
    int all_cert;                // no component of the level is uncertified: candidates need no certificate lookup
    uint32_t* pair_cnt_s;        // [np+1] accepted pairs of the parent at cell-order position t
    uint32_t* pair_cnt;          // [np+1] same, indexed by canonical rank (single GPU; NULL when sharded)
    uint2* pool; unsigned long long pool_cap;
    uint4* chunks; unsigned long long chunk_cap;   // (t, pool offset, count | segment tag << 8, prefix inside the parent / segment)
    unsigned long long* counters;
    // heavy parents: a parent whose search box has more than row_defer (y,z) rows of cells is not scanned by one warp;
    // k_gate<false> lists it as segments of GATE_SEG_ROWS rows (seg_info) and k_gate<true> shares the segments over the grid.
    int row_report;              // > 0: count parents above this many rows in *n_heavy_gate (tells the host to redo)
    int row_defer;               // > 0: leave such parents to k_gate<true>
    unsigned int* n_heavy_gate;  // control block
    uint32_t* seg_cnt;           // [segments] accepted pairs of a segment
    uint32_t* seg_off;           // [segments] exclusive prefix inside its parent (written by the parent's last segment)
    uint32_t* seg_done;          // [np] finished segments of a parent (self-resetting)
    uint4* seg_info;             // [seg_cap] (t, segment inside the parent, segments of the parent, its first global segment)
    uint32_t seg_cap;
    uint32_t* hv_hdr;            // [1] segments listed (may exceed seg_cap: overflow), [2] ticket counter of k_gate<true>
};

constexpr int GATE_SEG_ROWS = 64;            // rows of cells per work item of the heavy gate

// HEAVY = false: the ordinary kernel (one warp per parent, tickets of A.tb parents).
// HEAVY = true : only the parents listed in seg_info; one ticket = one segment of GATE_SEG_ROWS rows of one parent.
template <bool HEAVY>
__global__ void __launch_bounds__(GATE_WARPS * 32, HEM_GATE_MINBLOCKS)
k_gate(GateArgs A)
{
    pdl_enter();
    __shared__ float4 s_q[GATE_WARPS][64];                      // in-radius candidates: (mean, bits(slot)) - stage A needs nothing else of the row
    __shared__ uint32_t s_out[GATE_WARPS][GATE_OUTCAP];
    __shared__ uint2 s_q2[GATE_WARPS][64];
    // which parent of the ticket a queued candidate belongs to (the three queues run across the ticket's parents)
    __shared__ uint8_t s_qtag[GATE_WARPS][64], s_q2tag[GATE_WARPS][64], s_otag[GATE_WARPS][GATE_OUTCAP];
    // per parent of the ticket: (mean, weight), (colour, bits(t)), (ip00, ip01, ip02, ip11), (ip12, ip22, det, bits(rank | certified << 31))
    __shared__ float4 s_pc[GATE_WARPS][GATE_TB][4];
    __shared__ uint32_t s_tot[GATE_WARPS][GATE_TB];             // pairs of that parent already in the pool
    __shared__ uint32_t s_acc[GATE_WARPS][GATE_TB];             // pairs of that parent accepted so far (~0: left to k_gate<true>)
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* q = s_q[warp];
    uint32_t* ob = s_out[warp];
    uint8_t* otag = s_otag[warp];
    uint8_t* qtag = s_qtag[warp];
    uint8_t* q2tag = s_q2tag[warp];
    uint2* q2 = s_q2[warp];
    const float hc = A.g.h;
    const float INF = __uint_as_float(0x7f800000u);

    // persistent warps: parents are handed out in cell order through one ticket counter, so
    // neighbouring warps stream the same cell runs and a heavy parent delays nobody else
    for (;;) {
        // one ticket = A.tb (<= GATE_TB) consecutive parents: three same-address atomics per parent (ticket,
        // pool, chunk) serialised at the L2 and were 13 % of the kernel's stall samples
        int tfirst, tlast;
        uint32_t seg_id = 0, seg_local = 0, seg_first = 0, seg_n = 0;     // HEAVY only
        if (!HEAVY) {
            unsigned long long ticket = 0;
            if (lane == 0) ticket = atomicAdd(&A.counters[C_WORK], 1ull);
            tfirst = A.t_begin + (int)__shfl_sync(full, ticket, 0) * A.tb;
            if (tfirst >= A.t_end) break;
            tlast = min(tfirst + A.tb, A.t_end);
        } else {
            // one ticket = one segment of rows of one heavy parent; the parent is found by bisecting the segment bases
            uint32_t it = 0;
            if (lane == 0) it = atomicAdd(&A.hv_hdr[2], 1u);
            it = __shfl_sync(full, it, 0);
            if (it == 0u && lane == 0) A.counters[C_CHUNK] = A.hv_hdr[1];      // segments listed: the host checks it against seg_cap
            if (it >= min(A.hv_hdr[1], A.seg_cap)) break;
            const uint4 si = A.seg_info[it];
            tfirst = (int)si.x;
            tlast = tfirst + 1;
            seg_id = it;
            seg_local = si.y;
            seg_n = si.z;
            seg_first = si.w;
        }
      // The three queues (in radius -> stage A -> stage B -> likelihood) run across ALL the ticket's parents: a parent
      // has ~35 candidates in radius, ~15 after stage A and accepts ~8, so a drain per parent ran the stages on 55 % /
      // 47 % / 25 % of the lanes.  Entries carry the parent's position in the ticket and leave 32 at a time, each with
      // its parent's constants from shared memory; the pool gets one chunk record per parent present in a batch.
      int qn = 0, q2n = 0, on = 0;
      auto flush32 = [&]() {
          const int nb = min(on, 32);
          uint32_t jj = 0, tg = 0;
          if (lane < nb) { jj = ob[lane]; tg = otag[lane]; }
          const uint32_t tg_prev = __shfl_up_sync(full, tg, 1);
          const bool head = lane < nb && (lane == 0 || tg != tg_prev);
          const unsigned hb = __ballot_sync(full, head);
          unsigned long long base = 0, cid = 0;
          if (lane == 0) {
              const unsigned long long v = atomicAdd(&A.counters[C_POOL], ((unsigned long long)__popc(hb) << POOL_BITS) | (unsigned long long)nb);
              base = v & ((1ull << POOL_BITS) - 1ull);
              cid = v >> POOL_BITS;
          }
          base = __shfl_sync(full, base, 0);
          cid = __shfl_sync(full, cid, 0);
          const bool room = base + (unsigned long long)nb <= A.pool_cap;
          if (lane < nb) {
              const float4 q0 = s_pc[warp][tg][0], q1 = s_pc[warp][tg][1];
              const float4 c0 = A.sp0[jj], c1 = A.sp1[jj];
              const float cdet = A.sp3[jj].z;
              // hemLikelihoodOpacity, mixture.cpp:54-64; clamp + parent weight, :155-156
              const float dd = __fsqrt_rn(sqdist3(q0.x, q0.y, q0.z, c0.x, c0.y, c0.z));
              const float cdist = __fsqrt_rn(sqdist3(q1.x, q1.y, q1.z, c1.x, c1.y, c1.z));
              const float distWeight = expf(fd(fm(-dd, dd), A.decay2));
              const float colorInfluence = expf(fd(fm(-cdist, cdist), A.decay2));
              const float L = fm(fm(fm(distWeight, c1.w), colorInfluence), __fsqrt_rn(cdet));
              const float wl = fm(q0.w, clamp_ref(L, FLT_MIN, 1e8f));
              if (room) A.pool[base + lane] = make_uint2(jj, __float_as_uint(wl));      // the child's SLOT; k_chunk_copy turns it into its row
              if (head) {
                  const unsigned later = hb & ~((2u << lane) - 1u);
                  const uint32_t cnt = (uint32_t)((later ? __ffs(later) - 1 : nb) - lane);
                  const unsigned long long id = cid + (unsigned long long)__popc(hb & ((1u << lane) - 1u));
                  const uint32_t pre = s_tot[warp][tg];
                  if (id < A.chunk_cap)
                      A.chunks[id] = make_uint4(__float_as_uint(q1.w), (uint32_t)base + (uint32_t)lane, cnt | (HEAVY ? (seg_id + 1u) << 8 : 0u), pre);
                  s_tot[warp][tg] = pre + cnt;
              }
          }
          const uint32_t mv = (lane + 32 < on) ? ob[lane + 32] : 0u;
          const uint8_t mt = (lane + 32 < on) ? otag[lane + 32] : (uint8_t)0;
          __syncwarp();
          ob[lane] = mv; otag[lane] = mt;
          on -= nb;
          __syncwarp();
      };
      // stage B: candidates that survived the smd pre-test and the colour gate: trace, log-det, KLD gate
      auto stage_b = [&](int m) {
          bool acc = false;
          uint32_t jj = 0, tg = 0;
          if (lane < m) {
              const uint2 e = q2[lane];
              tg = q2tag[lane];
              jj = e.x;
              const float smd = __uint_as_float(e.y);
              const float4 i0 = s_pc[warp][tg][2], i1 = s_pc[warp][tg][3];       // (ip00, ip01, ip02, ip11), (ip12, ip22, det, .)
              const float4 c2 = A.sp2[jj], c3 = A.sp3[jj];
              const float t00 = fa(fa(fm(c2.x, i0.x), fm(c2.y, i0.y)), fm(c2.z, i0.z));
              const float t11 = fa(fa(fm(c2.y, i0.y), fm(c2.w, i0.w)), fm(c3.x, i1.x));
              const float t22 = fa(fa(fm(c2.z, i0.z), fm(c3.x, i1.x)), fm(c3.y, i1.y));
              const float tr = fa(fa(t00, t11), t22);
              // KLD(child || parent), gaussian.hpp:106-109
              const float kl = fm(0.5f, fs(fs(fa(smd, tr), 3.0f), logf_glibc(fd(c3.z, i1.z))));
              acc = !(kl > A.kthr);
          }
          const unsigned am = __ballot_sync(full, acc);
          if (acc) {
              const int pos = on + __popc(am & ((1u << lane) - 1u));
              ob[pos] = jj; otag[pos] = (uint8_t)tg;
          }
          // accepted pairs per parent: the first lane of each parent's run counts its run
          const uint32_t tg_prev = __shfl_up_sync(full, tg, 1);
          const bool head = lane < m && (lane == 0 || tg != tg_prev);
          const unsigned hb = __ballot_sync(full, head);
          if (head) {
              const unsigned later = hb & ~((2u << lane) - 1u);
              const int end = later ? __ffs(later) - 1 : m;
              const unsigned span = (end >= 32 ? 0xffffffffu : ((1u << end) - 1u)) & ~((1u << lane) - 1u);
              s_acc[warp][tg] += (uint32_t)__popc(am & span);
          }
          on += __popc(am);
          __syncwarp();
          if (on >= 32) flush32();
      };
      // stage A: candidates that passed the radius + parent-exclusion test: squared Mahalanobis distance with the
      // certified pre-test (hem_math.cuh), colour gate; survivors queue for stage B
      auto process = [&](int m) {
          bool keep = false;
          uint32_t jj = 0, tg = 0;
          float smd = 0.0f;
          if (lane < m) {
              const float4 c0 = q[lane];
              jj = __float_as_uint(c0.w);
              tg = qtag[lane];
              const float4 q0 = s_pc[warp][tg][0], i0 = s_pc[warp][tg][2], i1 = s_pc[warp][tg][3];
              // Both gates must pass, so their order is free: the Mahalanobis pre-test goes first (it rejects more than
              // half of what reaches this stage and needs nothing but the mean), the colour row is fetched for the rest.
              const float dx = fs(c0.x, q0.x), dy = fs(c0.y, q0.y), dz = fs(c0.z, q0.z);
              const float tx = fa(fa(fm(i0.x, dx), fm(i0.y, dy)), fm(i0.z, dz));
              const float ty = fa(fa(fm(i0.y, dx), fm(i0.w, dy)), fm(i1.x, dz));
              const float tz = fa(fa(fm(i0.z, dx), fm(i1.x, dy)), fm(i1.y, dz));
              smd = fa(fa(fm(dx, tx), fm(dy, ty)), fm(dz, tz));
              bool rej = false;
              if ((__float_as_uint(i1.w) >> 31) != 0u && smd > A.smd_rej)          // certified pre-test (hem_math.cuh): needs a certified candidate as well
                  rej = A.all_cert || __float_as_uint(A.sp3[jj].w) != 0u;
              if (!rej) {
                  // colour gate, mixture.cpp:121-124 / gaussian.hpp:111-114 (NaN passes)
                  const float4 c1 = A.sp1[jj];
                  const float4 q1 = s_pc[warp][tg][1];
                  const float cdist = __fsqrt_rn(sqdist3(c1.x, c1.y, c1.z, q1.x, q1.y, q1.z));
                  keep = !(cdist > A.cthr);
              }
          }
          const unsigned km = __ballot_sync(full, keep);
          if (keep) {
              const int pos = q2n + __popc(km & ((1u << lane) - 1u));
              q2[pos] = make_uint2(jj, __float_as_uint(smd)); q2tag[pos] = (uint8_t)tg;
          }
          q2n += __popc(km);
          __syncwarp();
          if (q2n >= 32) {
              stage_b(32);
              const uint2 mv = (lane + 32 < q2n) ? q2[lane + 32] : make_uint2(0u, 0u);
              const uint8_t mt = (lane + 32 < q2n) ? q2tag[lane + 32] : (uint8_t)0;
              __syncwarp();
              q2[lane] = mv; q2tag[lane] = mt;
              q2n -= 32;
              __syncwarp();
          }
      };
      // a full batch of in-radius candidates goes through stage A
      auto q_drain32 = [&]() {
          process(32);
          const float4 mv = (lane + 32 < qn) ? q[lane + 32] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          const uint8_t mt = (lane + 32 < qn) ? qtag[lane + 32] : (uint8_t)0;
          __syncwarp();
          q[lane] = mv; qtag[lane] = mt;
          qn -= 32;
          __syncwarp();
      };
      for (int t = tfirst; t < tlast; ++t) {

        const float4* pi = A.pinfo + (size_t)t * PINFO_F4;
        const float4 p0 = pi[0], p1 = pi[1], p3 = pi[3], p4 = pi[4], p5 = pi[5], p6 = pi[6], p7 = pi[7];
        const float rad = p4.x;
        const float r2 = fm(rad, rad);

        const uint32_t tslot = (uint32_t)(t - tfirst);
        if (lane == 0) {
            s_pc[warp][tslot][0] = make_float4(p0.x, p0.y, p0.z, p4.y);
            s_pc[warp][tslot][1] = make_float4(p1.x, p1.y, p1.z, __uint_as_float((uint32_t)t));
            s_pc[warp][tslot][2] = p5;
            s_pc[warp][tslot][3] = make_float4(p6.x, p6.y, p3.z, p4.z);
            s_tot[warp][tslot] = 0u;
            s_acc[warp][tslot] = 0u;
        }
        __syncwarp();

        if (rad > 0.0f) {        // NaN / zero radius: no child can pass `d2 < r*r` (pointindex.cpp:138)
            // conservative cell box (DESIGN.md "grid range"): the float test can only pass for
            // |d| <= rad*(1+5*2^-24) per axis; pads cover the rounding of the range ends themselves
            const float rr2 = p7.w;
            const int y0 = (int)(__float_as_uint(p6.z) & 0xffffu), y1 = (int)(__float_as_uint(p6.z) >> 16);
            const int z0 = (int)(__float_as_uint(p6.w) & 0xffffu), z1 = (int)(__float_as_uint(p6.w) >> 16);
            const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
            const int nrows = ny * nz;
            const float inv_ny = p7.z;
            // x range of the ellipsoid box (ptab_record; the whole grid when the parent searches its sphere)
            const int xlo = __float_as_int(p7.x), xhi = __float_as_int(p7.y);
            if (!HEAVY) {
                if (A.row_report > 0 && nrows > A.row_report && lane == 0) atomicAdd(A.n_heavy_gate, 1u);
                if (A.row_defer > 0 && nrows > A.row_defer) {          // uniform: left to k_gate<true>
                    // the parent's segments take consecutive ids; which ids depends on scheduling, the results do not
                    // (a segment's place in the parent's CSR range follows from its index inside the parent)
                    const uint32_t ns = (uint32_t)((nrows + GATE_SEG_ROWS - 1) / GATE_SEG_ROWS);
                    uint32_t sb = 0;
                    if (lane == 0) sb = atomicAdd(&A.hv_hdr[1], ns);
                    sb = __shfl_sync(full, sb, 0);
                    if (sb + ns <= A.seg_cap)
                        for (uint32_t k = lane; k < ns; k += 32) A.seg_info[sb + k] = make_uint4((uint32_t)t, k, ns, sb);
                    if (lane == 0) s_acc[warp][tslot] = 0xffffffffu;     // its pair count is k_gate<true>'s to write
                    __syncwarp();
                    continue;
                }
            }
            // The parent itself is its own first candidate (the grid holds non-parents only): the reference's radius
            // test sees distance 0 < r*r, the parent-exclusion test lets the parent through (mixture.cpp:131-133), and
            // the colour / KLD gates are evaluated on it like on any other candidate.
            // (the test is spelled out on the parent's own coordinates: a NaN / infinite mean fails it, as in the reference)
            if (sqdist3(p0.x, p0.y, p0.z, p0.x, p0.y, p0.z) < r2 && (!HEAVY || seg_local == 0u)) {
                if (lane == 0) { q[qn] = make_float4(p0.x, p0.y, p0.z, __uint_as_float(A.nc + (uint32_t)t)); qtag[qn] = (uint8_t)tslot; }
                qn += 1;
                __syncwarp();
                if (qn >= 32) q_drain32();
            }
            const int rb_begin = HEAVY ? (int)seg_local * GATE_SEG_ROWS : 0;
            const int rb_end = HEAVY ? min(nrows, rb_begin + GATE_SEG_ROWS) : nrows;

            for (int rb = rb_begin; rb < rb_end; rb += 32) {
                // one (y,z) row of cells per lane: cull rows the sphere misses, trim the x range of
                // the others to the chord, fetch the [begin,end) run of slots, then scan the lengths
                uint32_t ra = 0, len = 0;
                const int rrow = rb + lane;
                if (rrow < rb_end) {
                    // rrow = zq * ny + yr without the integer division: float estimate, corrected by one
                    int zq = (nrows < (1 << 22)) ? (int)(((float)rrow + 0.5f) * inv_ny) : rrow / ny;
                    int yr = rrow - zq * ny;
                    if (yr < 0) { --zq; yr += ny; } else if (yr >= ny) { ++zq; yr -= ny; }
                    const int y = y0 + yr, z = z0 + zq;
                    float lo = A.g.oy + (float)y * hc, hi = lo + hc;
                    float pad = (fabsf(A.g.oy) + fabsf(hi) + hc) * 4e-6f;
                    lo = (y == 0) ? -INF : lo - pad;
                    hi = (y == A.g.gy - 1) ? INF : hi + pad;
                    const float dy = fmaxf(0.0f, fmaxf(lo - p0.y, p0.y - hi));
                    lo = A.g.oz + (float)z * hc; hi = lo + hc;
                    pad = (fabsf(A.g.oz) + fabsf(hi) + hc) * 4e-6f;
                    lo = (z == 0) ? -INF : lo - pad;
                    hi = (z == A.g.gz - 1) ? INF : hi + pad;
                    const float dz = fmaxf(0.0f, fmaxf(lo - p0.z, p0.z - hi));
                    const float d2yz = dy * dy + dz * dz;
                    if (d2yz <= rr2) {
                        const float dxm = sqrtf(fmaxf(rr2 - d2yz, 0.0f)) * 1.00001f;
                        const float padx = (fabsf(p0.x) + dxm) * 2.4e-7f;
                        const int xa = max(cell_axis(p0.x - (dxm + padx), A.g.ox, A.g.inv_h, A.g.gx), xlo);
                        const int xb = min(cell_axis(p0.x + (dxm + padx), A.g.ox, A.g.inv_h, A.g.gx), xhi);
                        if (xa <= xb) {
                            const uint32_t base = (uint32_t)((z * A.g.gy + y) * A.g.gx);
                            ra = A.cell_start[base + xa];
                            len = A.cell_start[base + xb + 1] - ra;
                        }
                    }
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(full, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t total = __shfl_sync(full, incl, 31);
                const uint32_t excl = incl - len;

                // lanes walk the concatenation of the runs, 32 candidates per iteration; the slot and
                // the position of the NEXT window are fetched before the current one is tested
                auto locate = [&](uint32_t tt) -> uint32_t {
                    int row = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t v = __shfl_sync(full, incl, row + step - 1);
                        if (v <= tt) row += step;
                    }
                    row = min(row, 31);
                    return __shfl_sync(full, ra, row) + (tt - __shfl_sync(full, excl, row));
                };
                uint32_t jn = locate((uint32_t)lane);
                float4 cn = make_float4(0, 0, 0, 0);
                if ((uint32_t)lane < total) cn = A.sp0[jn];
                for (uint32_t t0 = 0; t0 < total; t0 += 32) {
                    const uint32_t tt = t0 + lane;
                    const uint32_t jj = jn;
                    const float4 c0 = cn;
                    if (t0 + 32 < total) {                      // uniform
                        jn = locate(tt + 32);
                        if (tt + 32 < total) cn = A.sp0[jn];
                    }
                    bool in = false;
                    if (tt < total) {
                        // radius test, pointindex.cpp:138 (query - point); the scanned slots hold non-parents only, so the
                        // parent exclusion of mixture.cpp:131-133 has nothing to reject here
                        const float d2 = sqdist3(p0.x, p0.y, p0.z, c0.x, c0.y, c0.z);
                        in = d2 < r2;
                    }
                    const unsigned m = __ballot_sync(full, in);
                    if (in) {
                        const int pos = qn + __popc(m & ((1u << lane) - 1u));
                        q[pos] = make_float4(c0.x, c0.y, c0.z, __uint_as_float(jj)); qtag[pos] = (uint8_t)tslot;
                    }
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) q_drain32();
                }
            }
        }
      }
      // end of the ticket: what its parents left in the queues
      if (qn > 0) { process(qn); qn = 0; }
      if (q2n > 0) { stage_b(q2n); q2n = 0; }
      if (on > 0) flush32();        // (< 32)
      __syncwarp();
      {
        const int t = tfirst;
        const uint32_t total_out = s_acc[warp][0];
        const uint32_t rank0 = __float_as_uint(s_pc[warp][0][3].w) & 0x7fffffffu;
        if (!HEAVY) {
            if (lane < tlast - tfirst) {
                const uint32_t cnt = s_acc[warp][lane];
                if (cnt != 0xffffffffu) {
                    A.pair_cnt_s[tfirst + lane] = cnt;
                    if (A.pair_cnt) A.pair_cnt[__float_as_uint(s_pc[warp][lane][3].w) & 0x7fffffffu] = cnt;
                }
            }
        } else {
            // the segment's count; the warp that finishes the parent's last segment turns the counts into offsets
            // (segment order = row order, so the parent's CSR order does not depend on scheduling)
            uint32_t prev = 0;
            if (lane == 0) {
                A.seg_cnt[seg_id] = total_out;
                __threadfence();
                prev = atomicAdd(&A.seg_done[t], 1u);
            }
            prev = __shfl_sync(full, prev, 0);
            if (prev == seg_n - 1) {
                __threadfence();
                uint32_t run = 0;
                for (uint32_t b = 0; b < seg_n; b += 32) {
                    const uint32_t k = b + lane;
                    const uint32_t c = k < seg_n ? __ldcg(&A.seg_cnt[seg_first + k]) : 0u;
                    uint32_t incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(full, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (k < seg_n) A.seg_off[seg_first + k] = run + incl - c;
                    run += __shfl_sync(full, incl, 31);
                }
                if (lane == 0) {
                    A.pair_cnt_s[t] = run;
                    if (A.pair_cnt) A.pair_cnt[rank0] = run;
                    A.seg_done[t] = 0u;
                }
            }
        }
        __syncwarp();
      }
    }
}

// pool -> cell-order CSR (what the moment kernel streams; a rank's slab of parents owns one
// contiguous segment of it, which is what the ranks exchange).  On one GPU the same pass also files
// every pair in its child's claim table (see k_claims below), saving a second sweep over the pairs.
constexpr int CLAIM_K1 = 4;      // first tier: 4 claims = one 32-byte sector per child, consecutive children consecutive sectors
constexpr int CLAIM_K = 16;      // second tier (children claimed by more than 4 parents), then the global overflow list

// file one accepted pair (parent rank, wL bits) under its child; slot order is arbitrary, k_sumlw_claims sorts by rank
__device__ __forceinline__ void claim_put(uint32_t c, uint32_t srank, uint32_t wl, uint32_t* __restrict__ cnt_c,
                                          uint2* __restrict__ claims, uint2* __restrict__ claims2, uint4* __restrict__ ovf,
                                          unsigned long long ovf_cap, unsigned long long* __restrict__ counters)
{
    const uint32_t slot = atomicAdd(&cnt_c[c], 1u);
    if (slot < (uint32_t)CLAIM_K1) {
        claims[(size_t)c * CLAIM_K1 + slot] = make_uint2(srank, wl);
    } else if (slot < (uint32_t)(CLAIM_K1 + CLAIM_K)) {
        claims2[(size_t)c * CLAIM_K + (slot - CLAIM_K1)] = make_uint2(srank, wl);
    } else {
        const unsigned long long o = atomicAdd(&counters[C_OVF], 1ull);
        if (o < ovf_cap) ovf[o] = make_uint4(c, srank, wl, 0u);
    }
}

__global__ void __launch_bounds__(256)
k_chunk_copy(const uint4* __restrict__ chunks, unsigned long long* __restrict__ counters,
             unsigned long long chunk_cap, unsigned long long pool_cap, const uint2* __restrict__ pool,
             const uint32_t* __restrict__ pair_off_s, long long dst_shift, uint2* __restrict__ csr_s,
             const float4* __restrict__ sp0, const uint2* __restrict__ ptab, uint32_t* __restrict__ cnt_c,
             uint2* __restrict__ claims, uint2* __restrict__ claims2, uint4* __restrict__ ovf,
             unsigned long long ovf_cap, const uint32_t* __restrict__ seg_off)
{
    pdl_enter();
    if (pool_overflowed(counters, pool_cap) || (counters[C_POOL] >> POOL_BITS) > chunk_cap) return;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // One lane per POOL ENTRY (a warp per chunk ran at a quarter of its lanes: a chunk holds ~8 pairs).  Chunk ids and pool
    // offsets come from the same packed atomic, so the chunk records are sorted by pool offset.  A warp takes a run of
    // CC_RUN consecutive 32-entry windows: it finds the chunk of the run's first entry by a 32-ary search (all lanes probe:
    // 4 dependent rounds for a million chunks, where a bisection took 20), the chunks that start inside a window mark their
    // start lanes, every lane counts the marks up to itself, and the chunk of the next window's first entry follows from the
    // number of marks - its records are in flight while this window is copied.  Runs are dealt round-robin so that the
    // whole grid sweeps the pool as one front.  The pool holds child SLOTS (cell order): the claim table is indexed by
    // slot, so that the claims of one child and of its neighbours - filed by parents a few cell rows apart - meet in L2
    // instead of each costing a 32-byte DRAM read-modify-write; the CSR gets the child's row index (sp0[slot].w).
    // The counts are read on the device; the launch is sized by the host's upper bound and capped.
    constexpr unsigned long long CC_RUN = 4;
    const unsigned long long packed = counters[C_POOL];
    const uint32_t nchunks = (uint32_t)min(packed >> POOL_BITS, chunk_cap);
    const unsigned long long npairs = packed & ((1ull << POOL_BITS) - 1ull);
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long nwin = (npairs + 31ull) >> 5;
    const unsigned long long nruns = (nwin + CC_RUN - 1ull) / CC_RUN;
    for (unsigned long long run = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; run < nruns; run += nwarps) {
        const unsigned long long wfirst = run * CC_RUN, wlast = min(nwin, wfirst + CC_RUN);
        uint32_t lo = 0;
        {
            const uint32_t e0 = (uint32_t)(wfirst * 32ull);
            uint32_t hi = nchunks;                              // chunks[lo].y <= e0 < chunks[hi].y (hi == nchunks: +inf)
            while (hi - lo > 1) {
                const uint32_t step = (hi - lo + 32u) / 33u;    // probes at lo + step, lo + 2 step, ..., lo + 32 step
                const uint32_t idx = lo + (uint32_t)(lane + 1) * step;
                const bool le = idx < hi && chunks[idx].y <= e0;
                const uint32_t cnt = (uint32_t)__popc(__ballot_sync(full, le));       // sorted: the lanes that pass are a prefix
                lo += cnt * step;
                hi = min(hi, lo + step);
            }
        }
        uint4 mine = make_uint4(0u, 0u, 0u, 0u);
        if (lo + (uint32_t)lane < nchunks) mine = chunks[lo + lane];
        for (unsigned long long w = wfirst; w < wlast; ++w) {
            const uint32_t e0 = (uint32_t)(w * 32ull);
            const bool inwin = lane > 0 && lo + (uint32_t)lane < nchunks && mine.y - e0 < 32u;      // starts inside the window (lane 0 = the chunk of e0)
            const unsigned marks = __reduce_or_sync(full, inwin ? (1u << (mine.y - e0)) : 0u);
            const int src = __popc(marks & ((2u << lane) - 1u));            // chunk of entry e0 + lane, relative to lo (< 32: each chunk has >= 1 pair)
            uint4 ch;
            ch.x = __shfl_sync(full, mine.x, src); ch.y = __shfl_sync(full, mine.y, src);
            ch.z = __shfl_sync(full, mine.z, src); ch.w = __shfl_sync(full, mine.w, src);
            if (w + 1 < wlast) {                                // uniform
                // the chunk of entry e0 + 32: the last one that starts inside this window, or its successor when that one starts at e0 + 32
                const uint32_t nst = (uint32_t)__popc(marks);
                uint32_t nlo = lo + nst;
                uint32_t ny = __shfl_sync(full, mine.y, (int)((nst + 1u) & 31u));
                if (nst + 1u >= 32u && nlo + 1u < nchunks) ny = chunks[nlo + 1u].y;
                if (nlo + 1u < nchunks && ny == e0 + 32u) ++nlo;
                lo = nlo;
                mine = make_uint4(0u, 0u, 0u, 0u);
                if (lo + (uint32_t)lane < nchunks) mine = chunks[lo + lane];
            }
            const unsigned long long e = (unsigned long long)e0 + lane;
            if (e >= npairs) continue;
            // a chunk of a heavy parent's segment carries the segment (index + 1) above the count: its prefix is relative to the segment
            if (ch.z >> 8) ch.w += seg_off[(ch.z >> 8) - 1u];
            const size_t dst = (size_t)((long long)pair_off_s[ch.x] + dst_shift) + ch.w + (size_t)(e - ch.y);    // dst_shift: offset into a rank's padded segment
            const uint2 p = pool[e];
            const uint32_t crow = __float_as_uint(reinterpret_cast<const float*>(sp0)[4 * (size_t)p.x + 3]) & 0x7fffffffu;
            csr_s[dst] = make_uint2(crow, p.y);
            if (claims) claim_put(p.x, ptab[ch.x].y, p.y, cnt_c, claims, claims2, ovf, ovf_cap, counters);
        }
    }
}

// per-parent pair counts from cell order to ascending canonical parent order
__global__ void k_cnt_canon(const uint2* __restrict__ ptab, const uint32_t* __restrict__ pair_cnt_s, int np,
                            uint32_t* __restrict__ pair_cnt)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < np) pair_cnt[ptab[t].y] = pair_cnt_s[t];
}

// cell-order CSR -> CSR in ascending canonical parent order: the input of the stable sort by
// child that fixes the sumLw summation order (mixture.cpp:143-163)
__global__ void k_csr_canon(const uint2* __restrict__ ptab, const uint32_t* __restrict__ pair_off_s,
                            const uint32_t* __restrict__ pair_off, const uint2* __restrict__ csr_s, int np,
                            uint32_t* __restrict__ csr_child, float* __restrict__ csr_wl)
{
    const int lane = threadIdx.x & 31;
    const int t = (int)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= np) return;
    const uint32_t a = pair_off_s[t], e = pair_off_s[t + 1];
    const size_t dst = pair_off[ptab[t].y];
    for (uint32_t k = a + lane; k < e; k += 32) {
        const uint2 p = csr_s[k];
        csr_child[dst + (k - a)] = p.x;
        csr_wl[dst + (k - a)] = __uint_as_float(p.y);
    }
}

// ---------------------------------------------------------------------------
// Per-child normaliser sumLw, claim-table path (sparse regime, <= 4 claims per child on
// average): every accepted pair drops (parent rank, wL) into one of CLAIM_K slots of its child
// (slot = atomic counter, order irrelevant), then one thread per child adds its claims in
// ASCENDING PARENT RANK — the reference's order (mixture.cpp:143-163) — so the result is
// bit-identical to the sort path and run-to-run deterministic.  Claims beyond CLAIM_K go to a
// global overflow list that the child scans (never reached at the benchmark densities).
// ---------------------------------------------------------------------------

__global__ void k_claims(const uint2* __restrict__ ptab, const uint32_t* __restrict__ pair_off_s,
                         const uint2* __restrict__ csr_s, int np, uint32_t* __restrict__ cnt_c,
                         uint2* __restrict__ claims, uint2* __restrict__ claims2, uint4* __restrict__ ovf,
                         unsigned long long ovf_cap, unsigned long long pool_cap, unsigned long long* __restrict__ counters,
                         uint32_t c_lo, uint32_t c_hi)      // only children in [c_lo, c_hi): this rank's share of the normaliser
{
    pdl_enter();
    if (pool_overflowed(counters, pool_cap)) return;
    const int lane = threadIdx.x & 31;
    const int t = (int)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= np) return;
    const uint32_t a = pair_off_s[t], e = pair_off_s[t + 1];
    const uint32_t srank = ptab[t].y;
    for (uint32_t k = a + lane; k < e; k += 32) {
        const uint2 p = csr_s[k];
        if (p.x < c_lo || p.x >= c_hi) continue;
        claim_put(p.x, srank, p.y, cnt_c, claims, claims2, ovf, ovf_cap, counters);
    }
}

// multi-GPU: the ranks agree on "the claim table gave up" (each rank sums its own range of children)
__global__ void k_status_put(const unsigned long long* __restrict__ counters, unsigned long long* __restrict__ slots, int rank)
{
    slots[rank] = counters[C_FAIL];
}
__global__ void k_status_or(unsigned long long* __restrict__ counters, const unsigned long long* __restrict__ slots, int world)
{
    unsigned long long v = 0ull;
    for (int r = 0; r < world; ++r) v |= slots[r];
    counters[C_FAIL] = v;
}

__global__ void __launch_bounds__(256)
k_sumlw_claims(const uint32_t* __restrict__ cnt_c, const uint2* __restrict__ claims,
               const uint2* __restrict__ claims2, const uint4* __restrict__ ovf,
               unsigned long long* __restrict__ counters, unsigned long long ovf_limit, int64_t c_lo, int64_t n,
               const float4* __restrict__ slot_sp0,      // table indexed by cell-order slot (fused chunk copy): sp0[slot].w = the child's row
               float* __restrict__ sumlw)
{
    pdl_enter();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t c = c_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // children [c_lo, n)
    const uint32_t m = c < n ? cnt_c[c] : 0u;                                     // 0: orphan, sumLw stays exactly 0
    const uint2* row = claims + (size_t)c * CLAIM_K1;
    const uint2* row2 = claims2 + (size_t)c * CLAIM_K;
    float s = 0.0f;
    bool have = false;
#define HEM_CSWAP(ka, wa, kb, wb) { if (kb < ka) { uint32_t t_ = ka; ka = kb; kb = t_; t_ = wa; wa = wb; wb = t_; } }
    if (m >= 1u && m <= 8u) {
        // the claims sit in registers (16-byte loads; slots beyond m were never written: masked), sorted by a network
        uint4 a, b = make_uint4(0u, 0u, 0u, 0u);
        if (m >= 2) a = reinterpret_cast<const uint4*>(row)[0];
        else { const uint2 e = row[0]; a = make_uint4(e.x, e.y, 0u, 0u); }
        if (m >= 4) b = reinterpret_cast<const uint4*>(row)[1];
        else if (m == 3) { const uint2 e = row[2]; b = make_uint4(e.x, e.y, 0u, 0u); }
        uint32_t k0 = a.x, k1 = m > 1 ? a.z : 0xffffffffu, k2 = m > 2 ? b.x : 0xffffffffu, k3 = m > 3 ? b.z : 0xffffffffu;
        uint32_t w0 = a.y, w1 = a.w, w2 = b.y, w3 = b.w;
        if (m <= 4u) {
            HEM_CSWAP(k0, w0, k1, w1) HEM_CSWAP(k2, w2, k3, w3) HEM_CSWAP(k0, w0, k2, w2) HEM_CSWAP(k1, w1, k3, w3) HEM_CSWAP(k1, w1, k2, w2)
            s = __uint_as_float(w0);                   // 0 + w0 == w0 exactly (w0 > 0)
            if (m > 1) s = fa(s, __uint_as_float(w1));
            if (m > 2) s = fa(s, __uint_as_float(w2));
            if (m > 3) s = fa(s, __uint_as_float(w3));
        } else {
            uint4 cc, d = make_uint4(0u, 0u, 0u, 0u);
            if (m >= 6) cc = reinterpret_cast<const uint4*>(row2)[0];
            else { const uint2 e = row2[0]; cc = make_uint4(e.x, e.y, 0u, 0u); }
            if (m == 8) d = reinterpret_cast<const uint4*>(row2)[1];
            else if (m == 7) { const uint2 e = row2[2]; d = make_uint4(e.x, e.y, 0u, 0u); }
            uint32_t k4 = cc.x, k5 = m > 5 ? cc.z : 0xffffffffu, k6 = m > 6 ? d.x : 0xffffffffu, k7 = m > 7 ? d.z : 0xffffffffu;
            uint32_t w4 = cc.y, w5 = cc.w, w6 = d.y, w7 = d.w;
            HEM_CSWAP(k0, w0, k1, w1) HEM_CSWAP(k2, w2, k3, w3) HEM_CSWAP(k4, w4, k5, w5) HEM_CSWAP(k6, w6, k7, w7)
            HEM_CSWAP(k0, w0, k2, w2) HEM_CSWAP(k1, w1, k3, w3) HEM_CSWAP(k4, w4, k6, w6) HEM_CSWAP(k5, w5, k7, w7)
            HEM_CSWAP(k1, w1, k2, w2) HEM_CSWAP(k5, w5, k6, w6) HEM_CSWAP(k0, w0, k4, w4) HEM_CSWAP(k3, w3, k7, w7)
            HEM_CSWAP(k1, w1, k5, w5) HEM_CSWAP(k2, w2, k6, w6)
            HEM_CSWAP(k1, w1, k4, w4) HEM_CSWAP(k3, w3, k6, w6)
            HEM_CSWAP(k2, w2, k4, w4) HEM_CSWAP(k3, w3, k5, w5)
            HEM_CSWAP(k3, w3, k4, w4)
            s = fa(fa(fa(fa(__uint_as_float(w0), __uint_as_float(w1)), __uint_as_float(w2)), __uint_as_float(w3)), __uint_as_float(w4));
            if (m > 5) s = fa(s, __uint_as_float(w5));
            if (m > 6) s = fa(s, __uint_as_float(w6));
            if (m > 7) s = fa(s, __uint_as_float(w7));
        }
        have = true;
    }
#undef HEM_CSWAP
    // 9 .. 20 claims (both tiers): the warp ranks the child's claims together, then adds them in rank order
    unsigned big = __ballot_sync(full, m > 8u && m <= (uint32_t)(CLAIM_K1 + CLAIM_K));
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int mm = (int)__shfl_sync(full, m, src);
        const size_t cc = (size_t)(c - lane + src);
        uint32_t key = 0xffffffffu, wv = 0u;
        if (lane < mm) {
            const uint2 e = lane < CLAIM_K1 ? claims[cc * CLAIM_K1 + lane] : claims2[cc * CLAIM_K + (lane - CLAIM_K1)];
            key = e.x; wv = e.y;
        }
        int rank = 0;
        for (int j = 0; j < mm; ++j) rank += __shfl_sync(full, key, j) < key ? 1 : 0;      // parent ranks are distinct
        float ss = 0.0f;
        for (int i = 0; i < mm; ++i) {
            const unsigned who = __ballot_sync(full, lane < mm && rank == i);
            ss = fa(ss, __uint_as_float(__shfl_sync(full, wv, __ffs(who) - 1)));
        }
        if (lane == src) { s = ss; have = true; }
    }
    if (m > (uint32_t)(CLAIM_K1 + CLAIM_K)) {
        // beyond both tiers: the rest sits in the global overflow list, which the child scans
        const unsigned long long nov = counters[C_OVF];
        // dense regime reached the claim table - too many overflow entries, or a hub child whose scan of the list would
        // be a long serial loop of ONE thread (m passes over it): give up, the host redoes the step by sorting
        if (nov > ovf_limit || (ovf_limit != ~0ull && (unsigned long long)m * (nov + (unsigned long long)(CLAIM_K1 + CLAIM_K)) > (1ull << 16))) {
            counters[C_FAIL] = 1ull;
        } else {
            long long last = -1;
            for (uint32_t i = 0; i < m; ++i) {                   // selection by ascending parent rank
                long long best = 0x7fffffffffffffffll;
                float bw = 0.0f;
                for (uint32_t k = 0; k < (uint32_t)(CLAIM_K1 + CLAIM_K); ++k) {
                    const uint2 e = k < (uint32_t)CLAIM_K1 ? row[k] : row2[k - CLAIM_K1];
                    if ((long long)e.x > last && (long long)e.x < best) { best = e.x; bw = __uint_as_float(e.y); }
                }
                for (unsigned long long k = 0; k < nov; ++k) {
                    const uint4 e = ovf[k];
                    if (e.x == (uint32_t)c && (long long)e.y > last && (long long)e.y < best) { best = e.y; bw = __uint_as_float(e.z); }
                }
                s = fa(s, bw);
                last = best;
            }
            have = true;
        }
    }
    if (have) {
        const size_t out = slot_sp0 ? (size_t)(__float_as_uint(reinterpret_cast<const float*>(slot_sp0)[4 * (size_t)c + 3]) & 0x7fffffffu) : (size_t)c;
        sumlw[out] = s;
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
// S4: fused parent moments.  mixture.cpp:168-247
//
// Latency-bound gather: every accepted pair pulls the child's 64-byte geometry row and its SH
// row (192 B at SH degree 3) from the canonical store.  Persistent warps take tickets of
// MOM_B consecutive parents in cell order, whose pairs are contiguous in the cell-order CSR:
//   phase 1 (lane = pair)  : child index, wL, sumLw[child], child weight -> w = wL/sumLw * weight
//   phase 2 (warp = row)   : 8 pairs at a time, all row loads issued first (geometry row by
//                            lanes 0..15, SH row by the whole warp, coalesced), then accumulated
//                            in list order; lane j owns element j of the geometry row
//   parent boundary        : normalise, det, flag, validity, one 64-byte + one SH row store
// ---------------------------------------------------------------------------
#ifndef HEM_MOM_U
#define HEM_MOM_U 4
#endif
#ifndef HEM_MOM_MINB
#define HEM_MOM_MINB 4
#endif
constexpr int MOM_WARPS = 8;
constexpr int MOM_B = HEM_MOM_B;
constexpr int MOM_U = HEM_MOM_U;          // child rows in flight per warp

struct MomArgs {
    const float4* geo; const float* sh; int Fp;
    const uint2* ptab;            // [np] (canonical index, canonical rank) in cell order
    const uint32_t* pair_off_s;   // [np+1] cell-order CSR offsets
    const uint2* csr_s;           // [P] (child, wL bits) in cell order
    const float* sumlw;
    int t_begin, t_end;           // this rank's slab of parents (cell order)
    int by_rank;                  // 1: row index = canonical rank (final place); 0: row index = t (staging for the exchange)
    unsigned long long pool_cap;  // the kernel does nothing if the pair pool overflowed (the host redoes the step)
    float4* ogeo; float* osh;
    unsigned long long* counters;
    // heavy parents (see k_moments_heavy): tickets holding a parent with more than heavy_thr pairs are not
    // accumulated here but flagged in hv_flag[ticket] (and counted in *n_heavy, control block) for k_moments_heavy;
    // which tickets are deferred depends on the pair counts only, never on scheduling
    uint8_t* hv_flag; uint32_t* hv_done; unsigned int* n_heavy; uint32_t heavy_thr;
};

constexpr uint32_t HEAVY_CHUNK = 2048;    // pairs per work item of the heavy kernel

// NSH = SH accumulators per lane: rows of up to 32*NSH floats (2 covers SH degree <= 3, the 3DGS standard)
// DEFER = false: the ordinary kernel; it only reports that a heavy parent exists (the host then redoes the step
// with DEFER = true, which leaves the tickets of heavy parents to k_moments_heavy).
// PART: 0 = geometry and SH rows in one pass (the resident path), 1 = geometry rows only, 2 = SH rows only.  The split
// passes serve the geometry-first pipeline of a build whose SH rows are still on their way from the host: every level's
// gates, claims and geometry moments need the 64-byte geometry rows only; the SH pass replays the pairs afterwards.
// Both passes add in the fused pass's order, so the rows are bit-identical to it.
template <int NSH, bool DEFER, int PART = 0>
__global__ void __launch_bounds__(MOM_WARPS * 32, HEM_MOM_MINB)
k_moments(MomArgs A)
{
    constexpr bool DO_GEO = PART != 2, DO_SH = PART != 1;
    pdl_enter();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int Fp = A.Fp;
    const float* geof = reinterpret_cast<const float*>(A.geo);
    float* ogeof = reinterpret_cast<float*>(A.ogeo);
    // lane j of the low half owns float j of a 64-byte geometry row:
    //  0 mx 1 my 2 mz 3 weight | 4 c00 5 c01 6 c02 7 c11 | 8 c12 9 c22 10 opacity 11 det | 12 r 13 g 14 b 15 flag
    const int j = lane & 15;
    const bool is_cov = j >= 4 && j <= 9;
    if (pool_overflowed(A.counters, A.pool_cap)) return;
    const int sa = (j <= 6) ? 0 : (j <= 8 ? 1 : 2);                 // first index of the outer-product term
    const int sb = (j == 4) ? 0 : ((j == 5 || j == 7) ? 1 : 2);     // second index

    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&A.counters[C_WORK2], 1ull);
        const int tb = A.t_begin + (int)__shfl_sync(full, ticket, 0) * MOM_B;
        if (tb >= A.t_end) break;
        const int nb = min(MOM_B, A.t_end - tb);
        const uint32_t off = (lane <= nb) ? A.pair_off_s[tb + lane] : 0xffffffffu;
        if (A.heavy_thr) {
            // one parent is one warp here; a parent with 10^5 pairs (a degenerate giant whose gates all pass) would
            // be the whole step's critical path: its ticket goes to k_moments_heavy, which splits it over the grid
            const uint32_t nxt = __shfl_down_sync(full, off, 1);
            const bool hv = lane < nb && (nxt - off) > A.heavy_thr;
            if (DEFER) {
                if (__ballot_sync(full, hv)) {
                    if (lane == 0) {
                        A.hv_flag[(tb - A.t_begin) / MOM_B] = 1;
                        atomicAdd(A.n_heavy, 1u);
                    }
                    continue;
                }
            } else if (hv) {
                atomicAdd(A.n_heavy, 1u);       // rare; tells the host to redo the step with the heavy path
            }
        }
        const uint2 pt = (lane < nb) ? A.ptab[tb + lane] : make_uint2(0u, 0u);
        float4 pq0 = make_float4(0, 0, 0, 0);
        if (lane < nb) pq0 = A.geo[4 * (size_t)pt.x];               // parent means (outer-product origin)
        const uint32_t pa = __shfl_sync(full, off, 0), pb = __shfl_sync(full, off, nb);

        int cur = 0;                                                 // parent of the batch being accumulated
        uint32_t nextb = __shfl_sync(full, off, 1);
        float pmj = 0.0f;                                            // parent mean component of lanes 0..2
        {
            const float x = __shfl_sync(full, pq0.x, 0), y = __shfl_sync(full, pq0.y, 0), z = __shfl_sync(full, pq0.z, 0);
            pmj = (j == 0) ? x : (j == 1 ? y : (j == 2 ? z : 0.0f));
        }
        float accg = 0.0f;
        float shacc[NSH];
#pragma unroll
        for (int q = 0; q < NSH; ++q) shacc[q] = 0.0f;

        auto finalize = [&]() {
            // normalise, mixture.cpp:219-226
            const float w_s = __shfl_sync(full, accg, 3);
            const float inv_w = fd(1.0f, w_s);
            const float scaled = fm(inv_w, accg);
            const float dm = fs(scaled, pmj);                        // mean - parent mean on lanes 0..2
            const float da = __shfl_sync(full, dm, sa), db = __shfl_sync(full, dm, sb);
            float outv = is_cov ? fs(scaled, fm(da, db)) : scaled;
            const uint32_t srank = __shfl_sync(full, pt.y, cur);
            if (j == 3) outv = w_s;
            if (j == 11 || j == 15) outv = 0.0f;      // det and parent flag are filled in by k_finish
            const size_t orow_i = A.by_rank ? (size_t)srank : (size_t)(tb + cur);
            if (DO_GEO && lane < 16) ogeof[16 * orow_i + lane] = outv;
            if (DO_SH) {
                float* orow = A.osh + orow_i * Fp;
#pragma unroll
                for (int q = 0; q < NSH; ++q)
                    if (lane + 32 * q < Fp) orow[lane + 32 * q] = fm(inv_w, shacc[q]);
            }
        };
        auto advance = [&]() {
            finalize();
            ++cur;
            accg = 0.0f;
#pragma unroll
            for (int q = 0; q < NSH; ++q) shacc[q] = 0.0f;
            if (cur < nb) {
                nextb = __shfl_sync(full, off, cur + 1);
                const float x = __shfl_sync(full, pq0.x, cur), y = __shfl_sync(full, pq0.y, cur), z = __shfl_sync(full, pq0.z, cur);
                pmj = (j == 0) ? x : (j == 1 ? y : (j == 2 ? z : 0.0f));
            }
        };

        for (uint32_t k0 = pa; k0 < pb; k0 += 32) {
            // phase 1: one pair per lane
            const uint32_t k = k0 + lane;
            uint32_t ch = 0;
            float w = 0.0f;
            if (k < pb) {
                const uint2 pr = A.csr_s[k];
                ch = pr.x;
                const float sl = A.sumlw[ch];
                const float cw = geof[16 * (size_t)ch + 3];
                w = fm(fd(__uint_as_float(pr.y), sl), cw);             // r_is * child.weight, mixture.cpp:196-197
            }
            const int cnt = (int)min(32u, pb - k0);
            // phase 2: rows, 8 pairs in flight
            for (int g = 0; g < cnt; g += MOM_U) {
                float gv[MOM_U], vs[NSH][MOM_U], wc[MOM_U];
#pragma unroll
                for (int u = 0; u < MOM_U; ++u) {
                    const int c = g + u;                  // <= 31; lanes past the window's last pair hold row 0 and weight 0
                    const uint32_t cc = __shfl_sync(full, ch, c);
                    wc[u] = __shfl_sync(full, w, c);
                    gv[u] = DO_GEO ? geof[16 * (size_t)cc + j] : 0.0f;
                    if (DO_SH) {
                        const float* row = A.sh + (size_t)cc * Fp;
#pragma unroll
                        for (int q = 0; q < NSH; ++q) vs[q][u] = lane + 32 * q < Fp ? row[lane + 32 * q] : 0.0f;
                    }
                }
#pragma unroll
                for (int u = 0; u < MOM_U; ++u) {
                    if (g + u < cnt) {                                  // uniform
                        while (k0 + g + u == nextb && cur < nb - 1) advance();   // parent boundary (empty parents too)
                        if (DO_GEO) {
                            const float dv = fs(gv[u], pmj);
                            const float da = __shfl_sync(full, dv, sa), db = __shfl_sync(full, dv, sb);
                            // tolerance-class sums (the reference's own summation order is not reproducible,
                            // DESIGN.md §2): fused multiply-adds
                            const float val = (j == 3) ? 1.0f : (is_cov ? fmaf(da, db, gv[u]) : gv[u]);
                            accg = fmaf(wc[u], val, accg);
                        } else {
                            accg = fmaf(wc[u], 1.0f, accg);       // the weight sum (lane 3 of the fused pass), on every lane
                        }
                        if (DO_SH) {
#pragma unroll
                            for (int q = 0; q < NSH; ++q) shacc[q] = fmaf(vs[q][u], wc[u], shacc[q]);
                        }
                    }
                }
            }
        }
        while (cur < nb - 1) advance();      // trailing parents without pairs
        finalize();
        __syncwarp();
    }
}

// Heavy parents.  Every deferred ticket's parents are cut into work items of HEAVY_CHUNK pairs that the whole grid
// shares; each item leaves an un-normalised partial row, and the warp that finishes a parent's last item adds the
// partial rows in item order and finalises the row exactly like k_moments does.  Item boundaries and the combine
// order depend only on the pair count, so the result is deterministic.
template <int NSH>
__global__ void __launch_bounds__(MOM_WARPS * 32)
k_moments_heavy(MomArgs A, float* __restrict__ part)
{
    pdl_enter();
    if (pool_overflowed(A.counters, A.pool_cap)) return;
    if (*A.n_heavy == 0) return;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = gridDim.x * MOM_WARPS;
    const int Fp = A.Fp;
    const float* geof = reinterpret_cast<const float*>(A.geo);
    float* ogeof = reinterpret_cast<float*>(A.ogeo);
    const int j = lane & 15;
    const bool is_cov = j >= 4 && j <= 9;
    const int sa = (j <= 6) ? 0 : (j <= 8 ? 1 : 2);
    const int sb = (j == 4) ? 0 : ((j == 5 || j == 7) ? 1 : 2);
    constexpr int PR = 16 + 32 * NSH;            // floats per partial row
    const int ntick = (A.t_end - A.t_begin + MOM_B - 1) / MOM_B;
    for (int tbase = 0; tbase < ntick; tbase += 32) {
      unsigned fm_ = __ballot_sync(full, tbase + lane < ntick && A.hv_flag[tbase + lane] != 0);
      while (fm_) {
        const int tb = A.t_begin + (tbase + __ffs(fm_) - 1) * MOM_B;
        fm_ &= fm_ - 1;
        const int nb = min(MOM_B, A.t_end - tb);
        for (int q = 0; q < nb; ++q) {
            const int t = tb + q;
            const uint32_t pa = A.pair_off_s[t], pb = A.pair_off_s[t + 1];
            const uint32_t nch = max(1u, (pb - pa + HEAVY_CHUNK - 1) / HEAVY_CHUNK);
            const uint2 pt = A.ptab[t];
            const float4 pq0 = A.geo[4 * (size_t)pt.x];
            const float pmj = (j == 0) ? pq0.x : (j == 1 ? pq0.y : (j == 2 ? pq0.z : 0.0f));
            for (uint32_t c = gwarp; c < nch; c += nwarps) {
                const uint32_t s0 = pa + c * HEAVY_CHUNK, s1 = min(pb, s0 + HEAVY_CHUNK);
                float accg = 0.0f;
                float shacc[NSH];
#pragma unroll
                for (int x = 0; x < NSH; ++x) shacc[x] = 0.0f;
                for (uint32_t k0 = s0; k0 < s1; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    uint32_t ch = 0;
                    float w = 0.0f;
                    if (k < s1) {
                        const uint2 pr = A.csr_s[k];
                        ch = pr.x;
                        w = fm(fd(__uint_as_float(pr.y), A.sumlw[ch]), geof[16 * (size_t)ch + 3]);
                    }
                    const int cnt = (int)min(32u, s1 - k0);
                    for (int g = 0; g < cnt; g += 8) {
                        float gv[8], vs[NSH][8], wc[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int cc_i = min(g + u, cnt - 1);
                            const uint32_t cc = __shfl_sync(full, ch, cc_i);
                            wc[u] = __shfl_sync(full, w, cc_i);
                            gv[u] = geof[16 * (size_t)cc + j];
                            const float* row = A.sh + (size_t)cc * Fp;
#pragma unroll
                            for (int x = 0; x < NSH; ++x) vs[x][u] = lane + 32 * x < Fp ? row[lane + 32 * x] : 0.0f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (g + u < cnt) {
                                const float dv = fs(gv[u], pmj);
                                const float da = __shfl_sync(full, dv, sa), db = __shfl_sync(full, dv, sb);
                                const float val = (j == 3) ? 1.0f : (is_cov ? fmaf(da, db, gv[u]) : gv[u]);
                                accg = fmaf(wc[u], val, accg);
#pragma unroll
                                for (int x = 0; x < NSH; ++x) shacc[x] = fmaf(vs[x][u], wc[u], shacc[x]);
                            }
                        }
                    }
                }
                // partial rows of consecutive parents never collide: item index = pa / CHUNK + c + t
                const size_t ibase = (size_t)(pa / HEAVY_CHUNK) + (size_t)t;
                float* pr = part + (ibase + c) * PR;
                if (lane < 16) __stcg(pr + lane, accg);
#pragma unroll
                for (int x = 0; x < NSH; ++x) __stcg(pr + 16 + 32 * x + lane, shacc[x]);
                __threadfence();
                uint32_t prev = 0;
                if (lane == 0) prev = atomicAdd(&A.hv_done[t], 1u);
                prev = __shfl_sync(full, prev, 0);
                if (prev == nch - 1) {
                    // last item of this parent: add the partial rows in item order, then finalise (mixture.cpp:219-226)
                    __threadfence();
                    float sg = 0.0f;
                    float ss[NSH];
#pragma unroll
                    for (int x = 0; x < NSH; ++x) ss[x] = 0.0f;
                    for (uint32_t cc = 0; cc < nch; ++cc) {
                        const float* r = part + (ibase + cc) * PR;
                        sg += __ldcg(r + j);
#pragma unroll
                        for (int x = 0; x < NSH; ++x) ss[x] += __ldcg(r + 16 + 32 * x + lane);
                    }
                    const float w_s = __shfl_sync(full, sg, 3);
                    const float inv_w = fd(1.0f, w_s);
                    const float scaled = fm(inv_w, sg);
                    const float dm = fs(scaled, pmj);
                    const float da = __shfl_sync(full, dm, sa), db = __shfl_sync(full, dm, sb);
                    float outv = is_cov ? fs(scaled, fm(da, db)) : scaled;
                    if (j == 3) outv = w_s;
                    if (j == 11 || j == 15) outv = 0.0f;
                    const size_t orow_i = A.by_rank ? (size_t)pt.y : (size_t)t;
                    if (lane < 16) ogeof[16 * orow_i + lane] = outv;
                    float* orow = A.osh + orow_i * Fp;
#pragma unroll
                    for (int x = 0; x < NSH; ++x)
                        if (lane + 32 * x < Fp) orow[lane + 32 * x] = fm(inv_w, ss[x]);
                    if (lane == 0) A.hv_done[t] = 0u;         // ready for the next step
                }
                __syncwarp();
            }
        }
      }
    }
}

// ---------------------------------------------------------------------------
// S5: orphans (mixture.cpp:251-253), their flags (:257-259) and validity (:262-282)
// ---------------------------------------------------------------------------
// orphan predicate, mixture.cpp:251-253: claimed by no parent <=> sumLw stayed exactly 0
struct IsOrphan {
    __host__ __device__ __forceinline__ uint32_t operator()(const float& v) const { return v == 0.0f ? 1u : 0u; }
};

__global__ void __launch_bounds__(256)
k_orphans(const float4* __restrict__ geo, const float* __restrict__ sh, int Fp, int64_t n,
          const float* __restrict__ sumlw, const uint32_t* __restrict__ orank, uint32_t np,
          float4* __restrict__ ogeo, float* __restrict__ osh, int do_geo, int do_sh)      // which half of a row to copy
{
    pdl_enter();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t wbase = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (wbase >= n) return;
    const int64_t i = wbase + lane;
    const bool orphan = i < n && sumlw[i] == 0.0f;
    unsigned m = __ballot_sync(full, orphan);
    const uint32_t myrank = orphan ? orank[i] : 0u;
    // Rows are copied ORPH_U at a time, all loads of the group issued before the first store: on the
    // coarse levels most of a warp's 32 rows are orphans and a row-at-a-time loop is one DRAM round
    // trip per row.
    constexpr int ORPH_U = 8;
    while (m) {
        int srcs[ORPH_U];
        size_t dsts[ORPH_U];
#pragma unroll
        for (int u = 0; u < ORPH_U; ++u) {
            srcs[u] = m ? __ffs(m) - 1 : -1;
            m &= m - 1;                                   // 0 stays 0
            dsts[u] = (size_t)np + __shfl_sync(full, myrank, srcs[u] < 0 ? 0 : srcs[u]);
        }
        if (Fp <= 64) {
            float4 g[ORPH_U];
            float a[ORPH_U], b[ORPH_U];
#pragma unroll
            for (int u = 0; u < ORPH_U; ++u) {
                if (srcs[u] < 0) continue;
                const size_t ci = (size_t)(wbase + srcs[u]);
                if (do_geo && lane < 4) g[u] = geo[4 * ci + lane];
                const float* srow = sh + ci * Fp;
                if (do_sh && lane < Fp) a[u] = srow[lane];
                if (do_sh && lane + 32 < Fp) b[u] = srow[lane + 32];
            }
#pragma unroll
            for (int u = 0; u < ORPH_U; ++u) {
                if (srcs[u] < 0) continue;
                if (do_geo && lane < 4) ogeo[4 * dsts[u] + lane] = g[u];
                float* drow = osh + dsts[u] * Fp;
                if (do_sh && lane < Fp) drow[lane] = a[u];
                if (do_sh && lane + 32 < Fp) drow[lane + 32] = b[u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < ORPH_U; ++u) {
                if (srcs[u] < 0) continue;
                const size_t ci = (size_t)(wbase + srcs[u]);
                if (do_geo && lane < 4) ogeo[4 * dsts[u] + lane] = geo[4 * ci + lane];
                const float* srow = sh + ci * Fp;
                float* drow = osh + dsts[u] * Fp;
                if (do_sh) for (int q = lane; q < Fp; q += 32) drow[q] = srow[q];
            }
        }
    }
}

// Per pre-filter output row j (merged parents [0,np) then orphans): determinant of the merged
// covariance, fresh parent flag drawn at the pre-filter position (mixture.cpp:257-259) and the
// validity predicate (mixture.cpp:262-270).  One THREAD per row: the Philox rounds and the
// determinant used to run warp-uniformly once per parent inside the moment kernel.
__global__ void __launch_bounds__(256)
k_finish(float4* __restrict__ geo, int64_t np, const uint32_t* __restrict__ orank,
         const float* __restrict__ sumlw, int64_t n_in, uint64_t seed, uint64_t rng_base,
         float pp, float distance_delta, uint8_t* __restrict__ valid,
         float* __restrict__ radius_n, uint8_t* __restrict__ psafe_n, uint32_t* __restrict__ pflag_n,
         unsigned long long pool_cap, Ctl* __restrict__ ctl)
{
    pdl_enter();
    __shared__ unsigned int s_red[8][10];
    __shared__ unsigned long long s_redd[8];
    if (pool_overflowed(ctl->counters, pool_cap)) return;     // merged rows were not produced: the host redoes the step
    // pre-filter row count = parents + orphans, read on the device (the host learns it afterwards)
    const int64_t norph = (int64_t)orank[n_in - 1] + (sumlw[n_in - 1] == 0.0f ? 1 : 0);
    const int64_t n = np + norph;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) ctl->counters[C_NORPH] = (unsigned long long)norph;
    bool bad = false;
    unsigned int bmax[3] = { 0u, 0u, 0u }, bminc[3] = { 0u, 0u, 0u };
    unsigned int rbits = 0, rcount = 0, pcount = 0, ccount = 0, ucount = 0;
    unsigned long long rsum = 0ull;
    if (j < n) {
        const float4 q0 = geo[4 * j], q1 = geo[4 * j + 1];
        float4 q2 = geo[4 * j + 2];
        float dt = q2.w;
        const Sym3 cc = { q1.x, q1.y, q1.z, q1.w, q2.x, q2.y };
        if (j < np) {               // merged row: determinant in the reference's operation order
            dt = det_sym3(cc);
            q2.w = dt;
            geo[4 * j + 2] = q2;
        }
        const bool flag = parent_flag(seed, rng_base + (uint64_t)j, pp);
        const bool ccert = child_cov_certified(cc, dt);
        ucount = ccert ? 0u : 1u;
        reinterpret_cast<float*>(geo)[16 * j + 15] = __uint_as_float((flag ? 1u : 0u) | (ccert ? 2u : 0u));
        bad = (q0.x != q0.x || q0.y != q0.y || q0.z != q0.z) || (dt != dt) || dt <= 0.0f;
        valid[j] = bad ? 0 : 1;
        // next level's S1 (mixture.cpp:70-95), valid as long as nothing gets erased by the filter
        float rad = 0.0f;
        bool cert = false;
        if (flag) {
            double emin, emax;
            rad = fm(distance_delta, __fsqrt_rn(eig_max_sym3(cc, &emin, &emax)));
            cert = parent_cov_certified(cc, dt, emin, emax);
            if (rad > 0.0f && rad < FLT_MAX) { rbits = __float_as_uint(rad); rcount = 1; rsum = radius_fx(rad); }
            pcount = 1; ccount = cert ? 1u : 0u;
        }
        radius_n[j] = rad; psafe_n[j] = cert ? 1 : 0; pflag_n[j] = flag ? 1u : 0u;
        if (q0.x == q0.x) { const unsigned o = f2u(q0.x); bmax[0] = o; bminc[0] = ~o; }
        if (q0.y == q0.y) { const unsigned o = f2u(q0.y); bmax[1] = o; bminc[1] = ~o; }
        if (q0.z == q0.z) { const unsigned o = f2u(q0.z); bmax[2] = o; bminc[2] = ~o; }
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(full, bad);
    if (lane == 0 && m) atomicAdd(&ctl->counters[C_INVALID], (unsigned long long)__popc(m));
    ucount = __reduce_add_sync(full, ucount);
    if (lane == 0 && ucount) atomicAdd(&ctl->nx_n_uncert, ucount);
#pragma unroll
    for (int a = 0; a < 3; ++a) { bmax[a] = __reduce_max_sync(full, bmax[a]); bminc[a] = __reduce_max_sync(full, bminc[a]); }
    rbits = __reduce_max_sync(full, rbits);
    rcount = __reduce_add_sync(full, rcount); pcount = __reduce_add_sync(full, pcount); ccount = __reduce_add_sync(full, ccount);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(full, rsum, o);
    if (lane == 0) {
        for (int a = 0; a < 3; ++a) { s_red[warp][a] = bminc[a]; s_red[warp][3 + a] = bmax[a]; }
        s_red[warp][6] = rbits; s_red[warp][7] = rcount; s_red[warp][8] = pcount; s_red[warp][9] = ccount;
        s_redd[warp] = rsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            for (int a = 0; a < 3; ++a) { bminc[a] = max(bminc[a], s_red[w][a]); bmax[a] = max(bmax[a], s_red[w][3 + a]); }
            rbits = max(rbits, s_red[w][6]); rcount += s_red[w][7]; pcount += s_red[w][8]; ccount += s_red[w][9];
            rsum += s_redd[w];
        }
        for (int a = 0; a < 3; ++a) {
            if (bmax[a]) atomicMax(&ctl->nx_bbmax_u[a], bmax[a]);
            if (bminc[a]) atomicMax(&ctl->nx_bbmin_c[a], bminc[a]);
        }
        if (pcount) atomicAdd(&ctl->nx_n_parents, pcount);
        if (ccount) atomicAdd(&ctl->nx_n_certified, ccount);
        if (rcount) {
            atomicMax(&ctl->nx_max_radius_bits, rbits);
            atomicAdd(&ctl->nx_n_radius, rcount);
            atomicAdd(&ctl->nx_sum_radius_fx, rsum);
        }
    }
}

// staged merged rows (index = cell-order position t) -> their final place (index = canonical rank)
__global__ void k_permute_rows(const uint2* __restrict__ ptab, int np, const float4* __restrict__ sgeo,
                               const float* __restrict__ ssh, int Fp, float4* __restrict__ ogeo,
                               float* __restrict__ osh, int do_geo, int do_sh)
{
    const int lane = threadIdx.x & 31;
    const int t = (int)(((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= np) return;
    const size_t d = ptab[t].y;
    if (do_geo && lane < 4) ogeo[4 * d + lane] = sgeo[4 * (size_t)t + lane];
    if (do_sh)
        for (int q = lane; q < Fp; q += 32) osh[d * Fp + q] = ssh[(size_t)t * Fp + q];
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
