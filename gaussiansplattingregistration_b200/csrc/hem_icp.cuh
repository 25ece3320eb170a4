// hem_icp.cuh — point-to-point ICP over pyramid levels on the GPU (sm_100a).
//
// "Next" row of SURVEY.md §8f (rank 4): the consumer of the HEM pyramid.  The reference calls
//   o3d.pipelines.registration.registration_icp(source, target, max_correspondence, init,
//       TransformationEstimationPointToPoint(), ICPConvergenceCriteria(rel_fitness, rel_rmse, max_iter))
// (src/utils/local_registration_util.py:76-100) once per pyramid level, coarsest first, chaining the
// transformation (src/gui/workers/registration/qt_multiscale_registrator.py:203-230).  Open3D
// (pinned 0.16.0, requirements.txt:3) is a third-party dependency that is not vendored in the
// reference; what is restated here is its published algorithm (Registration.cpp, RegistrationICP):
//   result_0 = evaluate(T_0);  repeat: U = umeyama(correspondences of result_k) (no scale),
//   T_{k+1} = U T_k, result_{k+1} = evaluate(T_{k+1}), stop when |d fitness| < rel_fitness and
//   |d rmse| < rel_rmse, or after max_iteration updates.
//   evaluate: nearest target point of each transformed source point with distance < max_correspondence;
//   fitness = #correspondences / #source; inlier_rmse = sqrt(sum d^2 / #correspondences).
//
// GPU formulation: the target cloud is binned once into a uniform grid (linear cell key, radix sort,
// cell-start table); each source thread walks Chebyshev shells of cells outwards until the best hit is
// provably the nearest (best <= (shell * cell)^2) or the search radius is covered.  All FP64.  One
// iteration = ONE k_icp_nn launch (fixed grid, per-block partial sums in a fixed order -> deterministic);
// its last block to finish reduces the partials, tests convergence and computes the rigid update with
// Horn's quaternion method (Jacobi on the 4x4 symmetric matrix), all on the device: the host enqueues
// every iteration up front and synchronises once; iterations after convergence exit at once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hem {
namespace icp {

constexpr int NN_THREADS = 128;
constexpr int NSUM = 17;          // n, sum d2, sum a[3], sum b[3], sum a b^T [9]

struct Grid {
    double lo[3];
    double h, inv_h;
    int g[3];
};

struct State {
    double T[16];                 // current transformation, row-major 4x4
    double fitness, rmse;         // of the last evaluation
    double prev_fitness, prev_rmse;
    long long n_corr;
    int iter;                     // updates applied so far
    int done;                     // 0 running, 1 converged, 2 max_iteration reached, 3 fewer than 3 correspondences
    int evals;                    // evaluations done
    int pad;
};

struct Mat16 { double m[16]; };

__device__ __forceinline__ unsigned long long enc_ordered(double d)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long e)
{
    const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

// bounding box of the finite points: box[0..2] = min (ordered encoding), box[3..5] = max
__global__ void __launch_bounds__(256) k_icp_bbox(const double* __restrict__ p, int64_t n, unsigned long long* __restrict__ box)
{
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
            lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
            lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
        }
    }
    __shared__ double s_lo[8][3], s_hi[8][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_lo[threadIdx.x >> 5][k] = lo[k]; s_hi[threadIdx.x >> 5][k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = s_lo[0][threadIdx.x], b = s_hi[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fmin(a, s_lo[w][threadIdx.x]); b = fmax(b, s_hi[w][threadIdx.x]); }
        atomicMin(&box[threadIdx.x], enc_ordered(a));
        atomicMax(&box[3 + threadIdx.x], enc_ordered(b));
    }
}

__device__ __forceinline__ int cell_of(double v, double lo, double inv_h, int g)
{
    const double c = floor((v - lo) * inv_h);
    return c < 0.0 ? 0 : (c >= (double)g ? g - 1 : (int)c);
}

__global__ void k_icp_cellkey(const double* __restrict__ p, int64_t n, Grid g, uint32_t* __restrict__ key,
                              uint32_t* __restrict__ val)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    uint32_t k = 0;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const int cx = cell_of(x, g.lo[0], g.inv_h, g.g[0]), cy = cell_of(y, g.lo[1], g.inv_h, g.g[1]),
                  cz = cell_of(z, g.lo[2], g.inv_h, g.g[2]);
        k = (uint32_t)cx + (uint32_t)g.g[0] * ((uint32_t)cy + (uint32_t)g.g[1] * (uint32_t)cz);
    }
    key[i] = k;
    val[i] = (uint32_t)i;
}

// sorted rows {x, y, z, original index} and the cell-start table (start[c] = first sorted position of a
// cell >= c; start[ncell] = n)
__global__ void k_icp_gather(const double* __restrict__ p, const uint32_t* __restrict__ skey,
                             const uint32_t* __restrict__ sidx, int64_t n, uint32_t ncell,
                             double4* __restrict__ rows, uint32_t* __restrict__ start)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = sidx[i];
    rows[i] = make_double4(p[3 * (int64_t)j], p[3 * (int64_t)j + 1], p[3 * (int64_t)j + 2],
                           __longlong_as_double((long long)j));
    const uint32_t k = skey[i];
    const uint32_t first = i ? skey[i - 1] + 1 : 0;
    for (uint32_t c = first; c <= k; ++c) start[c] = (uint32_t)i;
    if (i == n - 1)
        for (uint32_t c = k + 1; c <= ncell; ++c) start[c] = (uint32_t)n;
}

__global__ void k_icp_init(State* __restrict__ st, Mat16 init, unsigned int* __restrict__ block_counter)
{
    if (threadIdx.x || blockIdx.x) return;
    *block_counter = 0u;
    for (int k = 0; k < 16; ++k) st->T[k] = init.m[k];
    st->fitness = st->rmse = st->prev_fitness = st->prev_rmse = 0.0;
    st->n_corr = 0;
    st->iter = 0;
    st->done = 0;
    st->evals = 0;
    st->pad = 0;
}


// largest-eigenvalue eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations (FP64)
__host__ __device__ inline void jacobi4_max_eigvec(double A[4][4], double q[4])
{
    double V[4][4] = { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 4; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;   // eigenvector error ~ sqrt(off)/gap: at FP64 rounding level
        for (int p = 0; p < 3; ++p)
            for (int r = p + 1; r < 4; ++r) {
                const double apq = A[p][r];
                if (apq == 0.0) continue;
                const double theta = (A[r][r] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akq = A[k][r];
                    A[k][p] = c * akp - s * akq;
                    A[k][r] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], aqk = A[r][k];
                    A[p][k] = c * apk - s * aqk;
                    A[r][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][r];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][r] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > A[best][best]) best = i;
    double n2 = 0.0;
    for (int k = 0; k < 4; ++k) { q[k] = V[k][best]; n2 += q[k] * q[k]; }
    const double inv = 1.0 / sqrt(n2);
    for (int k = 0; k < 4; ++k) q[k] *= inv;
}

// Reduce the partial sums (fixed order), close the evaluation (fitness, rmse, convergence test) and,
// if the loop goes on, compute the rigid update and apply it to T.  Runs in the LAST block of k_icp_nn to
// finish (the order of the reduction does not depend on which block that is).
struct UpdateArgs {
    double rel_fitness, rel_rmse;
    int max_iteration;
    unsigned int* block_counter;     // zero before the first launch; the last block resets it
    State* host_copy;                // mapped host memory
};

__device__ __forceinline__ void icp_update(const double* __restrict__ partials, int nparts, int64_t ns, const UpdateArgs& U,
                                           double ox, double oy, double oz, State* __restrict__ st,
                                           double (*sm)[NSUM + 1])
{
    double acc[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) acc[k] = 0.0;
    for (int b = threadIdx.x; b < nparts; b += NN_THREADS)
#pragma unroll
        for (int k = 0; k < NSUM; ++k) acc[k] += __ldcg(&partials[(size_t)b * NSUM + k]);
#pragma unroll
    for (int k = 0; k < NSUM; ++k) sm[threadIdx.x][k] = acc[k];
    __syncthreads();
    for (int o = NN_THREADS / 2; o; o >>= 1) {
        if ((int)threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < NSUM; ++k) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
        __syncthreads();
    }
    const double rel_fitness = U.rel_fitness, rel_rmse = U.rel_rmse;
    const int max_iteration = U.max_iteration;
    State* host_copy = U.host_copy;
    if (threadIdx.x) return;
    const double* S = sm[0];
    const double n = S[0];
    const double fitness = ns > 0 ? n / (double)ns : 0.0;
    const double rmse = n > 0.0 ? sqrt(S[1] / n) : 0.0;
    const int it = st->iter;
    const bool conv = it > 0 && fabs(st->prev_fitness - fitness) < rel_fitness && fabs(st->prev_rmse - rmse) < rel_rmse;
    st->fitness = fitness;
    st->rmse = rmse;
    st->n_corr = (long long)n;
    st->evals += 1;
    if (conv) st->done = 1;
    else if (it >= max_iteration) st->done = 2;
    else if (n < 3.0) st->done = 3;
    else {
        const double inv = 1.0 / n;
        const double ca[3] = { S[2] * inv, S[3] * inv, S[4] * inv }, cb[3] = { S[5] * inv, S[6] * inv, S[7] * inv };
        double H[3][3];   // H[i][j] = mean (a_i - ca_i)(b_j - cb_j)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) H[i][j] = S[8 + 3 * i + j] * inv - ca[i] * cb[j];
        // Horn 1987: the rotation maximising sum b . (R a) is the top eigenvector of N(H)
        double N[4][4] = {
            { H[0][0] + H[1][1] + H[2][2], H[1][2] - H[2][1], H[2][0] - H[0][2], H[0][1] - H[1][0] },
            { H[1][2] - H[2][1], H[0][0] - H[1][1] - H[2][2], H[0][1] + H[1][0], H[2][0] + H[0][2] },
            { H[2][0] - H[0][2], H[0][1] + H[1][0], -H[0][0] + H[1][1] - H[2][2], H[1][2] + H[2][1] },
            { H[0][1] - H[1][0], H[2][0] + H[0][2], H[1][2] + H[2][1], -H[0][0] - H[1][1] + H[2][2] } };
        double q[4];
        jacobi4_max_eigvec(N, q);
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        const double R[3][3] = { { 1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y) },
                                 { 2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x) },
                                 { 2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y) } };
        // b - o = R (a - o) + t'  ->  b = R a + (t' + o - R o)
        const double o[3] = { ox, oy, oz };
        double t[3];
        for (int i = 0; i < 3; ++i) {
            const double tp = cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]);
            t[i] = tp + o[i] - (R[i][0] * o[0] + R[i][1] * o[1] + R[i][2] * o[2]);
        }
        double Tn[16];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 4; ++j)
                Tn[4 * i + j] = R[i][0] * st->T[j] + R[i][1] * st->T[4 + j] + R[i][2] * st->T[8 + j] + (j == 3 ? t[i] : 0.0);
        }
        Tn[12] = 0.0; Tn[13] = 0.0; Tn[14] = 0.0; Tn[15] = 1.0;
        for (int k = 0; k < 16; ++k) st->T[k] = Tn[k];
        st->prev_fitness = fitness;
        st->prev_rmse = rmse;
        st->iter = it + 1;
    }
    if (host_copy) *host_copy = *st;
}

// distance from coordinate q to the slab of cell c along one axis (0 inside; a hair of slack for the
// rounding of the cell assignment)
__device__ __forceinline__ double slab_dist(double q, double lo, double h, int c)
{
    const double a = lo + (double)c * h;
    return fmax(0.0, fmax(a - q, q - (a + h)) - 1e-9 * h);
}

__device__ __forceinline__ void scan_range(const double4* __restrict__ rows, uint32_t s, uint32_t e, double px,
                                           double py, double pz, double& best, double4& hit)
{
    for (uint32_t j = s; j < e; ++j) {
        const double4 r = rows[j];
        const double dx = r.x - px, dy = r.y - py, dz = r.z - pz;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { best = d2; hit = r; }
    }
}

// One evaluation: correspondences of T*source in the target grid + the sums the update needs.
// Coordinates enter the sums relative to `origin` (the target box centre) so that the centred
// cross-covariance is not formed from large cancelling terms.
__global__ void __launch_bounds__(NN_THREADS)
k_icp_nn(const double* __restrict__ src, int64_t ns, const double4* __restrict__ rows,
         const uint32_t* __restrict__ start, Grid g, double max_corr, double ox, double oy, double oz,
         State* __restrict__ st, double* __restrict__ partials, int32_t* __restrict__ corr, UpdateArgs U)
{
    if (st->done) return;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->T[k];
    double acc[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) acc[k] = 0.0;
    const double r2max = max_corr * max_corr;
    const int gmax = max(g.g[0], max(g.g[1], g.g[2]));

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += (int64_t)gridDim.x * blockDim.x) {
        const double sx = src[3 * i], sy = src[3 * i + 1], sz = src[3 * i + 2];
        const double px = T[0] * sx + T[1] * sy + T[2] * sz + T[3];
        const double py = T[4] * sx + T[5] * sy + T[6] * sz + T[7];
        const double pz = T[8] * sx + T[9] * sy + T[10] * sz + T[11];
        double best = r2max;
        double4 hit = make_double4(0.0, 0.0, 0.0, 0.0);
        if (isfinite(px) && isfinite(py) && isfinite(pz)) {
            const int cx = cell_of(px, g.lo[0], g.inv_h, g.g[0]), cy = cell_of(py, g.lo[1], g.inv_h, g.g[1]),
                      cz = cell_of(pz, g.lo[2], g.inv_h, g.g[2]);
            for (int r = 0; r < gmax; ++r) {
                const int z0 = cz - r, z1 = cz + r, y0 = cy - r, y1 = cy + r, x0 = cx - r, x1 = cx + r;
                const int xa = max(x0, 0), xb = min(x1, g.g[0] - 1);
                for (int z = max(z0, 0); z <= min(z1, g.g[2] - 1); ++z) {
                    const bool zedge = (z == z0) || (z == z1);
                    const double dz = slab_dist(pz, g.lo[2], g.h, z);
                    if (dz * dz >= best) continue;
                    for (int y = max(y0, 0); y <= min(y1, g.g[1] - 1); ++y) {
                        // rows of cells the current best sphere misses are skipped, the others trimmed to its chord
                        const double dy = slab_dist(py, g.lo[1], g.h, y);
                        const double rem = best - (dz * dz + dy * dy);
                        if (rem <= 0.0) continue;
                        const uint32_t base = (uint32_t)g.g[0] * ((uint32_t)y + (uint32_t)g.g[1] * (uint32_t)z);
                        if (zedge || y == y0 || y == y1) {
                            // a whole x-run of the shell: consecutive keys, one contiguous range of rows
                            const double half = sqrt(rem) + 1e-9 * g.h;
                            const int ca = max(xa, cell_of(px - half, g.lo[0], g.inv_h, g.g[0]));
                            const int cb = min(xb, cell_of(px + half, g.lo[0], g.inv_h, g.g[0]));
                            if (ca <= cb) scan_range(rows, start[base + ca], start[base + cb + 1], px, py, pz, best, hit);
                        } else {
                            if (x0 >= 0) {
                                const double dx = slab_dist(px, g.lo[0], g.h, x0);
                                if (dx * dx < rem) scan_range(rows, start[base + x0], start[base + x0 + 1], px, py, pz, best, hit);
                            }
                            if (x1 < g.g[0]) {
                                const double dx = slab_dist(px, g.lo[0], g.h, x1);
                                if (dx * dx < rem) scan_range(rows, start[base + x1], start[base + x1 + 1], px, py, pz, best, hit);
                            }
                        }
                    }
                }
                // every point outside the block of Chebyshev radius r is farther than r*h
                const double reach = (double)r * g.h;
                if (best <= reach * reach) break;
            }
        }
        const bool ok = best < r2max;
        if (corr) corr[i] = ok ? (int32_t)__double_as_longlong(hit.w) : -1;
        if (ok) {
            const double ax = px - ox, ay = py - oy, az = pz - oz;
            const double bx = hit.x - ox, by = hit.y - oy, bz = hit.z - oz;
            acc[0] += 1.0;
            acc[1] += best;
            acc[2] += ax; acc[3] += ay; acc[4] += az;
            acc[5] += bx; acc[6] += by; acc[7] += bz;
            acc[8] += ax * bx; acc[9] += ax * by; acc[10] += ax * bz;
            acc[11] += ay * bx; acc[12] += ay * by; acc[13] += ay * bz;
            acc[14] += az * bx; acc[15] += az * by; acc[16] += az * bz;
        }
    }
    __shared__ double sm[NN_THREADS][NSUM + 1];
    __shared__ bool s_last;
#pragma unroll
    for (int k = 0; k < NSUM; ++k)
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NSUM; ++k) sm[w][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < NSUM) {
        double s = 0.0;
        for (int q = 0; q < NN_THREADS / 32; ++q) s += sm[q][threadIdx.x];
        __stcg(&partials[(size_t)blockIdx.x * NSUM + threadIdx.x], s);
    }
    // last block to finish closes the evaluation (threadfence-reduction pattern)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done_blocks = atomicAdd(U.block_counter, 1u);
        s_last = (done_blocks == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    icp_update(partials, (int)gridDim.x, ns, U, ox, oy, oz, st, sm);
    if (threadIdx.x == 0) *U.block_counter = 0u;
}

} // namespace icp
} // namespace hem
