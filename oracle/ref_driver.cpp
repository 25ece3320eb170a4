// TEST INFRASTRUCTURE — not product code.
//
// White-box driver around the UNMODIFIED reference HEM sources, which are
// compiled where they lie under /root/reference/src/cpp_ext by oracle/Makefile
// (outputs only into oracle/_ref/).  Nothing here re-implements the algorithm:
// every clustering step is executed by the reference's own
// hem::Mixture::createClusterLevel (reference src/cpp_ext/src/mixture.cpp:66-285)
// and the child lists are re-derived with the reference's own
// PointIndex::radiusSearch / ColorDelta / KLD (pointindex.cpp:120-143,
// gaussian.hpp:106-114) in the loop shape of mixture.cpp:106-137, because the
// reference keeps them as locals and never exports them.
//
// The only intervention is the RNG redirect of oracle/ref_rng_shim.h, so that
// the reference and the CUDA path see the same Philox-derived parent mask.
//
// C ABI (consumed through ctypes by oracle/refhem.py):
//   refhem_run / refhem_run_level / refhem_num_levels / refhem_level_size /
//   refhem_level_copy / refhem_children_count / refhem_children_copy /
//   refhem_level_seconds / refhem_free / refhem_philox_u32

#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <chrono>
#include <iostream>
#include <sstream>
#include <unordered_map>
#include <cassert>
#include <omp.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#define private public
#define protected public
#include "mixture.hpp"
#undef private
#undef protected

// ---------------------------------------------------------------------------
// Counter-based RNG shared with the CUDA path: Philox4x32-10, key = seed,
// counter = (k, 0, 0, 0), first output word.
// ---------------------------------------------------------------------------
static inline uint32_t philox_u32(uint64_t seed, uint64_t k)
{
    uint32_t c0 = (uint32_t)k, c1 = (uint32_t)(k >> 32), c2 = 0, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

static uint64_t g_seed = 0;
static uint64_t g_counter = 0;   // k: running (pre-filter) component index
static int g_nibble = 0;
static uint32_t g_word = 0;
static uint64_t g_calls = 0;

namespace std {
int hemoracle_rand()
{
    if (g_nibble == 0) g_word = philox_u32(g_seed, g_counter);
    int v = (int)((g_word >> (4 * g_nibble)) & 0xFu);
    if (++g_nibble == 8) { g_nibble = 0; ++g_counter; }
    ++g_calls;
    return v;
}
}

namespace {

struct Run
{
    hem::Mixture* mix = nullptr;
    hem::MixtureLevel input;
    std::vector<double> seconds;      // per clustering level
    int F = 0;
    float dd = 0, cd = 0;
};

void fill_level(hem::MixtureLevel& ml, const float* xyz, const float* col, const float* op,
                const float* cov6, const float* sh, int64_t n, int F)
{
    ml.pointSet.resize(n); ml.colorSet.resize(n); ml.covarianceSet.resize(n);
    ml.opacities.resize(n); ml.features.resize(n);
    for (int64_t i = 0; i < n; ++i)
    {
        ml.pointSet[i] = hem::vec3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]);
        ml.colorSet[i] = hem::vec3(col[3*i], col[3*i+1], col[3*i+2]);
        ml.covarianceSet[i] = hem::smat3(cov6[6*i], cov6[6*i+1], cov6[6*i+2], cov6[6*i+3], cov6[6*i+4], cov6[6*i+5]);
        ml.opacities[i] = op[i];
        ml.features[i] = hem::FeatureVector(std::vector<float>(sh + (size_t)i*F, sh + (size_t)(i+1)*F));
    }
}

} // namespace

extern "C" {

uint32_t refhem_philox_u32(uint64_t seed, uint64_t k) { return philox_u32(seed, k); }

// Whole pyramid, free-running, exactly the reference's own driver sequence
// (mixture_wrapper.cpp:10-18 -> Mixture ctor/initMixture -> CreateMixture).
void* refhem_run(const float* xyz, const float* col, const float* op, const float* cov6, const float* sh,
                 int64_t n, int F, unsigned levels, float red, float dd, float cd, float decay,
                 uint64_t seed, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Run* r = new Run();
    r->F = F; r->dd = dd; r->cd = cd;
    fill_level(r->input, xyz, col, op, cov6, sh, n, F);
    g_seed = seed; g_counter = 0; g_nibble = 0; g_calls = 0;
    std::cout.setstate(std::ios_base::failbit);        // silence "clustering level l"
    r->mix = new hem::Mixture(r->input, red, dd, cd, decay);
    r->mix->mixtureList.clear();
    r->mix->mixtureList.push_back(r->mix->initialMixture);
    for (unsigned l = 0; l < levels; ++l)
    {
        auto t0 = std::chrono::steady_clock::now();
        r->mix->createClusterLevel(r->mix->mixtureList.back());
        auto t1 = std::chrono::steady_clock::now();
        r->seconds.push_back(std::chrono::duration<double>(t1 - t0).count());
    }
    std::cout.clear();
    return r;
}

// One teacher-forced clustering step on caller-supplied rows, weights and
// parent flags.  rng_base = counter of the first output component's flag.
void* refhem_run_level(const float* xyz, const float* col, const float* op, const float* cov6, const float* sh,
                       const float* weight, const uint8_t* is_parent,
                       int64_t n, int F, float red, float dd, float cd, float decay,
                       uint64_t seed, uint64_t rng_base, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Run* r = new Run();
    r->F = F; r->dd = dd; r->cd = cd;
    fill_level(r->input, xyz, col, op, cov6, sh, n, F);
    g_seed = seed; g_counter = 0; g_nibble = 0; g_calls = 0;
    std::cout.setstate(std::ios_base::failbit);
    r->mix = new hem::Mixture(r->input, red, dd, cd, decay);
    for (int64_t i = 0; i < n; ++i)
    {
        r->mix->initialMixture[i].weight = weight ? weight[i] : 1.0f;
        r->mix->initialMixture[i].is_parent = is_parent[i] != 0;
    }
    r->mix->mixtureList.clear();
    r->mix->mixtureList.push_back(r->mix->initialMixture);
    g_counter = rng_base; g_nibble = 0;
    auto t0 = std::chrono::steady_clock::now();
    r->mix->createClusterLevel(r->mix->mixtureList.back());
    auto t1 = std::chrono::steady_clock::now();
    r->seconds.push_back(std::chrono::duration<double>(t1 - t0).count());
    std::cout.clear();
    return r;
}

int refhem_num_levels(void* h) { return (int)((Run*)h)->mix->mixtureList.size(); }

int64_t refhem_level_size(void* h, int level) { return (int64_t)((Run*)h)->mix->mixtureList.at(level).size(); }

double refhem_level_seconds(void* h, int level) { return ((Run*)h)->seconds.at(level); }

uint64_t refhem_rng_calls(void) { return g_calls; }

void refhem_level_copy(void* h, int level, float* xyz, float* col, float* op, float* cov6, float* sh,
                       float* weight, uint8_t* is_parent)
{
    Run* r = (Run*)h;
    const std::vector<hem::Component>& L = r->mix->mixtureList.at(level);
    const int F = r->F;
    for (size_t i = 0; i < L.size(); ++i)
    {
        const hem::Component& c = L[i];
        if (xyz) { xyz[3*i] = c.mean.x; xyz[3*i+1] = c.mean.y; xyz[3*i+2] = c.mean.z; }
        if (col) { col[3*i] = c.color.x; col[3*i+1] = c.color.y; col[3*i+2] = c.color.z; }
        if (op) op[i] = c.opacity;
        if (cov6) { cov6[6*i] = c.cov.e00; cov6[6*i+1] = c.cov.e01; cov6[6*i+2] = c.cov.e02;
                    cov6[6*i+3] = c.cov.e11; cov6[6*i+4] = c.cov.e12; cov6[6*i+5] = c.cov.e22; }
        if (sh) { const std::vector<float>& v = c.featureVector.GetVector();
                  for (int k = 0; k < F && k < (int)v.size(); ++k) sh[(size_t)i*F + k] = v[k]; }
        if (weight) weight[i] = c.weight;
        if (is_parent) is_parent[i] = c.is_parent ? 1 : 0;
    }
}

// Child lists of the clustering step whose INPUT is mixtureList[level].
// Pass child == nullptr to obtain only counts.  off has nparents+1 entries;
// parent_idx has nparents entries; children are sorted ascending per parent.
// Returns the total number of accepted (parent, child) pairs, fills
// *nparents_out.
int64_t refhem_children(void* h, int level, int64_t* nparents_out, int64_t* off, uint32_t* parent_idx,
                        uint32_t* child, int64_t cap)
{
    using namespace hem;
    Run* r = (Run*)h;
    std::vector<Component>& M = r->mix->mixtureList.at(level);
    const float distanceDelta = r->dd, colorDelta = r->cd;
    size_t n = M.size();
    std::vector<vec3> centers(n);
    std::vector<uint> parents;
    std::vector<float> radii;
    float maxR = 0;
    for (uint i = 0; i < n; ++i)
    {
        centers[i] = M[i].mean;
        if (M[i].is_parent)
        {
            parents.push_back(i);
            float q = distanceDelta * sqrtf(M[i].cov.eigenvalues().z);
            radii.push_back(q);
            if (q > maxR) maxR = q;
        }
    }
    *nparents_out = (int64_t)parents.size();
    if (parents.empty()) { if (off) off[0] = 0; return 0; }
    PointIndex index(centers, maxR);
    std::vector<std::vector<uint>> lists(parents.size());
    #pragma omp parallel for schedule(dynamic, 64)
    for (int s_ = 0; s_ < (int)parents.size(); ++s_)
    {
        int s = parents[s_];
        const Gaussian& parent = M[s];
        std::vector<uint> res;
        index.radiusSearch(parent.mean, radii[s_], res);
        for (size_t k = 0; k < res.size(); ++k)
        {
            int i = res[k];
            const Component& c = M[i];
            if (ColorDelta(c, parent) > colorDelta * colorDelta * 0.5f) continue;
            if (KLD(c, parent) > distanceDelta * distanceDelta * 0.5f) continue;
            if (c.is_parent && s != i) continue;
            lists[s_].push_back(i);
        }
        std::sort(lists[s_].begin(), lists[s_].end());
    }
    int64_t total = 0;
    for (size_t s_ = 0; s_ < parents.size(); ++s_)
    {
        if (off) off[s_] = total;
        if (parent_idx) parent_idx[s_] = parents[s_];
        if (child)
            for (size_t k = 0; k < lists[s_].size() && total + (int64_t)k < cap; ++k)
                child[total + k] = lists[s_][k];
        total += (int64_t)lists[s_].size();
    }
    if (off) off[parents.size()] = total;
    return total;
}

// Query radius of every component of a level as the reference computes it
// (mixture.cpp:88, vec.hpp:736-768); used to pin the radius arithmetic.
void refhem_radii(void* h, int level, float dd, float* out)
{
    Run* r = (Run*)h;
    std::vector<hem::Component>& M = r->mix->mixtureList.at(level);
    for (size_t i = 0; i < M.size(); ++i) out[i] = dd * sqrtf(M[i].cov.eigenvalues().z);
}

void refhem_free(void* h)
{
    Run* r = (Run*)h;
    delete r->mix;
    delete r;
}

} // extern "C"
