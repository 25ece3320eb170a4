// hem_icp_host.inl — host orchestration + C ABI of the ICP consumer (kernels: hem_icp.cuh).
// Included at the end of hem_b200.cu (shares hem_ctx, ensure, cub_call, to_device).

namespace {

struct IcpParams {
    double max_corr, rel_fitness, rel_rmse;
    int max_iteration;
};

// d_src / d_tgt are device arrays [n*3] FP64.  result = HEM_ICP_RESULT_LEN doubles (host).
int icp_run(hem_ctx* ctx, const double* d_src, int64_t ns, const double* d_tgt, int64_t nt, const double* init,
            const IcpParams& P, double* result, int32_t* correspondence)
{
    using namespace hem::icp;
    cudaStream_t st = ctx->stream;
    if (ns < 0 || nt < 0 || ns > 0x7fffffffll || nt > 0x7fffffffll) { ctx->err = "icp: point count out of range"; return HEM_ERR_INVALID; }
    if (!(P.max_corr > 0.0) || P.max_iteration < 0) { ctx->err = "icp: max_correspondence must be > 0 and max_iteration >= 0"; return HEM_ERR_INVALID; }
    Mat16 T0;
    for (int k = 0; k < 16; ++k) T0.m[k] = init ? init[k] : ((k % 5) == 0 ? 1.0 : 0.0);
    for (int k = 0; k < HEM_ICP_RESULT_LEN; ++k) result[k] = 0.0;
    for (int k = 0; k < 16; ++k) result[k] = T0.m[k];
    if (ns == 0 || nt == 0) {
        // nothing to match: fitness 0, rmse 0, transformation = init (Open3D returns the same for empty input)
        result[20] = 3.0;
        if (correspondence && ns > 0) {
            if (is_device_ptr(correspondence)) CK(cudaMemsetAsync(correspondence, 0xff, (size_t)ns * 4, st));
            else memset(correspondence, 0xff, (size_t)ns * 4);
            CK(cudaStreamSynchronize(st));
        }
        return HEM_OK;
    }
    if (!ctx->h_icp) {
        CK(cudaHostAlloc((void**)&ctx->h_icp, sizeof(State), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&ctx->h_icp_dev, ctx->h_icp, 0));
        CK(cudaEventCreate(&ctx->ev_icp0));
        CK(cudaEventCreate(&ctx->ev_icp1));
    }
    CK(cudaEventRecord(ctx->ev_icp0, st));

    // ---- target grid -------------------------------------------------------------------------
    CKR(ensure(ctx, ctx->icp_box, 6 * sizeof(unsigned long long)));
    unsigned long long hbox[6];
    for (int k = 0; k < 3; ++k) { hbox[k] = ~0ull; hbox[3 + k] = 0ull; }
    CK(cudaMemcpyAsync(ctx->icp_box.p, hbox, sizeof hbox, cudaMemcpyHostToDevice, st));
    k_icp_bbox<<<std::min<unsigned>(blocks_for(nt, 256), (unsigned)ctx->sm_count * 8), 256, 0, st>>>(
        d_tgt, nt, ctx->icp_box.as<unsigned long long>());
    CK(cudaGetLastError()); ctx->n_launches++;
    CK(cudaMemcpyAsync(hbox, ctx->icp_box.p, sizeof hbox, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    Grid g;
    double ext[3];
    bool any = hbox[0] != ~0ull;
    for (int k = 0; k < 3; ++k) {
        const double lo = any ? dec_ordered(hbox[k]) : 0.0, hi = any ? dec_ordered(hbox[3 + k]) : 0.0;
        g.lo[k] = lo;
        ext[k] = std::max(hi - lo, 0.0);
    }
    // about two target points per occupied cell for a volume-filling cloud; never more than ~4 cells per point
    const double emax = std::max(ext[0], std::max(ext[1], ext[2]));
    double vol = 1.0;
    for (int k = 0; k < 3; ++k) vol *= std::max(ext[k], emax * 1e-3);
    double h = std::cbrt(vol / std::max(0.5 * (double)nt, 1.0));
    if (!(h > 0.0) || !std::isfinite(h)) h = 1.0;
    const double cell_limit = std::min(std::max(4.0 * (double)nt, 4096.0), 268435456.0);
    for (;;) {
        double cells = 1.0;
        for (int k = 0; k < 3; ++k) cells *= std::floor(ext[k] / h) + 1.0;
        if (cells <= cell_limit) break;
        h *= 1.25;
    }
    g.h = h;
    g.inv_h = 1.0 / h;
    uint64_t ncell64 = 1;
    for (int k = 0; k < 3; ++k) { g.g[k] = (int)(std::floor(ext[k] / h) + 1.0); ncell64 *= (uint64_t)g.g[k]; }
    const uint32_t ncell = (uint32_t)ncell64;
    const double ox = g.lo[0] + 0.5 * ext[0], oy = g.lo[1] + 0.5 * ext[1], oz = g.lo[2] + 0.5 * ext[2];

    CKR(ensure(ctx, ctx->icp_key, (size_t)nt * 4));
    CKR(ensure(ctx, ctx->icp_key2, (size_t)nt * 4));
    CKR(ensure(ctx, ctx->icp_val, (size_t)nt * 4));
    CKR(ensure(ctx, ctx->icp_val2, (size_t)nt * 4));
    CKR(ensure(ctx, ctx->icp_rows, (size_t)nt * sizeof(double4)));
    CKR(ensure(ctx, ctx->icp_start, ((size_t)ncell + 1) * 4));
    k_icp_cellkey<<<blocks_for(nt, 256), 256, 0, st>>>(d_tgt, nt, g, ctx->icp_key.as<uint32_t>(), ctx->icp_val.as<uint32_t>());
    CK(cudaGetLastError()); ctx->n_launches++;
    {
        const int bits = bits_for(ncell ? ncell - 1 : 0);
        uint32_t *k1 = ctx->icp_key.as<uint32_t>(), *k2 = ctx->icp_key2.as<uint32_t>(), *v1 = ctx->icp_val.as<uint32_t>(),
                 *v2 = ctx->icp_val2.as<uint32_t>();
        CKR(cub_call(ctx, [&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, k1, k2, v1, v2, (int)nt, 0, bits, st);
        }));
        ctx->n_launches += 2 + (uint64_t)((bits + 7) / 8);
    }
    k_icp_gather<<<blocks_for(nt, 256), 256, 0, st>>>(d_tgt, ctx->icp_key2.as<uint32_t>(), ctx->icp_val2.as<uint32_t>(), nt, ncell,
                                                       ctx->icp_rows.as<double4>(), ctx->icp_start.as<uint32_t>());
    CK(cudaGetLastError()); ctx->n_launches++;

    // ---- iterations, all enqueued up front; the device decides when to stop -------------------
    const int nblocks = (int)std::min<int64_t>((ns + NN_THREADS - 1) / NN_THREADS, (int64_t)ctx->sm_count * 16);
    CKR(ensure(ctx, ctx->icp_part, (size_t)nblocks * NSUM * sizeof(double)));
    CKR(ensure(ctx, ctx->icp_state, sizeof(State) + 16));
    int32_t* d_corr = nullptr;
    if (correspondence) {
        if (is_device_ptr(correspondence)) d_corr = correspondence;
        else { CKR(ensure(ctx, ctx->icp_corr, (size_t)ns * 4)); d_corr = ctx->icp_corr.as<int32_t>(); }
    }
    State* dst = ctx->icp_state.as<State>();
    UpdateArgs U;
    U.rel_fitness = P.rel_fitness; U.rel_rmse = P.rel_rmse; U.max_iteration = P.max_iteration;
    U.block_counter = (unsigned int*)(ctx->icp_state.as<char>() + sizeof(State));
    U.host_copy = ctx->h_icp_dev;
    k_icp_init<<<1, 32, 0, st>>>(dst, T0, U.block_counter);
    CK(cudaGetLastError()); ctx->n_launches++;
    for (int k = 0; k <= P.max_iteration; ++k) {
        k_icp_nn<<<nblocks, NN_THREADS, 0, st>>>(d_src, ns, ctx->icp_rows.as<double4>(), ctx->icp_start.as<uint32_t>(), g, P.max_corr,
                                                 ox, oy, oz, dst, ctx->icp_part.as<double>(), d_corr, U);
        ctx->n_launches++;
    }
    CK(cudaGetLastError());
    if (correspondence && d_corr != correspondence) {
        CK(cudaMemcpyAsync(correspondence, d_corr, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        ctx->bytes_d2h += (size_t)ns * 4;
    }
    CK(cudaEventRecord(ctx->ev_icp1, st));
    CK(cudaStreamSynchronize(st));
    const State& S = *ctx->h_icp;
    for (int k = 0; k < 16; ++k) result[k] = S.T[k];
    result[16] = S.fitness;
    result[17] = S.rmse;
    result[18] = (double)S.iter;
    result[19] = (double)S.n_corr;
    result[20] = (double)S.done;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_icp0, ctx->ev_icp1);
    result[21] = (double)ms;
    result[22] = (double)ncell;
    result[23] = h;
    return HEM_OK;
}

} // namespace

extern "C" {

int hem_icp_point_to_point(hem_ctx* ctx, const double* source, int64_t n_source, const double* target, int64_t n_target,
                           const double* init, double max_correspondence, double relative_fitness, double relative_rmse,
                           int max_iteration, double* result, int32_t* correspondence)
{
    if (!ctx) return HEM_ERR_INVALID;
    if (!result || (n_source > 0 && !source) || (n_target > 0 && !target)) { ctx->err = "icp: null argument"; return HEM_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const void *ds = nullptr, *dt = nullptr;
    CKR(to_device(ctx, source, (size_t)std::max<int64_t>(n_source, 0) * 24, ctx->icp_src, &ds));
    CKR(to_device(ctx, target, (size_t)std::max<int64_t>(n_target, 0) * 24, ctx->icp_tgt, &dt));
    IcpParams P{ max_correspondence, relative_fitness, relative_rmse, max_iteration };
    return icp_run(ctx, (const double*)ds, n_source, (const double*)dt, n_target, init, P, result, correspondence);
}

int hem_icp_multiscale(hem_ctx* source_ctx, hem_ctx* target_ctx, int n_scales, const double* max_correspondence,
                       const int32_t* max_iteration, double relative_fitness, double relative_rmse, const double* init,
                       double* results)
{
    hem_ctx* ctx = source_ctx;
    if (!ctx || !target_ctx) return HEM_ERR_INVALID;
    if (!results || !max_correspondence || !max_iteration || n_scales < 1) { ctx->err = "icp: null argument"; return HEM_ERR_INVALID; }
    if (ctx->device != target_ctx->device) { ctx->err = "icp: both pyramids must live on the same device"; return HEM_ERR_INVALID; }
    const int nl_s = hem_num_levels(ctx), nl_t = hem_num_levels(target_ctx);
    // the reference's own checks (qt_multiscale_registrator.py:170-194)
    if (nl_s != nl_t) { ctx->err = "The two point cloud lists differ in size."; return HEM_ERR_STATE; }
    if (nl_s <= 1) { ctx->err = "There are no downscaled mixtures."; return HEM_ERR_STATE; }
    if (n_scales != nl_s) { ctx->err = "The number of iterations and the mixture levels do not match."; return HEM_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    // both pyramids are read on ctx's stream: make sure the target's builder stream is idle
    CK(cudaStreamSynchronize(target_ctx->stream));
    double T[16];
    for (int k = 0; k < 16; ++k) T[k] = init ? init[k] : ((k % 5) == 0 ? 1.0 : 0.0);
    for (int i = 0; i < n_scales; ++i) {
        const int level = nl_s - 1 - i;                       // pc_list[-(i+1)], coarsest first
        Level& Ls = ctx->levels[level];
        Level& Lt = target_ctx->levels[level];
        CKR(ensure(ctx, ctx->icp_src, (size_t)std::max<int64_t>(Ls.n, 1) * 24));
        CKR(ensure(ctx, ctx->icp_tgt, (size_t)std::max<int64_t>(Lt.n, 1) * 24));
        if (Ls.n) {
            k_registration_inputs<<<blocks_for(Ls.n, 256), 256, 0, ctx->stream>>>(Ls.geo.as<float4>(), Ls.n, ctx->icp_src.as<double>(),
                                                                                 nullptr, nullptr, nullptr);
            ctx->n_launches++;
        }
        if (Lt.n) {
            k_registration_inputs<<<blocks_for(Lt.n, 256), 256, 0, ctx->stream>>>(Lt.geo.as<float4>(), Lt.n, ctx->icp_tgt.as<double>(),
                                                                                 nullptr, nullptr, nullptr);
            ctx->n_launches++;
        }
        CK(cudaGetLastError());
        IcpParams P{ max_correspondence[i], relative_fitness, relative_rmse, (int)max_iteration[i] };
        double* r = results + (size_t)i * HEM_ICP_RESULT_LEN;
        CKR(icp_run(ctx, ctx->icp_src.as<double>(), Ls.n, ctx->icp_tgt.as<double>(), Lt.n, T, P, r, nullptr));
        for (int k = 0; k < 16; ++k) T[k] = r[k];
    }
    return HEM_OK;
}

} // extern "C"
