// hpf_step_batch: one minibatch update (included at the end of hpf_engine.cu; same translation unit).
// Mirrors hpfrec/cython_loops.pxi 275-325 (user epoch body), 329-377 (item epoch body) and
// 423-473 (Cython partial_fit).

extern "C" int hpf_step_batch(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t nnz,
                              const void* users, int64_t n_users, const void* items, int64_t n_items,
                              int32_t index_bytes, int32_t user_batch, double rho, double mult,
                              int32_t blend_all_rates) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->state_loaded || !h->mat_valid) return fail(HPF_ESTATE, "no state loaded (call hpf_load_state first)");
    if (nnz < 0 || nnz >= (1ll << 31) || n_users < 0 || n_items < 0) return fail(HPF_EINVAL, "bad sizes");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (nnz > 0 && (!ix_u || !ix_i || !Y)) return fail(HPF_EINVAL, "NULL triple array");
    if ((n_users > 0 && !users) || (n_items > 0 && !items)) return fail(HPF_EINVAL, "NULL id list");
    if (!(rho >= 0.0 && rho <= 1.0)) return fail(HPF_EINVAL, "step size must be in [0, 1]");
    DeviceGuard guard(h->device);
    if (!h->stamp_u) {
        CK(cudaMalloc(&h->stamp_u, sizeof(int) * (size_t)(h->nU > 0 ? h->nU : 1)));
        CK(cudaMalloc(&h->stamp_i, sizeof(int) * (size_t)(h->nI > 0 ? h->nI : 1)));
        CK(cudaMemsetAsync(h->stamp_u, 0, sizeof(int) * (size_t)(h->nU > 0 ? h->nU : 1), h->stream));
        CK(cudaMemsetAsync(h->stamp_i, 0, sizeof(int) * (size_t)(h->nI > 0 ? h->nI : 1), h->stream));
        h->batch_step = 0;
    }
    hpf_engine* sc = h;
    h->batch_step += 1;
    const int step = h->batch_step;

    int *u32 = nullptr, *i32 = nullptr, *ulist = nullptr, *ilist = nullptr, *d_bad = nullptr;
    void* yfree = nullptr;
    const void* yv = nullptr;
    int rc = HPF_OK;
    auto alloc_i = [&](int** p, int64_t n) {
        if (rc == HPF_OK && cudaMalloc(p, sizeof(int) * (size_t)(n > 0 ? n : 1)) != cudaSuccess)
            rc = fail(HPF_ENOMEM, "device allocation failed in hpf_step_batch");
    };
    alloc_i(&u32, nnz);
    alloc_i(&i32, nnz);
    alloc_i(&ulist, n_users);
    alloc_i(&ilist, n_items);
    alloc_i(&d_bad, 1);
    if (rc == HPF_OK) cudaMemsetAsync(d_bad, 0, 4, h->stream);
    if (rc == HPF_OK) rc = stage_index(h, ix_u, nnz, index_bytes, h->nU, u32, d_bad);
    if (rc == HPF_OK) rc = stage_index(h, ix_i, nnz, index_bytes, h->nI, i32, d_bad);
    if (rc == HPF_OK) rc = stage_index(h, users, n_users, index_bytes, h->nU, ulist, d_bad);
    if (rc == HPF_OK) rc = stage_index(h, items, n_items, index_bytes, h->nI, ilist, d_bad);
    if (rc == HPF_OK) rc = stage_in(h, Y, (size_t)nnz * h->rb, &yv, &yfree);
    if (rc == HPF_OK) {
        int bad = 0;
        cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        if (bad) rc = fail(HPF_EINVAL, "index out of range in minibatch");
    }
    if (rc == HPF_OK) {
        h->x_valid = false;  // per-row factors / accumulators are only valid for the batch rows from here on
        drop_graphs(h);
        rc = dispatch(h->rb, h->ld, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            const int ld = h->ld, k = h->k;
            const size_t smem = sizeof(double) * ld;
            // roles
            const bool ub = user_batch != 0;
            const int64_t nM = ub ? h->nU : h->nI, nm = ub ? h->nI : h->nU;
            real *shpM = (real*)(ub ? h->Gshp : h->Lshp), *rteM = (real*)(ub ? h->Grte : h->Lrte);
            real *shpm = (real*)(ub ? h->Lshp : h->Gshp), *rtem = (real*)(ub ? h->Lrte : h->Grte);
            real *xM = (real*)(ub ? h->xu : h->xi), *xm = (real*)(ub ? h->xi : h->xu);
            real *accM = (real*)(ub ? h->accU : h->accI), *accm = (real*)(ub ? h->accI : h->accU);
            real *rateM = (real*)(ub ? h->krte : h->trte), *ratem = (real*)(ub ? h->trte : h->krte);
            int *stampM = ub ? sc->stamp_u : sc->stamp_i, *stampm = ub ? sc->stamp_i : sc->stamp_u;
            const int *listM = ub ? ulist : ilist, *listm = ub ? ilist : ulist;
            const int64_t nlM = ub ? n_users : n_items, nlm = ub ? n_items : n_users;
            double *csM = ub ? h->Tsum : h->Bsum, *csm = ub ? h->Bsum : h->Tsum;
            const real priorM = (real)(ub ? h->a : h->c), priorm = (real)(ub ? h->c : h->a);
            const real k_shp = (real)h->k_shp, t_shp = (real)h->t_shp;
            const real add_k = (real)h->add_k, add_t = (real)h->add_t;
            const real srM = ub ? k_shp : t_shp, srm = ub ? t_shp : k_shp;
            const real addM = ub ? add_k : add_t, addm = ub ? add_t : add_k;

            // 1. softmax factors of the participating rows from the current state; zero their sums
            if (nlM > 0) {
                hpf::batch_prepare_kernel<real, C::lpg, C::vpl><<<row_grid(nlM, C::lpg), 256, 0, h->stream>>>(
                    (int)nlM, listM, ld, k, shpM, rteM, xM, accM, stampM, step);
                h->launches++;
            }
            if (nlm > 0) {
                hpf::batch_prepare_kernel<real, C::lpg, C::vpl><<<row_grid(nlm, C::lpg), 256, 0, h->stream>>>(
                    (int)nlm, listm, ld, k, shpm, rtem, xm, accm, stampm, step);
                h->launches++;
            }
            CKK();
            // 2. phi + scatter over the batch triples (update_phi + update_G_n_L_sh)
            TRY(launch_sweep_coo<C>(h, u32, i32, yv, nnz, h->xu, h->xi, h->accU, h->accI, ld, nullptr, k, h->stream));
            // 3. column sums of the minor side's current expectation (Beta.sum(0) at pxi:300, Theta.sum(0) at 352)
            CK(cudaMemsetAsync(csm, 0, sizeof(double) * ld, h->stream));
            CK(cudaMemsetAsync(csM, 0, sizeof(double) * ld, h->stream));
            if (nm > 0) {
                long long want = (nm * (long long)ld + 255) / 256;
                if (want > 148 * 8) want = 148 * 8;
                hpf::colsum_kernel<real><<<(unsigned)want, 256, smem, h->stream>>>(nm, ld, k, shpm, rtem, csm);
                h->launches++;
                CKK();
            }
            // 4. major side, all rows
            if (nM > 0) {
                hpf::batch_major_kernel<real, C::lpg, C::vpl><<<row_grid(nM, C::lpg), 256, smem, h->stream>>>(
                    (int)nM, ld, k, xM, accM, shpM, rteM, rateM, stampM, step, csm, csM, priorM, srM, addM,
                    (real)rho, blend_all_rates);
                h->launches++;
                CKK();
            }
            // 5. minor side: batch rows (SVI) or all rows (partial_fit)
            const int64_t nrows5 = blend_all_rates ? nm : nlm;
            if (nrows5 > 0) {
                hpf::batch_minor_kernel<real, C::lpg, C::vpl><<<row_grid(nrows5, C::lpg), 256, 0, h->stream>>>(
                    (int)nrows5, blend_all_rates ? nullptr : listm, ld, k, xm, accm, shpm, rtem, ratem, stampm, step,
                    csM, priorm, srm, addm, (real)rho, (real)mult);
                h->launches++;
                CKK();
            }
            return HPF_OK;
        });
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (rc == HPF_OK && e != cudaSuccess) rc = fail(HPF_ECUDA, "minibatch kernels failed: %s", cudaGetErrorString(e));
    cudaFree(u32);
    cudaFree(i32);
    cudaFree(ulist);
    cudaFree(ilist);
    cudaFree(d_bad);
    cudaFree(yfree);
    return rc;
}
