// Minibatch steps (included at the end of hpf_engine.cu; same translation unit).
// Mirrors hpfrec/cython_loops.pxi 275-325 (user epoch body), 329-377 (item epoch body) and
// 423-473 (Cython partial_fit).  Two front ends share one core:
//   hpf_step_batch      explicit COO triples + unique id lists from the caller (partial_fit)
//   hpf_step_batch_ids  only the batched row ids; the triples come from the resident CSR/CSC
//                       orderings (SVI epochs of fit_hpf; replaces get_unique_items_batch pxi:27-42)

namespace {

int ensure_stamps(hpf_engine* h) {
    if (h->stamp_u) return HPF_OK;
    CK(hpf_malloc(&h->stamp_u, sizeof(int) * (size_t)(h->nU > 0 ? h->nU : 1)));
    CK(hpf_malloc(&h->stamp_i, sizeof(int) * (size_t)(h->nI > 0 ? h->nI : 1)));
    CK(cudaMemsetAsync(h->stamp_u, 0, sizeof(int) * (size_t)(h->nU > 0 ? h->nU : 1), h->stream));
    CK(cudaMemsetAsync(h->stamp_i, 0, sizeof(int) * (size_t)(h->nI > 0 ? h->nI : 1), h->stream));
    h->batch_step = 0;
    return HPF_OK;
}

// Core of one minibatch update.  iu/ii/yv: device triples of the batch (int32 ids).  list_major: the
// batched ids (device int32).  list_minor: unique opposite-side ids, or nullptr when they are already
// stamped for `step` (device-assembled batches).  The minor side is walked by list when one is
// given and rates are blended on batch rows only; otherwise all rows are walked under the stamp.
int batch_core(hpf_engine* h, const int* iu, const int* ii, const void* yv, int64_t nnz, const int* list_major,
               int64_t n_major_ids, const int* list_minor, int64_t n_minor_ids, bool user_batch, double rho,
               double mult, bool blend_all, int step, bool grouped_padded = false) {
    h->x_valid = false;  // per-row factors / accumulators are only valid for the batch rows from here on
    drop_graphs(h);
    TRY(resolve_robust(h));
    return dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        using real = typename C::real;
        const int ld = h->ld, k = h->k;
        const size_t smem = sizeof(double) * ld;
        const bool ub = user_batch;
        const int64_t nM = ub ? h->nU : h->nI, nm = ub ? h->nI : h->nU;
        real *shpM = (real*)(ub ? h->Gshp : h->Lshp), *rteM = (real*)(ub ? h->Grte : h->Lrte);
        real *shpm = (real*)(ub ? h->Lshp : h->Gshp), *rtem = (real*)(ub ? h->Lrte : h->Grte);
        real *xM = (real*)(ub ? h->xu : h->xi), *xm = (real*)(ub ? h->xi : h->xu);
        real *accM = (real*)(ub ? h->accU : h->accI), *accm = (real*)(ub ? h->accI : h->accU);
        real *rateM = (real*)(ub ? h->krte : h->trte), *ratem = (real*)(ub ? h->trte : h->krte);
        int *stampM = ub ? h->stamp_u : h->stamp_i, *stampm = ub ? h->stamp_i : h->stamp_u;
        double *csM = ub ? h->Tsum : h->Bsum, *csm = ub ? h->Bsum : h->Tsum;
        const real priorM = (real)(ub ? h->a : h->c), priorm = (real)(ub ? h->c : h->a);
        const real k_shp = (real)h->k_shp, t_shp = (real)h->t_shp;
        const real add_k = (real)h->add_k, add_t = (real)h->add_t;
        const real srM = ub ? k_shp : t_shp, srm = ub ? t_shp : k_shp;
        const real addM = ub ? add_k : add_t, addm = ub ? add_t : add_k;
        real* dirM = h->robust_on ? (real*)(ub ? h->dirU : h->dirI) : nullptr;
        real* dirm = h->robust_on ? (real*)(ub ? h->dirI : h->dirU) : nullptr;

        // 1. softmax factors of the participating rows from the current state; zero their sums
        if (n_major_ids > 0) {
            hpf::batch_prepare_kernel<real, C::lpg, C::vpl><<<row_grid(n_major_ids, C::lpg), 256, 0, h->stream>>>(
                (int)n_major_ids, list_major, ld, k, shpM, rteM, xM, accM, dirM, stampM, step);
            h->launches++;
        }
        const int64_t n_prep_minor = list_minor ? n_minor_ids : nm;
        if (n_prep_minor > 0) {
            hpf::batch_prepare_kernel<real, C::lpg, C::vpl><<<row_grid(n_prep_minor, C::lpg), 256, 0, h->stream>>>(
                (int)n_prep_minor, list_minor, ld, k, shpm, rtem, xm, accm, dirm, stampm, step);
            h->launches++;
        }
        CKK();
        // 2. phi + scatter over the batch triples (update_phi + update_G_n_L_sh).  Device-assembled batches are grouped
        //    by the batched side and padded: their batched side accumulates in registers (one RED per row, not per nnz)
        if (grouped_padded && !h->robust_on)
            TRY(launch_sweep_batch<C>(h, ub ? iu : ii, ub ? ii : iu, yv, nnz, xM, xm, accM, accm));
        else
            TRY(launch_sweep_coo<C>(h, iu, ii, yv, nnz, h->xu, h->xi, h->accU, h->accI, ld, nullptr, k, h->stream));
        // 3. column sums of the minor side's current expectation (Beta.sum(0) at pxi:300, Theta.sum(0) at 352).
        //    Between consecutive minibatch steps they are carried: the major kernel recomputes its side's sums in
        //    full, the minor kernel adds the change of its batch rows (in double), so only the first step after
        //    anything else touched the state pays the full pass.
        CK(cudaMemsetAsync(csM, 0, sizeof(double) * ld, h->stream));
        if (!h->batch_colsums_valid) CK(cudaMemsetAsync(csm, 0, sizeof(double) * ld, h->stream));
        if (nm > 0 && !h->batch_colsums_valid) {
            long long want = (nm * (long long)ld + 255) / 256;
            if (want > 148 * 8) want = 148 * 8;
            hpf::colsum_kernel<real><<<(unsigned)want, 256, smem, h->stream>>>(nm, ld, k, shpm, rtem, csm);
            h->launches++;
            CKK();
        }
        // 4. major side, all rows
        if (nM > 0) {
            hpf::batch_major_kernel<real, C::lpg, C::vpl><<<row_grid(nM, C::lpg), 256, smem, h->stream>>>(
                (int)nM, ld, k, xM, accM, dirM, shpM, rteM, rateM, stampM, step, csm, csM, priorM, srM, addM, (real)rho,
                blend_all ? 1 : 0);
            h->launches++;
            CKK();
        }
        // 5. minor side: the listed rows (SVI with a caller list), or all rows under the stamp
        const bool walk_all = blend_all || list_minor == nullptr;
        const int64_t nrows5 = walk_all ? nm : n_minor_ids;
        if (nrows5 > 0) {
            hpf::batch_minor_kernel<real, C::lpg, C::vpl><<<row_grid(nrows5, C::lpg), 256, smem, h->stream>>>(
                (int)nrows5, walk_all ? nullptr : list_minor, ld, k, xm, accm, dirm, shpm, rtem, ratem, stampm, step, csM,
                csm, priorm, srm, addm, (real)rho, (real)mult, blend_all ? 1 : 0);
            h->launches++;
            CKK();
        }
        h->batch_colsums_valid = true;
        return HPF_OK;
    });
}

// grow-only device scratch
int grow_bytes(void** p, int64_t* cap_elems, int64_t need_elems, size_t elem) {
    if (*p != nullptr && need_elems <= *cap_elems) return HPF_OK;
    hpf_free(*p);
    *p = nullptr;
    const int64_t n = need_elems + need_elems / 4 + 16;
    CK(hpf_malloc(p, (size_t)n * elem));
    *cap_elems = n;
    return HPF_OK;
}

}  // namespace

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
    TRY(ensure_stamps(h));
    h->batch_step += 1;
    const int step = h->batch_step;

    int *u32 = nullptr, *i32 = nullptr, *ulist = nullptr, *ilist = nullptr, *d_bad = nullptr;
    void* yfree = nullptr;
    const void* yv = nullptr;
    int rc = HPF_OK;
    auto alloc_i = [&](int** p, int64_t n) {
        if (rc == HPF_OK && hpf_malloc(p, sizeof(int) * (size_t)(n > 0 ? n : 1)) != cudaSuccess)
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
        const bool ub = user_batch != 0;
        rc = batch_core(h, u32, i32, yv, nnz, ub ? ulist : ilist, ub ? n_users : n_items, ub ? ilist : ulist,
                        ub ? n_items : n_users, ub, rho, mult, blend_all_rates != 0, step);
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (rc == HPF_OK && e != cudaSuccess) rc = fail(HPF_ECUDA, "minibatch kernels failed: %s", cudaGetErrorString(e));
    hpf_free(u32);
    hpf_free(i32);
    hpf_free(ulist);
    hpf_free(ilist);
    hpf_free(d_bad);
    hpf_free(yfree);
    return rc;
}

extern "C" int hpf_step_batch_ids(hpf_engine* h, const void* ids, int64_t n_ids, int32_t index_bytes,
                                  int32_t user_batch, double rho, double mult, int32_t blend_all_rates) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->state_loaded || !h->mat_valid) return fail(HPF_ESTATE, "no state loaded (call hpf_load_state first)");
    if (!h->data_loaded || !h->A_ptr || !h->B_ptr)
        return fail(HPF_ESTATE, "hpf_step_batch_ids needs triples loaded with a single L2 panel per side "
                                "(set option panel_mb large enough before hpf_load_coo)");
    if (n_ids < 0 || n_ids >= (1ll << 31)) return fail(HPF_EINVAL, "bad n_ids");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (n_ids > 0 && !ids) return fail(HPF_EINVAL, "NULL id list");
    if (!(rho >= 0.0 && rho <= 1.0)) return fail(HPF_EINVAL, "step size must be in [0, 1]");
    DeviceGuard guard(h->device);
    TRY(ensure_stamps(h));
    h->batch_step += 1;
    const int step = h->batch_step;
    const bool ub = user_batch != 0;
    const int64_t n_major = ub ? h->nU : h->nI;
    const int* ptr = ub ? h->A_ptr : h->B_ptr;
    const int* src_minor = ub ? h->A_col : h->B_col;
    const void* src_val = ub ? h->A_val : h->B_val;
    int* stamp_minor = ub ? h->stamp_i : h->stamp_u;

    // ids -> device int32, per-row counts, exclusive scan, total
    if (n_ids + 1 > h->bt_cap_ids || !h->bt_ids) {
        int64_t c1 = 0, c2 = 0, c3 = 0;
        TRY(grow_bytes((void**)&h->bt_ids, &c1, n_ids + 1, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_cnt, &c2, n_ids + 1, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_off, &c3, n_ids + 1, sizeof(int)));
        h->bt_cap_ids = c1;
    }
    int* d_bad = nullptr;
    CK(hpf_malloc(&d_bad, 4));
    cudaMemsetAsync(d_bad, 0, 4, h->stream);
    int rc = stage_index(h, ids, n_ids, index_bytes, n_major, h->bt_ids, d_bad);
    int total = 0;
    if (rc == HPF_OK && n_ids > 0) {
        int bad = 0;
        cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        if (bad) rc = fail(HPF_EINVAL, "batch id out of range");
    }
    if (rc == HPF_OK && n_ids > 0) {
        cudaMemsetAsync(h->bt_cnt + n_ids, 0, sizeof(int), h->stream);
        hpf::batch_count_kernel<<<nblk(n_ids), 256, 0, h->stream>>>((int)n_ids, h->bt_ids, ptr, h->bt_cnt);
        h->launches++;
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, h->bt_cnt, h->bt_off, (int)n_ids + 1, h->stream);
        if (need > h->bt_scan_bytes) {
            hpf_free(h->bt_scan_tmp);
            h->bt_scan_tmp = nullptr;
            h->bt_scan_bytes = 0;
            if (hpf_malloc(&h->bt_scan_tmp, need + 256) != cudaSuccess)
                rc = fail(HPF_ENOMEM, "scan scratch allocation failed");
            else
                h->bt_scan_bytes = need + 256;
        }
        if (rc == HPF_OK) {
            size_t bytes = h->bt_scan_bytes;
            cub::DeviceScan::ExclusiveSum(h->bt_scan_tmp, bytes, h->bt_cnt, h->bt_off, (int)n_ids + 1, h->stream);
            h->launches++;
            cudaMemcpyAsync(&total, h->bt_off + n_ids, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(HPF_ECUDA, "batch scan failed: %s", cudaGetErrorString(e));
        }
    }
    hpf_free(d_bad);
    if (rc != HPF_OK) return rc;

    // compact triples of the batch (row after row, like get_i_batch_pass2) + stamps of the minor ids
    if (total > h->bt_cap_nnz || !h->bt_major) {
        int64_t c1 = 0, c2 = 0, c3 = 0;
        TRY(grow_bytes((void**)&h->bt_major, &c1, total + kPadEntries, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_minor, &c2, total + kPadEntries, sizeof(int)));
        TRY(grow_bytes(&h->bt_val, &c3, total + kPadEntries, (size_t)h->rb));
        h->bt_cap_nnz = c1 - kPadEntries;
    }
    if (total > 0) {
        if (h->rb == 4)
            hpf::batch_expand_kernel<float><<<nblk(total), 256, 0, h->stream>>>(
                (int)n_ids, h->bt_ids, ptr, h->bt_off, total, src_minor, (const float*)src_val, h->bt_major, h->bt_minor,
                (float*)h->bt_val, stamp_minor, step);
        else
            hpf::batch_expand_kernel<double><<<nblk(total), 256, 0, h->stream>>>(
                (int)n_ids, h->bt_ids, ptr, h->bt_off, total, src_minor, (const double*)src_val, h->bt_major, h->bt_minor,
                (double*)h->bt_val, stamp_minor, step);
        if (h->rb == 4) hpf::pad_order_kernel<float><<<nblk(kPadEntries), 256, 0, h->stream>>>(h->bt_major, h->bt_minor, (float*)h->bt_val, total, kPadEntries);
        else hpf::pad_order_kernel<double><<<nblk(kPadEntries), 256, 0, h->stream>>>(h->bt_major, h->bt_minor, (double*)h->bt_val, total, kPadEntries);
        h->launches += 2;
        CKK();
    }
    const int* iu = ub ? h->bt_major : h->bt_minor;
    const int* ii = ub ? h->bt_minor : h->bt_major;
    return batch_core(h, iu, ii, h->bt_val, total, h->bt_ids, n_ids, nullptr, 0, ub, rho, mult, blend_all_rates != 0, step, true);
}


namespace {

// host copies of the CSR / CSC row pointers: minibatch sizes are then known without a device round trip
int ensure_host_ptrs(hpf_engine* h) {
    if (!h->hA_ptr.empty() || !h->hB_ptr.empty()) return HPF_OK;
    h->hA_ptr.resize((size_t)h->nU + 1);
    h->hB_ptr.resize((size_t)h->nI + 1);
    CK(cudaMemcpyAsync(h->hA_ptr.data(), h->A_ptr, sizeof(int) * ((size_t)h->nU + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->hB_ptr.data(), h->B_ptr, sizeof(int) * ((size_t)h->nI + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return HPF_OK;
}

}  // namespace

// One whole SVI epoch (pxi:275-325 or 329-377): the shuffled id list is cut into consecutive minibatches of
// `batch_rows` ids, multiplier n / |batch| (pxi:282 / 334), batch rows only for the hierarchical rates.  The
// list crosses to the device once, minibatch sizes come from host copies of the row pointers, and nothing
// in the loop waits for the device: the call returns with the epoch's kernels in flight (stream-ordered
// with everything that follows), so the caller's shuffle of the next epoch overlaps them.
extern "C" int hpf_step_epoch_ids(hpf_engine* h, const void* ids, int64_t n_ids, int32_t index_bytes,
                                  int64_t batch_rows, int32_t user_batch, double rho) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->state_loaded || !h->mat_valid) return fail(HPF_ESTATE, "no state loaded (call hpf_load_state first)");
    if (!h->data_loaded || !h->A_ptr || !h->B_ptr)
        return fail(HPF_ESTATE, "hpf_step_epoch_ids needs triples loaded with a single L2 panel per side "
                                "(set option panel_mb large enough before hpf_load_coo)");
    const bool ub = user_batch != 0;
    const int64_t n_major = ub ? h->nU : h->nI;
    if (n_ids < 0 || n_ids > n_major) return fail(HPF_EINVAL, "bad n_ids");
    if (batch_rows <= 0) return fail(HPF_EINVAL, "batch_rows must be positive");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (n_ids > 0 && !ids) return fail(HPF_EINVAL, "NULL id list");
    if (!(rho >= 0.0 && rho <= 1.0)) return fail(HPF_EINVAL, "step size must be in [0, 1]");
    if (n_ids == 0) return HPF_OK;
    DeviceGuard guard(h->device);
    TRY(ensure_stamps(h));
    TRY(ensure_host_ptrs(h));
    // host view of the ids (range-checked here, so the device never sees a bad one) in one of two pinned
    // staging buffers: a pageable source would make cudaMemcpyAsync wait for the stream (the previous epoch)
    const int flip = h->ep_flip;
    h->ep_flip ^= 1;
    if (n_ids > h->ep_pin_cap[flip]) {
        if (h->ep_pin[flip]) {
            CK(cudaEventSynchronize(h->ep_ev[flip]));
            CK(cudaFreeHost(h->ep_pin[flip]));
            h->ep_pin[flip] = nullptr;
        }
        CK(cudaHostAlloc((void**)&h->ep_pin[flip], sizeof(int) * (size_t)(n_ids + n_ids / 4 + 16), cudaHostAllocDefault));
        h->ep_pin_cap[flip] = n_ids + n_ids / 4 + 16;
        if (!h->ep_ev[flip]) CK(cudaEventCreateWithFlags(&h->ep_ev[flip], cudaEventDisableTiming));
    } else {
        CK(cudaEventSynchronize(h->ep_ev[flip]));  // the copy that last used this buffer (two epochs ago) is done
    }
    int* hid = h->ep_pin[flip];
    {
        std::vector<char> tmp;
        const void* src = ids;
        if (is_device_ptr(ids)) {
            tmp.resize((size_t)n_ids * index_bytes);
            CK(cudaMemcpy(tmp.data(), ids, tmp.size(), cudaMemcpyDeviceToHost));
            src = tmp.data();
        }
        for (int64_t q = 0; q < n_ids; ++q) {
            const long long v = index_bytes == 8 ? ((const long long*)src)[q] : (long long)((const int*)src)[q];
            if (v < 0 || v >= n_major) return fail(HPF_EINVAL, "batch id out of range");
            hid[(size_t)q] = (int)v;
        }
    }
    const std::vector<int>& hp = ub ? h->hA_ptr : h->hB_ptr;
    const int64_t nb = (n_ids + batch_rows - 1) / batch_rows;
    std::vector<int64_t> totals((size_t)nb, 0);
    int64_t max_total = 0;
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t q0 = b * batch_rows, q1 = q0 + batch_rows < n_ids ? q0 + batch_rows : n_ids;
        int64_t t = 0;
        for (int64_t q = q0; q < q1; ++q) t += hp[(size_t)hid[(size_t)q] + 1] - hp[(size_t)hid[(size_t)q]];
        totals[(size_t)b] = t;
        if (t > max_total) max_total = t;
    }
    // grow-only scratch; growing frees blocks that kernels of an earlier call may still use -> drain first
    const int64_t rows_cap = batch_rows < n_ids ? batch_rows : n_ids;
    const bool grow = n_ids > h->ep_cap_ids || !h->ep_ids || rows_cap + 1 > h->bt_cap_ids || !h->bt_cnt ||
                      max_total > h->bt_cap_nnz || !h->bt_major;
    if (grow) CK(cudaStreamSynchronize(h->stream));
    if (n_ids > h->ep_cap_ids || !h->ep_ids) TRY(grow_bytes((void**)&h->ep_ids, &h->ep_cap_ids, n_ids, sizeof(int)));
    if (rows_cap + 1 > h->bt_cap_ids || !h->bt_cnt) {
        int64_t c1 = 0, c2 = 0, c3 = 0;
        TRY(grow_bytes((void**)&h->bt_ids, &c1, rows_cap + 1, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_cnt, &c2, rows_cap + 1, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_off, &c3, rows_cap + 1, sizeof(int)));
        h->bt_cap_ids = c1;
    }
    if (max_total > h->bt_cap_nnz || !h->bt_major) {
        int64_t c1 = 0, c2 = 0, c3 = 0;
        TRY(grow_bytes((void**)&h->bt_major, &c1, max_total + kPadEntries, sizeof(int)));
        TRY(grow_bytes((void**)&h->bt_minor, &c2, max_total + kPadEntries, sizeof(int)));
        TRY(grow_bytes(&h->bt_val, &c3, max_total + kPadEntries, (size_t)h->rb));
        h->bt_cap_nnz = c1 - kPadEntries;
    }
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, h->bt_cnt, h->bt_off, (int)rows_cap + 1, h->stream);
    if (need > h->bt_scan_bytes) {
        CK(cudaStreamSynchronize(h->stream));
        hpf_free(h->bt_scan_tmp);
        h->bt_scan_tmp = nullptr;
        h->bt_scan_bytes = 0;
        CK(hpf_malloc(&h->bt_scan_tmp, need + 256));
        h->bt_scan_bytes = need + 256;
    }
    CK(cudaMemcpyAsync(h->ep_ids, hid, sizeof(int) * (size_t)n_ids, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ep_ev[flip], h->stream));

    const int* ptr = ub ? h->A_ptr : h->B_ptr;
    const int* src_minor = ub ? h->A_col : h->B_col;
    const void* src_val = ub ? h->A_val : h->B_val;
    int* stamp_minor = ub ? h->stamp_i : h->stamp_u;
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t q0 = b * batch_rows, q1 = q0 + batch_rows < n_ids ? q0 + batch_rows : n_ids;
        const int64_t nq = q1 - q0, total = totals[(size_t)b];
        const int* d_ids = h->ep_ids + q0;
        h->batch_step += 1;
        const int step = h->batch_step;
        CK(cudaMemsetAsync(h->bt_cnt + nq, 0, sizeof(int), h->stream));
        hpf::batch_count_kernel<<<nblk(nq), 256, 0, h->stream>>>((int)nq, d_ids, ptr, h->bt_cnt);
        size_t bytes = h->bt_scan_bytes;
        cub::DeviceScan::ExclusiveSum(h->bt_scan_tmp, bytes, h->bt_cnt, h->bt_off, (int)nq + 1, h->stream);
        h->launches += 2;
        if (total > 0) {
            if (h->rb == 4)
                hpf::batch_expand_kernel<float><<<nblk(total), 256, 0, h->stream>>>(
                    (int)nq, d_ids, ptr, h->bt_off, total, src_minor, (const float*)src_val, h->bt_major, h->bt_minor,
                    (float*)h->bt_val, stamp_minor, step);
            else
                hpf::batch_expand_kernel<double><<<nblk(total), 256, 0, h->stream>>>(
                    (int)nq, d_ids, ptr, h->bt_off, total, src_minor, (const double*)src_val, h->bt_major, h->bt_minor,
                    (double*)h->bt_val, stamp_minor, step);
            if (h->rb == 4) hpf::pad_order_kernel<float><<<nblk(kPadEntries), 256, 0, h->stream>>>(h->bt_major, h->bt_minor, (float*)h->bt_val, total, kPadEntries);
            else hpf::pad_order_kernel<double><<<nblk(kPadEntries), 256, 0, h->stream>>>(h->bt_major, h->bt_minor, (double*)h->bt_val, total, kPadEntries);
            h->launches += 2;
        }
        CKK();
        const int* iu = ub ? h->bt_major : h->bt_minor;
        const int* ii = ub ? h->bt_minor : h->bt_major;
        const double mult = (double)n_major / (double)nq;
        TRY(batch_core(h, iu, ii, h->bt_val, total, d_ids, nq, nullptr, 0, ub, rho, mult, false, step, true));
    }
    return HPF_OK;
}
