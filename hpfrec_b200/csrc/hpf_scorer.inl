// Device-resident scoring of a fitted model (included at the end of hpf_engine.cu; same translation unit).
// A scorer holds only the two expectation matrices Theta (nU x k) and Beta (nI x k) in the engine's padded row
// layout, so predict / eval_llk / topN of the reference's HPF class (hpfrec/__init__.py:1198-1446) run without
// re-uploading them per call:
//   hpf_scorer_predict   predict_multiple   cython_loops.pxi:803-810
//   hpf_scorer_llk       llk_plus_rmse + sum_prediction   pxi:627-658, 816-825 (behind calc_llk, pxi:525-534)
//   hpf_scorer_topn      HPF.topN   hpfrec/__init__.py:1296-1396 (numpy argpartition / setdiff1d / argsort there)

struct hpf_scorer {
    int device = 0;
    int64_t nU = 0, nI = 0;
    int k = 0, ld = 0, rb = 4;
    void *theta = nullptr, *beta = nullptr;
    // topN scratch (grow-only)
    void *scores = nullptr, *scores_sorted = nullptr;
    int *ids = nullptr, *ids_sorted = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_bytes = 0;
    int64_t cap = 0;
};

namespace hpf {

// scores[q] = -(theta_row . beta[pool ? pool[q] : q]); ids[q] = the item row.  Negated so that an ascending
// radix sort lists the best items first and equal scores keep ascending item order.
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
topn_scores_kernel(const real* __restrict__ theta_row, const real* __restrict__ beta, int ld, int64_t n,
                   const int* __restrict__ pool, real* __restrict__ scores, int* __restrict__ ids) {
    constexpr int EPV = Pack<real>::N;
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int64_t g0 = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int64_t gstride = (int64_t)gridDim.x * groups_per_block;
    Pack<real> t[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        act[v] = (gl + LPG * v) * EPV < ld;
        t[v] = act[v] ? ldg_pack(theta_row + (gl + LPG * v) * EPV) : pack_zero<real>();
    }
    const int64_t trips = (n + gstride - 1) / gstride;
    for (int64_t it = 0; it < trips; ++it) {
        const int64_t q = g0 + it * gstride;
        const bool live = q < n;
        const int r = live ? (pool ? pool[q] : (int)q) : 0;
        real s = real(0);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const Pack<real> b = ldg_pack(beta + (size_t)r * ld + (gl + LPG * v) * EPV);
#pragma unroll
            for (int e = 0; e < EPV; ++e) s = fma(t[v].v[e], b.v[e], s);
        }
        s = group_sum<LPG>(s, gmask);
        if (live && gl == 0) {
            scores[q] = -s;
            ids[q] = r;
        }
    }
}

// items the user has already seen (sorted or not) drop to the end of the ranking
template <typename real>
__global__ void topn_mask_seen_kernel(const int* __restrict__ seen, int64_t n_seen, const int* __restrict__ pool_pos,
                                      real* __restrict__ scores) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_seen) return;
    const int at = pool_pos ? pool_pos[q] : seen[q];
    if (at >= 0) scores[at] = INFINITY;
}

}  // namespace hpf

namespace {

template <typename F>
int scorer_dispatch(hpf_scorer* s, F&& f) {
    return dispatch(s->rb, s->ld, f);
}

int scorer_grow(hpf_scorer* s, int64_t n) {
    if (n <= s->cap && s->scores) return HPF_OK;
    void** ps[] = {&s->scores, &s->scores_sorted, (void**)&s->ids, (void**)&s->ids_sorted, &s->sort_tmp};
    for (void** p : ps) {
        hpf_free(*p);
        *p = nullptr;
    }
    const int64_t c = n + n / 4 + 16;
    CK(hpf_malloc(&s->scores, (size_t)c * s->rb));
    CK(hpf_malloc(&s->scores_sorted, (size_t)c * s->rb));
    CK(hpf_malloc(&s->ids, sizeof(int) * (size_t)c));
    CK(hpf_malloc(&s->ids_sorted, sizeof(int) * (size_t)c));
    size_t need = 0;
    if (s->rb == 4)
        cub::DeviceRadixSort::SortPairs(nullptr, need, (const float*)s->scores, (float*)s->scores_sorted, s->ids, s->ids_sorted, (int)c);
    else
        cub::DeviceRadixSort::SortPairs(nullptr, need, (const double*)s->scores, (double*)s->scores_sorted, s->ids, s->ids_sorted, (int)c);
    CK(hpf_malloc(&s->sort_tmp, need + 256));
    s->sort_bytes = need + 256;
    s->cap = c;
    return HPF_OK;
}

// a view of the scorer's matrices as an engine-shaped object for the shared scoring kernels
struct ScoreView {
    int ld, k, rb;
};

}  // namespace

extern "C" {

int hpf_scorer_create(hpf_scorer** out, const void* Theta, const void* Beta, int64_t nU, int64_t nI, int32_t k,
                      int32_t real_bytes, int32_t device) {
    if (!out) return fail(HPF_EINVAL, "out is NULL");
    *out = nullptr;
    if (!Theta || !Beta) return fail(HPF_EINVAL, "NULL factor matrix");
    if (k <= 0 || (real_bytes != 4 && real_bytes != 8)) return fail(HPF_EINVAL, "bad k / real_bytes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(HPF_ECUDA, "no such CUDA device %d (this library has no CPU path)", device);
    }
    int ld = 0, kw = 0;
    TRY(row_layout(k, real_bytes, &ld, &kw));   // the engine's row layout
    TRY(dispatch(real_bytes, ld, [](auto) { return HPF_OK; }));
    if (nU < 0 || nI < 0 || nU >= (1ll << 31) || nI >= (1ll << 31)) return fail(HPF_EINVAL, "bad shape");
    DeviceGuard guard(device);
    hpf_scorer* s = new hpf_scorer();
    s->device = device;
    s->nU = nU;
    s->nI = nI;
    s->k = k;
    s->ld = ld;
    s->rb = real_bytes;
    int rc = HPF_OK;
    if (hpf_malloc(&s->theta, (size_t)(nU > 0 ? nU : 1) * ld * real_bytes) != cudaSuccess ||
        hpf_malloc(&s->beta, (size_t)(nI > 0 ? nI : 1) * ld * real_bytes) != cudaSuccess)
        rc = fail(HPF_ENOMEM, "device allocation failed in hpf_scorer_create");
    if (rc == HPF_OK) {
        hpf_engine tmp;  // only the fields upload_matrix reads
        tmp.device = device;
        tmp.stream = nullptr;
        if (real_bytes == 4) {
            rc = upload_matrix<float>(&tmp, Theta, s->theta, nU, k, ld, 0.0f);
            if (rc == HPF_OK) rc = upload_matrix<float>(&tmp, Beta, s->beta, nI, k, ld, 0.0f);
        } else {
            rc = upload_matrix<double>(&tmp, Theta, s->theta, nU, k, ld, 0.0);
            if (rc == HPF_OK) rc = upload_matrix<double>(&tmp, Beta, s->beta, nI, k, ld, 0.0);
        }
        if (rc == HPF_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(HPF_ECUDA, "upload failed");
    }
    if (rc != HPF_OK) {
        std::string keep = g_err;
        hpf_scorer_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return HPF_OK;
}

int hpf_scorer_destroy(hpf_scorer* s) {
    if (!s) return HPF_OK;
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {s->theta, s->beta, s->scores, s->scores_sorted, s->ids, s->ids_sorted, s->sort_tmp};
    for (void* p : ptrs) hpf_free(p);
    delete s;
    return HPF_OK;
}

// shared by predict and llk: out3 = {sum Y log yhat [- lgamma(Y+1)], sum (Y - yhat)^2, sum yhat}
static int scorer_score(hpf_scorer* s, const void* ix_u, const void* ix_i, const void* Y, int64_t n, int32_t index_bytes,
                        int full_llk, double* out3, void* pred_out) {
    if (n < 0 || n >= (1ll << 31)) return fail(HPF_EINVAL, "n out of range");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    DeviceGuard guard(s->device);
    hpf_engine tmp;  // stage_index / stage_in only use device, stream, launches
    tmp.device = s->device;
    tmp.stream = nullptr;
    int *u32 = nullptr, *i32 = nullptr, *d_bad = nullptr;
    void *yfree = nullptr, *pred_dev = nullptr;
    double* d_sums = nullptr;
    const void* yv = nullptr;
    const size_t n1 = (size_t)(n > 0 ? n : 1);
    int rc = HPF_OK;
    if (hpf_malloc(&u32, 4 * n1) != cudaSuccess || hpf_malloc(&i32, 4 * n1) != cudaSuccess || hpf_malloc(&d_bad, 4) != cudaSuccess ||
        hpf_malloc((void**)&d_sums, sizeof(double) * 4) != cudaSuccess)
        rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK) {
        cudaMemsetAsync(d_bad, 0, 4, nullptr);
        cudaMemsetAsync(d_sums, 0, sizeof(double) * 4, nullptr);
        rc = stage_index(&tmp, ix_u, n, index_bytes, s->nU, u32, d_bad);
    }
    if (rc == HPF_OK) rc = stage_index(&tmp, ix_i, n, index_bytes, s->nI, i32, d_bad);
    if (rc == HPF_OK && Y) rc = stage_in(&tmp, Y, (size_t)n * s->rb, &yv, &yfree);
    if (rc == HPF_OK) {
        int bad = 0;
        cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
        if (bad) rc = fail(HPF_EINVAL, "index out of range");
    }
    const bool pred_is_dev = pred_out && is_device_ptr(pred_out);
    if (rc == HPF_OK && pred_out && !pred_is_dev && hpf_malloc(&pred_dev, n1 * s->rb) != cudaSuccess)
        rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK && n > 0) {
        rc = scorer_dispatch(s, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            const int gpb = 256 / C::lpg;
            long long want = (n + gpb - 1) / gpb;
            if (want > 148 * 16) want = 148 * 16;
            hpf::score_kernel<real, C::lpg, C::vpl><<<(unsigned)want, 256>>>(
                u32, i32, (const real*)yv, n, (const real*)s->theta, (const real*)s->beta, s->ld, full_llk,
                out3 ? d_sums : nullptr, (real*)(pred_out ? (pred_is_dev ? pred_out : pred_dev) : nullptr));
            CKK();
            return HPF_OK;
        });
    }
    if (rc == HPF_OK && out3 && cudaMemcpy(out3, d_sums, sizeof(double) * 3, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(HPF_ECUDA, "D2H failed");
    if (rc == HPF_OK && pred_dev && n > 0 && cudaMemcpy(pred_out, pred_dev, (size_t)n * s->rb, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(HPF_ECUDA, "D2H of predictions failed");
    if (rc == HPF_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(HPF_ECUDA, "scoring kernels failed");
    hpf_free(u32);
    hpf_free(i32);
    hpf_free(d_bad);
    hpf_free(d_sums);
    hpf_free(yfree);
    hpf_free(pred_dev);
    return rc;
}

int hpf_scorer_predict(hpf_scorer* s, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes, void* out) {
    if (!s || (n > 0 && (!ix_u || !ix_i || !out))) return fail(HPF_EINVAL, "NULL argument");
    return scorer_score(s, ix_u, ix_i, nullptr, n, index_bytes, 0, nullptr, out);
}

int hpf_scorer_llk(hpf_scorer* s, const void* ix_u, const void* ix_i, const void* Y, int64_t n, int32_t index_bytes,
                   int32_t full_llk, double out[3]) {
    if (!s || !out || (n > 0 && (!ix_u || !ix_i || !Y))) return fail(HPF_EINVAL, "NULL argument");
    return scorer_score(s, ix_u, ix_i, Y, n, index_bytes, full_llk, out, nullptr);
}

int hpf_scorer_topn(hpf_scorer* s, int64_t user, int32_t n, const void* pool, int64_t n_pool, const void* seen,
                    int64_t n_seen, int32_t index_bytes, int64_t* out_ids, void* out_scores, int32_t* n_out) {
    if (!s || !out_ids || !n_out) return fail(HPF_EINVAL, "NULL argument");
    if (user < 0 || user >= s->nU) return fail(HPF_EINVAL, "user row out of range");
    if (n < 0 || n_pool < 0 || n_seen < 0) return fail(HPF_EINVAL, "negative count");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if ((n_pool > 0 && !pool) || (n_seen > 0 && !seen)) return fail(HPF_EINVAL, "NULL id list");
    DeviceGuard guard(s->device);
    const int64_t m = pool ? n_pool : s->nI;
    *n_out = 0;
    if (m == 0 || n == 0) return HPF_OK;
    TRY(scorer_grow(s, m));
    hpf_engine tmp;
    tmp.device = s->device;
    tmp.stream = nullptr;
    int *d_pool = nullptr, *d_seen = nullptr, *d_pos = nullptr, *d_bad = nullptr;
    int rc = HPF_OK;
    std::vector<int> pos;
    if (hpf_malloc(&d_bad, 4) != cudaSuccess) rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK) cudaMemsetAsync(d_bad, 0, 4, nullptr);
    if (rc == HPF_OK && pool) {
        if (hpf_malloc(&d_pool, sizeof(int) * (size_t)n_pool) != cudaSuccess) rc = fail(HPF_ENOMEM, "device allocation failed");
        if (rc == HPF_OK) rc = stage_index(&tmp, pool, n_pool, index_bytes, s->nI, d_pool, d_bad);
    }
    if (rc == HPF_OK && n_seen > 0) {
        if (hpf_malloc(&d_seen, sizeof(int) * (size_t)n_seen) != cudaSuccess) rc = fail(HPF_ENOMEM, "device allocation failed");
        if (rc == HPF_OK) rc = stage_index(&tmp, seen, n_seen, index_bytes, s->nI, d_seen, d_bad);
    }
    if (rc == HPF_OK) {
        int bad = 0;
        cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
        if (bad) rc = fail(HPF_EINVAL, "item id out of range");
    }
    // with a pool, a seen item masks its POSITIONS in the pool: resolved on the host (both lists are small
    // next to the model: the pool is what the caller just handed over)
    if (rc == HPF_OK && pool && n_seen > 0) {
        std::vector<long long> hp((size_t)n_pool), hs((size_t)n_seen);
        auto fetch = [&](const void* src, int64_t cnt, std::vector<long long>& dst) {
            std::vector<char> raw((size_t)cnt * index_bytes);
            cudaMemcpy(raw.data(), src, raw.size(), cudaMemcpyDefault);
            for (int64_t q = 0; q < cnt; ++q)
                dst[(size_t)q] = index_bytes == 8 ? ((const long long*)raw.data())[q] : (long long)((const int*)raw.data())[q];
        };
        fetch(pool, n_pool, hp);
        fetch(seen, n_seen, hs);
        std::unordered_map<long long, char> is_seen;
        for (long long v : hs) is_seen[v] = 1;
        for (int64_t q = 0; q < n_pool; ++q)
            if (is_seen.count(hp[(size_t)q])) pos.push_back((int)q);
        if (!pos.empty()) {
            if (hpf_malloc(&d_pos, sizeof(int) * pos.size()) != cudaSuccess) rc = fail(HPF_ENOMEM, "device allocation failed");
            else cudaMemcpy(d_pos, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice);
        }
    }
    if (rc == HPF_OK) {
        rc = scorer_dispatch(s, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            const int gpb = 256 / C::lpg;
            long long want = (m + gpb - 1) / gpb;
            if (want > 148 * 8) want = 148 * 8;
            hpf::topn_scores_kernel<real, C::lpg, C::vpl><<<(unsigned)want, 256>>>(
                (const real*)s->theta + (size_t)user * s->ld, (const real*)s->beta, s->ld, m, d_pool, (real*)s->scores, s->ids);
            if (n_seen > 0 && !pool)
                hpf::topn_mask_seen_kernel<real><<<nblk(n_seen), 256>>>(d_seen, n_seen, nullptr, (real*)s->scores);
            else if (!pos.empty())
                hpf::topn_mask_seen_kernel<real><<<nblk((long long)pos.size()), 256>>>(nullptr, (int64_t)pos.size(), d_pos, (real*)s->scores);
            CKK();
            size_t bytes = s->sort_bytes;
            cudaError_t e = cub::DeviceRadixSort::SortPairs(s->sort_tmp, bytes, (const real*)s->scores, (real*)s->scores_sorted,
                                                            s->ids, s->ids_sorted, (int)m);
            if (e != cudaSuccess) return fail(HPF_ECUDA, "radix sort failed: %s", cudaGetErrorString(e));
            return HPF_OK;
        });
    }
    if (rc == HPF_OK) {
        const int64_t take = n < m ? n : m;
        std::vector<int> hid((size_t)take);
        std::vector<char> hsc((size_t)take * s->rb);
        if (cudaMemcpy(hid.data(), s->ids_sorted, sizeof(int) * (size_t)take, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(hsc.data(), s->scores_sorted, hsc.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(HPF_ECUDA, "D2H of the ranking failed");
        int cnt = 0;
        for (int64_t q = 0; rc == HPF_OK && q < take; ++q) {
            const double sc = s->rb == 4 ? (double)((const float*)hsc.data())[q] : ((const double*)hsc.data())[q];
            if (!(sc < INFINITY)) break;  // masked (seen) items sort last
            out_ids[cnt] = hid[(size_t)q];
            if (out_scores) {
                if (s->rb == 4) ((float*)out_scores)[cnt] = (float)(-sc);
                else ((double*)out_scores)[cnt] = -sc;
            }
            ++cnt;
        }
        *n_out = cnt;
    }
    hpf_free(d_pool);
    hpf_free(d_seen);
    hpf_free(d_pos);
    hpf_free(d_bad);
    return rc;
}

}  // extern "C"
