// Device-side ingest of the training triples (included at the end of hpf_engine.cu; same translation unit): the two
// pandas / scipy steps HPF._process_data and HPF._store_metadata run on the host before a fit,
//   hpf_factorize       pd.factorize of an integer id column           hpfrec/__init__.py:478-479
//   hpf_csr_metadata    coo_array(...).tocsr() -> indptr / indices     hpfrec/__init__.py:587-606
// as radix sorts (CUB) plus a few one-pass kernels.  Stateless: plain pointers in, plain pointers out, host or device.

namespace hpf {

__global__ void ingest_keys_kernel(const void* __restrict__ values, int value_bytes, long long n,
                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ pos) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    keys[j] = value_bytes == 8 ? ((const unsigned long long*)values)[j] : (unsigned long long)((const unsigned*)values)[j];
    pos[j] = (unsigned)j;
}
// head[j] = 1 where a run of equal keys starts
__global__ void ingest_heads_kernel(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ head) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}
// per run: the key and (stable sort => smallest) original position of its first element
__global__ void ingest_runs_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ pos,
                                   const int* __restrict__ head, const int* __restrict__ run1, long long n,
                                   unsigned long long* __restrict__ run_key, unsigned* __restrict__ run_first,
                                   unsigned* __restrict__ run_id) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !head[j]) return;
    const int r = run1[j] - 1;
    run_key[r] = keys[j];
    if (run_first) run_first[r] = pos[j];
    if (run_id) run_id[r] = (unsigned)r;
}
// runs ordered by first appearance: code of run order[c] is c; uniques[c] = its key
__global__ void ingest_rank_kernel(const unsigned* __restrict__ order, const unsigned long long* __restrict__ run_key,
                                   int nruns, int value_bytes, unsigned* __restrict__ rank, void* __restrict__ uniques) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nruns) return;
    const unsigned r = order[c];
    rank[r] = (unsigned)c;
    if (value_bytes == 8) ((unsigned long long*)uniques)[c] = run_key[r];
    else ((unsigned*)uniques)[c] = (unsigned)run_key[r];
}
__global__ void ingest_codes_kernel(const unsigned* __restrict__ pos, const int* __restrict__ run1,
                                    const unsigned* __restrict__ rank, long long n, int code_bytes, void* __restrict__ codes) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned c = rank[run1[j] - 1];
    if (code_bytes == 8) ((long long*)codes)[pos[j]] = (long long)c;
    else ((int*)codes)[pos[j]] = (int)c;
}
template <typename IT>
__global__ void ingest_pair_keys_kernel(const IT* __restrict__ iu, const IT* __restrict__ ii, long long n, long long nU,
                                        long long nI, unsigned long long* __restrict__ keys, int* __restrict__ bad) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long u = (long long)iu[j], i = (long long)ii[j];
    if (u < 0 || u >= nU || i < 0 || i >= nI) {
        *bad = 1;
        keys[j] = 0;
        return;
    }
    keys[j] = (unsigned long long)u * (unsigned long long)nI + (unsigned long long)i;
}
__global__ void ingest_split_kernel(const unsigned long long* __restrict__ run_key, int nruns, long long nI,
                                    int* __restrict__ urow, void* __restrict__ indices, int out_bytes) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const unsigned long long key = run_key[r];
    urow[r] = (int)(key / (unsigned long long)nI);
    const long long i = (long long)(key % (unsigned long long)nI);
    if (out_bytes == 8) ((long long*)indices)[r] = i;
    else ((int*)indices)[r] = (int)i;
}
__global__ void ingest_widen_ptr_kernel(const int* __restrict__ ptr, long long n, long long* __restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = ptr[j];
}

}  // namespace hpf

namespace {

// scratch of one ingest call: everything is released by the destructor, whatever path the call takes
struct IngestScratch {
    std::vector<void*> blocks;
    template <typename T>
    bool get(T** p, size_t bytes) {
        void* q = nullptr;
        if (hpf_malloc(&q, bytes > 0 ? bytes : 16) != cudaSuccess) return false;
        blocks.push_back(q);
        *p = (T*)q;
        return true;
    }
    ~IngestScratch() {
        for (void* q : blocks) hpf_free(q);
    }
};

// host or device input -> device pointer (staged copy when it is host memory)
bool ingest_in(IngestScratch& s, const void* src, size_t bytes, const void** dev) {
    if (bytes == 0 || is_device_ptr(src)) {
        *dev = src;
        return true;
    }
    void* tmp = nullptr;
    if (!s.get(&tmp, bytes)) return false;
    if (cudaMemcpy(tmp, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    *dev = tmp;
    return true;
}
bool ingest_out(void* dst, const void* dev_src, size_t bytes) {
    if (bytes == 0) return true;
    return cudaMemcpy(dst, dev_src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost) == cudaSuccess;
}

// runs of equal keys in a sorted key array: head flags + their inclusive scan (run number + 1); returns the run count
int ingest_runs(IngestScratch& s, const unsigned long long* keys, int64_t n, int** head, int** run1, int* nruns) {
    if (!s.get(head, 4 * (size_t)n) || !s.get(run1, 4 * (size_t)n)) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    hpf::ingest_heads_kernel<<<nblk(n), 256>>>(keys, n, *head);
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, *head, *run1, (int)n);
    void* tmp = nullptr;
    if (!s.get(&tmp, tb)) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    CK(cub::DeviceScan::InclusiveSum(tmp, tb, *head, *run1, (int)n));
    CK(cudaMemcpy(nruns, *run1 + (n - 1), 4, cudaMemcpyDeviceToHost));
    return HPF_OK;
}

}  // namespace

extern "C" int hpf_factorize(int32_t device, const void* values, int64_t n, int32_t value_bytes, void* codes_out,
                             int32_t code_bytes, void* uniques_out, int64_t* n_unique) {
    if (n < 0 || n >= (1ll << 31)) return fail(HPF_EINVAL, "n must be in [0, 2^31)");
    if (value_bytes != 4 && value_bytes != 8) return fail(HPF_EINVAL, "value_bytes must be 4 or 8");
    if (code_bytes != 4 && code_bytes != 8) return fail(HPF_EINVAL, "code_bytes must be 4 or 8");
    if (!n_unique || (n > 0 && (!values || !codes_out || !uniques_out))) return fail(HPF_EINVAL, "NULL argument");
    *n_unique = 0;
    if (n == 0) return HPF_OK;
    DeviceGuard guard(device);
    IngestScratch s;
    const void* d_vals = nullptr;
    if (!ingest_in(s, values, (size_t)n * value_bytes, &d_vals)) return fail(HPF_ENOMEM, "staging the id column failed");
    unsigned long long *k_in, *k_out, *run_key;
    unsigned *p_in, *p_out;
    if (!s.get(&k_in, 8 * (size_t)n) || !s.get(&k_out, 8 * (size_t)n) || !s.get(&p_in, 4 * (size_t)n) || !s.get(&p_out, 4 * (size_t)n))
        return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    hpf::ingest_keys_kernel<<<nblk(n), 256>>>(d_vals, value_bytes, n, k_in, p_in);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, p_in, p_out, (int)n, 0, value_bytes * 8);
    void* tmp = nullptr;
    if (!s.get(&tmp, tb)) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, p_in, p_out, (int)n, 0, value_bytes * 8));  // stable
    int *head, *run1, nruns = 0;
    TRY(ingest_runs(s, k_out, n, &head, &run1, &nruns));
    unsigned *run_first, *run_id, *first_sorted, *order, *rank;
    if (!s.get(&run_key, 8 * (size_t)nruns) || !s.get(&run_first, 4 * (size_t)nruns) || !s.get(&run_id, 4 * (size_t)nruns) ||
        !s.get(&first_sorted, 4 * (size_t)nruns) || !s.get(&order, 4 * (size_t)nruns) || !s.get(&rank, 4 * (size_t)nruns))
        return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    hpf::ingest_runs_kernel<<<nblk(n), 256>>>(k_out, p_out, head, run1, n, run_key, run_first, run_id);
    // order of first appearance: sort the runs by the position of their first element
    size_t tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, run_first, first_sorted, run_id, order, nruns, 0, 32);
    void* tmp2 = nullptr;
    if (!s.get(&tmp2, tb2)) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    CK(cub::DeviceRadixSort::SortPairs(tmp2, tb2, run_first, first_sorted, run_id, order, nruns, 0, 32));
    void *d_uniq, *d_codes;
    if (!s.get(&d_uniq, (size_t)nruns * value_bytes) || !s.get(&d_codes, (size_t)n * code_bytes))
        return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    hpf::ingest_rank_kernel<<<nblk(nruns), 256>>>(order, run_key, nruns, value_bytes, rank, d_uniq);
    hpf::ingest_codes_kernel<<<nblk(n), 256>>>(p_out, run1, rank, n, code_bytes, d_codes);
    CK(cudaGetLastError());
    if (!ingest_out(codes_out, d_codes, (size_t)n * code_bytes) || !ingest_out(uniques_out, d_uniq, (size_t)nruns * value_bytes))
        return fail(HPF_ECUDA, "copying the factorize result out failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaDeviceSynchronize());
    *n_unique = nruns;
    return HPF_OK;
}

extern "C" int hpf_csr_metadata(int32_t device, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes, int64_t nU,
                                int64_t nI, int64_t* indptr_out, void* indices_out, int32_t out_index_bytes, int64_t* n_out) {
    if (n < 0 || n >= (1ll << 31)) return fail(HPF_EINVAL, "n must be in [0, 2^31)");
    if (nU < 0 || nI <= 0 || nU >= (1ll << 31) || nI >= (1ll << 31)) return fail(HPF_EINVAL, "bad matrix shape");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (out_index_bytes != 4 && out_index_bytes != 8) return fail(HPF_EINVAL, "out_index_bytes must be 4 or 8");
    if (!indptr_out || !n_out || (n > 0 && (!ix_u || !ix_i || !indices_out))) return fail(HPF_EINVAL, "NULL argument");
    DeviceGuard guard(device);
    IngestScratch s;
    *n_out = 0;
    long long* d_ptr64;
    if (!s.get(&d_ptr64, 8 * (size_t)(nU + 1))) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    if (n == 0) {
        CK(cudaMemset(d_ptr64, 0, 8 * (size_t)(nU + 1)));
        if (!ingest_out(indptr_out, d_ptr64, 8 * (size_t)(nU + 1))) return fail(HPF_ECUDA, "copy out failed");
        return HPF_OK;
    }
    const void *d_u = nullptr, *d_i = nullptr;
    if (!ingest_in(s, ix_u, (size_t)n * index_bytes, &d_u) || !ingest_in(s, ix_i, (size_t)n * index_bytes, &d_i))
        return fail(HPF_ENOMEM, "staging the index columns failed");
    unsigned long long *k_in, *k_out, *run_key;
    int* d_bad;
    if (!s.get(&k_in, 8 * (size_t)n) || !s.get(&k_out, 8 * (size_t)n) || !s.get(&d_bad, 4))
        return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    CK(cudaMemset(d_bad, 0, 4));
    if (index_bytes == 8)
        hpf::ingest_pair_keys_kernel<long long><<<nblk(n), 256>>>((const long long*)d_u, (const long long*)d_i, n, nU, nI, k_in, d_bad);
    else
        hpf::ingest_pair_keys_kernel<int><<<nblk(n), 256>>>((const int*)d_u, (const int*)d_i, n, nU, nI, k_in, d_bad);
    int bad = 0;
    CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    if (bad) return fail(HPF_EINVAL, "index out of range in the triples");
    int end_bit = 1;
    while (end_bit < 64 && (((unsigned long long)nU * (unsigned long long)nI - 1ull) >> end_bit)) ++end_bit;
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, k_in, k_out, (int)n, 0, end_bit);
    void* tmp = nullptr;
    if (!s.get(&tmp, tb)) return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    CK(cub::DeviceRadixSort::SortKeys(tmp, tb, k_in, k_out, (int)n, 0, end_bit));
    int *head, *run1, nruns = 0;
    TRY(ingest_runs(s, k_out, n, &head, &run1, &nruns));  // duplicates of a (user, item) pair collapse, like sum_duplicates
    int *urow, *d_ptr;
    void* d_idx;
    if (!s.get(&run_key, 8 * (size_t)nruns) || !s.get(&urow, 4 * (size_t)nruns) || !s.get(&d_ptr, 4 * (size_t)(nU + 1)) ||
        !s.get(&d_idx, (size_t)nruns * out_index_bytes))
        return fail(HPF_ENOMEM, "ingest scratch allocation failed");
    hpf::ingest_runs_kernel<<<nblk(n), 256>>>(k_out, nullptr, head, run1, n, run_key, nullptr, nullptr);
    hpf::ingest_split_kernel<<<nblk(nruns), 256>>>(run_key, nruns, nI, urow, d_idx, out_index_bytes);
    hpf::row_ptr_kernel<<<nblk((int64_t)nruns + 1), 256>>>(urow, nruns, (int)nU, d_ptr);
    hpf::ingest_widen_ptr_kernel<<<nblk(nU + 1), 256>>>(d_ptr, nU + 1, d_ptr64);
    CK(cudaGetLastError());
    if (!ingest_out(indptr_out, d_ptr64, 8 * (size_t)(nU + 1)) || !ingest_out(indices_out, d_idx, (size_t)nruns * out_index_bytes))
        return fail(HPF_ECUDA, "copying the CSR arrays out failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaDeviceSynchronize());
    *n_out = nruns;
    return HPF_OK;
}
