// Host side of the HPF engine + the extern "C" boundary declared in include/hpf_b200.h.
// Built by hpfrec_b200/build.py:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared ...
// No torch, no CPU compute path: every entry point either runs CUDA kernels or fails with a code.
#include "../../include/hpf_b200.h"
#include "hpf_kernels.cuh"
#include "hpf_batch.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? HPF_ENOMEM : HPF_ECUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                        \
    } while (0)
#define CKK() CK(cudaGetLastError())
#define TRY(expr)                      \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != HPF_OK) return rc_; \
    } while (0)

// -------------------------------------------------------------------------------------------------
// Measured defaults (B200, 1M x 380K x 48M nnz; profiles/).  Everything a tuning run can change lives
// in this block; each value is also an option (hpf_set_option / HPF_OPTIONS / HPF_ROW_ALIGN), so the
// test-suite and the bench can be run under a candidate configuration before it becomes the default.
// -------------------------------------------------------------------------------------------------
constexpr int kDefaultRowAlign = 128;    // bytes; row stride rule in hpf_create (cache-line aligned rows)
constexpr double kDefaultPanelMb = 96.0;  // L2 panel of the gathered factor side
constexpr int kDefaultChunk = 256;       // nnz walked by one lane group
constexpr int kMaxChunk = 4096;
constexpr int kDefaultSweepMode = 0;     // 0 two-pass sweep_rows_kernel, 1 single-pass sweep_coo_kernel (cross-check)
// The triple arrays of an ordering are padded with zero-count entries so that every lane group of a launch
// owns a whole chunk and every warp whole groups (at most 8 groups per warp): < 9 chunks of padding.
constexpr int kPadEntries = 9 * kMaxChunk;

template <typename real_, int LPG, int VPL>
struct Cfg {
    using real = real_;
    static constexpr int lpg = LPG, vpl = VPL;
};

// lane-group shape from the padded row length (16-byte packs per row)
template <typename F>
int dispatch(int real_bytes, int ld, F&& f) {
    const int packs = ld * real_bytes / 16;
    if (real_bytes == 4) {
        if (packs <= 8) return f(Cfg<float, 8, 1>{});
        if (packs <= 16) return f(Cfg<float, 8, 2>{});
        if (packs <= 32) return f(Cfg<float, 16, 2>{});
        if (packs <= 64) return f(Cfg<float, 32, 2>{});
        if (packs <= 128) return f(Cfg<float, 32, 4>{});
    } else if (real_bytes == 8) {
        if (packs <= 8) return f(Cfg<double, 8, 1>{});
        if (packs <= 16) return f(Cfg<double, 8, 2>{});
        if (packs <= 32) return f(Cfg<double, 16, 2>{});
        if (packs <= 64) return f(Cfg<double, 32, 2>{});
        if (packs <= 128) return f(Cfg<double, 32, 4>{});
    }
    return fail(HPF_EINVAL, "unsupported row length: k too large for this build (ld=%d, real_bytes=%d)", ld,
                real_bytes);
}

// Row stride: rows are padded so that every row starts on a boundary of `row_align` bytes.  32 = whole
// sectors (smallest footprint); 128 = whole cache lines, so a lane group's 128-byte load or RED never
// straddles two lines (fewer L1 tag look-ups and L2 requests per gathered row; measured in
// profiles/).  Pad packs beyond kw are never read or written by the sweep.
int row_layout(int k, int real_bytes, int* ld_out, int* kw_out) {
    int row_align = kDefaultRowAlign;
    if (const char* env = getenv("HPF_ROW_ALIGN")) row_align = atoi(env);
    if (row_align != 32 && row_align != 64 && row_align != 128 && row_align != 256)
        return fail(HPF_EINVAL, "HPF_ROW_ALIGN must be 32, 64, 128 or 256 (got %d)", row_align);
    const int per_pack = 16 / real_bytes;
    const int kw = (k + per_pack - 1) / per_pack * per_pack;
    int ld = kw;
    if ((size_t)kw * real_bytes > 32) {  // rows of one sector or less gain nothing from wider alignment
        const int per_align = row_align / real_bytes;
        // never pad a row to more than the next power of two of its size (a 36-byte row is not worth 128)
        int cap = per_pack;
        while (cap < kw) cap *= 2;
        ld = (kw + per_align - 1) / per_align * per_align;
        if (ld > cap) ld = cap;
    }
    const int per32 = 32 / real_bytes;  // at least whole 32-byte sectors
    ld = (ld + per32 - 1) / per32 * per32;
    *ld_out = ld;
    *kw_out = kw;
    return HPF_OK;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// -------------------------------------------------------------------------------------------------
// Caching device allocator.  cudaMalloc / cudaFree of the engine's multi-hundred-MB buffers cost tens
// of milliseconds per fit (cudaFree also synchronises the device); repeated fits in one process
// (hyper-parameter sweeps, partial_fit loops, the bench's end-to-end call) reuse blocks instead.
// Blocks are matched by exact (device, rounded size); the cache is capped (HPF_CACHE_MB, default a
// quarter of the device's memory) and can be released with hpf_trim_cache().  Buffers that were exported
// over CUDA IPC are never cached (a remote rank may still have them mapped): hpf_uncached_free.  All frees in this file happen after the stream
// that used the block has been synchronised, so immediate reuse is safe.
// -------------------------------------------------------------------------------------------------
struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;
    std::unordered_map<void*, std::pair<int, size_t>> live;
    size_t cached_bytes = 0;
    size_t cap_bytes = 0;
    bool cap_init = false;
};
DevCache g_cache;

inline size_t round_block(size_t bytes) { return (bytes + 511) & ~(size_t)511; }

cudaError_t hpf_malloc_impl(void** p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t rb = round_block(bytes > 0 ? bytes : 1);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto it = g_cache.free_blocks.find({dev, rb});
        if (it != g_cache.free_blocks.end()) {
            *p = it->second;
            g_cache.free_blocks.erase(it);
            g_cache.cached_bytes -= rb;
            g_cache.live[*p] = {dev, rb};
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, rb);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back to the driver and retry once
        cudaGetLastError();
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lk(g_cache.mu);
            for (auto& kv : g_cache.free_blocks) drop.push_back(kv.second);
            g_cache.free_blocks.clear();
            g_cache.cached_bytes = 0;
        }
        for (void* q : drop) cudaFree(q);
        e = cudaMalloc(p, rb);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.live[*p] = {dev, rb};
    }
    return e;
}
template <typename T>
cudaError_t hpf_malloc(T** p, size_t bytes) {
    return hpf_malloc_impl((void**)p, bytes);
}

// frees a block without offering it to the cache (IPC-exported buffers)
void hpf_uncached_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.live.erase(p);
    }
    cudaFree(p);
}

void hpf_free(void* p) {
    if (!p) return;
    std::pair<int, size_t> info;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        if (!g_cache.cap_init) {
            // default cap: a quarter of the device's memory (HPF_CACHE_MB overrides); everything beyond it goes
            // straight back to the driver, so other allocators in the process are never starved by idle blocks
            size_t free_b = 0, total_b = 0;
            double cap_mb = 8192.0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) cap_mb = (double)total_b / 1048576.0 / 4.0;
            else cudaGetLastError();
            if (const char* env = getenv("HPF_CACHE_MB")) cap_mb = atof(env);
            g_cache.cap_bytes = (size_t)cap_mb * 1048576ull;
            g_cache.cap_init = true;
        }
        auto it = g_cache.live.find(p);
        if (it == g_cache.live.end()) {  // not ours (should not happen)
            cudaFree(p);
            return;
        }
        info = it->second;
        g_cache.live.erase(it);
        if (g_cache.cached_bytes + info.second <= g_cache.cap_bytes) {
            g_cache.free_blocks.insert({info, p});
            g_cache.cached_bytes += info.second;
            return;
        }
    }
    cudaFree(p);
}

}  // namespace

// -------------------------------------------------------------------------------------------------
struct hpf_engine {
    int device = 0;
    int64_t nU = 0, nI = 0;
    int k = 0, ld = 0, rb = 4;
    int kw = 0;  // active row width: k rounded up to whole 16-byte packs (<= ld, the row stride)
    cudaStream_t stream = nullptr;
    // constants of the updates, already rounded to `real` the way the reference's typed locals are
    // (cdef real_t k_shp = a_prime + k*a, pxi:173-174; add_k_rte = a_prime/b_prime, pxi:209-210)
    double a = 0.3, c = 0.3, k_shp = 0.0, t_shp = 0.0, add_k = 0.3, add_t = 0.3;
    // variational state, padded (n x ld)
    void *Gshp = nullptr, *Grte = nullptr, *Lshp = nullptr, *Lrte = nullptr, *krte = nullptr, *trte = nullptr;
    // engine-private: per-row softmax factors and sweep accumulators
    void *xu = nullptr, *xi = nullptr, *accU = nullptr, *accI = nullptr;
    double *Tsum = nullptr, *Bsum = nullptr;  // column sums of Theta / Beta (ld doubles each)
    bool state_loaded = false;
    bool x_valid = false;    // xu/xi/Bsum consistent with (shp, rte) and acc buffers zero
    bool mat_valid = false;  // Gshp/Grte/Lshp/Lrte hold the current state
    // data: two orderings of the same triples
    int64_t nnz = 0;
    int *A_row = nullptr, *A_col = nullptr;  // user-major: row = user, col = item
    void* A_val = nullptr;
    int *B_row = nullptr, *B_col = nullptr;  // item-major: row = item, col = user
    void* B_val = nullptr;
    bool data_loaded = false;
    int last_panels = 1, panelsA = 1, panelsB = 1;
    int *A_ptr = nullptr, *B_ptr = nullptr;  // CSR / CSC row pointers (only when both orderings are single-panel)
    // grow-only scratch of the device-assembled minibatch (hpf_step_batch_ids)
    int *bt_major = nullptr, *bt_minor = nullptr, *bt_cnt = nullptr, *bt_off = nullptr, *bt_ids = nullptr;
    void* bt_val = nullptr;
    void* bt_scan_tmp = nullptr;
    size_t bt_scan_bytes = 0;
    int64_t bt_cap_nnz = 0, bt_cap_ids = 0;
    int* ep_ids = nullptr;  // the id list of a whole epoch (hpf_step_epoch_ids)
    int64_t ep_cap_ids = 0;
    int* ep_pin[2] = {nullptr, nullptr};  // pinned staging of the id list, double-buffered
    int64_t ep_pin_cap[2] = {0, 0};
    cudaEvent_t ep_ev[2] = {nullptr, nullptr};
    int ep_flip = 0;
    std::vector<int> hA_ptr, hB_ptr;     // host copies of the row pointers (minibatch sizes without a device round trip)
    bool batch_colsums_valid = false;    // Tsum / Bsum hold the current column sums of Theta / Beta (carried between minibatch steps)
    // options
    double panel_mb = kDefaultPanelMb;
    int chunk = kDefaultChunk;
    int sweep_mode = kDefaultSweepMode;
    int use_graph = 0;
    int v_lpg = 0, v_block = 0, v_depth = 0, v_minb = 0;  // sweep-kernel shape overrides: lanes per row, CTA size, rows in
                                                          // flight per lane group, resident CTAs per SM (0 = default)
    int v_hint = -1;              // L2 policies of the sweep (-1 default; see sweep_rows_kernel)
    double keep_frac = 0.5;       // hint=1: fraction of the gathered lines loaded with evict_last
    int v_smem_gather = -1;       // gathered rows staged in a shared-memory ring with cp.async (-1 default: when the row
                                  // stride is the whole class width) or loaded straight into registers (0)
    int v_prefetch = 0;           // 1: L2 prefetch of the gathered rows two batches ahead (measurement variant)
    int v_fullrow = -1;           // copy whole row strides instead of zero-filling pad packs (-1 default = off)
    int v_rs_ctas = 0;            // CTAs of the stand-alone reduce-scatter kernel (0 = default 48)
    int v_overlap_update = -1;    // user update under the item-major pass: -1 measured default, 0 off, N > 0 on with N CTAs
    void* xu_alt = nullptr;       // second user-factor buffer for that overlap (the pass gathers the old one)
    cudaStream_t side = nullptr;  // its stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int v_robust = -1;            // rescue path for underflowing normalisers: -1 auto (tiny shape priors), 0 off, 1 on
    bool robust_on = false;       // resolved from v_robust and (a, c) when a step starts
    void *dirU = nullptr, *dirI = nullptr;  // robust mode: phi sums that bypass the row factor (nU x ld, nI x ld)
    int strict = 0;               // unknown shape = error instead of falling back to the default
    int64_t launches = 0;
    cudaGraphExec_t graph_lean = nullptr, graph_mat = nullptr;
    // optional per-kernel timing of full-batch iterations
    int timing = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_ms[4] = {0, 0, 0, 0};
    int64_t phase_iters = 0;
    // multi-GPU peer memory (hpf_peer_attach): every rank's item-side buffers, opened through CUDA IPC
    hpf::PeerTable peers;
    bool peer_attached = false;
    bool peer_multicast = false;  // peers.mc_* are valid: the exchange kernel uses multimem.ld_reduce / multimem.st
    bool ipc_exported = false;    // the item-side buffers were handed out with hpf_peer_export
    bool items_pre_reduced = false;  // hpf_reduce_items_peer has left the all-rank item sums of the owned slice in accI
    bool items_adopted = false;   // the five item-side buffers belong to the caller (symmetric memory): never freed here
    std::vector<void*> ipc_opened;
    // minibatch membership stamps (allocated at the first hpf_step_batch)
    int *stamp_u = nullptr, *stamp_i = nullptr;
    int batch_step = 0;

    size_t mat_bytes(int64_t n) const { return (size_t)n * ld * rb; }
};

namespace {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        want = dev;
    }
    ~DeviceGuard() {
        if (prev != want) cudaSetDevice(prev);
    }
    int want;
};

int free_data(hpf_engine* h) {
    hpf_free(h->A_row);
    hpf_free(h->A_col);
    hpf_free(h->A_val);
    hpf_free(h->B_row);
    hpf_free(h->B_col);
    hpf_free(h->B_val);
    hpf_free(h->A_ptr);
    hpf_free(h->B_ptr);
    h->A_ptr = h->B_ptr = nullptr;
    h->hA_ptr.clear();
    h->hB_ptr.clear();
    h->A_row = h->A_col = h->B_row = h->B_col = nullptr;
    h->A_val = h->B_val = nullptr;
    h->data_loaded = false;
    h->nnz = 0;
    return HPF_OK;
}

void drop_graphs(hpf_engine* h) {
    if (h->graph_lean) cudaGraphExecDestroy(h->graph_lean);
    if (h->graph_mat) cudaGraphExecDestroy(h->graph_mat);
    h->graph_lean = h->graph_mat = nullptr;
}

// Stages a caller buffer ([h|d]) on the device.  Returns the device pointer to read from and, if a
// staging copy was made, the allocation to free afterwards.
int stage_in(hpf_engine* h, const void* src, size_t bytes, const void** dev, void** to_free) {
    *to_free = nullptr;
    if (bytes == 0 || is_device_ptr(src)) {
        *dev = src;
        return HPF_OK;
    }
    void* tmp = nullptr;
    CK(hpf_malloc(&tmp, bytes));
    cudaError_t e = cudaMemcpyAsync(tmp, src, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) {
        hpf_free(tmp);
        return fail(HPF_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    }
    *dev = tmp;
    *to_free = tmp;
    return HPF_OK;
}

// caller index array ([h|d], 4 or 8 bytes per entry) -> validated int32 device array
int stage_index(hpf_engine* h, const void* src, int64_t n, int index_bytes, int64_t limit, int* out,
                int* d_bad) {
    const void* dev;
    void* tmp;
    TRY(stage_in(h, src, (size_t)n * index_bytes, &dev, &tmp));
    if (n > 0) {
        if (index_bytes == 8)
            hpf::convert_index_kernel<long long><<<nblk(n), 256, 0, h->stream>>>((const long long*)dev, out, n, limit, d_bad);
        else
            hpf::convert_index_kernel<int><<<nblk(n), 256, 0, h->stream>>>((const int*)dev, out, n, limit, d_bad);
        h->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (tmp) {
        cudaStreamSynchronize(h->stream);
        hpf_free(tmp);
    }
    if (e != cudaSuccess) return fail(HPF_ECUDA, "index conversion failed: %s", cudaGetErrorString(e));
    return HPF_OK;
}

template <typename real>
int upload_matrix(hpf_engine* h, const void* src, void* dst, int64_t nrows, int k, int ld, real fill) {
    const void* dev;
    void* tmp;
    TRY(stage_in(h, src, (size_t)nrows * k * sizeof(real), &dev, &tmp));
    if (nrows > 0) {
        hpf::pad_rows_kernel<real><<<nblk(nrows * ld), 256, 0, h->stream>>>((const real*)dev, (real*)dst, nrows, k, ld, fill);
        h->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (tmp) {
        cudaStreamSynchronize(h->stream);
        hpf_free(tmp);
    }
    if (e != cudaSuccess) return fail(HPF_ECUDA, "pad_rows failed: %s", cudaGetErrorString(e));
    return HPF_OK;
}

// padded engine matrix (optionally divided elementwise by `denom`) -> caller buffer [h|d]
template <typename real>
int download_matrix(hpf_engine* h, const void* src, const void* denom, void* dst, int64_t nrows, int k,
                    int ld) {
    if (dst == nullptr || nrows == 0) return HPF_OK;
    const size_t bytes = (size_t)nrows * k * sizeof(real);
    const bool dev_dst = is_device_ptr(dst);
    void* tmp = nullptr;
    real* out = (real*)dst;
    if (!dev_dst) {
        CK(hpf_malloc(&tmp, bytes));
        out = (real*)tmp;
    }
    hpf::unpad_rows_kernel<real><<<nblk(nrows * k), 256, 0, h->stream>>>((const real*)src, (const real*)denom, out, nrows, k, ld);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !dev_dst) e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (tmp) hpf_free(tmp);
    if (e != cudaSuccess) return fail(HPF_ECUDA, "export failed: %s", cudaGetErrorString(e));
    return HPF_OK;
}

// ---- ordering build: sort triples by (panel of minor id, major id) ---------------------------------
template <typename real>
int build_order(hpf_engine* h, const int* major, const int* minor, const real* val, int64_t n,
                int64_t n_major, int64_t n_minor, int** o_row, int** o_col, void** o_val) {
    // kPadEntries spare entries (zero counts, last row / column repeated): see sweep_rows_kernel
    CK(hpf_malloc(o_row, sizeof(int) * (size_t)(n + kPadEntries)));
    CK(hpf_malloc(o_col, sizeof(int) * (size_t)(n + kPadEntries)));
    CK(hpf_malloc(o_val, sizeof(real) * (size_t)(n + kPadEntries)));
    if (n == 0) return HPF_OK;
    // panels: the gathered (minor) factor matrix is cut so one panel stays L2-resident
    const double minor_bytes = (double)n_minor * h->ld * h->rb;
    int panels = (int)((minor_bytes + h->panel_mb * 1048576.0 - 1.0) / (h->panel_mb * 1048576.0));
    if (panels < 1) panels = 1;
    h->last_panels = panels;
    const int per_panel = (int)((n_minor + panels - 1) / panels);
    const unsigned long long span = (unsigned long long)(n_major > 0 ? n_major : 1);
    int end_bit = 1;
    while (end_bit < 64 && ((unsigned long long)panels * span - 1ull) >> end_bit) ++end_bit;

    unsigned long long *k_in = nullptr, *k_out = nullptr;
    unsigned *p_in = nullptr, *p_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = HPF_OK;
    cudaError_t e;
#define OCK(call)                                                                        \
    if (rc == HPF_OK && (e = (call)) != cudaSuccess)                                     \
    rc = fail(e == cudaErrorMemoryAllocation ? HPF_ENOMEM : HPF_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e))
    OCK(hpf_malloc(&k_in, 8 * (size_t)n));
    OCK(hpf_malloc(&k_out, 8 * (size_t)n));
    OCK(hpf_malloc(&p_in, 4 * (size_t)n));
    OCK(hpf_malloc(&p_out, 4 * (size_t)n));
    if (rc == HPF_OK) {
        hpf::make_keys_kernel<<<nblk(n), 256, 0, h->stream>>>(major, minor, n, per_panel, span, k_in, p_in);
        h->launches++;
    }
    OCK(cudaGetLastError());
    OCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, p_in, p_out, (int)n, 0, end_bit, h->stream));
    OCK(hpf_malloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 16));
    OCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, p_in, p_out, (int)n, 0, end_bit, h->stream));
    if (rc == HPF_OK) {
        hpf::apply_order_kernel<real><<<nblk(n), 256, 0, h->stream>>>(k_out, p_out, n, span, minor, val, *o_row, *o_col, (real*)*o_val);
        hpf::pad_order_kernel<real><<<nblk(kPadEntries), 256, 0, h->stream>>>(*o_row, *o_col, (real*)*o_val, n, kPadEntries);
        h->launches += 5;
    }
    OCK(cudaGetLastError());
    OCK(cudaStreamSynchronize(h->stream));
#undef OCK
    hpf_free(k_in);
    hpf_free(k_out);
    hpf_free(p_in);
    hpf_free(p_out);
    hpf_free(tmp);
    return rc;
}

#include "hpf_sweep_dispatch.inl"

template <typename C>
int launch_sweep_coo(hpf_engine* h, const int* iu, const int* ii, const void* val, int64_t n, const void* xu,
                     const void* xi, void* accU, void* accI, int ld, void* phi, int k, cudaStream_t st) {
    using real = typename C::real;
    if (n == 0) return HPF_OK;
    const int chunk = 64;
    const long long groups = (n + chunk - 1) / chunk;
    if (h && h->robust_on) {  // rescue path: needs the materialised state of an engine
        hpf::RescueArgs<real> rs{(const real*)h->Gshp, (const real*)h->Grte, (const real*)h->Lshp, (const real*)h->Lrte,
                                 (real*)h->dirU, (real*)h->dirI, k};
        hpf::sweep_coo_kernel<real, C::lpg, C::vpl, true><<<nblk(groups * C::lpg), 256, 0, st>>>(
            iu, ii, (const real*)val, n, chunk, (const real*)xu, (const real*)xi, (real*)accU, (real*)accI, ld,
            (real*)phi, k, rs);
    } else {
        hpf::sweep_coo_kernel<real, C::lpg, C::vpl, false><<<nblk(groups * C::lpg), 256, 0, st>>>(
            iu, ii, (const real*)val, n, chunk, (const real*)xu, (const real*)xi, (real*)accU, (real*)accI, ld,
            (real*)phi, k, hpf::RescueArgs<real>{});
    }
    if (h) h->launches++;
    CKK();
    return HPF_OK;
}

int row_grid(int64_t nrows, int lpg) {
    const int gpb = 256 / lpg;
    long long want = (nrows + gpb - 1) / gpb;
    const long long cap = 148 * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

// x_out: where the new per-row factors go (nullptr = in place).  side_ctas > 0: the launch that runs UNDER a sweep
// pass on h->side, as that many 128-thread CTAs (one per SM fits beside the pass's three CTAs).
template <typename C>
int launch_update_rows(hpf_engine* h, bool users, bool mat, void* x_out = nullptr, int side_ctas = 0) {
    using real = typename C::real;
    const int64_t n = users ? h->nU : h->nI;
    if (n == 0) return HPF_OK;
    const int block = side_ctas > 0 ? 128 : 256;
    const int grid = side_ctas > 0 ? side_ctas : row_grid(n, C::lpg);
    cudaStream_t st = side_ctas > 0 ? h->side : h->stream;
    const size_t smem = sizeof(double) * h->ld;
    real* x = (real*)(users ? h->xu : h->xi);
    real* xo = x_out ? (real*)x_out : x;
    real* acc = (real*)(users ? h->accU : h->accI);
    real* shp = (real*)(users ? h->Gshp : h->Lshp);
    real* rte = (real*)(users ? h->Grte : h->Lrte);
    real* rate = (real*)(users ? h->krte : h->trte);
    const double* other = users ? h->Bsum : h->Tsum;
    double* out = users ? h->Tsum : h->Bsum;
    const real prior = (real)(users ? h->a : h->c);
    const real shp_rate = (real)(users ? h->k_shp : h->t_shp);
    const real add_rate = (real)(users ? h->add_k : h->add_t);
    real* direct = h->robust_on ? (real*)(users ? h->dirU : h->dirI) : nullptr;
    if (mat)
        hpf::update_rows_kernel<real, C::lpg, C::vpl, true><<<grid, block, smem, st>>>(
            (int)n, h->ld, h->k, x, xo, acc, direct, shp, rte, rate, other, out, prior, shp_rate, add_rate);
    else
        hpf::update_rows_kernel<real, C::lpg, C::vpl, false><<<grid, block, smem, st>>>(
            (int)n, h->ld, h->k, x, xo, acc, direct, shp, rte, rate, other, out, prior, shp_rate, add_rate);
    h->launches++;
    CKK();
    return HPF_OK;
}

template <typename C>
int launch_rows_to_x(hpf_engine* h, int64_t nrows, const int* rows, const void* shp, const void* rte, void* x,
                     double* colsum, int ld, int k, cudaStream_t st) {
    using real = typename C::real;
    if (nrows == 0) return HPF_OK;
    hpf::rows_to_x_kernel<real, C::lpg, C::vpl><<<row_grid(nrows, C::lpg), 256, sizeof(double) * ld, st>>>(
        (int)nrows, rows, ld, k, (const real*)shp, (const real*)rte, (real*)x, colsum);
    if (h) h->launches++;
    CKK();
    return HPF_OK;
}

// make xu/xi/Bsum consistent with the materialised state and zero the accumulators
int ensure_x(hpf_engine* h) {
    h->batch_colsums_valid = false;  // full-batch work owns Tsum / Bsum from here on
    if (h->x_valid) return HPF_OK;
    if (!h->state_loaded) return fail(HPF_ESTATE, "no state loaded (call hpf_load_state first)");
    CK(cudaMemsetAsync(h->Bsum, 0, sizeof(double) * h->ld, h->stream));
    CK(cudaMemsetAsync(h->accU, 0, h->mat_bytes(h->nU), h->stream));
    CK(cudaMemsetAsync(h->accI, 0, h->mat_bytes(h->nI), h->stream));
    TRY(dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        TRY(launch_rows_to_x<C>(h, h->nU, nullptr, h->Gshp, h->Grte, h->xu, nullptr, h->ld, h->k, h->stream));
        TRY(launch_rows_to_x<C>(h, h->nI, nullptr, h->Lshp, h->Lrte, h->xi, h->Bsum, h->ld, h->k, h->stream));
        return HPF_OK;
    }));
    h->x_valid = true;
    return HPF_OK;
}

// phase marks for the optional per-kernel timing ("timing" option): 0 start, 1 after pass B,
// 2 after pass A, 3 after the user update, 4 after the item update
void mark(hpf_engine* h, int which) {
    if (h->timing && h->ev[0]) cudaEventRecord(h->ev[which], h->stream);
}

// Resolves robust mode (rescue of underflowing normalisers, see sweep_rescue) and allocates its two
// "direct" sum matrices on first use.  Auto rule: psi(x) ~ -1/x for small x, so a row's E[log] entries can
// spread by ~1/prior; its exponentials keep full support in `real` while that spread stays below the
// exponent range (87 for float, 708 for double).
int resolve_robust(hpf_engine* h) {
    const double lo = h->a < h->c ? h->a : h->c;
    const bool want = h->v_robust >= 0 ? h->v_robust != 0 : lo < (h->rb == 4 ? 0.05 : 0.004);
    if (want && !h->dirU) {
        CK(hpf_malloc(&h->dirU, h->mat_bytes(h->nU > 0 ? h->nU : 1)));
        CK(hpf_malloc(&h->dirI, h->mat_bytes(h->nI > 0 ? h->nI : 1)));
        CK(cudaMemsetAsync(h->dirU, 0, h->mat_bytes(h->nU > 0 ? h->nU : 1), h->stream));
        CK(cudaMemsetAsync(h->dirI, 0, h->mat_bytes(h->nI > 0 ? h->nI : 1), h->stream));
    }
    if (want != h->robust_on) drop_graphs(h);
    h->robust_on = want;
    return HPF_OK;
}

// sides: bit 0 = item-major pass (item-side sums), bit 1 = user-major pass (user-side sums).  The
// single-pass cross-check mode (sweep=1: COO atomics) produces both sides in the "user" call.
int do_sweep(hpf_engine* h, int sides = 3) {
    return dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        if (sides & 1) mark(h, 0);
        if (h->sweep_mode == 1) {
            if (sides & 1) mark(h, 1);
            if (sides & 2) {
                TRY(launch_sweep_coo<C>(h, h->A_row, h->A_col, h->A_val, h->nnz, h->xu, h->xi, h->accU, h->accI, h->ld,
                                        nullptr, h->k, h->stream));
                mark(h, 2);
            }
            return HPF_OK;
        }
        // item-major pass first: its output (item-side partial sums) is what a multi-GPU caller
        // exchanges, so the exchange can overlap the user-major pass
        if (sides & 1) {
            TRY(launch_sweep_major<C>(h, h->B_row, h->B_col, h->B_val, h->xi, h->xu, h->accI, false));
            mark(h, 1);
        }
        if (sides & 2) {
            TRY(launch_sweep_major<C>(h, h->A_row, h->A_col, h->A_val, h->xu, h->xi, h->accU, true));
            mark(h, 2);
        }
        return HPF_OK;
    });
}

int do_update(hpf_engine* h, bool users, bool mat) {
    // the side being updated accumulates its column sums into a zeroed buffer
    CK(cudaMemsetAsync(users ? h->Tsum : h->Bsum, 0, sizeof(double) * h->ld, h->stream));
    return dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        return launch_update_rows<C>(h, users, mat);
    });
}

// The user update only needs the user-major pass (and last iteration's Beta column sums), so it can run UNDER the
// item-major pass: user-major pass -> { item-major pass on the main stream || user update on a second stream, writing the
// new factors into a second buffer because the pass is still gathering the old ones } -> item update.  The update is
// issue / DRAM bound, the pass is bound by the L1TEX data pipe, and one 128-thread update CTA fits beside the pass's
// three CTAs on every SM.
// Measured at H (profiles/r02_overlap_update.txt): 296 CTAs 2.26-2.28 ms per iteration against 2.36 sequential; 148 and
// 222 leave the update as the long pole, 370-592 take too much from the pass; 64-thread CTAs and one row of look-ahead in
// the update kernel were both slower.  Small problems keep the sequential order (the fork / join is not free).
constexpr int kDefaultOverlapUpdate = 296;
int overlap_update_ctas(hpf_engine* h) {
    const int dflt = (h->nnz >= (1 << 22) && h->nU >= 100000) ? kDefaultOverlapUpdate : 0;
    const int v = h->v_overlap_update >= 0 ? h->v_overlap_update : dflt;
    if (v <= 0 || h->timing || h->robust_on || h->sweep_mode != 0 || h->nnz == 0 || h->nU == 0) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (h->stream != nullptr && h->stream != cudaStreamLegacy) cudaStreamIsCapturing(h->stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) return 0;  // a captured iteration would bake the buffer swap in
    return v;
}

// item-major pass on the main stream || user update on the second stream (new factors into xu_alt), then the swap
int item_pass_with_user_update(hpf_engine* h, bool mat, int ctas) {
    if (!h->xu_alt) {
        const size_t mu = h->mat_bytes(h->nU);
        CK(hpf_malloc(&h->xu_alt, mu));
        CK(cudaMemsetAsync(h->xu_alt, 0, mu, h->stream));  // pad packs are read by whole-stride copies and never written
    }
    if (!h->side) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    CK(cudaMemsetAsync(h->Tsum, 0, sizeof(double) * h->ld, h->side));
    TRY(dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        return launch_update_rows<C>(h, true, mat, h->xu_alt, ctas);
    }));
    CK(cudaEventRecord(h->ev_join, h->side));
    TRY(do_sweep(h, 1));  // item-major pass: gathers the OLD user factors
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    void* t = h->xu;
    h->xu = h->xu_alt;
    h->xu_alt = t;
    drop_graphs(h);  // a graph captured earlier holds the old buffer roles
    return HPF_OK;
}

int one_iteration_overlapped(hpf_engine* h, bool mat, int ctas) {
    TRY(do_sweep(h, 2));  // user-major pass
    TRY(item_pass_with_user_update(h, mat, ctas));
    TRY(do_update(h, false, mat));
    return HPF_OK;
}

int one_iteration(hpf_engine* h, bool mat) {
    if (h->robust_on) mat = true;  // the rescue path recomputes E[log] from the materialised state
    if (const int ctas = overlap_update_ctas(h)) return one_iteration_overlapped(h, mat, ctas);
    TRY(do_sweep(h));
    TRY(do_update(h, true, mat));
    mark(h, 3);
    TRY(do_update(h, false, mat));
    mark(h, 4);
    if (h->timing && h->ev[0]) {  // profiling mode only: serialises iterations
        CK(cudaEventSynchronize(h->ev[4]));
        for (int p = 0; p < 4; ++p) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev[p], h->ev[p + 1]));
            h->phase_ms[p] += ms;
        }
        h->phase_iters++;
    }
    return HPF_OK;
}

int capture_iteration(hpf_engine* h, bool mat, cudaGraphExec_t* out) {
    cudaGraph_t g = nullptr;
    cudaStream_t cs = h->stream;
    cudaStream_t own = nullptr;
    if (cs == nullptr || cs == cudaStreamLegacy) {  // the legacy stream cannot be captured
        CK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
        cs = own;
    }
    cudaStream_t saved = h->stream;
    h->stream = cs;
    int64_t l0 = h->launches;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    int rc = HPF_OK;
    if (e == cudaSuccess) {
        rc = one_iteration(h, mat);
        cudaError_t e2 = cudaStreamEndCapture(cs, &g);
        if (rc == HPF_OK && e2 != cudaSuccess) rc = fail(HPF_ECUDA, "graph capture failed: %s", cudaGetErrorString(e2));
    } else {
        rc = fail(HPF_ECUDA, "cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
    }
    h->launches = l0;
    h->stream = saved;
    if (rc == HPF_OK) {
        e = cudaGraphInstantiate(out, g, 0);
        if (e != cudaSuccess) rc = fail(HPF_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    if (g) cudaGraphDestroy(g);
    if (own) cudaStreamDestroy(own);
    return rc;
}

constexpr int kLaunchesPerIteration = 4;  // 2 sweep passes + 2 row updates (memsets not counted)

}  // namespace

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int hpf_abi_version(void) { return 1; }
const char* hpf_last_error(void) { return g_err.c_str(); }

int hpf_create(hpf_engine** out, int64_t nU, int64_t nI, int32_t k, int32_t real_bytes, int32_t device) {
    if (!out) return fail(HPF_EINVAL, "out is NULL");
    *out = nullptr;
    if (nU < 0 || nI < 0 || k <= 0) return fail(HPF_EINVAL, "bad shape nU=%lld nI=%lld k=%d", (long long)nU, (long long)nI, k);
    if (nU >= (1ll << 31) || nI >= (1ll << 31)) return fail(HPF_EINVAL, "row counts must be < 2^31");
    if (real_bytes != 4 && real_bytes != 8) return fail(HPF_EINVAL, "real_bytes must be 4 or 8");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HPF_ECUDA, "no CUDA device available (this engine has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(HPF_EINVAL, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    int ld = 0, kw = 0;
    TRY(row_layout(k, real_bytes, &ld, &kw));
    TRY(dispatch(real_bytes, ld, [](auto) { return HPF_OK; }));
    hpf_engine* h = new hpf_engine();
    h->device = device;
    h->nU = nU;
    h->nI = nI;
    h->k = k;
    h->ld = ld;
    h->kw = kw;
    h->rb = real_bytes;
    hpf_set_hyper(h, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0);  // the reference's defaults (hpfrec/__init__.py:205-206)
    const size_t mu = h->mat_bytes(nU > 0 ? nU : 1), mi = h->mat_bytes(nI > 0 ? nI : 1);
    cudaError_t e = cudaSuccess;
    void** mats_u[] = {&h->Gshp, &h->Grte, &h->xu, &h->accU};
    void** mats_i[] = {&h->Lshp, &h->Lrte, &h->xi, &h->accI};
    for (auto p : mats_u)
        if (e == cudaSuccess) e = hpf_malloc(p, mu);
    for (auto p : mats_i)
        if (e == cudaSuccess) e = hpf_malloc(p, mi);
    if (e == cudaSuccess) e = hpf_malloc(&h->krte, (size_t)(nU > 0 ? nU : 1) * real_bytes);
    if (e == cudaSuccess) e = hpf_malloc(&h->trte, (size_t)(nI > 0 ? nI : 1) * real_bytes);
    if (e == cudaSuccess) e = hpf_malloc((void**)&h->Tsum, sizeof(double) * ld);
    if (e == cudaSuccess) e = hpf_malloc((void**)&h->Bsum, sizeof(double) * ld);
    if (e == cudaSuccess) e = cudaMemset(h->Tsum, 0, sizeof(double) * ld);
    if (e == cudaSuccess) e = cudaMemset(h->Bsum, 0, sizeof(double) * ld);
    // pad packs of the factor buffers are skipped by the kernels: define them once
    if (e == cudaSuccess) e = cudaMemsetAsync(h->xu, 0, mu, nullptr);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->xi, 0, mi, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);  // the engine's stream may not be ordered after stream 0
    if (e != cudaSuccess) {
        int rc = fail(e == cudaErrorMemoryAllocation ? HPF_ENOMEM : HPF_ECUDA, "device allocation failed: %s", cudaGetErrorString(e));
        hpf_destroy(h);
        return rc;
    }
    // HPF_OPTIONS="name=value,name=value": defaults for hpf_set_option applied to every new engine (used by
    // the tuning / measurement scripts to run the unmodified test-suite and bench under a candidate
    // configuration).  Unknown names are an error.
    if (const char* env = getenv("HPF_OPTIONS")) {
        std::string all(env);
        size_t pos = 0;
        while (pos < all.size()) {
            size_t end = all.find(',', pos);
            if (end == std::string::npos) end = all.size();
            const std::string item = all.substr(pos, end - pos);
            pos = end + 1;
            const size_t eq = item.find('=');
            if (item.empty()) continue;
            if (eq == std::string::npos) {
                hpf_destroy(h);
                return fail(HPF_EINVAL, "HPF_OPTIONS: expected name=value, got '%s'", item.c_str());
            }
            const std::string name = item.substr(0, eq);
            const int rc = hpf_set_option(h, name.c_str(), atof(item.c_str() + eq + 1));
            if (rc != HPF_OK) {
                const std::string keep = g_err;
                hpf_destroy(h);
                g_err = keep;
                return rc;
            }
        }
    }
    *out = h;
    return HPF_OK;
}

int hpf_destroy(hpf_engine* h) {
    if (!h) return HPF_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    drop_graphs(h);
    free_data(h);
    for (auto e : h->ev)
        if (e) cudaEventDestroy(e);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    for (int b = 0; b < 2; ++b) {
        if (h->ep_pin[b]) cudaFreeHost(h->ep_pin[b]);
        if (h->ep_ev[b]) cudaEventDestroy(h->ep_ev[b]);
    }
    if (h->items_adopted) h->accI = h->xi = h->trte = h->Lshp = h->Lrte = nullptr;
    if (h->ipc_exported) {  // another process may still map these: give them back to the driver, not to the cache
        void* shared[] = {h->accI, h->xi, h->trte, h->Lshp, h->Lrte};
        for (void* p : shared) hpf_uncached_free(p);
        h->accI = h->xi = h->trte = h->Lshp = h->Lrte = nullptr;
    }
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    hpf_free(h->xu_alt);
    void* ptrs[] = {h->Gshp, h->Grte, h->Lshp, h->Lrte, h->krte, h->trte, h->xu, h->xi, h->accU, h->accI, h->Tsum, h->Bsum, h->stamp_u, h->stamp_i,
                    h->bt_major, h->bt_minor, h->bt_cnt, h->bt_off, h->bt_ids, h->bt_val, h->bt_scan_tmp, h->dirU, h->dirI, h->ep_ids};
    for (void* p : ptrs) hpf_free(p);
    delete h;
    return HPF_OK;
}

int hpf_set_hyper(hpf_engine* h, double a, double a_prime, double b_prime, double c, double c_prime, double d_prime) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!(a > 0 && a_prime > 0 && b_prime > 0 && c > 0 && c_prime > 0 && d_prime > 0))
        return fail(HPF_EINVAL, "hyper-parameters must be positive");
    if (h->rb == 4) {
        const float af = (float)a, apf = (float)a_prime, bpf = (float)b_prime, cf = (float)c, cpf = (float)c_prime, dpf = (float)d_prime;
        h->a = af;
        h->c = cf;
        h->k_shp = (float)(apf + (float)h->k * af);
        h->t_shp = (float)(cpf + (float)h->k * cf);
        h->add_k = (float)(apf / bpf);
        h->add_t = (float)(cpf / dpf);
    } else {
        h->a = a;
        h->c = c;
        h->k_shp = a_prime + (double)h->k * a;
        h->t_shp = c_prime + (double)h->k * c;
        h->add_k = a_prime / b_prime;
        h->add_t = c_prime / d_prime;
    }
    drop_graphs(h);
    return HPF_OK;
}

int hpf_set_constants(hpf_engine* h, double a, double c, double k_shp, double t_shp, double add_k_rte, double add_t_rte) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!(a > 0 && c > 0 && k_shp > 0 && t_shp > 0 && add_k_rte > 0 && add_t_rte > 0))
        return fail(HPF_EINVAL, "constants must be positive");
    h->a = a;
    h->c = c;
    h->k_shp = k_shp;
    h->t_shp = t_shp;
    h->add_k = add_k_rte;
    h->add_t = add_t_rte;
    drop_graphs(h);
    return HPF_OK;
}

int hpf_set_stream(hpf_engine* h, void* cuda_stream) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    h->stream = (cudaStream_t)cuda_stream;
    return HPF_OK;
}

int hpf_set_option(hpf_engine* h, const char* name, double value) {
    if (!h || !name) return fail(HPF_EINVAL, "engine or name is NULL");
    if (!strcmp(name, "panel_mb")) {
        if (!(value > 0)) return fail(HPF_EINVAL, "panel_mb must be > 0");
        h->panel_mb = value;  // takes effect at the next hpf_load_coo
    } else if (!strcmp(name, "chunk")) {
        if (value < 32 || value > kMaxChunk || ((int)value % 32) != 0)
            return fail(HPF_EINVAL, "chunk must be a multiple of 32 in [32, %d]", kMaxChunk);
        h->chunk = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "sweep")) {
        if (value != 0 && value != 1) return fail(HPF_EINVAL, "sweep must be 0 (two-pass) or 1 (single-pass COO cross-check)");
        h->sweep_mode = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "lpg") || !strcmp(name, "block") || !strcmp(name, "depth") || !strcmp(name, "minb")) {
        if (value < 0 || value > 1024) return fail(HPF_EINVAL, "%s out of range", name);
        if (!strcmp(name, "lpg")) h->v_lpg = (int)value;      // 0 = default of the row class
        if (!strcmp(name, "block")) h->v_block = (int)value;  // 0 = default
        if (!strcmp(name, "depth")) h->v_depth = (int)value;  // 0 = default
        if (!strcmp(name, "minb")) h->v_minb = (int)value;    // 0 = default
        drop_graphs(h);
    } else if (!strcmp(name, "keep_frac")) {
        if (!(value >= 0.0 && value <= 1.0)) return fail(HPF_EINVAL, "keep_frac must be in [0, 1]");
        h->keep_frac = value;
        drop_graphs(h);
    } else if (!strcmp(name, "hint")) {
        if (value != -1 && value != 0 && value != 1 && value != 2) return fail(HPF_EINVAL, "hint must be -1 (default), 0, 1 or 2");
        h->v_hint = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "smem_gather")) {
        if (value != -1 && value != 0 && value != 1) return fail(HPF_EINVAL, "smem_gather must be -1 (default), 0 or 1");
        h->v_smem_gather = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "prefetch")) {
        h->v_prefetch = value > 0 ? (int)value : 0;
        drop_graphs(h);
    } else if (!strcmp(name, "fullrow") || !strcmp(name, "robust")) {
        if (value != -1 && value != 0 && value != 1) return fail(HPF_EINVAL, "%s must be -1 (default), 0 or 1", name);
        if (!strcmp(name, "fullrow")) h->v_fullrow = (int)value;
        if (!strcmp(name, "robust")) h->v_robust = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "overlap_update")) {
        if (value < -1 || value > 148 * 16) return fail(HPF_EINVAL, "overlap_update out of range");
        h->v_overlap_update = (int)value;
        drop_graphs(h);
    } else if (!strcmp(name, "rs_ctas")) {
        if (value < 0 || value > 148 * 8) return fail(HPF_EINVAL, "rs_ctas out of range");
        h->v_rs_ctas = (int)value;
    } else if (!strcmp(name, "strict")) {
        h->strict = (int)value;
    } else if (!strcmp(name, "use_graph")) {
        h->use_graph = (int)value;
    } else if (!strcmp(name, "timing")) {
        h->timing = (int)value;
        if (h->timing && !h->ev[0]) {
            DeviceGuard guard(h->device);
            for (auto& e : h->ev) CK(cudaEventCreate(&e));
        }
        for (double& v : h->phase_ms) v = 0.0;
        h->phase_iters = 0;
    } else {
        return fail(HPF_EINVAL, "unknown option '%s'", name);
    }
    return HPF_OK;
}

int hpf_load_state(hpf_engine* h, const void* Gamma_shp, const void* Gamma_rte, const void* Lambda_shp,
                   const void* Lambda_rte, const void* k_rte, const void* t_rte) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!Gamma_shp || !Gamma_rte || !Lambda_shp || !Lambda_rte || !k_rte || !t_rte)
        return fail(HPF_EINVAL, "all six state arrays are required");
    DeviceGuard guard(h->device);
    auto up = [&](auto one) {
        using real = decltype(one);
        TRY(upload_matrix<real>(h, Gamma_shp, h->Gshp, h->nU, h->k, h->ld, real(0)));
        TRY(upload_matrix<real>(h, Gamma_rte, h->Grte, h->nU, h->k, h->ld, real(1)));
        TRY(upload_matrix<real>(h, Lambda_shp, h->Lshp, h->nI, h->k, h->ld, real(0)));
        TRY(upload_matrix<real>(h, Lambda_rte, h->Lrte, h->nI, h->k, h->ld, real(1)));
        return HPF_OK;
    };
    TRY(h->rb == 4 ? up(0.0f) : up(0.0));
    CK(cudaMemcpyAsync(h->krte, k_rte, (size_t)h->nU * h->rb, cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(h->trte, t_rte, (size_t)h->nI * h->rb, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->state_loaded = true;
    h->mat_valid = true;
    h->x_valid = false;
    h->batch_colsums_valid = false;
    return HPF_OK;
}

int hpf_export_state(hpf_engine* h, void* Gamma_shp, void* Gamma_rte, void* Lambda_shp, void* Lambda_rte,
                     void* k_rte, void* t_rte, void* Theta, void* Beta) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->state_loaded) return fail(HPF_ESTATE, "no state loaded");
    if (!h->mat_valid) return fail(HPF_ESTATE, "internal: state matrices not materialised");
    DeviceGuard guard(h->device);
    auto down = [&](auto one) {
        using real = decltype(one);
        TRY(download_matrix<real>(h, h->Gshp, nullptr, Gamma_shp, h->nU, h->k, h->ld));
        TRY(download_matrix<real>(h, h->Grte, nullptr, Gamma_rte, h->nU, h->k, h->ld));
        TRY(download_matrix<real>(h, h->Lshp, nullptr, Lambda_shp, h->nI, h->k, h->ld));
        TRY(download_matrix<real>(h, h->Lrte, nullptr, Lambda_rte, h->nI, h->k, h->ld));
        TRY(download_matrix<real>(h, h->Gshp, h->Grte, Theta, h->nU, h->k, h->ld));
        TRY(download_matrix<real>(h, h->Lshp, h->Lrte, Beta, h->nI, h->k, h->ld));
        return HPF_OK;
    };
    TRY(h->rb == 4 ? down(0.0f) : down(0.0));
    if (k_rte) CK(cudaMemcpyAsync(k_rte, h->krte, (size_t)h->nU * h->rb, cudaMemcpyDefault, h->stream));
    if (t_rte) CK(cudaMemcpyAsync(t_rte, h->trte, (size_t)h->nI * h->rb, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return HPF_OK;
}

int hpf_load_coo(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t nnz, int32_t index_bytes) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (nnz < 0 || nnz >= (1ll << 31)) return fail(HPF_EINVAL, "nnz must be in [0, 2^31)");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (nnz > 0 && (!ix_u || !ix_i || !Y)) return fail(HPF_EINVAL, "NULL triple array");
    DeviceGuard guard(h->device);
    drop_graphs(h);
    free_data(h);
    h->nnz = nnz;
    int *u32 = nullptr, *i32 = nullptr, *d_bad = nullptr;
    void* to_free = nullptr;
    const void* yv = nullptr;
    int rc = HPF_OK;
    auto cleanup = [&]() {
        hpf_free(u32);
        hpf_free(i32);
        hpf_free(d_bad);
        if (to_free) hpf_free(to_free);
    };
    const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
    if (hpf_malloc(&u32, 4 * n1) != cudaSuccess || hpf_malloc(&i32, 4 * n1) != cudaSuccess ||
        hpf_malloc(&d_bad, 4) != cudaSuccess) {
        cleanup();
        h->nnz = 0;
        return fail(HPF_ENOMEM, "device allocation failed in hpf_load_coo");
    }
    cudaMemsetAsync(d_bad, 0, 4, h->stream);
    rc = stage_index(h, ix_u, nnz, index_bytes, h->nU, u32, d_bad);
    if (rc == HPF_OK) rc = stage_index(h, ix_i, nnz, index_bytes, h->nI, i32, d_bad);
    if (rc == HPF_OK) rc = stage_in(h, Y, (size_t)nnz * h->rb, &yv, &to_free);
    int bad = 0;
    if (rc == HPF_OK && cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess)
        cudaStreamSynchronize(h->stream);
    if (rc == HPF_OK && bad) rc = fail(HPF_EINVAL, "index out of range in ix_u/ix_i (nU=%lld, nI=%lld)", (long long)h->nU, (long long)h->nI);
    if (rc == HPF_OK) {
        if (h->rb == 4) {
            rc = build_order<float>(h, u32, i32, (const float*)yv, nnz, h->nU, h->nI, &h->A_row, &h->A_col, &h->A_val);
            h->panelsA = h->last_panels;
            if (rc == HPF_OK) rc = build_order<float>(h, i32, u32, (const float*)yv, nnz, h->nI, h->nU, &h->B_row, &h->B_col, &h->B_val);
            h->panelsB = h->last_panels;
        } else {
            rc = build_order<double>(h, u32, i32, (const double*)yv, nnz, h->nU, h->nI, &h->A_row, &h->A_col, &h->A_val);
            h->panelsA = h->last_panels;
            if (rc == HPF_OK) rc = build_order<double>(h, i32, u32, (const double*)yv, nnz, h->nI, h->nU, &h->B_row, &h->B_col, &h->B_val);
            h->panelsB = h->last_panels;
        }
    }
    // single-panel orderings are plain CSR / CSC: keep their row pointers for device-side minibatch
    // assembly (hpf_step_batch_ids)
    if (rc == HPF_OK && h->panelsA == 1 && h->panelsB == 1) {
        if (hpf_malloc(&h->A_ptr, sizeof(int) * (size_t)(h->nU + 1)) != cudaSuccess ||
            hpf_malloc(&h->B_ptr, sizeof(int) * (size_t)(h->nI + 1)) != cudaSuccess) {
            rc = fail(HPF_ENOMEM, "device allocation of row pointers failed");
        } else {
            hpf::row_ptr_kernel<<<nblk(nnz + 1), 256, 0, h->stream>>>(h->A_row, nnz, (int)h->nU, h->A_ptr);
            hpf::row_ptr_kernel<<<nblk(nnz + 1), 256, 0, h->stream>>>(h->B_row, nnz, (int)h->nI, h->B_ptr);
            h->launches += 2;
            if (cudaGetLastError() != cudaSuccess) rc = fail(HPF_ECUDA, "row_ptr_kernel failed");
        }
    }
    cudaStreamSynchronize(h->stream);
    cleanup();
    if (rc != HPF_OK) {
        std::string keep = g_err;
        free_data(h);
        g_err = keep;
        return rc;
    }
    h->data_loaded = true;
    return HPF_OK;
}

int hpf_sweep(hpf_engine* h) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->data_loaded) return fail(HPF_ESTATE, "no data loaded (call hpf_load_coo first)");
    DeviceGuard guard(h->device);
    TRY(resolve_robust(h));
    TRY(ensure_x(h));
    return do_sweep(h);
}

int hpf_sweep_side(hpf_engine* h, int32_t side) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (side != 0 && side != 1) return fail(HPF_EINVAL, "side must be 0 (items) or 1 (users)");
    if (!h->data_loaded) return fail(HPF_ESTATE, "no data loaded (call hpf_load_coo first)");
    DeviceGuard guard(h->device);
    TRY(resolve_robust(h));
    if (h->robust_on) return fail(HPF_EINVAL, "robust mode (tiny shape priors) is not available for sharded sweeps");
    TRY(ensure_x(h));
    return do_sweep(h, side == 0 ? 1 : 2);
}

int hpf_update_users(hpf_engine* h) { return hpf_update_users_ex(h, 1); }

int hpf_update_users_ex(hpf_engine* h, int32_t materialize) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->x_valid) return fail(HPF_ESTATE, "hpf_update_users must follow hpf_sweep");
    DeviceGuard guard(h->device);
    return do_update(h, true, materialize != 0);
}

int hpf_item_pass_with_user_update(hpf_engine* h, int32_t materialize) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->data_loaded) return fail(HPF_ESTATE, "no data loaded (call hpf_load_coo first)");
    if (!h->x_valid) return fail(HPF_ESTATE, "hpf_item_pass_with_user_update must follow the user-major pass");
    DeviceGuard guard(h->device);
    TRY(resolve_robust(h));
    if (h->robust_on) return fail(HPF_EINVAL, "robust mode (tiny shape priors) is not available for sharded sweeps");
    const int ctas = h->v_overlap_update >= 0 ? h->v_overlap_update : kDefaultOverlapUpdate;
    if (ctas <= 0 || h->nnz == 0 || h->nU == 0 || h->sweep_mode != 0) {  // sequential form of the same two steps
        TRY(do_sweep(h, 1));
        return do_update(h, true, materialize != 0);
    }
    return item_pass_with_user_update(h, materialize != 0, ctas);
}

int hpf_update_items(hpf_engine* h) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->x_valid) return fail(HPF_ESTATE, "hpf_update_items must follow hpf_sweep");
    DeviceGuard guard(h->device);
    // Tsum was produced by hpf_update_users (and all-reduced by the caller): do not zero it here
    CK(cudaMemsetAsync(h->Bsum, 0, sizeof(double) * h->ld, h->stream));
    TRY(dispatch(h->rb, h->ld, [&](auto cfg) {
        using C = decltype(cfg);
        return launch_update_rows<C>(h, false, true);
    }));
    h->mat_valid = true;
    return HPF_OK;
}

int hpf_partials(hpf_engine* h, void** item_sums, int64_t* item_sums_count, void** theta_colsum, int64_t* theta_colsum_count) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (item_sums) *item_sums = h->accI;
    if (item_sums_count) *item_sums_count = h->nI * h->ld;
    if (theta_colsum) *theta_colsum = h->Tsum;
    if (theta_colsum_count) *theta_colsum_count = h->k;
    return HPF_OK;
}

int hpf_step_full(hpf_engine* h, int32_t niter) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (niter < 0) return fail(HPF_EINVAL, "niter must be >= 0");
    if (!h->data_loaded) return fail(HPF_ESTATE, "no data loaded (call hpf_load_coo first)");
    if (!h->state_loaded) return fail(HPF_ESTATE, "no state loaded (call hpf_load_state first)");
    if (niter == 0) return HPF_OK;
    DeviceGuard guard(h->device);
    TRY(resolve_robust(h));
    TRY(ensure_x(h));
    h->mat_valid = false;
    const bool graph = h->use_graph && !h->timing && h->stream != nullptr && h->stream != cudaStreamLegacy;
    if (graph) {
        if (!h->graph_lean) TRY(capture_iteration(h, false, &h->graph_lean));
        if (!h->graph_mat) TRY(capture_iteration(h, true, &h->graph_mat));
    }
    for (int it = 0; it < niter; ++it) {
        const bool mat = (it == niter - 1);  // only the last iteration of a call stores shp/rte
        if (graph) {
            CK(cudaGraphLaunch(mat ? h->graph_mat : h->graph_lean, h->stream));
            h->launches += kLaunchesPerIteration;
        } else {
            TRY(one_iteration(h, mat));
        }
    }
    h->mat_valid = true;
    return HPF_OK;
}

// ---- multi-GPU peer memory ----------------------------------------------------------------------------
int hpf_peer_export(hpf_engine* h, void* handles) {
    if (!h || !handles) return fail(HPF_EINVAL, "NULL argument");
    DeviceGuard guard(h->device);
    void* bufs[HPF_PEER_BUFFERS] = {h->accI, h->xi, h->trte, h->Lshp, h->Lrte};
    for (int b = 0; b < HPF_PEER_BUFFERS; ++b)
        CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)((char*)handles + (size_t)b * HPF_IPC_HANDLE_BYTES), bufs[b]));
    h->ipc_exported = true;
    return HPF_OK;
}

int hpf_peer_attach(hpf_engine* h, int32_t rank, int32_t world, const void* all_handles) {
    if (!h || !all_handles) return fail(HPF_EINVAL, "NULL argument");
    if (world < 1 || world > hpf::kMaxPeers || rank < 0 || rank >= world)
        return fail(HPF_EINVAL, "bad rank/world (%d/%d, at most %d peers)", rank, world, hpf::kMaxPeers);
    static_assert(sizeof(cudaIpcMemHandle_t) <= HPF_IPC_HANDLE_BYTES, "IPC handle larger than the ABI slot");
    DeviceGuard guard(h->device);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    h->peer_attached = false;
    void* own[HPF_PEER_BUFFERS] = {h->accI, h->xi, h->trte, h->Lshp, h->Lrte};
    for (int p = 0; p < world; ++p) {
        void* ptr[HPF_PEER_BUFFERS];
        for (int b = 0; b < HPF_PEER_BUFFERS; ++b) {
            if (p == rank) {
                ptr[b] = own[b];
                continue;
            }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char*)all_handles + ((size_t)p * HPF_PEER_BUFFERS + b) * HPF_IPC_HANDLE_BYTES, sizeof(hd));
            CK(cudaIpcOpenMemHandle(&ptr[b], hd, cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(ptr[b]);
        }
        h->peers.acc[p] = ptr[0];
        h->peers.x[p] = ptr[1];
        h->peers.rate[p] = ptr[2];
        h->peers.shp[p] = ptr[3];
        h->peers.rte[p] = ptr[4];
    }
    h->peers.mc_acc = h->peers.mc_x = h->peers.mc_rate = h->peers.mc_shp = h->peers.mc_rte = nullptr;
    h->peer_multicast = false;
    h->peers.world = world;
    h->peers.rank = rank;
    h->peer_attached = true;
    return HPF_OK;
}

int hpf_item_buffer_bytes(hpf_engine* h, int64_t out[HPF_PEER_BUFFERS]) {
    if (!h || !out) return fail(HPF_EINVAL, "NULL argument");
    const int64_t mi = (int64_t)h->mat_bytes(h->nI > 0 ? h->nI : 1);
    out[0] = mi;                                            // item_sums
    out[1] = mi;                                            // item softmax factors
    out[2] = (int64_t)(h->nI > 0 ? h->nI : 1) * h->rb;      // t_rte
    out[3] = mi;                                            // Lambda_shp
    out[4] = mi;                                            // Lambda_rte
    return HPF_OK;
}

int hpf_adopt_item_buffers(hpf_engine* h, void* const bufs[HPF_PEER_BUFFERS]) {
    if (!h || !bufs) return fail(HPF_EINVAL, "NULL argument");
    if (h->state_loaded || h->data_loaded) return fail(HPF_ESTATE, "hpf_adopt_item_buffers must precede hpf_load_state / hpf_load_coo");
    for (int b = 0; b < HPF_PEER_BUFFERS; ++b)
        if (!bufs[b] || !is_device_ptr(bufs[b]) || ((uintptr_t)bufs[b] & 255u))
            return fail(HPF_EINVAL, "item buffer %d must be a 256-byte aligned device pointer", b);
    DeviceGuard guard(h->device);
    if (!h->items_adopted) {
        void* mine[] = {h->accI, h->xi, h->trte, h->Lshp, h->Lrte};
        for (void* p : mine) hpf_free(p);
    }
    h->accI = bufs[0];
    h->xi = bufs[1];
    h->trte = bufs[2];
    h->Lshp = bufs[3];
    h->Lrte = bufs[4];
    h->items_adopted = true;
    // same initial contents as hpf_create gives its own buffers: pad packs of the factor buffer defined, sums zero
    const size_t mi = h->mat_bytes(h->nI > 0 ? h->nI : 1);
    CK(cudaMemsetAsync(h->xi, 0, mi, h->stream));
    CK(cudaMemsetAsync(h->accI, 0, mi, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->x_valid = false;
    return HPF_OK;
}

int hpf_peer_attach_ptrs(hpf_engine* h, int32_t rank, int32_t world, void* const* peer_ptrs, void* const* mc_ptrs) {
    if (!h || !peer_ptrs) return fail(HPF_EINVAL, "NULL argument");
    if (world < 1 || world > hpf::kMaxPeers || rank < 0 || rank >= world)
        return fail(HPF_EINVAL, "bad rank/world (%d/%d, at most %d peers)", rank, world, hpf::kMaxPeers);
    void* own[HPF_PEER_BUFFERS] = {h->accI, h->xi, h->trte, h->Lshp, h->Lrte};
    for (int b = 0; b < HPF_PEER_BUFFERS; ++b)
        if (peer_ptrs[(size_t)rank * HPF_PEER_BUFFERS + b] != own[b])
            return fail(HPF_EINVAL, "peer table entry of this rank is not the engine's own buffer %d (adopt the buffers first)", b);
    for (int p = 0; p < world; ++p) {
        void* const* row = peer_ptrs + (size_t)p * HPF_PEER_BUFFERS;
        h->peers.acc[p] = row[0];
        h->peers.x[p] = row[1];
        h->peers.rate[p] = row[2];
        h->peers.shp[p] = row[3];
        h->peers.rte[p] = row[4];
    }
    h->peer_multicast = mc_ptrs != nullptr && mc_ptrs[0] != nullptr;
    h->peers.mc_acc = h->peer_multicast ? mc_ptrs[0] : nullptr;
    h->peers.mc_x = h->peer_multicast ? mc_ptrs[1] : nullptr;
    h->peers.mc_rate = h->peer_multicast ? mc_ptrs[2] : nullptr;
    h->peers.mc_shp = h->peer_multicast ? mc_ptrs[3] : nullptr;
    h->peers.mc_rte = h->peer_multicast ? mc_ptrs[4] : nullptr;
    h->peers.world = world;
    h->peers.rank = rank;
    h->peer_attached = true;
    return HPF_OK;
}

int hpf_update_items_peer(hpf_engine* h, int32_t materialize) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->peer_attached) return fail(HPF_ESTATE, "hpf_peer_attach has not been called");
    if (!h->x_valid) return fail(HPF_ESTATE, "hpf_update_items_peer must follow hpf_sweep");
    DeviceGuard guard(h->device);
    const int r0 = (int)(h->nI * (int64_t)h->peers.rank / h->peers.world);
    const int r1 = (int)(h->nI * (int64_t)(h->peers.rank + 1) / h->peers.world);
    CK(cudaMemsetAsync(h->Bsum, 0, sizeof(double) * h->ld, h->stream));
    if (r1 > r0) {
        TRY(dispatch(h->rb, h->ld, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            const int grid = row_grid(r1 - r0, C::lpg);
            const size_t smem = sizeof(double) * h->ld;
            const real prior = (real)h->c, shp_rate = (real)h->t_shp, add_rate = (real)h->add_t;
            const int pre = h->items_pre_reduced ? 1 : 0;
            if (h->peer_multicast) {
                if (materialize)
                    hpf::update_items_peer_kernel<real, C::lpg, C::vpl, true, true><<<grid, 256, smem, h->stream>>>(
                        r0, r1, h->ld, h->k, h->peers, h->Tsum, h->Bsum, prior, shp_rate, add_rate, pre);
                else
                    hpf::update_items_peer_kernel<real, C::lpg, C::vpl, false, true><<<grid, 256, smem, h->stream>>>(
                        r0, r1, h->ld, h->k, h->peers, h->Tsum, h->Bsum, prior, shp_rate, add_rate, pre);
            } else if (materialize)
                hpf::update_items_peer_kernel<real, C::lpg, C::vpl, true, false><<<grid, 256, smem, h->stream>>>(
                    r0, r1, h->ld, h->k, h->peers, h->Tsum, h->Bsum, prior, shp_rate, add_rate, pre);
            else
                hpf::update_items_peer_kernel<real, C::lpg, C::vpl, false, false><<<grid, 256, smem, h->stream>>>(
                    r0, r1, h->ld, h->k, h->peers, h->Tsum, h->Bsum, prior, shp_rate, add_rate, pre);
            h->launches++;
            CKK();
            return HPF_OK;
        }));
    }
    h->items_pre_reduced = false;
    h->mat_valid = materialize != 0;
    return HPF_OK;
}

int hpf_reduce_items_peer(hpf_engine* h, void* stream) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    if (!h->peer_attached) return fail(HPF_ESTATE, "hpf_peer_attach has not been called");
    if (!h->x_valid) return fail(HPF_ESTATE, "hpf_reduce_items_peer must follow the item-major pass");
    DeviceGuard guard(h->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int r0 = (int)(h->nI * (int64_t)h->peers.rank / h->peers.world);
    const int r1 = (int)(h->nI * (int64_t)(h->peers.rank + 1) / h->peers.world);
    if (r1 > r0) {
        TRY(dispatch(h->rb, h->ld, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            // NVLink-latency-bound, and it shares the SMs with the user-major pass (measured at 2 GPUs: 296 CTAs slowed
            // that pass by 0.08-0.28 ms): a few CTAs, each thread keeping 4 rows' packs in flight
            int grid = row_grid(r1 - r0, C::lpg);
            const int cap = h->v_rs_ctas > 0 ? h->v_rs_ctas : 48;
            if (grid > cap) grid = cap;
            if (h->peer_multicast)
                hpf::reduce_items_peer_kernel<real, C::lpg, C::vpl, true><<<grid, 256, 0, st>>>(r0, r1, h->ld, h->k, h->peers);
            else
                hpf::reduce_items_peer_kernel<real, C::lpg, C::vpl, false><<<grid, 256, 0, st>>>(r0, r1, h->ld, h->k, h->peers);
            h->launches++;
            CKK();
            return HPF_OK;
        }));
    }
    h->items_pre_reduced = true;
    return HPF_OK;
}

int hpf_peer_finish(hpf_engine* h) {
    if (!h) return fail(HPF_EINVAL, "engine is NULL");
    DeviceGuard guard(h->device);
    // (re-zeroing the owned rows in every replica with multicast stores from the exchange kernel was measured:
    // correct, but slower than this local memset at 2 GPUs and no faster at 8)
    CK(cudaMemsetAsync(h->accI, 0, h->mat_bytes(h->nI), h->stream));
    return HPF_OK;
}

int hpf_beta_colsum(hpf_engine* h, void** ptr, int64_t* count) {
    if (!h || !ptr) return fail(HPF_EINVAL, "NULL argument");
    *ptr = h->Bsum;
    if (count) *count = h->k;
    return HPF_OK;
}

int hpf_trim_cache(void) {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        for (auto& kv : g_cache.free_blocks) drop.push_back(kv.second);
        g_cache.free_blocks.clear();
        g_cache.cached_bytes = 0;
    }
    for (void* q : drop) cudaFree(q);
    return HPF_OK;
}

int hpf_launch_count(hpf_engine* h, int64_t* out) {
    if (!h || !out) return fail(HPF_EINVAL, "NULL argument");
    *out = h->launches;
    return HPF_OK;
}

int hpf_phase_ms(hpf_engine* h, double out[4], int64_t* iterations) {
    if (!h || !out) return fail(HPF_EINVAL, "NULL argument");
    for (int p = 0; p < 4; ++p) out[p] = h->phase_ms[p];
    if (iterations) *iterations = h->phase_iters;
    return HPF_OK;
}

int hpf_describe(hpf_engine* h, char* buf, int64_t n) {
    if (!h || !buf || n <= 0) return fail(HPF_EINVAL, "engine or buffer is NULL");
    const int packs_row = h->ld * h->rb / 16;
    int cls = 8;
    while (cls < packs_row) cls *= 2;
    const double lo = h->a < h->c ? h->a : h->c;
    const int robust = h->v_robust >= 0 ? h->v_robust : (lo < (h->rb == 4 ? 0.05 : 0.004) ? 1 : 0);
    const bool full_ok = h->ld * h->rb / 16 == cls;
    const int smem_gather = h->v_smem_gather >= 0 ? h->v_smem_gather : ((full_ok && !robust) ? 1 : 0);
    const int fullrow = full_ok ? (h->v_fullrow >= 0 ? h->v_fullrow : smem_gather) : 0;
    const RowsShape d = default_rows_shape(cls, h->rb, smem_gather != 0);
    int lpg = h->v_lpg ? h->v_lpg : d.lpg, block = h->v_block ? h->v_block : d.block;
    int depth = h->v_depth ? h->v_depth : d.depth, minb = h->v_minb ? h->v_minb : d.minb;
    if (lpg == 0) {  // generic shape: the row class's lane-group width, 128-thread CTAs, 2 rows in flight
        lpg = cls <= 16 ? 8 : (cls <= 32 ? 16 : 32);
        block = 128;
        depth = 2;
        minb = 2;
    }
    const int hint = h->v_hint >= 0 ? h->v_hint : kDefaultHint;
    // CTAs of the user update that runs under the item-major pass in hpf_step_full (0 = the two run in sequence)
    int ovu = h->v_overlap_update >= 0 ? h->v_overlap_update : ((h->nnz >= (1 << 22) && h->nU >= 100000) ? kDefaultOverlapUpdate : 0);
    if (robust || h->sweep_mode != 0) ovu = 0;
    snprintf(buf, (size_t)n,
             "real_bytes=%d k=%d kw=%d ld=%d sweep=%d kernel=%s lpg=%d depth=%d block=%d minb=%d smem_gather=%d hint=%d fullrow=%d robust=%d chunk=%d "
             "panel_mb=%g panels_user_major=%d panels_item_major=%d launches_per_iteration=%d overlap_update=%d",
             h->rb, h->k, h->kw, h->ld, h->sweep_mode, h->sweep_mode == 1 ? "sweep_coo_kernel" : "sweep_rows_kernel", lpg,
             depth, block, minb, smem_gather, hint, fullrow, robust, h->chunk, h->panel_mb, h->panelsA, h->panelsB, h->sweep_mode == 0 ? 4 : 3,
             ovu);
    return HPF_OK;
}

int hpf_ld(hpf_engine* h, int32_t* out) {
    if (!h || !out) return fail(HPF_EINVAL, "NULL argument");
    *out = h->ld;
    return HPF_OK;
}

// ---- scoring ---------------------------------------------------------------------------------------
static int score_common(hpf_engine* h, const int* iu, const int* ii, const void* val, int64_t n, int full_llk,
                        double* out4, void* pred) {
    // E[x] matrices are materialised into the (otherwise idle between iterations) accumulators'
    // sibling buffers: reuse xu/xi would destroy sweep state, so use scratch allocations
    void *theta = nullptr, *beta = nullptr;
    double* d_sums = nullptr;
    int rc = HPF_OK;
    cudaError_t e = hpf_malloc(&theta, h->mat_bytes(h->nU > 0 ? h->nU : 1));
    if (e == cudaSuccess) e = hpf_malloc(&beta, h->mat_bytes(h->nI > 0 ? h->nI : 1));
    if (e == cudaSuccess) e = hpf_malloc((void**)&d_sums, sizeof(double) * (4 + 2 * (size_t)h->ld));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_sums, 0, sizeof(double) * (4 + 2 * (size_t)h->ld), h->stream);
    if (e != cudaSuccess) rc = fail(HPF_ENOMEM, "scratch allocation failed: %s", cudaGetErrorString(e));
    if (rc == HPF_OK) {
        rc = dispatch(h->rb, h->ld, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            if (h->nU > 0) hpf::ratio_rows_kernel<real><<<nblk(h->nU * h->ld), 256, 0, h->stream>>>(h->nU, h->ld, h->k, (const real*)h->Gshp, (const real*)h->Grte, (real*)theta);
            if (h->nI > 0) hpf::ratio_rows_kernel<real><<<nblk(h->nI * h->ld), 256, 0, h->stream>>>(h->nI, h->ld, h->k, (const real*)h->Lshp, (const real*)h->Lrte, (real*)beta);
            h->launches += 2;
            CKK();
            if (n > 0) {
                const int gpb = 256 / C::lpg;
                long long want = (n + gpb - 1) / gpb;
                if (want > 148 * 16) want = 148 * 16;
                hpf::score_kernel<real, C::lpg, C::vpl><<<(unsigned)want, 256, 0, h->stream>>>(
                    iu, ii, (const real*)val, n, (const real*)theta, (const real*)beta, h->ld, full_llk,
                    out4 ? d_sums : nullptr, (real*)pred);
                h->launches++;
                CKK();
            }
            if (out4) {
                const size_t smem = sizeof(double) * h->ld;
                if (h->nU > 0) hpf::colsum_kernel<real><<<148 * 4, 256, smem, h->stream>>>(h->nU, h->ld, h->k, (const real*)theta, (const real*)nullptr, d_sums + 4);
                if (h->nI > 0) hpf::colsum_kernel<real><<<148 * 4, 256, smem, h->stream>>>(h->nI, h->ld, h->k, (const real*)beta, (const real*)nullptr, d_sums + 4 + h->ld);
                hpf::dot_cols_kernel<<<1, 32, 0, h->stream>>>(h->k, d_sums + 4, d_sums + 4 + h->ld, d_sums + 3);
                h->launches += 3;
                CKK();
            }
            return HPF_OK;
        });
    }
    if (rc == HPF_OK && out4) {
        e = cudaMemcpyAsync(out4, d_sums, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) rc = fail(HPF_ECUDA, "D2H failed: %s", cudaGetErrorString(e));
    }
    e = cudaStreamSynchronize(h->stream);
    if (rc == HPF_OK && e != cudaSuccess) rc = fail(HPF_ECUDA, "score kernels failed: %s", cudaGetErrorString(e));
    hpf_free(theta);
    hpf_free(beta);
    hpf_free(d_sums);
    return rc;
}

int hpf_llk_train(hpf_engine* h, int32_t full_llk, double out[4]) {
    if (!h || !out) return fail(HPF_EINVAL, "NULL argument");
    if (!h->data_loaded || !h->state_loaded || !h->mat_valid) return fail(HPF_ESTATE, "needs loaded data and state");
    DeviceGuard guard(h->device);
    return score_common(h, h->A_row, h->A_col, h->A_val, h->nnz, full_llk, out, nullptr);
}

static int score_external(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t n,
                          int32_t index_bytes, int full_llk, double* out4, void* pred_out) {
    if (n < 0 || n >= (1ll << 31)) return fail(HPF_EINVAL, "n out of range");
    if (index_bytes != 4 && index_bytes != 8) return fail(HPF_EINVAL, "index_bytes must be 4 or 8");
    if (!h->state_loaded || !h->mat_valid) return fail(HPF_ESTATE, "no state loaded");
    DeviceGuard guard(h->device);
    int *u32 = nullptr, *i32 = nullptr, *d_bad = nullptr;
    void *yfree = nullptr, *pred_dev = nullptr;
    const void* yv = nullptr;
    const size_t n1 = (size_t)(n > 0 ? n : 1);
    int rc = HPF_OK;
    if (hpf_malloc(&u32, 4 * n1) != cudaSuccess || hpf_malloc(&i32, 4 * n1) != cudaSuccess || hpf_malloc(&d_bad, 4) != cudaSuccess)
        rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK) {
        cudaMemsetAsync(d_bad, 0, 4, h->stream);
        rc = stage_index(h, ix_u, n, index_bytes, h->nU, u32, d_bad);
    }
    if (rc == HPF_OK) rc = stage_index(h, ix_i, n, index_bytes, h->nI, i32, d_bad);
    if (rc == HPF_OK && Y) rc = stage_in(h, Y, (size_t)n * h->rb, &yv, &yfree);
    int bad = 0;
    if (rc == HPF_OK) {
        cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        if (bad) rc = fail(HPF_EINVAL, "index out of range");
    }
    const bool pred_is_dev = pred_out && is_device_ptr(pred_out);
    if (rc == HPF_OK && pred_out && !pred_is_dev && hpf_malloc(&pred_dev, n1 * h->rb) != cudaSuccess)
        rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK) rc = score_common(h, u32, i32, yv, n, full_llk, out4, pred_out ? (pred_is_dev ? pred_out : pred_dev) : nullptr);
    if (rc == HPF_OK && pred_dev && n > 0) {
        if (cudaMemcpy(pred_out, pred_dev, (size_t)n * h->rb, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(HPF_ECUDA, "D2H of predictions failed");
    }
    hpf_free(u32);
    hpf_free(i32);
    hpf_free(d_bad);
    hpf_free(yfree);
    hpf_free(pred_dev);
    return rc;
}

int hpf_llk(hpf_engine* h, const void* ix_u, const void* ix_i, const void* Y, int64_t nnz, int32_t index_bytes,
            int32_t full_llk, double out[4]) {
    if (!h || !out || (nnz > 0 && (!ix_u || !ix_i || !Y))) return fail(HPF_EINVAL, "NULL argument");
    return score_external(h, ix_u, ix_i, Y, nnz, index_bytes, full_llk, out, nullptr);
}

int hpf_predict(hpf_engine* h, const void* ix_u, const void* ix_i, int64_t n, int32_t index_bytes, void* out) {
    if (!h || (n > 0 && (!ix_u || !ix_i || !out))) return fail(HPF_EINVAL, "NULL argument");
    return score_external(h, ix_u, ix_i, nullptr, n, index_bytes, 0, nullptr, out);
}

// ---- stateless L1 forms -------------------------------------------------------------------------------
int hpf_update_shapes(int32_t real_bytes, int32_t index_bytes, int32_t device, void* G_sh, const void* G_rt,
                      void* L_sh, const void* L_rt, void* phi, const void* Y, const void* ix_u, const void* ix_i,
                      int64_t nU, int64_t nI, int64_t nY, int32_t k, double a, double c) {
    if (!G_sh || !G_rt || !L_sh || !L_rt) return fail(HPF_EINVAL, "NULL state array");
    hpf_engine* h = nullptr;
    TRY(hpf_create(&h, nU, nI, k, real_bytes, device));
    DeviceGuard guard(device);
    int rc = HPF_OK;
    void *kr = nullptr, *tr = nullptr, *phi_dev = nullptr;
    int *u32 = nullptr, *i32 = nullptr, *d_bad = nullptr;
    void* yfree = nullptr;
    const void* yv = nullptr;
    const size_t n1 = (size_t)(nY > 0 ? nY : 1);
    do {
        // rate vectors are not touched by these two loops; upload dummies
        std::vector<char> ones((size_t)(nU > nI ? nU : nI) * real_bytes + 8, 0);
        if ((rc = hpf_load_state(h, G_sh, G_rt, L_sh, L_rt, ones.data(), ones.data())) != HPF_OK) break;
        h->a = a;
        h->c = c;
        if ((rc = resolve_robust(h)) != HPF_OK) break;
        if ((rc = ensure_x(h)) != HPF_OK) break;
        if (hpf_malloc(&u32, 4 * n1) != cudaSuccess || hpf_malloc(&i32, 4 * n1) != cudaSuccess || hpf_malloc(&d_bad, 4) != cudaSuccess) {
            rc = fail(HPF_ENOMEM, "device allocation failed");
            break;
        }
        cudaMemsetAsync(d_bad, 0, 4, h->stream);
        if ((rc = stage_index(h, ix_u, nY, index_bytes, nU, u32, d_bad)) != HPF_OK) break;
        if ((rc = stage_index(h, ix_i, nY, index_bytes, nI, i32, d_bad)) != HPF_OK) break;
        if ((rc = stage_in(h, Y, (size_t)nY * real_bytes, &yv, &yfree)) != HPF_OK) break;
        int bad = 0;
        cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
        if (bad) {
            rc = fail(HPF_EINVAL, "index out of range");
            break;
        }
        const bool phi_dev_dst = phi && is_device_ptr(phi);
        if (phi && !phi_dev_dst && hpf_malloc(&phi_dev, n1 * (size_t)k * real_bytes) != cudaSuccess) {
            rc = fail(HPF_ENOMEM, "device allocation of phi failed");
            break;
        }
        void* phi_target = phi ? (phi_dev_dst ? phi : phi_dev) : nullptr;
        rc = dispatch(real_bytes, h->ld, [&](auto cfg) {
            using C = decltype(cfg);
            using real = typename C::real;
            TRY(launch_sweep_coo<C>(h, u32, i32, yv, nY, h->xu, h->xi, h->accU, h->accI, h->ld, phi_target, k, h->stream));
            if (nU > 0) hpf::finish_shapes_kernel<real><<<nblk(nU * h->ld), 256, 0, h->stream>>>(nU * h->ld, (const real*)h->xu, (const real*)h->accU, (const real*)(h->robust_on ? h->dirU : nullptr), (real*)h->Gshp, (real)a);
            if (nI > 0) hpf::finish_shapes_kernel<real><<<nblk(nI * h->ld), 256, 0, h->stream>>>(nI * h->ld, (const real*)h->xi, (const real*)h->accI, (const real*)(h->robust_on ? h->dirI : nullptr), (real*)h->Lshp, (real)c);
            CKK();
            return HPF_OK;
        });
        if (rc != HPF_OK) break;
        if ((rc = hpf_export_state(h, G_sh, nullptr, L_sh, nullptr, nullptr, nullptr, nullptr, nullptr)) != HPF_OK) break;
        if (phi_dev && nY > 0 && cudaMemcpy(phi, phi_dev, (size_t)nY * k * real_bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(HPF_ECUDA, "D2H of phi failed");
    } while (0);
    std::string keep = g_err;
    hpf_free(kr);
    hpf_free(tr);
    hpf_free(phi_dev);
    hpf_free(u32);
    hpf_free(i32);
    hpf_free(d_bad);
    hpf_free(yfree);
    hpf_destroy(h);
    g_err = keep;
    return rc;
}

int hpf_digamma(int32_t real_bytes, int32_t device, const void* x, void* out, int64_t n) {
    if (real_bytes != 4 && real_bytes != 8) return fail(HPF_EINVAL, "real_bytes must be 4 or 8");
    if (n < 0 || (n > 0 && (!x || !out))) return fail(HPF_EINVAL, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(HPF_ECUDA, "no such CUDA device %d", device);
    }
    if (n == 0) return HPF_OK;
    DeviceGuard guard(device);
    void *dx = nullptr, *dout = nullptr;
    const size_t bytes = (size_t)n * real_bytes;
    int rc = HPF_OK;
    if (hpf_malloc(&dx, bytes) != cudaSuccess || hpf_malloc(&dout, bytes) != cudaSuccess) rc = fail(HPF_ENOMEM, "device allocation failed");
    if (rc == HPF_OK && cudaMemcpy(dx, x, bytes, cudaMemcpyDefault) != cudaSuccess) rc = fail(HPF_ECUDA, "copy in failed");
    if (rc == HPF_OK) {
        if (real_bytes == 4)
            hpf::digamma_kernel<float><<<nblk(n), 256>>>((const float*)dx, (float*)dout, n);
        else
            hpf::digamma_kernel<double><<<nblk(n), 256>>>((const double*)dx, (double*)dout, n);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpy(out, dout, bytes, cudaMemcpyDefault) != cudaSuccess)
            rc = fail(HPF_ECUDA, "digamma kernel failed");
    }
    hpf_free(dx);
    hpf_free(dout);
    return rc;
}

}  // extern "C"

#include "hpf_batch_host.inl"
#include "hpf_scorer.inl"
#include "hpf_ingest.inl"
