// Minibatch (SVI / partial_fit) kernels and small reductions of the HPF engine (sm_100a).
// Reference semantics: hpfrec/cython_loops.pxi ("pxi") user-epoch body 275-325, item-epoch body
// 329-377, Cython partial_fit 423-473.  "major" = the batched side, "minor" = the opposite side.
#pragma once
#include "hpf_device.cuh"

namespace hpf {

template <typename real>
__global__ void digamma_kernel(const real* __restrict__ x, real* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = digamma(x[i]);
}

// column sums (double) of a padded (nrows x ld) matrix, optionally of the elementwise ratio num/den
template <typename real>
__global__ void __launch_bounds__(256)
colsum_kernel(long long nrows, int ld, int k, const real* __restrict__ num, const real* __restrict__ den,
              double* __restrict__ out) {
    extern __shared__ double s_col[];
    for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
    __syncthreads();
    // thread t owns column (t % ld) of rows t/ld, t/ld + rows_per_pass, ... ; blockDim is 256 and the
    // host guarantees ld <= 256 here, wider rows loop over column tiles
    for (int c0 = 0; c0 < ld; c0 += blockDim.x) {
        const int rows_per_pass = (ld - c0 >= (int)blockDim.x) ? 1 : blockDim.x / (ld - c0);
        const int width = (ld - c0 >= (int)blockDim.x) ? blockDim.x : (ld - c0);
        const int j = c0 + threadIdx.x % width;
        const int sub = threadIdx.x / width;
        if (sub >= rows_per_pass || j >= k) continue;
        double s = 0.0;
        for (long long r = (long long)blockIdx.x * rows_per_pass + sub; r < nrows;
             r += (long long)gridDim.x * rows_per_pass) {
            const real v = num[r * ld + j];
            s += den ? (double)(v / den[r * ld + j]) : (double)v;
        }
        atomicAdd(&s_col[j], s);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) atomicAdd(out + j, s_col[j]);
}

// out[0] = sum_j a[j] * b[j]   (the all-pairs shortcut term of pxi:78)
__global__ void dot_cols_kernel(int k, const double* __restrict__ a, const double* __restrict__ b,
                                double* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += a[j] * b[j];
        out[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// batch prepare: for every listed row compute the softmax factor x from the CURRENT (shp, rte)
// (what update_phi reads at pxi:292-298 / 438-440), zero the row's accumulator, stamp membership.
// ---------------------------------------------------------------------------------------------
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
batch_prepare_kernel(int nrows, const int* __restrict__ rows, int ld, int k, const real* __restrict__ shp,
                     const real* __restrict__ rte, real* __restrict__ x, real* __restrict__ acc,
                     real* __restrict__ direct, int* __restrict__ stamp, int step) {
    constexpr int EPV = Pack<real>::N;
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
    }
    for (int q = g0; q < nrows; q += gstride) {
        // rows == nullptr: walk ALL rows and prepare those already stamped for this step (the unique
        // opposite-side ids of a device-assembled batch, see batch_expand_kernel)
        const int r = rows ? rows[q] : q;
        if (!rows && stamp[r] != step) continue;  // uniform inside the lane group
        Pack<real> E[VPL];
        real m = -INFINITY;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const Pack<real> sv = ld_pack(shp + (size_t)r * ld + off[v]);
            const Pack<real> tv = ld_pack(rte + (size_t)r * ld + off[v]);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const real lg = (off[v] + e < k) ? elog(sv.v[e], tv.v[e]) : -INFINITY;
                E[v].v[e] = lg;
                m = lg > m ? lg : m;
            }
        }
        m = group_max<LPG>(m, gmask);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            Pack<real> xn;
#pragma unroll
            for (int e = 0; e < EPV; ++e) xn.v[e] = (off[v] + e < k) ? rexp(E[v].v[e] - m) : real(0);
            st_pack(x + (size_t)r * ld + off[v], xn);
            st_pack(acc + (size_t)r * ld + off[v], pack_zero<real>());
            if (direct != nullptr) st_pack(direct + (size_t)r * ld + off[v], pack_zero<real>());
        }
        if (rows && gl == 0) stamp[r] = step;
    }
}

// ---------------------------------------------------------------------------------------------
// major (batched) side, ALL rows:
//   rte[r,:]  = shp_rate/rate[r] + colsum_minor            pxi:300 / 352 / 443 / 446 (full overwrite)
//   shp[r,:]  = prior + x*acc            for batch rows     pxi:304-314 (local rows are replaced)
//   colsum_major += shp/rte                                 (Theta.sum(axis=0) of pxi:320 / Beta.sum of 372)
//   rate[r]   = rho*(add + sum_j shp/rte) + (1-rho)*rate[r]   batch rows (pxi:324/377) or all rows
//                                                            when blend_all (partial_fit, pxi:472-473)
// ---------------------------------------------------------------------------------------------
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
batch_major_kernel(int nrows, int ld, int k, const real* __restrict__ x, const real* __restrict__ acc,
                   const real* __restrict__ direct, real* __restrict__ shp, real* __restrict__ rte, real* __restrict__ rate,
                   const int* __restrict__ stamp, int step, const double* __restrict__ colsum_minor,
                   double* __restrict__ colsum_major, real prior, real shp_rate, real add_rate, real rho,
                   int blend_all) {
    constexpr int EPV = Pack<real>::N;
    extern __shared__ double s_col[];
    for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
    __syncthreads();
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    const real prev = real(1) - rho;
    int off[VPL];
    bool act[VPL];
    real other[VPL][EPV], csum[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            other[v][e] = (act[v] && off[v] + e < k) ? (real)colsum_minor[off[v] + e] : real(0);
            csum[v][e] = real(0);
        }
    }
    for (int r = g0; r < nrows; r += gstride) {
        const bool inb = stamp[r] == step;
        const real old_rate = rate[r];
        const real inv = shp_rate / old_rate;
        real rowsum = real(0);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            Pack<real> sv;
            if (inb) {
                const Pack<real> xv = ld_pack(x + (size_t)r * ld + off[v]);
                const Pack<real> av = ld_pack(acc + (size_t)r * ld + off[v]);
                const Pack<real> dv = direct ? ld_pack(direct + (size_t)r * ld + off[v]) : pack_zero<real>();
#pragma unroll
                for (int e = 0; e < EPV; ++e) sv.v[e] = (off[v] + e < k) ? fma(xv.v[e], av.v[e], prior) + dv.v[e] : real(0);
                st_pack(shp + (size_t)r * ld + off[v], sv);
            } else {
                sv = ld_pack(shp + (size_t)r * ld + off[v]);
            }
            Pack<real> tv;
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const bool real_col = off[v] + e < k;
                tv.v[e] = real_col ? inv + other[v][e] : real(1);
                const real th = real_col ? sv.v[e] / tv.v[e] : real(0);
                rowsum += th;
                csum[v][e] += th;
            }
            st_pack(rte + (size_t)r * ld + off[v], tv);
        }
        rowsum = group_sum<LPG>(rowsum, gmask);
        if (gl == 0 && (inb || blend_all)) rate[r] = rho * (add_rate + rowsum) + prev * old_rate;
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        if (!act[v]) continue;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            if (off[v] + e < k) atomicAdd(&s_col[off[v] + e], (double)csum[v][e]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) atomicAdd(colsum_major + j, s_col[j]);
}

// ---------------------------------------------------------------------------------------------
// minor (opposite) side; rows = the unique minor ids of the batch, or ALL rows when blend_all:
//   batch rows:  shp = rho*mult*(prior + x*acc) + (1-rho)*shp                 pxi:316 / 368
//                rte = rho*(shp_rate/rate[r] + colsum_major) + (1-rho)*rte    pxi:320 / 372
//   rate[r] = rho*(add + sum_j shp/rte) + (1-rho)*rate[r]    batch rows, or all rows when blend_all
// ---------------------------------------------------------------------------------------------
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
batch_minor_kernel(int nrows, const int* __restrict__ rows, int ld, int k, const real* __restrict__ x,
                   const real* __restrict__ acc, const real* __restrict__ direct, real* __restrict__ shp, real* __restrict__ rte,
                   real* __restrict__ rate, const int* __restrict__ stamp, int step,
                   const double* __restrict__ colsum_major, double* __restrict__ colsum_minor, real prior,
                   real shp_rate, real add_rate, real rho, real mult, int blend_all) {
    constexpr int EPV = Pack<real>::N;
    // colsum_minor += (new - old) expectation of the rows this step changes, so that the next step does not
    // have to re-sum the whole side (double accumulation; shared memory first, one global atomic per column)
    extern __shared__ double s_col[];
    for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
    __syncthreads();
    double dsum[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int e = 0; e < EPV; ++e) dsum[v][e] = 0.0;
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    const real prev = real(1) - rho;
    const real rm = rho * mult;
    int off[VPL];
    bool act[VPL];
    real other[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            other[v][e] = (act[v] && off[v] + e < k) ? (real)colsum_major[off[v] + e] : real(0);
    }
    for (int q = g0; q < nrows; q += gstride) {
        const int r = rows ? rows[q] : q;
        const bool inb = rows ? true : (stamp[r] == step);
        if (!inb && !blend_all) continue;  // all-rows walk of an SVI step: only stamped rows change
        const real old_rate = rate[r];
        const real inv = shp_rate / old_rate;
        real rowsum = real(0);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            Pack<real> sv = ld_pack(shp + (size_t)r * ld + off[v]);
            Pack<real> tv = ld_pack(rte + (size_t)r * ld + off[v]);
            if (inb) {
                const Pack<real> xv = ld_pack(x + (size_t)r * ld + off[v]);
                const Pack<real> av = ld_pack(acc + (size_t)r * ld + off[v]);
                const Pack<real> dv = direct ? ld_pack(direct + (size_t)r * ld + off[v]) : pack_zero<real>();
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    if (off[v] + e < k) {
                        const real old_e = sv.v[e] / tv.v[e];
                        sv.v[e] = rm * (fma(xv.v[e], av.v[e], prior) + dv.v[e]) + prev * sv.v[e];
                        tv.v[e] = rho * (inv + other[v][e]) + prev * tv.v[e];
                        dsum[v][e] += (double)(sv.v[e] / tv.v[e]) - (double)old_e;
                    }
                }
                st_pack(shp + (size_t)r * ld + off[v], sv);
                st_pack(rte + (size_t)r * ld + off[v], tv);
            }
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                if (off[v] + e < k) rowsum += sv.v[e] / tv.v[e];
        }
        rowsum = group_sum<LPG>(rowsum, gmask);
        if (gl == 0) rate[r] = rho * (add_rate + rowsum) + prev * old_rate;
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        if (!act[v]) continue;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            if (off[v] + e < k && dsum[v][e] != 0.0) atomicAdd(&s_col[off[v] + e], dsum[v][e]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x)
        if (s_col[j] != 0.0) atomicAdd(colsum_minor + j, s_col[j]);
}

// ---------------------------------------------------------------------------------------------
// K5  device-side minibatch assembly (replaces get_i_batch_pass1/2 + np.unique, pxi:27-42, 774-797):
//   count:   cnt[q] = ptr[ids[q]+1] - ptr[ids[q]]            (then an exclusive scan gives off[])
//   expand:  for every nnz of every listed row, copy the triple into compact (major, minor, y) arrays
//            in row-after-row order and stamp the minor id as a member of this step's batch
// ---------------------------------------------------------------------------------------------
__global__ void batch_count_kernel(int n_ids, const int* __restrict__ ids, const int* __restrict__ ptr,
                                   int* __restrict__ cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_ids) cnt[q] = ptr[ids[q] + 1] - ptr[ids[q]];
}

template <typename real>
__global__ void __launch_bounds__(256)
batch_expand_kernel(int n_ids, const int* __restrict__ ids, const int* __restrict__ ptr,
                    const int* __restrict__ off, long long total, const int* __restrict__ src_minor,
                    const real* __restrict__ src_val, int* __restrict__ out_major, int* __restrict__ out_minor,
                    real* __restrict__ out_val, int* __restrict__ stamp_minor, int step) {
    // one thread per OUTPUT triple (balanced whatever the row degrees: an item batch holds rows of 10^5 nnz next to
    // rows of one): binary search of the position in the exclusive offsets gives the listed row it belongs to
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int lo = 0, hi = n_ids;  // off[lo] <= t < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= (int)t) lo = mid;
        else hi = mid;
    }
    const int id = __ldg(ids + lo);
    const int src = __ldg(ptr + id) + ((int)t - __ldg(off + lo));
    const int m = __ldg(src_minor + src);
    out_major[t] = id;
    out_minor[t] = m;
    out_val[t] = __ldg(src_val + src);
    stamp_minor[m] = step;  // benign race: every writer stores the same value
}

// row pointer array of a sorted ordering: ptr[r] = first position with row >= r
__global__ void row_ptr_kernel(const int* __restrict__ row, long long nnz, int nrows, int* __restrict__ ptr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    const int cur = i < nnz ? row[i] : nrows;
    const int prv = i > 0 ? row[i - 1] : -1;
    for (int r = prv + 1; r <= cur; ++r) ptr[r] = (int)i;
}

}  // namespace hpf
