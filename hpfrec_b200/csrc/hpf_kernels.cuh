// Kernels of the HPF coordinate-ascent engine (sm_100a).  See DESIGN.md for the data layout and the
// roofline of each kernel.  Reference = david-cortes/hpfrec, hpfrec/cython_loops.pxi ("pxi").
//
// Key identity used throughout: the reference's multinomial parameter (update_phi, pxi:551-577)
//     phi[n,j] = Y[n] * softmax_j( Eu[u,j] + Ei[i,j] ),   Eu = psi(G_sh) - log(G_rt),  Ei likewise
// factorises into per-ROW quantities:  with  xu[u,j] = exp(Eu[u,j] - max_j Eu[u,:])  and xi likewise,
//     phi[n,j] = Y[n] * xu[u,j]*xi[i,j] / dot(xu[u,:], xi[i,:]).
// So psi/log/exp are evaluated once per factor ROW per iteration ((nU+nI)*k times) instead of once
// per nnz (nnz*k times), phi is never materialised, and the scatter (update_G_n_L_sh, pxi:613-621)
// becomes  G_sh[u,:] = a + xu[u,:] * sum_{n in u} w_n xi[i_n,:],   w_n = Y[n]/dot  (mirror for L_sh).
#pragma once
#include "hpf_device.cuh"
#include "hpf_sweep.cuh"

namespace hpf {

// =============================================================================================
// K1+K3 fused row update for full-batch CAVI (one lane group per factor row):
//   shp  = prior + x * acc                       pxi:239-249 (scatter result) ; acc is re-zeroed
//   rte  = shp_rate / rate[r] + colsum_other[j]  pxi:236 (users) / pxi:255 (items)
//   E[x] = shp / rte                             pxi:251 / 256  -> column sums (double) for the other side
//   rate[r] = add_rate + sum_j E[x]              pxi:258 / 259
//   x    = exp(psi(shp) - log(rte) - rowmax)     the per-row factor of next iteration's update_phi; written to x_out,
//                                                which is x itself, or a second buffer when the update runs UNDER the
//                                                other side's pass, which is still gathering the old factors
// MAT=true also stores shp and rte (needed for export / minibatch steps); lean iterations skip that.
// `direct` (robust mode, else NULL): phi sums of the nnz the sweep's rescue path handled, added to the
// shape as they are (not scaled by x) and re-zeroed.
// =============================================================================================
template <typename real, int LPG, int VPL, bool MAT>
__global__ void __launch_bounds__(256)
update_rows_kernel(int nrows, int ld, int k, const real* x, real* x_out, real* __restrict__ acc, real* __restrict__ direct,
                   real* __restrict__ shp_out, real* __restrict__ rte_out, real* __restrict__ rate,
                   const double* __restrict__ colsum_other, double* __restrict__ colsum_out,
                   real prior, real shp_rate, real add_rate) {
    constexpr int EPV = Pack<real>::N;
    extern __shared__ double s_col[];  // ld doubles
    for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
    __syncthreads();

    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;

    int off[VPL];
    bool act[VPL];
    real other[VPL][EPV];
    real csum[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            const int j = off[v] + e;
            other[v][e] = (act[v] && j < k) ? (real)colsum_other[j] : real(0);
            csum[v][e] = real(0);
        }
    }

    for (int r = g0; r < nrows; r += gstride) {
        const real inv = shp_rate / rate[r];
        Pack<real> shp[VPL], rte[VPL], E[VPL];
        real rowsum = real(0);
        real m = -INFINITY;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const Pack<real> xv = ld_pack(x + (size_t)r * ld + off[v]);
            const Pack<real> av = ld_pack(acc + (size_t)r * ld + off[v]);
            Pack<real> dv = pack_zero<real>();
            if (direct != nullptr) {
                dv = ld_pack(direct + (size_t)r * ld + off[v]);
                st_pack(direct + (size_t)r * ld + off[v], pack_zero<real>());
            }
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const bool real_col = off[v] + e < k;
                const real s_ = fma(xv.v[e], av.v[e], prior) + dv.v[e];
                const real t_ = inv + other[v][e];
                shp[v].v[e] = real_col ? s_ : real(0);
                rte[v].v[e] = real_col ? t_ : real(1);
                const real th = real_col ? rratio(s_, t_) : real(0);
                rowsum += th;
                csum[v][e] += th;
                const real lg = real_col ? elog(s_, t_) : -INFINITY;
                E[v].v[e] = lg;
                m = lg > m ? lg : m;
            }
        }
        rowsum = group_sum<LPG>(rowsum, gmask);
        m = group_max<LPG>(m, gmask);
        if (gl == 0) rate[r] = add_rate + rowsum;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            Pack<real> xn;
#pragma unroll
            for (int e = 0; e < EPV; ++e) xn.v[e] = (off[v] + e < k) ? rexp(E[v].v[e] - m) : real(0);
            st_pack(x_out + (size_t)r * ld + off[v], xn);
            st_pack(acc + (size_t)r * ld + off[v], pack_zero<real>());
            if (MAT) {
                st_pack(shp_out + (size_t)r * ld + off[v], shp[v]);
                st_pack(rte_out + (size_t)r * ld + off[v], rte[v]);
            }
        }
    }
    // column sums: registers -> shared (double) -> one global atomic per column per block
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        if (!act[v]) continue;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            if (off[v] + e < k) atomicAdd(&s_col[off[v] + e], (double)csum[v][e]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) atomicAdd(colsum_out + j, s_col[j]);
}

// =============================================================================================
// Fused reduce-scatter + item update + all-gather over NVLink peer memory (multi-GPU).
// Rank `rank` owns item rows [r0, r1).  For each owned row it (1) loads the row's partial sums from
// EVERY rank's item_sums buffer (P2P LDG.128 over NVLink, fixed rank order => deterministic and the
// same bits on every replica), (2) applies the item update of update_rows_kernel (pxi:255-259 + next
// iteration's softmax factor), (3) stores the new factor row and rate (and, when MAT, shape/rate
// matrices) into EVERY rank's replica (P2P STG.128).  Column sums of Beta over the owned rows are
// accumulated locally (the caller all-reduces those k doubles, which is also the barrier that
// publishes the peer writes).
// =============================================================================================
constexpr int kMaxPeers = 16;
struct PeerTable {
    void* acc[kMaxPeers];
    void* x[kMaxPeers];
    void* rate[kMaxPeers];
    void* shp[kMaxPeers];
    void* rte[kMaxPeers];
    // NVSwitch multicast views of the same five buffers (NULL without multicast support): one multimem.ld_reduce on
    // mc_acc returns the sum over every rank's copy, computed inside the switch; a store to mc_x lands in every replica
    void *mc_acc, *mc_x, *mc_rate, *mc_shp, *mc_rte;
    int world;
    int rank;
};

// multimem (NVLS) accessors: SASS LDGMC.E.ADD / STG on a multicast address
__device__ __forceinline__ Pack<float> mc_ld_reduce(const float* p) {
    Pack<float> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ Pack<double> mc_ld_reduce(const double* p) {
    Pack<double> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(r.v[1]) : "l"(p + 1) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st(float* p, const Pack<float>& r) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                 "f"(r.v[3])
                 : "memory");
}
__device__ __forceinline__ void mc_st(double* p, const Pack<double>& r) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(r.v[0]) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p + 1), "d"(r.v[1]) : "memory");
}
__device__ __forceinline__ void mc_st_scalar(float* p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_st_scalar(double* p, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// MC: the reduce-scatter is ONE multimem.ld_reduce per pack (the NVSwitch adds the ranks' copies: this GPU receives
// 1/world of the bytes the pull form moves) and the all-gather ONE multimem.st per pack.
template <typename real, int LPG, int VPL, bool MAT, bool MC>
__global__ void __launch_bounds__(256)
update_items_peer_kernel(int r0, int r1, int ld, int k, PeerTable pt, const double* __restrict__ colsum_other,
                         double* __restrict__ colsum_out, real prior, real shp_rate, real add_rate, int pre_reduced) {
    constexpr int EPV = Pack<real>::N;
    extern __shared__ double s_col[];
    for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
    __syncthreads();
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    const real* x_local = (const real*)pt.x[pt.rank];
    const real* rate_local = (const real*)pt.rate[pt.rank];
    int off[VPL];
    bool act[VPL];
    real other[VPL][EPV], csum[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            const int j = off[v] + e;
            other[v][e] = (act[v] && j < k) ? (real)colsum_other[j] : real(0);
            csum[v][e] = real(0);
        }
    }
    for (int r = r0 + g0; r < r1; r += gstride) {
        const real inv = shp_rate / rate_local[r];
        Pack<real> shp[VPL], rte[VPL], E[VPL], asum[VPL];
        real rowsum = real(0);
        real m = -INFINITY;
        // all peer loads of a batch of 8 ranks are issued before the first add: NVLink reads have
        // ~2 us latency, so memory-level parallelism per thread is what fills the links
#pragma unroll
        for (int v = 0; v < VPL; ++v) asum[v] = pack_zero<real>();
        if (pre_reduced) {  // reduce_items_peer_kernel already left the all-rank sums in this rank's own buffer
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v]) asum[v] = ld_pack((const real*)pt.acc[pt.rank] + (size_t)r * ld + off[v]);
        } else if (MC) {
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v]) asum[v] = mc_ld_reduce((const real*)pt.mc_acc + (size_t)r * ld + off[v]);
        }
        for (int p0 = 0; !MC && !pre_reduced && p0 < pt.world; p0 += 8) {
            Pack<real> pv[VPL][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool have = p0 + q < pt.world;
                const real* base = (const real*)pt.acc[have ? p0 + q : pt.rank];
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    pv[v][q] = (have && act[v]) ? ld_pack(base + (size_t)r * ld + off[v]) : pack_zero<real>();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)  // fixed rank order: every replica receives identical bits
#pragma unroll
                for (int v = 0; v < VPL; ++v)
#pragma unroll
                    for (int e = 0; e < EPV; ++e) asum[v].v[e] += pv[v][q].v[e];
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const size_t at = (size_t)r * ld + off[v];
            const Pack<real> av = asum[v];
            const Pack<real> xv = ld_pack(x_local + at);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const bool real_col = off[v] + e < k;
                const real s_ = fma(xv.v[e], av.v[e], prior);
                const real t_ = inv + other[v][e];
                shp[v].v[e] = real_col ? s_ : real(0);
                rte[v].v[e] = real_col ? t_ : real(1);
                const real th = real_col ? rratio(s_, t_) : real(0);
                rowsum += th;
                csum[v][e] += th;
                const real lg = real_col ? elog(s_, t_) : -INFINITY;
                E[v].v[e] = lg;
                m = lg > m ? lg : m;
            }
        }
        rowsum = group_sum<LPG>(rowsum, gmask);
        m = group_max<LPG>(m, gmask);
        const real new_rate = add_rate + rowsum;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const size_t at = (size_t)r * ld + off[v];
            Pack<real> xn;
#pragma unroll
            for (int e = 0; e < EPV; ++e) xn.v[e] = (off[v] + e < k) ? rexp(E[v].v[e] - m) : real(0);
            if (MC) {
                mc_st((real*)pt.mc_x + at, xn);
                if (MAT) {
                    mc_st((real*)pt.mc_shp + at, shp[v]);
                    mc_st((real*)pt.mc_rte + at, rte[v]);
                }
            }
            for (int p = 0; !MC && p < pt.world; ++p) {
                st_pack((real*)pt.x[p] + at, xn);
                if (MAT) {
                    st_pack((real*)pt.shp[p] + at, shp[v]);
                    st_pack((real*)pt.rte[p] + at, rte[v]);
                }
            }
        }
        if (gl == 0) {
            if (MC) mc_st_scalar((real*)pt.mc_rate + r, new_rate);
            for (int p = 0; !MC && p < pt.world; ++p) ((real*)pt.rate[p])[r] = new_rate;
        }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        if (!act[v]) continue;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            if (off[v] + e < k) atomicAdd(&s_col[off[v] + e], (double)csum[v][e]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) atomicAdd(colsum_out + j, s_col[j]);
}

// =============================================================================================
// The reduce-scatter half of the exchange on its own, so that it can run on a second stream UNDER the user-major
// pass and the user update (it only needs every rank's item-major pass to be finished): rank `rank` sums its slice
// of item rows over all ranks' partial sums -- one multimem.ld_reduce per pack (MC) or P2P loads in fixed rank order
// -- and leaves the totals in its OWN buffer, in place (no other rank ever reads those rows of this buffer).
// update_items_peer_kernel then runs with pre_reduced = 1.
// =============================================================================================
template <typename real, int LPG, int VPL, bool MC>
__global__ void __launch_bounds__(256)
reduce_items_peer_kernel(int r0, int r1, int ld, int k, PeerTable pt) {
    constexpr int EPV = Pack<real>::N;
    constexpr int U = 4;  // rows per lane group per trip: the kernel runs on a FEW CTAs (it shares the SMs with the
                          // user-major pass), so the loads in flight that fill NVLink have to come from each thread
    const int gl = (threadIdx.x & 31) % LPG;
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    real* mine = (real*)pt.acc[pt.rank];
    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
    }
    for (int rb = r0 + g0; rb < r1; rb += gstride * U) {
        Pack<real> asum[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VPL; ++v) asum[u][v] = pack_zero<real>();
        if (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = rb + u * gstride;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (r < r1 && act[v]) asum[u][v] = mc_ld_reduce((const real*)pt.mc_acc + (size_t)r * ld + off[v]);
            }
        } else {
            for (int p = 0; p < pt.world; ++p) {  // fixed rank order, as in the fused kernel: identical bits everywhere
                const real* base = (const real*)pt.acc[p];
                Pack<real> pv[U][VPL];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = rb + u * gstride;
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
                        pv[u][v] = (r < r1 && act[v]) ? ld_pack(base + (size_t)r * ld + off[v]) : pack_zero<real>();
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
#pragma unroll
                        for (int e = 0; e < EPV; ++e) asum[u][v].v[e] += pv[u][v].v[e];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = rb + u * gstride;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (r < r1 && act[v]) st_pack(mine + (size_t)r * ld + off[v], asum[u][v]);
        }
    }
}

// =============================================================================================
// K1 alone: x rows from materialised (shp, rte); optional row list (minibatch) and optional column
// sums of shp/rte (used to seed Beta.sum(axis=0) after a state upload).
// =============================================================================================
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
rows_to_x_kernel(int nrows, const int* __restrict__ rows, int ld, int k, const real* __restrict__ shp,
                 const real* __restrict__ rte, real* __restrict__ x, double* __restrict__ colsum_out) {
    constexpr int EPV = Pack<real>::N;
    extern __shared__ double s_col[];
    if (colsum_out != nullptr) {
        for (int j = threadIdx.x; j < ld; j += blockDim.x) s_col[j] = 0.0;
        __syncthreads();
    }
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const int g0 = blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const int gstride = gridDim.x * groups_per_block;
    int off[VPL];
    bool act[VPL];
    real csum[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
#pragma unroll
        for (int e = 0; e < EPV; ++e) csum[v][e] = real(0);
    }
    for (int q = g0; q < nrows; q += gstride) {
        const int r = rows ? rows[q] : q;
        Pack<real> E[VPL];
        real m = -INFINITY;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const Pack<real> sv = ld_pack(shp + (size_t)r * ld + off[v]);
            const Pack<real> tv = ld_pack(rte + (size_t)r * ld + off[v]);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const bool real_col = off[v] + e < k;
                const real lg = real_col ? elog(sv.v[e], tv.v[e]) : -INFINITY;
                E[v].v[e] = lg;
                m = lg > m ? lg : m;
                if (real_col) csum[v][e] += sv.v[e] / tv.v[e];
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
        }
    }
    if (colsum_out != nullptr) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                if (off[v] + e < k) atomicAdd(&s_col[off[v] + e], (double)csum[v][e]);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < k; j += blockDim.x) atomicAdd(colsum_out + j, s_col[j]);
    }
}

// shp = prior + x * acc (+ direct) for every row (finishes a stand-alone hpf_update_shapes call)
template <typename real>
__global__ void finish_shapes_kernel(long long n_elems, const real* __restrict__ x, const real* __restrict__ acc,
                                     const real* __restrict__ direct, real* __restrict__ shp, real prior) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_elems) shp[i] = fma(x[i], acc[i], prior) + (direct ? direct[i] : real(0));
}

// =============================================================================================
// ingest helpers
// =============================================================================================
template <typename IDX>
__global__ void convert_index_kernel(const IDX* __restrict__ in, int* __restrict__ out, long long n,
                                     long long limit, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = (long long)in[i];
    if (v < 0 || v >= limit) atomicExch(bad, 1);
    out[i] = (int)v;
}

// sort key of one ordering: (panel of the minor id, major id)
__global__ void make_keys_kernel(const int* __restrict__ major, const int* __restrict__ minor,
                                 long long n, int minor_per_panel, unsigned long long major_span,
                                 unsigned long long* __restrict__ keys, unsigned* __restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long panel = (unsigned long long)(minor[i] / minor_per_panel);
    keys[i] = panel * major_span + (unsigned long long)major[i];
    perm[i] = (unsigned)i;
}

template <typename real>
__global__ void apply_order_kernel(const unsigned long long* __restrict__ keys_sorted,
                                   const unsigned* __restrict__ perm, long long n,
                                   unsigned long long major_span, const int* __restrict__ minor,
                                   const real* __restrict__ val, int* __restrict__ out_row,
                                   int* __restrict__ out_col, real* __restrict__ out_val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned p = perm[i];
    out_row[i] = (int)(keys_sorted[i] % major_span);
    out_col[i] = minor[p];
    out_val[i] = val[p];
}

// zero-count padding behind the n sorted triples of an ordering: repeats the last (row, col) so that the
// sweep sees neither a row change nor an invalid address, and contributes exactly 0
template <typename real>
__global__ void pad_order_kernel(int* __restrict__ row, int* __restrict__ col, real* __restrict__ val, long long n,
                                 int pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pad) return;
    row[n + i] = row[n - 1];
    col[n + i] = col[n - 1];
    val[n + i] = real(0);
}

// strided copy helpers between packed (n x k) caller layout and padded (n x ld) engine layout
template <typename real>
__global__ void pad_rows_kernel(const real* __restrict__ in, real* __restrict__ out, long long nrows,
                                int k, int ld, real fill) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ld) return;
    const long long r = i / ld;
    const int j = (int)(i - r * ld);
    out[i] = j < k ? in[r * k + j] : fill;
}
template <typename real>
__global__ void unpad_rows_kernel(const real* __restrict__ in, const real* __restrict__ denom,
                                  real* __restrict__ out, long long nrows, int k, int ld) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * k) return;
    const long long r = i / k;
    const int j = (int)(i - r * k);
    const real v = in[r * ld + j];
    out[i] = denom ? v / denom[r * ld + j] : v;
}

// =============================================================================================
// K6/K7  llk_plus_rmse (pxi:627-658) + sum_prediction (pxi:816-825) + predict_multiple (pxi:803-810)
//   yhat = sum_j (G_sh/G_rt)[u,j] * (L_sh/L_rt)[i,j]   evaluated from the materialised state
//   out[0] += Y log yhat [- lgamma(Y+1)],  out[1] += (Y-yhat)^2,  out[2] += yhat   (double sums)
// =============================================================================================
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
score_kernel(const int* __restrict__ iu, const int* __restrict__ ii, const real* __restrict__ val,
             long long n, const real* __restrict__ theta, const real* __restrict__ beta, int ld,
             int full_llk, double* __restrict__ sums, real* __restrict__ pred) {
    constexpr int EPV = Pack<real>::N;
    __shared__ double s_part[3];
    if (threadIdx.x < 3) s_part[threadIdx.x] = 0.0;
    __syncthreads();
    const int gl = (threadIdx.x & 31) % LPG;
    const unsigned gmask = group_mask<LPG>();
    const int groups_per_block = blockDim.x / LPG;
    const long long g0 = (long long)blockIdx.x * groups_per_block + threadIdx.x / LPG;
    const long long gstride = (long long)gridDim.x * groups_per_block;
    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < ld;
    }
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
    // every group of the warp runs the same number of trips (shuffles need converged groups only,
    // but keeping warps converged is cheaper); out-of-range trips are masked
    const long long trips = (n + gstride - 1) / gstride;
    for (long long t = 0; t < trips; ++t) {
        const long long q = g0 + t * gstride;
        const bool live = q < n;
        const int u = live ? iu[q] : 0;
        const int i = live ? ii[q] : 0;
        real s = real(0);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (!act[v]) continue;
            const Pack<real> a = ldg_pack(theta + (size_t)u * ld + off[v]);
            const Pack<real> b = ldg_pack(beta + (size_t)i * ld + off[v]);
#pragma unroll
            for (int e = 0; e < EPV; ++e) s = fma(a.v[e], b.v[e], s);
        }
        s = group_sum<LPG>(s, gmask);
        if (live && gl == 0) {
            if (pred != nullptr) pred[q] = s;
            if (sums != nullptr) {
                const double y = (double)val[q];
                const double yh = (double)s;
                l0 += full_llk ? y * log(yh) - lgamma(y + 1.0) : (double)((real)y * rlog(s));
                const real d = (real)y - s;
                l1 += (double)(d * d);
                l2 += yh;
            }
        }
    }
    if (sums != nullptr) {
        if (gl == 0) {
            atomicAdd(&s_part[0], l0);
            atomicAdd(&s_part[1], l1);
            atomicAdd(&s_part[2], l2);
        }
        __syncthreads();
        if (threadIdx.x < 3) atomicAdd(sums + threadIdx.x, s_part[threadIdx.x]);
    }
}

// materialise E[x] = shp / rte (padded layout, pad columns -> 0) and its column sums
template <typename real>
__global__ void ratio_rows_kernel(long long nrows, int ld, int k, const real* __restrict__ shp,
                                  const real* __restrict__ rte, real* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ld) return;
    const int j = (int)(i % ld);
    out[i] = j < k ? shp[i] / rte[i] : real(0);
}

}  // namespace hpf
