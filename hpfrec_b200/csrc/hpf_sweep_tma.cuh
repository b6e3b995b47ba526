// Staged-gather variant of the CAVI sweep (sm_100a): the gathered factor rows are brought into shared
// memory by 1-D bulk async copies (cp.async.bulk.shared::cluster.global, the TMA engine; SASS UBLKCP)
// completing on mbarriers, eight rows deep per lane group, so gather latency is decoupled from
// registers and from the math.  Same contract as sweep_major_kernel (hpf_kernels.cuh):
//   acc[r, :] += sum_{n in segment} (Y[n] / dot(xown[r], xgat[c_n])) * xgat[c_n, :]
// Work split: 8 lanes own one nnz at a time (4 groups per warp); a group walks a contiguous chunk of
// the (panel, major)-sorted triples in batches of 8 (one coalesced triple per lane).  Stage t of a
// group's 8-deep ring always holds the row of the batch's t-th nnz; the lane that holds that nnz's
// column id arms the stage's mbarrier with the row size and issues the copy; after the group has
// consumed stage t it is immediately re-armed with the t-th nnz of the NEXT batch.
#pragma once
#include "hpf_device.cuh"

namespace hpf {

template <typename real, int VPL, int MINB>
__global__ void __launch_bounds__(256, MINB)
sweep_tma_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                 long long nnz, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                 real* __restrict__ acc, int ld) {
    constexpr int LPG = 8, NSTG = 8, EPV = Pack<real>::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (LPG - 1), g = lane / LPG;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const long long beg = group * (long long)chunk;
    if (beg >= nnz) return;  // whole groups leave together; everything below is group-scoped
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const unsigned gmask = group_mask<LPG>();

    const uint32_t row_bytes = (uint32_t)ld * (uint32_t)sizeof(real);
    const uint32_t warp_bytes = 32u * row_bytes + 32u * 8u;  // 4 groups x 8 stages of rows, then their mbarriers
    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * warp_bytes;
    const uint32_t rows0 = wbase + (uint32_t)(g * NSTG) * row_bytes;
    const uint32_t bars0 = wbase + 32u * row_bytes + (uint32_t)(g * NSTG) * 8u;
    mbar_init(bars0 + (uint32_t)gl * 8u, 1);
    mbar_fence_init();
    __syncwarp(gmask);

    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < ld;
    }
    Pack<real> own[VPL], sum[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        own[v] = pack_zero<real>();
        sum[v] = pack_zero<real>();
    }
    int cur = -1;

    // triples of the current batch, the next batch (its column ids feed the re-arming) and one more
    int r = -1, c = 0, rn = -1, cn = 0;
    real y = real(0), yn = real(0);
    if (beg + gl < end) {
        r = __ldg(row + beg + gl);
        c = __ldg(col + beg + gl);
        y = __ldg(val + beg + gl);
    }
    if (beg + LPG + gl < end) {
        rn = __ldg(row + beg + LPG + gl);
        cn = __ldg(col + beg + LPG + gl);
        yn = __ldg(val + beg + LPG + gl);
    }
    if (r >= 0) {  // prologue: every lane stages the row of its own nnz
        mbar_expect_tx(bars0 + (uint32_t)gl * 8u, row_bytes);
        bulk_g2s(rows0 + (uint32_t)gl * row_bytes, xgat + (size_t)c * ld, row_bytes, bars0 + (uint32_t)gl * 8u);
    }
    uint32_t phase = 0;
    for (long long base = beg; base < end; base += LPG) {
        int r2 = -1, c2 = 0;
        real y2 = real(0);
        const long long idx2 = base + 2 * LPG + gl;
        if (idx2 < end) {
            r2 = __ldg(row + idx2);
            c2 = __ldg(col + idx2);
            y2 = __ldg(val + idx2);
        }
#pragma unroll
        for (int t = 0; t < NSTG; ++t) {
            if (base + t >= end) break;  // uniform inside the group
            const int rr = __shfl_sync(gmask, r, t, LPG);
            const real yy = __shfl_sync(gmask, y, t, LPG);
            mbar_wait(bars0 + (uint32_t)t * 8u, phase);
            Pack<real> gth[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                gth[v] = act[v] ? lds_pack<real>(rows0 + (uint32_t)t * row_bytes + (uint32_t)off[v] * (uint32_t)sizeof(real))
                                : pack_zero<real>();
            __syncwarp(gmask);  // all lanes have read stage t: it may be overwritten
            if (gl == t && rn >= 0) {  // re-arm stage t with the t-th nnz of the next batch (this lane holds it)
                mbar_expect_tx(bars0 + (uint32_t)t * 8u, row_bytes);
                bulk_g2s(rows0 + (uint32_t)t * row_bytes, xgat + (size_t)cn * ld, row_bytes, bars0 + (uint32_t)t * 8u);
            }
            if (rr != cur) {
                if (cur >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
                        if (act[v]) red_add_pack(acc + (size_t)cur * ld + off[v], sum[v]);
                }
                cur = rr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = act[v] ? ldg_pack(xown + (size_t)cur * ld + off[v]) : pack_zero<real>();
                    sum[v] = pack_zero<real>();
                }
            }
            real s = real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) s = fma(own[v].v[e], gth[v].v[e], s);
            s = group_sum<LPG>(s, gmask);
            const real w = rdiv_fast(yy, s);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) sum[v].v[e] = fma(w, gth[v].v[e], sum[v].v[e]);
        }
        phase ^= 1u;
        r = rn;
        c = cn;
        y = yn;
        rn = r2;
        cn = c2;
        yn = y2;
    }
    (void)c;
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * ld + off[v], sum[v]);
    }
}

}  // namespace hpf
