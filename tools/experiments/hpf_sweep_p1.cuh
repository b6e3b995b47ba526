// sweep_rows_p1_kernel: the full-batch sweep for rows of LPG * VM + 1 sixteen-byte packs (sm_100a) -- k = 49..52 in
// fp32 (13 packs = 4 * 3 + 1), k = 49..50 in fp64 and k = 97..100 in fp32 (25 packs = 8 * 3 + 1): the headline row
// lengths.  Same contract as sweep_rows_kernel (hpf_sweep.cuh): one launch is one direction of update_phi +
// update_G_n_L_sh (pxi:551-621) without materialising phi,
//
//       acc[r, :] += sum_{n : row(n) = r}  (Y[n] / dot(xown[r, :], xgat[col(n), :])) * xgat[col(n), :]
//
// Why a second form.  ncu on sweep_rows_kernel's shared-memory ring (profiles/r02_ncu_full_sweep_rows_sg_fullrow.csv):
// the L1TEX data pipe is the limiter (80 %), because every gathered byte crosses it twice (LDGSTS in, LDS out) and a
// 128-bit shared-memory access costs one wavefront per quarter-warp that has ANY active lane.  With 8 lanes per row
// and 2 packs per lane a 13-pack row costs 2 + 2 wavefronts whether the copy moves 16 packs (whole stride) or 13
// (5 of 8 lanes active in every quarter).  Here the row is split into LPG * VM "main" packs, copied and read by the
// row's own lane group with every lane active, and ONE "extra" pack per row, which the first NG lanes of the warp
// handle for all NG rows of a step in one quarter-warp: 4 * VM + 1 wavefronts per NG rows instead of 8 per 4 --
// 1.625 instead of 2 per row per direction for 13 packs, and 7 L2 sectors per row instead of 8.  The price: the extra
// pack's partial dot product and the row's weight cross lanes with two shuffles per step, and lane x < NG keeps a
// second "current row" (that of lane group x) with its own flush.
//
// MEASURED AND REJECTED (B200, H workload, profiles/r02_tune_p1_form.jsonl; every shape verified against the COO
// kernel first, and the form passed the oracle parity tests in all shapes): fp32 (4 lanes x 3 packs + 1) best 1.19 /
// 1.17 ms per pass against 1.00 / 0.98 ms for sweep_rows_kernel; fp64 (8 x 3 + 1) 2.79 / 2.73 against 2.65 / 2.58.
// Fewer shared-memory wavefronts and 12-22 % fewer L2 sectors did not pay: with 4 lanes per row every LDGSTS touches
// 8 rows (twice the L1 tag look-ups and L2 requests of half the size), and the extra pack's divergent blocks add issue
// slots.  Not compiled into the library.
#pragma once
#include "hpf_sweep.cuh"

namespace hpf {

template <int VM, int NG, int D>
struct SweepP1Smem {
    static constexpr uint32_t TRIP = 1024u;                  // two batches of staged triples, first: 512-byte aligned
    static constexpr uint32_t SLOT = VM * 512u + NG * 16u;   // one step: NG rows, main packs lane-linear, then NG extras
    static constexpr uint32_t RING = D * SLOT;               // gathered rows in flight
    static constexpr uint32_t OWN = D * SLOT;                // own rows, staged as far ahead as a gathered row
    static constexpr uint32_t WARP = (TRIP + RING + OWN + 511u) / 512u * 512u;
};

// own-row staging as real branches (see stage_own_row in hpf_sweep.cuh for why these are not inlined)
template <int LPG, int VM>
__device__ __noinline__ void p1_stage_own_main(uint32_t dst, const char* src) {
#pragma unroll
    for (int v = 0; v < VM; ++v) cp_async16(dst + (uint32_t)v * 512u, src + v * (LPG * 16));
}
__device__ __noinline__ void p1_stage_own_extra(uint32_t dst, const char* src) { cp_async16(dst, src); }

// ngroups: lane groups of the launch = padded nnz / chunk, a multiple of NG.  chunk: nnz per lane group, a multiple
// of LPG.  D: steps in flight (divides LPG).  row_bytes: stride of the factor matrices in bytes (>= (LPG*VM+1)*16).
template <typename real, int LPG, int VM, int D, int MINB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_rows_p1_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val, long long ngroups,
                     int chunk, const real* __restrict__ xown, const real* __restrict__ xgat, real* __restrict__ acc, int ld) {
    constexpr int EPV = Pack<real>::N;
    constexpr int NG = 32 / LPG;
    constexpr int PX = LPG * VM;  // index of the extra pack
    using SM = SweepP1Smem<VM, NG, D>;
    static_assert(LPG % D == 0 && D >= 1, "steps in flight must divide the batch length");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG, g = lane / LPG;
    const int x = lane % NG;       // the lane group whose extra pack this lane handles (lanes < NG only)
    const bool isx = lane < NG;
    const long long wg0 = ((long long)blockIdx.x * (BLOCK / 32) + warp) * NG;
    if (wg0 >= ngroups) return;  // warp-uniform
    const long long beg = (wg0 + g) * (long long)chunk;
    const int nbatch = chunk / LPG;
    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);

    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * SM::WARP;
    // staged triples: swizzled as in sweep_rows_kernel; tripx = the region of lane group x
    constexpr int SWZ_MASK = (LPG >= 8 ? 7 : LPG - 1);
    const uint32_t swz = (uint32_t)((LPG >= 8 ? g : (g >> 1)) & SWZ_MASK);
    const uint32_t swzx = (uint32_t)((LPG >= 8 ? x : (x >> 1)) & SWZ_MASK);
    const uint32_t trip0 = (wbase + (uint32_t)(g * LPG) * 16u) ^ (swz * 16u);
    const uint32_t tripx0 = (wbase + (uint32_t)(x * LPG) * 16u) ^ (swzx * 16u);
    const uint32_t ring_m = wbase + SM::TRIP + (uint32_t)lane * 16u;                        // + slot * SLOT + v * 512
    const uint32_t ring_x = wbase + SM::TRIP + (uint32_t)VM * 512u + (uint32_t)x * 16u;     // + slot * SLOT
    const uint32_t own_m = ring_m + SM::RING, own_x = ring_x + SM::RING;

    const char* gat_m = reinterpret_cast<const char*>(xgat) + (unsigned)gl * 16u;
    const char* gat_x = reinterpret_cast<const char*>(xgat) + (unsigned)PX * 16u;
    const char* own_src_m = reinterpret_cast<const char*>(xown) + (unsigned)gl * 16u;
    const char* own_src_x = reinterpret_cast<const char*>(xown) + (unsigned)PX * 16u;
    asm volatile("" : "+l"(gat_m));
    asm volatile("" : "+l"(gat_x));

    auto load_triple = [&](int b, int& r, int& c, real& y) {
        const long long idx = beg + (long long)b * LPG + gl;
        r = __ldg(row + idx);
        c = __ldg(col + idx);
        y = __ldg(val + idx);
    };

    int r_staged = -1, r_staged_x = -1;
    // put step s in flight: the gathered row of every group (main packs by the group, extra pack by lane x), and the
    // own row of a group whose row id changes at that step
    auto stage = [&](int s, uint32_t tm, uint32_t tx) {
        const uint32_t slot = (uint32_t)(s % D) * SM::SLOT;
        int ra, ca;
        lds64(tm, ra, ca);
        const char* src = gat_m + (uint64_t)(unsigned)ca * row_bytes;
#pragma unroll
        for (int v = 0; v < VM; ++v) cp_async16(ring_m + slot + (uint32_t)v * 512u, src + v * (LPG * 16));
        if (ra != r_staged) {
            p1_stage_own_main<LPG, VM>(own_m + slot, own_src_m + (uint64_t)(unsigned)ra * row_bytes);
            r_staged = ra;
        }
        if (isx) {
            int rax, cax;
            lds64(tx, rax, cax);
            cp_async16(ring_x + slot, gat_x + (uint64_t)(unsigned)cax * row_bytes);
            if (rax != r_staged_x) {
                p1_stage_own_extra(own_x + slot, own_src_x + (uint64_t)(unsigned)rax * row_bytes);
                r_staged_x = rax;
            }
        }
        cp_async_commit();
    };

    Pack<real> own[VM], sum[VM], ownx = pack_zero<real>(), sumx = pack_zero<real>();
#pragma unroll
    for (int v = 0; v < VM; ++v) {
        own[v] = pack_zero<real>();
        sum[v] = pack_zero<real>();
    }
    int cur = -1, curx = -1;

    {
        int r, c;
        real y;
        load_triple(0, r, c, y);
        sts_triple(trip0 ^ ((uint32_t)gl * 16u), r, c, y);
        load_triple(nbatch > 1 ? 1 : 0, r, c, y);
        sts_triple((trip0 + 512u) ^ ((uint32_t)gl * 16u), r, c, y);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < D; ++t) stage(t, trip0 ^ ((uint32_t)t * 16u), tripx0 ^ ((uint32_t)t * 16u));

    uint32_t cb = 0u, nb = 512u;  // byte offsets of the current / next triple buffer
    for (int b = 0; b < nbatch; ++b) {
        int r2, c2;
        real y2;
        load_triple(b + 2 < nbatch ? b + 2 : nbatch - 1, r2, c2, y2);
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            const uint32_t slot = (uint32_t)(t % D) * SM::SLOT;
            int rr, rrx = 0;
            real yy, yx;
            lds_row_y((trip0 + cb) ^ ((uint32_t)t * 16u), rr, yy);
            if (isx) lds_row_y((tripx0 + cb) ^ ((uint32_t)t * 16u), rrx, yx);
            cp_async_wait<D - 1>();  // step t has landed
            if (rr != cur) {  // divergent between groups
                if (cur >= 0) {
#pragma unroll
                    for (int v = 0; v < VM; ++v) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
                }
                cur = rr;
#pragma unroll
                for (int v = 0; v < VM; ++v) {
                    own[v] = lds_pack<real>(own_m + slot + (uint32_t)v * 512u);
                    sum[v] = pack_zero<real>();
                }
            }
            if (isx && rrx != curx) {
                if (curx >= 0) red_add_pack(acc + (size_t)curx * ld + PX * EPV, sumx);
                curx = rrx;
                ownx = lds_pack<real>(own_x + slot);
                sumx = pack_zero<real>();
            }
            Pack<real> gv[VM], gvx = pack_zero<real>();
#pragma unroll
            for (int v = 0; v < VM; ++v) gv[v] = lds_pack<real>(ring_m + slot + (uint32_t)v * 512u);
            if (isx) gvx = lds_pack<real>(ring_x + slot);
            typename DotOf<real>::type d0, d1, dx;
#pragma unroll
            for (int v = 0; v < VM; ++v) {
                if (v & 1) d1.add(own[v], gv[v]);
                else d0.add(own[v], gv[v]);
            }
            dx.add(ownx, gvx);  // zero in lanes >= NG (gvx is zero there)
            real s = VM > 1 ? d0.total() + d1.total() : d0.total();
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            s += __shfl_sync(FULL, dx.total(), g);                  // lane g holds the extra pack's share of row g
            const real w = rdiv_rcp(yy, s);
#pragma unroll
            for (int v = 0; v < VM; ++v) axpy_pack(sum[v], w, gv[v]);
            const real wx = __shfl_sync(FULL, w, x * LPG);          // weight of lane group x, from its first lane
            axpy_pack(sumx, wx, gvx);                               // gvx is zero in lanes >= NG
            // ---- put step t + D in flight
            if (t + D < LPG) stage(t + D, (trip0 + cb) ^ ((uint32_t)(t + D) * 16u), (tripx0 + cb) ^ ((uint32_t)(t + D) * 16u));
            else stage(t + D, (trip0 + nb) ^ ((uint32_t)(t + D - LPG) * 16u), (tripx0 + nb) ^ ((uint32_t)(t + D - LPG) * 16u));
        }
        __syncwarp();
        sts_triple((trip0 + cb) ^ ((uint32_t)gl * 16u), r2, c2, y2);
        __syncwarp();
        const uint32_t tmp = cb;
        cb = nb;
        nb = tmp;
    }
    cp_async_wait<0>();
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VM; ++v) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
    }
    if (isx && curx >= 0) red_add_pack(acc + (size_t)curx * ld + PX * EPV, sumx);
}

}  // namespace hpf
