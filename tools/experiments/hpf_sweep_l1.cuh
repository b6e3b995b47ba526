// L1-staged form of the full-batch sweep (sm_100a).  Same contract as sweep_rows_kernel (hpf_sweep.cuh): one
// launch is one direction of update_phi + update_G_n_L_sh (pxi:551-621) without materialising phi,
//
//       acc[r, :] += sum_{n : row(n) = r}  (Y[n] / dot(xown[r, :], xgat[col(n), :])) * xgat[col(n), :]
//
// Why another form.  ncu on the shared-memory ring (profiles/r02_ncu_full_sweep_rows_sg_fullrow.csv) shows the L1TEX
// data pipe at 80 % with 4.7 wavefronts per nnz: every gathered byte crosses the pipe twice (LDGSTS into shared memory,
// LDS out of it).  Loading the gathered row straight into registers crosses it once, but then the rows in flight live
// in registers (32 per lane for 4 rows) and occupancy pays for the latency.  Here the rows in flight live in the L1
// data cache instead: D steps ahead every lane of a group TOUCHES one 32-byte sector of the row its group will need
// (prefetch.global.L1 -> CCTL.E.PF1, or a 4-byte load whose value is never needed), and the step itself loads the row
// with plain 128-bit loads that hit L1 (one crossing of the data pipe, ~40 cycles instead of an L2/DRAM round trip).
// Shared memory only holds the staged triples (1 KB per warp), so almost the whole 256 KB array is L1.
//
// MEASURED AND REJECTED (B200, H workload, profiles/r02_tune_l1_staged.jsonl): best shape 1.26 / 1.21 ms per pass
// against 1.00 / 0.98 ms for the shared-memory ring; without any touch the same kernel takes 1.32 / 1.24 ms, so the
// L1 prefetch buys 5 % and the form as a whole loses 25 %.  4-byte touch loads are worse still (1.7-2.0 ms).  An
// L1-allocating load crosses the L1 data array twice as well (fill + read), so the wavefront saving this form was
// built for does not exist.  Not compiled into the library; kept here with the dispatch fragment that drove it.
//
// Needs whole-stride rows (row stride == LPG * VPL * 16 bytes): the pad packs of a row are zero, so no per-pack
// predicates exist outside the flush.
#pragma once
#include "hpf_sweep.cuh"

namespace hpf {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ int touch_b32(const void* p) {
    int v;
    asm volatile("ld.global.ca.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <typename T>
__device__ __forceinline__ T ldg_noalloc(const T* p);
template <>
__device__ __forceinline__ int ldg_noalloc<int>(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ float ldg_noalloc<float>(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ double ldg_noalloc<double>(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// the step's real load: L1-cached; EF = mark the line evict_first once it has been consumed
template <typename real, bool EF>
__device__ __forceinline__ Pack<real> ldg_pack_l1(const void* p) {
    Pack<real> r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    if (EF)
        asm volatile("ld.global.nc.L1::evict_first.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    else
        asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    return r;
}
// whole staged triple with one LDS.128: float {row, col, y, row}; double {row, col, y}
__device__ __forceinline__ void lds_triple(uint32_t addr, int& r, int& c, float& y) {
    int yb, r2;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r), "=r"(c), "=r"(yb), "=r"(r2) : "r"(addr));
    y = __int_as_float(yb);
}
__device__ __forceinline__ void lds_triple(uint32_t addr, int& r, int& c, double& y) {
    int lo, hi;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r), "=r"(c), "=r"(lo), "=r"(hi) : "r"(addr));
    y = __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}

// D:     touch distance in steps (1..LPG).
// LA:    1 = the step's real loads are issued one step ahead into a second register set; 0 = at the step itself.
// TOUCH: 0 none (plain register form, the baseline of the measurement), 1 prefetch.global.L1, 2 4-byte loads.
// EF:    real loads carry L1::evict_first.
template <typename real, int LPG, int VPL, int D, int LA, int MINB, int BLOCK, int TOUCH, bool EF>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_rows_l1_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val, long long ngroups,
                     int chunk, const real* __restrict__ xown, const real* __restrict__ xgat, real* __restrict__ acc, int kw) {
    constexpr int EPV = Pack<real>::N;
    constexpr int NG = 32 / LPG;
    constexpr unsigned ROWB = LPG * VPL * 16u;        // bytes of one row (whole stride)
    constexpr int LD = (int)(ROWB / sizeof(real));
    constexpr int NSECT = (int)(ROWB / 32u);          // 32-byte sectors of one row
    constexpr int TPL = (NSECT + LPG - 1) / LPG;      // sectors a lane touches
    static_assert(D >= 1 && D <= LPG && (LA == 0 || LA == 1), "touch distance within one batch");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG, g = lane / LPG;
    const long long wg0 = ((long long)blockIdx.x * (BLOCK / 32) + warp) * NG;
    if (wg0 >= ngroups) return;  // warp-uniform
    const long long beg = (wg0 + g) * (long long)chunk;
    const int nbatch = chunk / LPG;
    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * 1024u;
    // staged triples, swizzled as in sweep_rows_kernel
    constexpr int SWZ_MASK = (LPG >= 8 ? 7 : LPG - 1);
    const uint32_t swz = (uint32_t)((LPG >= 8 ? g : (g >> 1)) & SWZ_MASK);
    const uint32_t trip0 = (wbase + (uint32_t)(g * LPG) * 16u) ^ (swz * 16u);

    const int last_pack = (kw - 1) / EPV;
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) act[v] = gl + LPG * v <= last_pack;
    const char* gat_lane = reinterpret_cast<const char*>(xgat) + (unsigned)gl * 16u;
    const char* own_lane = reinterpret_cast<const char*>(xown) + (unsigned)gl * 16u;
    const char* gat_sect = reinterpret_cast<const char*>(xgat) + (unsigned)gl * 32u;
    const char* own_sect = reinterpret_cast<const char*>(xown) + (unsigned)gl * 32u;
    asm volatile("" : "+l"(gat_lane));
    asm volatile("" : "+l"(gat_sect));

    auto load_triple = [&](int b, int& r, int& c, real& y) {
        const long long idx = beg + (long long)b * LPG + gl;
        r = ldg_noalloc(row + idx);
        c = ldg_noalloc(col + idx);
        y = ldg_noalloc(val + idx);
    };

    int tq[TOUCH == 2 ? D : 1][TPL], oq[TOUCH == 2 ? D : 1][TPL];
    int sink = 0;
#pragma unroll
    for (int s = 0; s < (TOUCH == 2 ? D : 1); ++s)
#pragma unroll
        for (int j = 0; j < TPL; ++j) tq[s][j] = oq[s][j] = 0;
    int r_touched = -1;
    auto touch = [&](int s, uint32_t trip_addr) {
        if (TOUCH == 0) return;
        int ra, ca;
        lds64(trip_addr, ra, ca);
        const char* p = gat_sect + (uint64_t)(unsigned)ca * ROWB;
#pragma unroll
        for (int j = 0; j < TPL; ++j) {
            if (NSECT % LPG != 0 && gl + j * LPG >= NSECT) continue;
            if (TOUCH == 1) prefetch_l1(p + j * (LPG * 32));
            else tq[s % D][j] = touch_b32(p + j * (LPG * 32));
        }
        if (ra != r_touched) {
            const char* q = own_sect + (uint64_t)(unsigned)ra * ROWB;
#pragma unroll
            for (int j = 0; j < TPL; ++j) {
                if (NSECT % LPG != 0 && gl + j * LPG >= NSECT) continue;
                if (TOUCH == 1) prefetch_l1(q + j * (LPG * 32));
                else oq[s % D][j] = touch_b32(q + j * (LPG * 32));
            }
            r_touched = ra;
        }
    };

    Pack<real> gq[LA + 1][VPL];
    auto load_row_at = [&](int slot, int ca) {
        const char* src = gat_lane + (uint64_t)(unsigned)ca * ROWB;
#pragma unroll
        for (int v = 0; v < VPL; ++v) gq[slot][v] = ldg_pack_l1<real, EF>(src + v * (LPG * 16));
    };
    auto load_row = [&](int slot, uint32_t trip_addr) {
        int ra, ca;
        lds64(trip_addr, ra, ca);
        load_row_at(slot, ca);
    };

    Pack<real> own[VPL], sum[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        own[v] = pack_zero<real>();
        sum[v] = pack_zero<real>();
    }
    int cur = -1;
    auto flush = [&]() {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * LD + (gl + LPG * v) * EPV, sum[v]);
    };

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
    for (int t = 0; t < D; ++t) touch(t, trip0 ^ ((uint32_t)t * 16u));
    if (LA) load_row(0, trip0);

    uint32_t tb_cur = trip0, tb_nxt = trip0 + 512u;
    for (int b = 0; b < nbatch; ++b) {
        int r2, c2;
        real y2;
        load_triple(b + 2 < nbatch ? b + 2 : nbatch - 1, r2, c2, y2);
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            // touch step t + D, load step t + LA (this batch or the next one)
            if (TOUCH == 2) {
#pragma unroll
                for (int j = 0; j < TPL; ++j) sink ^= tq[t % D][j] ^ oq[t % D][j];
            }
            if (t + D < LPG) touch(t + D, tb_cur ^ ((uint32_t)(t + D) * 16u));
            else touch(t + D, tb_nxt ^ ((uint32_t)(t + D - LPG) * 16u));
            int rr;
            real yy;
            if (LA) {
                if (t + 1 < LPG) load_row((t + 1) & 1, tb_cur ^ ((uint32_t)(t + 1) * 16u));
                else load_row((t + 1) & 1, tb_nxt);
                lds_row_y(tb_cur ^ ((uint32_t)t * 16u), rr, yy);
            } else {
                int cc;
                lds_triple(tb_cur ^ ((uint32_t)t * 16u), rr, cc, yy);
                load_row_at(0, cc);
            }
            if (rr != cur) {  // divergent between groups, no shuffles inside
                if (cur >= 0) flush();
                cur = rr;
                const char* src = own_lane + (uint64_t)(unsigned)rr * ROWB;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = ldg_pack_l1<real, EF>(src + v * (LPG * 16));
                    sum[v] = pack_zero<real>();
                }
            }
            Pack<real> gv[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) gv[v] = gq[LA ? (t & 1) : 0][v];
            typename DotOf<real>::type d0, d1;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if ((v & 1) && VPL > 2) d1.add(own[v], gv[v]);
                else d0.add(own[v], gv[v]);
            }
            real s = VPL > 2 ? d0.total() + d1.total() : d0.total();
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            const real w = rdiv_rcp(yy, s);
#pragma unroll
            for (int v = 0; v < VPL; ++v) axpy_pack(sum[v], w, gv[v]);
        }
        __syncwarp();
        sts_triple(tb_cur ^ ((uint32_t)gl * 16u), r2, c2, y2);
        __syncwarp();
        const uint32_t tmp = tb_cur;
        tb_cur = tb_nxt;
        tb_nxt = tmp;
    }
    if (cur >= 0) flush();
    if (TOUCH == 2 && sink == 0x5bd1e995) asm volatile("st.shared.b32 [%0], %1;" ::"r"(wbase), "r"(sink) : "memory");
}

}  // namespace hpf
