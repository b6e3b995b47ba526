// Sweep kernels of the HPF coordinate-ascent engine (sm_100a): sweep_rows_kernel (full-batch CAVI, one
// direction per launch) and sweep_coo_kernel (minibatches, any nnz order, both directions at once).
//
// sweep_rows_kernel: one launch is one DIRECTION of the per-nnz part of a CAVI iteration: it replaces update_phi
// (pxi:551-591) + update_G_n_L_sh (pxi:613-621) of david-cortes/hpfrec for ONE of the two shape
// matrices, without materialising phi (hpf_kernels.cuh explains the per-row factorisation):
//
//       acc[r, :] += sum_{n : row(n) = r}  (Y[n] / dot(xown[r, :], xgat[col(n), :])) * xgat[col(n), :]
//
// The triples are sorted by (L2 panel of col, row).  LPG consecutive lanes form a lane group that walks
// one contiguous chunk of the list; the row a group currently owns lives in registers together with
// its running sum, which is flushed with one vector RED per 16-byte pack when the row id changes.
//
// What this revision changes against the round-1 kernel (17 issued instructions per nnz, issue-bound):
//   * the triples of a batch go through shared memory (one STS.128 per lane per batch), so a step reads
//     them with two broadcast LDS.64 instead of three SHFL plus register queues;
//   * the arrays are padded by the host to whole chunks with zero-count entries, so the loop has no
//     "past the end" predicates at all;
//   * packs beyond the row's active width are handled by cp.async's src-size operand (0 = zero-fill,
//     no global read) instead of per-lane predicates and duplicated address chains, or -- FULLROW --
//     by copying the zero padding of a full-stride row;
//   * the dot product and the accumulation use the packed fp32 FMA of sm_100 (FFMA2: fma.rn.f32x2),
//     halving the FP32 instruction count;
//   * every global access carries an L2 policy: gathered rows evict_last (they are the panel that has
//     to stay resident), the streamed triples / own rows / REDs evict_first.
// ROBUST adds the rescue path for normalisers that underflow with tiny shape hyper-parameters
// (see sweep_rescue below).
#pragma once
#include "hpf_device.cuh"

namespace hpf {

// ---- small device helpers local to the sweep --------------------------------------------------------
__device__ __forceinline__ void cp_async16_pol(uint32_t dst, const void* src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
// src_size = 16 copies, src_size = 0 zero-fills the 16 bytes without reading global memory
__device__ __forceinline__ void cp_async16_sz_pol(uint32_t dst, const void* src, uint32_t src_size, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_size), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_sz(uint32_t dst, const void* src, uint32_t src_size) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_size) : "memory");
}
__device__ __forceinline__ void lds64(uint32_t addr, int& a, int& b) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ void red_add_pack_pol(float* p, const Pack<float>& r, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]),
                 "f"(r.v[2]), "f"(r.v[3]), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void red_add_pack_pol(double* p, const Pack<double>& r, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(r.v[0]), "l"(pol) : "memory");
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p + 1), "d"(r.v[1]), "l"(pol) : "memory");
}

// two-lane partial dot product / scaled accumulation of one pack; float uses FFMA2 (fma.rn.f32x2)
struct Dot2f {
    float2 d;
    __device__ __forceinline__ Dot2f() : d(make_float2(0.f, 0.f)) {}
    __device__ __forceinline__ void add(const Pack<float>& a, const Pack<float>& b) {
        d = __ffma2_rn(make_float2(a.v[0], a.v[1]), make_float2(b.v[0], b.v[1]), d);
        d = __ffma2_rn(make_float2(a.v[2], a.v[3]), make_float2(b.v[2], b.v[3]), d);
    }
    __device__ __forceinline__ float total() const { return d.x + d.y; }
};
struct Dot2d {
    double d0, d1;
    __device__ __forceinline__ Dot2d() : d0(0.0), d1(0.0) {}
    __device__ __forceinline__ void add(const Pack<double>& a, const Pack<double>& b) {
        d0 = fma(a.v[0], b.v[0], d0);
        d1 = fma(a.v[1], b.v[1], d1);
    }
    __device__ __forceinline__ double total() const { return d0 + d1; }
};
template <typename real> struct DotOf;
template <> struct DotOf<float> { using type = Dot2f; };
template <> struct DotOf<double> { using type = Dot2d; };

__device__ __forceinline__ void axpy_pack(Pack<float>& s, float w, const Pack<float>& g) {
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(ww, make_float2(g.v[0], g.v[1]), make_float2(s.v[0], s.v[1]));
    const float2 hi = __ffma2_rn(ww, make_float2(g.v[2], g.v[3]), make_float2(s.v[2], s.v[3]));
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = hi.x; s.v[3] = hi.y;
}
__device__ __forceinline__ void axpy_pack(Pack<double>& s, double w, const Pack<double>& g) {
    s.v[0] = fma(w, g.v[0], s.v[0]);
    s.v[1] = fma(w, g.v[1], s.v[1]);
}

// normaliser below which the per-row factorisation is no longer trusted (ROBUST instantiations only):
// every x is exp(E - rowmax) <= 1 with one entry equal to 1 per row, so a normaliser this small means
// the two rows' exponentials have (almost) disjoint support in `real`
template <typename real> __device__ __forceinline__ real rescue_threshold();
template <> __device__ __forceinline__ float rescue_threshold<float>() { return 1e-25f; }
template <> __device__ __forceinline__ double rescue_threshold<double>() { return 1e-250; }

// Materialised state the rescue path recomputes E[log] from, and where it puts its result: the
// multinomial of the reference itself, phi[n, :] = Y[n] * softmax(Eown + Egat) with the JOINT maximum
// subtracted (pxi:560-577), added to `direct` -- a sum the row update adds to the shape without
// multiplying it by the row factor x (which is exactly what underflowed).
template <typename real>
struct RescueArgs {
    const real* shp_own;
    const real* rte_own;
    const real* shp_gat;
    const real* rte_gat;
    real* direct_own;
    real* direct_gat;  // NULL in the two-pass sweep (the other pass owns that side)
    int k;
};

template <typename real, int LPG, int VPL>
__device__ __noinline__ void sweep_rescue(const RescueArgs<real>& ra, int r_own, int c_gat, real y, int ld, int gl,
                                          real* phi_row = nullptr) {
    constexpr int EPV = Pack<real>::N;
    const unsigned gmask = group_mask<LPG>();
    real l[VPL][EPV];
    real m = -INFINITY;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int off = (gl + LPG * v) * EPV;
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            const int j = off + e;
            real lj = -INFINITY;
            if (j < ra.k) {
                const size_t ao = (size_t)r_own * ld + j, ag = (size_t)c_gat * ld + j;
                lj = elog(ra.shp_own[ao], ra.rte_own[ao]) + elog(ra.shp_gat[ag], ra.rte_gat[ag]);
            }
            l[v][e] = lj;
            m = lj > m ? lj : m;
        }
    }
    m = group_max<LPG>(m, gmask);
    real s = real(0);
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            l[v][e] = rexp(l[v][e] - m);  // exp(-inf) = 0 for pad columns
            s += l[v][e];
        }
    s = group_sum<LPG>(s, gmask);
    const real w = y / s;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int off = (gl + LPG * v) * EPV;
#pragma unroll
        for (int e = 0; e < EPV; ++e)
            if (off + e < ra.k) {
                atomicAdd(ra.direct_own + (size_t)r_own * ld + off + e, w * l[v][e]);
                if (ra.direct_gat != nullptr) atomicAdd(ra.direct_gat + (size_t)c_gat * ld + off + e, w * l[v][e]);
                if (phi_row != nullptr) phi_row[off + e] = w * l[v][e];
            }
    }
}

// shared memory of one warp: own-row ring (OWN_DEPTH slots of VPL x 512 B) and two batches of staged triples
// (32 x 16 B each).  The GATHERED rows never touch shared memory: they are loaded straight into registers.
// GD = slots of the optional gathered-row ring (0: the gathered rows go straight into registers).
template <int VPL, int GD = 0>
struct SweepSmem {
    static constexpr int OWN_DEPTH = GD > 4 ? GD : 4;  // an own row may be staged as far ahead as a gathered one
    static constexpr uint32_t SLOT = VPL * 512u;
    static constexpr uint32_t RING = OWN_DEPTH * SLOT;
    static constexpr uint32_t GRING = GD * SLOT;
    static constexpr uint32_t TRIP = 2u * 512u;
    static constexpr uint32_t WARP = RING + GRING + TRIP;
};

// staged triple, 16 bytes: float {row, col, y, row}; double {row, col, y}
__device__ __forceinline__ void sts_triple(uint32_t addr, int r, int c, float y) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r), "r"(c), "r"(__float_as_int(y)), "r"(r) : "memory");
}
__device__ __forceinline__ void sts_triple(uint32_t addr, int r, int c, double y) {
    const long long yb = __double_as_longlong(y);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r), "r"(c), "r"((int)(yb & 0xffffffffll)),
                 "r"((int)(yb >> 32))
                 : "memory");
}
__device__ __forceinline__ void lds_row_y(uint32_t addr, int& r, float& y) {
    int yb;
    lds64(addr + 8u, yb, r);
    y = __int_as_float(yb);
}
__device__ __forceinline__ void lds_row_y(uint32_t addr, int& r, double& y) {
    int lo, hi;
    asm volatile("{\n\t.reg .b32 c;\n\tld.shared.v4.b32 {%0, c, %1, %2}, [%3];\n\t}" : "=r"(r), "=r"(lo), "=r"(hi) : "r"(addr));
    y = __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}

// Own-row staging, deliberately NOT inlined: inlined, ptxas if-converts the rare "row id changed" block into
// predicated LDGSTS instructions that are issued at EVERY step -- and a predicated-off LDGSTS still costs
// shared-memory wavefronts in the L1TEX data pipe (ncu source page of the inlined form: 2 LDGSTS + 3 filler LDS
// per step, 164 M wavefronts per launch, more than all real shared-memory traffic together).  A call is a
// real branch.  act_mask: bit v set = pack v of this lane is inside the row's active width (else zero-fill).
template <int LPG, int VPL, int HINT>
__device__ __noinline__ void stage_own_row(uint32_t dst, const char* src, uint32_t act_mask, uint64_t pol) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const uint32_t szv = (act_mask >> v) & 1u ? 16u : 0u;
        if (HINT) cp_async16_sz_pol(dst + (uint32_t)v * 512u, src + v * (LPG * 16), szv, pol);
        else cp_async16_sz(dst + (uint32_t)v * 512u, src + v * (LPG * 16), szv);
    }
}

// =============================================================================================
// ngroups: number of lane groups of the launch = padded nnz / chunk, a multiple of 32 / LPG.
// chunk:   nnz per lane group, a multiple of LPG.  row/col/val hold ngroups * chunk entries.
// D:       gathered rows in flight per lane group (register pipeline depth; divides LPG, <= 4).
// HINT:    0 no L2 policies; 1 streamed data (triples, own rows, REDs) evict_first + gathered rows evict_last
//          for a fraction `keep_frac` of the lines; 2 streamed data evict_first only.
//
// Why registers and not shared memory for the gathered rows: ncu on the shared-memory-staged revision of
// this kernel (profiles/r02_ncu_full_smem_staged_lpg8.csv) shows the L1TEX data pipe at 89 % with 6.2
// shared-memory wavefronts per nnz -- every gathered byte crossed shared memory twice (LDGSTS in, LDS out)
// and the 128 B/clk/SM port was the bound (1.16-1.30 ms per pass whatever the shape, L2 panel or hint).
// The bare staging pattern tops out at 60 G rows/s with LDGSTS and 46 G rows/s with TMA tile::gather4
// (tools/tma_gather_probe.cu, profiles/r02_tma_gather4_probe.jsonl), against 80 G rows/s for plain
// 128-bit loads into registers (tools/gather_probe.cu).
// =============================================================================================
// PF:      1 = every lane asks L2 for the gathered row of its own triple two batches ahead (prefetch.global.L2):
//          a gather that would miss L2 becomes a hit by the time it is issued; 2 = the same into L1.
// SG:      1 = the gathered rows are staged in a shared-memory ring of D slots with cp.async (LDGSTS: no registers
//          and no scoreboard per row in flight, but every gathered byte then crosses shared memory twice);
//          0 = 128-bit loads straight into registers.
// FUSE:    1 = one-pass form for minibatches (triples grouped by the batched side): the walk also pushes
//          w_n * xown[r, :] into the OTHER side's sums (acc_minor) with one vector RED per pack per nnz, so both
//          shape matrices come out of one launch (update_phi_csr + update_G_n_L_sh_csr, pxi:666-768).
template <typename real, int LPG, int VPL, int D, int MINB, int BLOCK, int HINT, bool FULLROW, bool ROBUST, int PF = 0, int SG = 0,
          int FUSE = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_rows_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                  long long ngroups, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                  real* __restrict__ acc, int ld, int kw, float keep_frac, RescueArgs<real> rescue,
                  real* __restrict__ acc_minor) {
    constexpr int EPV = Pack<real>::N;
    constexpr int NG = 32 / LPG;  // lane groups per warp
    using SM = SweepSmem<VPL, SG ? D : 0>;
    constexpr int OD = SM::OWN_DEPTH;
    static_assert(LPG % D == 0 && LPG % OD == 0 && (SG || D <= OD), "pipeline depth must divide the lane-group width");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG, g = lane / LPG;
    const long long wg0 = ((long long)blockIdx.x * (BLOCK / 32) + warp) * NG;
    if (wg0 >= ngroups) return;  // warp-uniform
    const long long beg = (wg0 + g) * (long long)chunk;
    const int nbatch = chunk / LPG;

    uint64_t pol_keep = 0, pol_stream = 0;
    if (HINT) pol_stream = l2_policy_stream();
    if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol_keep) : "f"(keep_frac));
    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);
    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * SM::WARP;
    const uint32_t ring_o = wbase + (uint32_t)lane * 16u;
    const uint32_t ring_g = ring_o + SM::RING;
    // staged triples: group g's batch is LPG consecutive 16-byte slots; step t sits at slot t ^ swz so that the
    // lanes of a quarter-warp store to 128 contiguous bytes and the broadcast reads of the NG groups at one step
    // hit different banks (the un-swizzled [t][g] layout cost 16 wavefronts per STS.128)
    constexpr int SWZ_MASK = (LPG >= 8 ? 7 : LPG - 1);
    const uint32_t swz = (uint32_t)((LPG >= 8 ? g : (g >> 1)) & SWZ_MASK);
    // the group's region is aligned to its size, so base + ((t ^ swz) * 16) == (base ^ swz * 16) ^ (t * 16): one LOP3
    const uint32_t trip0 = (wbase + SM::RING + SM::GRING + (uint32_t)(g * LPG) * 16u) ^ (swz * 16u);  // + buffer*512, ^ (t * 16)
    // Packs beyond the row's active width: the lane re-reads the row's LAST active pack instead (same
    // 32-byte sector as its neighbour: no extra traffic, no predicate); its own-row registers are zero, so
    // the duplicate never reaches the normaliser, and its sums are never flushed.
    const int last_pack = (kw - 1) / EPV;
    bool act[VPL];
    uint32_t act_mask = 0;
    const char* gat_lane[VPL];
    const char* own_lane = reinterpret_cast<const char*>(xown) + (unsigned)(gl * EPV) * (unsigned)sizeof(real);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int p = gl + LPG * v;
        act[v] = p <= last_pack;
        act_mask |= act[v] ? (1u << v) : 0u;
        const int pc = (FULLROW || act[v]) ? p : last_pack;
        gat_lane[v] = reinterpret_cast<const char*>(xgat) + (unsigned)pc * 16u;
        // opaque to the optimiser: keeps the lane's base as ONE 64-bit register pair, so a row address is a single
        // IMAD.WIDE (col * row_bytes + base) instead of a 32-bit offset chain plus a 64-bit add of the uniform base
        asm volatile("" : "+l"(gat_lane[v]));
    }

    auto load_triple = [&](int b, int& r, int& c, real& y) {
        const long long idx = beg + (long long)b * LPG + gl;
        if (HINT) {
            r = ldg_stream(row + idx, pol_stream);
            c = ldg_stream(col + idx, pol_stream);
            y = ldg_stream(val + idx, pol_stream);
        } else {
            r = __ldg(row + idx);
            c = __ldg(col + idx);
            y = __ldg(val + idx);
        }
    };
    Pack<real> gq[SG ? 1 : D][VPL];  // gathered rows in flight (register form)
    int r_staged = -1;
    // put one step in flight: its gathered row into registers, its own row (when the row id changes at
    // that step) into the shared-memory ring
    auto stage = [&](int s, uint32_t trip_addr) {
        int ra, ca;
        lds64(trip_addr, ra, ca);
        const uint64_t roff = (uint64_t)(unsigned)ca * row_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const char* srcb = (FULLROW ? gat_lane[0] + v * (LPG * 16) : gat_lane[v]) + roff;
            if (SG) {
                // inactive lanes re-copy the row's last active pack (see gat_lane): no predicate, no zero-fill
                const uint32_t dst = ring_g + (uint32_t)(s % D) * SM::SLOT + (uint32_t)v * 512u;
                if (HINT == 1) cp_async16_pol(dst, srcb, pol_keep);
                else cp_async16(dst, srcb);
            } else {
                const real* src = reinterpret_cast<const real*>(srcb);
                if (HINT == 1) gq[SG ? 0 : s % D][v] = ldg_pack_hint(src, pol_keep);
                else gq[SG ? 0 : s % D][v] = ldg_pack(src);
            }
        }
        if (ra != r_staged) {
            stage_own_row<LPG, VPL, HINT>(ring_o + (uint32_t)(s % OD) * SM::SLOT, own_lane + (uint64_t)(unsigned)ra * row_bytes,
                                          act_mask, pol_stream);
            r_staged = ra;
        }
        cp_async_commit();
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
            if (act[v]) {
                real* dst = acc + (size_t)cur * ld + (gl + LPG * v) * EPV;
                if (HINT) red_add_pack_pol(dst, sum[v], pol_stream);
                else red_add_pack(dst, sum[v]);
            }
    };

    // ---- prologue: batches 0 and 1 into the two triple buffers, first D steps in flight ------------
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
    for (int t = 0; t < D; ++t) stage(t, trip0 ^ ((uint32_t)t * 16u));

    uint32_t tb_cur = trip0, tb_nxt = trip0 + 512u;
    for (int b = 0; b < nbatch; ++b) {
        // triples of batch b+2 (clamped: the tail re-reads the last batch; those steps are never consumed)
        int r2, c2;
        real y2;
        load_triple(b + 2 < nbatch ? b + 2 : nbatch - 1, r2, c2, y2);
        if (PF) {
            const char* pf = reinterpret_cast<const char*>(xgat) + (uint64_t)(unsigned)c2 * row_bytes;
            if (PF == 2) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 128));
            } else {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 128));
            }
        }
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            // ---- consume step t
            int rr;
            real yy;
            lds_row_y(tb_cur ^ ((uint32_t)t * 16u), rr, yy);
            if (SG) cp_async_wait<D - 1>();  // all but the newest D-1 groups have landed: step t is in
            if (rr != cur) {  // divergent between groups, no shuffles inside
                if (cur >= 0) flush();
                cur = rr;
                if (!SG) cp_async_wait<D - 1>();  // own rows of all steps up to t have landed
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = lds_pack<real>(ring_o + (uint32_t)(t % OD) * SM::SLOT + (uint32_t)v * 512u);
                    sum[v] = pack_zero<real>();
                }
            }
            Pack<real> gv[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (SG) gv[v] = lds_pack<real>(ring_g + (uint32_t)(t % D) * SM::SLOT + (uint32_t)v * 512u);
                else gv[v] = gq[SG ? 0 : t % D][v];
            }
            typename DotOf<real>::type d0, d1;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if ((v & 1) && VPL > 2) d1.add(own[v], gv[v]);
                else d0.add(own[v], gv[v]);
            }
            real s = VPL > 2 ? d0.total() + d1.total() : d0.total();
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            real w = rdiv_rcp(yy, s);
            if (ROBUST) {
                if (!(s >= rescue_threshold<real>())) {  // group-uniform (s is the group's sum)
                    if (yy > real(0)) {
                        int ra_, ca_;
                        lds64(tb_cur ^ ((uint32_t)t * 16u), ra_, ca_);
                        sweep_rescue<real, LPG, VPL>(rescue, cur, ca_, yy, ld, gl);
                    }
                    w = real(0);
                }
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) axpy_pack(sum[v], w, gv[v]);
            if (FUSE) {
                int rf, cf;
                lds64(tb_cur ^ ((uint32_t)t * 16u), rf, cf);
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if (!act[v]) continue;
                    Pack<real> p;
#pragma unroll
                    for (int e = 0; e < EPV; ++e) p.v[e] = w * own[v].v[e];
                    red_add_pack(acc_minor + (size_t)cf * ld + (gl + LPG * v) * EPV, p);
                }
            }
            // ---- put step t + D in flight (this batch or the next one), into the registers just consumed
            if (t + D < LPG) stage(t + D, tb_cur ^ ((uint32_t)(t + D) * 16u));
            else stage(t + D, tb_nxt ^ ((uint32_t)(t + D - LPG) * 16u));
        }
        // batch b is consumed: its buffer takes batch b+2
        __syncwarp();
        sts_triple(tb_cur ^ ((uint32_t)gl * 16u), r2, c2, y2);
        __syncwarp();
        const uint32_t tmp = tb_cur;
        tb_cur = tb_nxt;
        tb_nxt = tmp;
    }
    cp_async_wait<0>();
    if (cur >= 0) flush();
}

// =============================================================================================
// K2'  single-pass COO sweep with atomics on both sides: any nnz order, used for minibatches
//      (partial_fit pxi:438-459, SVI pxi:292-314) and as the cross-check of the two-pass sweep.
//      accU[u,:] += w_n * xi[i,:]    accI[i,:] += w_n * xu[u,:]     (optionally phi[n,:] written)
//      ROBUST: normalisers below rescue_threshold take the reference's own per-nnz softmax (sweep_rescue).
// =============================================================================================
template <typename real, int LPG, int VPL, bool ROBUST>
__global__ void __launch_bounds__(256)
sweep_coo_kernel(const int* __restrict__ iu, const int* __restrict__ ii, const real* __restrict__ val,
                 long long nnz, int chunk, const real* __restrict__ xu, const real* __restrict__ xi,
                 real* __restrict__ accU, real* __restrict__ accI, int ld,
                 real* __restrict__ phi, int k, RescueArgs<real> rescue) {
    constexpr int EPV = Pack<real>::N;
    const int gl = (threadIdx.x & 31) % LPG;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const long long beg = group * (long long)chunk;
    if (beg >= nnz) return;
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const unsigned gmask = group_mask<LPG>();
    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < k;
    }
    for (long long base = beg; base < end; base += LPG) {
        const long long idx = base + gl;
        int u = 0, i = 0;
        real y = real(0);
        if (idx < end) {
            u = __ldg(iu + idx);
            i = __ldg(ii + idx);
            y = __ldg(val + idx);
        }
        const int cnt = (end - base < LPG) ? (int)(end - base) : LPG;
        for (int t = 0; t < cnt; ++t) {
            const int uu = __shfl_sync(gmask, u, t, LPG);
            const int it = __shfl_sync(gmask, i, t, LPG);
            const real yy = __shfl_sync(gmask, y, t, LPG);
            Pack<real> gu[VPL], gi[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                gu[v] = act[v] ? ldg_pack(xu + (size_t)uu * ld + off[v]) : pack_zero<real>();
                gi[v] = act[v] ? ldg_pack(xi + (size_t)it * ld + off[v]) : pack_zero<real>();
            }
            real s = real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) s = fma(gu[v].v[e], gi[v].v[e], s);
            s = group_sum<LPG>(s, gmask);
            real w = rdiv_fast(yy, s);
            bool rescued = false;
            if (ROBUST) {
                if (!(s >= rescue_threshold<real>())) {  // group-uniform
                    sweep_rescue<real, LPG, VPL>(rescue, uu, it, yy, ld, gl, phi ? phi + (size_t)(base + t) * k : nullptr);
                    w = real(0);
                    rescued = true;
                }
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (!act[v]) continue;
                Pack<real> pu, pi;
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    pu.v[e] = w * gi[v].v[e];
                    pi.v[e] = w * gu[v].v[e];
                }
                red_add_pack(accU + (size_t)uu * ld + off[v], pu);
                red_add_pack(accI + (size_t)it * ld + off[v], pi);
                if (phi != nullptr && !rescued) {
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
                        if (off[v] + e < k)
                            phi[(size_t)(base + t) * k + off[v] + e] = pu.v[e] * gu[v].v[e];
                }
            }
        }
    }
}

}  // namespace hpf
