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

// shared memory of one warp: gather ring, own-row ring (DEPTH slots of VPL x 512 B each), two batches of
// staged triples (32 x 16 B each)
template <int VPL>
struct SweepSmem {
    static constexpr int DEPTH = 4;
    static constexpr uint32_t SLOT = VPL * 512u;
    static constexpr uint32_t RING = DEPTH * SLOT;
    static constexpr uint32_t TRIP = 2u * 512u;
    static constexpr uint32_t WARP = 2u * RING + TRIP;
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
    int c, lo, hi;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r), "=r"(c), "=r"(lo), "=r"(hi) : "r"(addr));
    y = __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}

// =============================================================================================
// ngroups: number of lane groups of the launch = padded nnz / chunk, a multiple of 32 / LPG.
// chunk:   nnz per lane group, a multiple of LPG.  row/col/val hold ngroups * chunk entries.
// =============================================================================================
template <typename real, int LPG, int VPL, int MINB, int BLOCK, int HINT, bool FULLROW, bool ROBUST>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_rows_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                  long long ngroups, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                  real* __restrict__ acc, int ld, int kw, RescueArgs<real> rescue) {
    constexpr int EPV = Pack<real>::N;
    constexpr int NG = 32 / LPG;  // lane groups per warp
    using SM = SweepSmem<VPL>;
    constexpr int DEPTH = SM::DEPTH, LOOK = DEPTH - 1;
    static_assert(LPG % DEPTH == 0, "lane-group width must be a multiple of the ring depth");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG, g = lane / LPG;
    const long long wg0 = ((long long)blockIdx.x * (BLOCK / 32) + warp) * NG;
    if (wg0 >= ngroups) return;  // warp-uniform
    const long long beg = (wg0 + g) * (long long)chunk;
    const int nbatch = chunk / LPG;

    uint64_t pol_keep = 0, pol_stream = 0;
    if (HINT) {
        pol_keep = l2_policy_keep();
        pol_stream = l2_policy_stream();
    }
    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);
    const uint32_t wbase = smem_u32(smem_raw) + (uint32_t)warp * SM::WARP;
    const uint32_t ring_g = wbase + (uint32_t)lane * 16u;
    const uint32_t ring_o = ring_g + SM::RING;
    const uint32_t trip0 = wbase + 2u * SM::RING + (uint32_t)g * 16u;  // + buffer*512 + t*NG*16
    const unsigned off0 = (unsigned)(gl * EPV) * (unsigned)sizeof(real);
    const char* gat_lane = reinterpret_cast<const char*>(xgat) + off0;
    const char* own_lane = reinterpret_cast<const char*>(xown) + off0;
    // src-size of each pack's copy: 16, or 0 (zero-fill) beyond the active width
    uint32_t sz[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        act[v] = (gl + LPG * v) * EPV < kw;
        sz[v] = (FULLROW || act[v]) ? 16u : 0u;
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
    auto copy_row = [&](uint32_t dst, const char* src, uint64_t pol) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (FULLROW) {
                if (HINT) cp_async16_pol(dst + (uint32_t)v * 512u, src + v * (LPG * 16), pol);
                else cp_async16(dst + (uint32_t)v * 512u, src + v * (LPG * 16));
            } else {
                if (HINT) cp_async16_sz_pol(dst + (uint32_t)v * 512u, src + v * (LPG * 16), sz[v], pol);
                else cp_async16_sz(dst + (uint32_t)v * 512u, src + v * (LPG * 16), sz[v]);
            }
        }
    };
    int r_staged = -1;
    // stage one step: the gathered row always, the own row when the row id changes at that step
    auto stage = [&](int slot, uint32_t trip_addr) {
        int ra, ca;
        lds64(trip_addr, ra, ca);
        copy_row(ring_g + (uint32_t)slot * SM::SLOT, gat_lane + (uint64_t)(unsigned)ca * row_bytes, pol_keep);
        if (ra != r_staged) {
            copy_row(ring_o + (uint32_t)slot * SM::SLOT, own_lane + (uint64_t)(unsigned)ra * row_bytes, pol_stream);
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

    // ---- prologue: batches 0 and 1 into the two triple buffers, first LOOK steps staged ------------
    {
        int r, c;
        real y;
        load_triple(0, r, c, y);
        sts_triple(trip0 + (uint32_t)gl * (NG * 16u), r, c, y);
        load_triple(nbatch > 1 ? 1 : 0, r, c, y);
        sts_triple(trip0 + 512u + (uint32_t)gl * (NG * 16u), r, c, y);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < LOOK; ++t) stage(t, trip0 + (uint32_t)t * (NG * 16u));

    uint32_t tb_cur = trip0, tb_nxt = trip0 + 512u;
    for (int b = 0; b < nbatch; ++b) {
        // triples of batch b+2 (clamped: the tail re-reads the last batch, whose staging is never consumed)
        int r2, c2;
        real y2;
        load_triple(b + 2 < nbatch ? b + 2 : nbatch - 1, r2, c2, y2);
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            // ---- stage step t + LOOK (this batch or the next one)
            if (t + LOOK < LPG) stage((t + LOOK) % DEPTH, tb_cur + (uint32_t)(t + LOOK) * (NG * 16u));
            else stage((t + LOOK) % DEPTH, tb_nxt + (uint32_t)(t + LOOK - LPG) * (NG * 16u));
            cp_async_wait<LOOK>();  // all but the newest LOOK groups have landed: step t is in
            // ---- consume step t
            int rr;
            real yy;
            lds_row_y(tb_cur + (uint32_t)t * (NG * 16u), rr, yy);
            const uint32_t slot_off = (uint32_t)(t % DEPTH) * SM::SLOT;
            if (rr != cur) {  // divergent between groups, no shuffles inside
                if (cur >= 0) flush();
                cur = rr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = lds_pack<real>(ring_o + slot_off + (uint32_t)v * 512u);
                    sum[v] = pack_zero<real>();
                }
            }
            Pack<real> gv[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) gv[v] = lds_pack<real>(ring_g + slot_off + (uint32_t)v * 512u);
            typename DotOf<real>::type d0, d1;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (v & 1) d1.add(own[v], gv[v]);
                else d0.add(own[v], gv[v]);
            }
            real s = VPL > 1 ? d0.total() + d1.total() : d0.total();
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            real w = rdiv_rcp(yy, s);
            if (ROBUST) {
                if (!(s >= rescue_threshold<real>())) {  // group-uniform (s is the group's sum)
                    if (yy > real(0)) {
                        int ra_, ca_;
                        lds64(tb_cur + (uint32_t)t * (NG * 16u), ra_, ca_);
                        sweep_rescue<real, LPG, VPL>(rescue, cur, ca_, yy, ld, gl);
                    }
                    w = real(0);
                }
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) axpy_pack(sum[v], w, gv[v]);
        }
        // batch b is consumed: its buffer takes batch b+2
        __syncwarp();
        sts_triple(tb_cur + (uint32_t)gl * (NG * 16u), r2, c2, y2);
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
