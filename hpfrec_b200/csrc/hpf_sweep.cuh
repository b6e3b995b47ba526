// Sweep kernels of the HPF coordinate-ascent engine (sm_100a): the per-nnz part of one CAVI iteration,
// replacing update_phi (pxi:551) + update_G_n_L_sh (pxi:613) of david-cortes/hpfrec without ever
// materialising phi.  Four generations of the two-pass kernel are kept selectable (option "kernel"):
// each one's header says what the previous one was measured to be bound by.  See hpf_kernels.cuh for the
// identity that turns the per-nnz softmax into per-row factors, DESIGN.md section 5 for the numbers.
#pragma once
#include "hpf_device.cuh"

namespace hpf {

// =============================================================================================
// K2  cavi sweep, one direction ("major" side = rows owned by consecutive nnz, "minor" = gathered)
//     replaces update_phi (pxi:551) + update_G_n_L_sh (pxi:613) for ONE of the two shape matrices.
//     nnz are sorted by (L2 panel of the minor id, major id); every lane group walks a contiguous
//     chunk, keeps the major row's x and the running sum in registers, gathers the minor row with
//     128-bit loads, reduces the normaliser with shuffles inside the group, and flushes the running
//     sum with one vector RED per pack when the major id changes (so atomics happen once per
//     (row, chunk) segment, not once per nnz).
//       acc[r, :] += sum_{n in segment} (Y[n] / dot(xown[r], xgat[c_n])) * xgat[c_n, :]
// =============================================================================================
//     HINT: 0 plain loads; 1 triples evict_first + gathers evict_last; 2 triples evict_first only;
//     3 triples evict_first + gathers L1::no_allocate.
//     FUSE=1 ("one-pass" mode): the same walk also pushes w_n * xown[r,:] into the MINOR side's sums
//     with one vector RED per pack per nnz, so a single user-major pass produces both shape matrices;
//     gathers ride the L2->SM response path and the REDs the SM->L2 request path.
template <typename real, int LPG, int VPL, int UNROLL, int MINB, int HINT, int FUSE>
__global__ void __launch_bounds__(256, MINB)
sweep_major_kernel(const int* __restrict__ row, const int* __restrict__ col,
                   const real* __restrict__ val, long long nnz, int chunk,
                   const real* __restrict__ xown, const real* __restrict__ xgat,
                   real* __restrict__ acc, real* __restrict__ acc_minor, int ld, int kw) {
    constexpr int EPV = Pack<real>::N;
    static_assert(LPG % UNROLL == 0, "UNROLL must divide LPG");
    const int gl = (threadIdx.x & 31) % LPG;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const long long beg = group * (long long)chunk;
    if (beg >= nnz) return;  // whole groups leave together; shuffles below use the group mask
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const unsigned gmask = group_mask<LPG>();
    uint64_t pol_keep = 0, pol_stream = 0;
    if (HINT) {
        pol_keep = l2_policy_keep();
        pol_stream = l2_policy_stream();
    }

    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        off[v] = (gl + LPG * v) * EPV;
        act[v] = off[v] < kw;  // packs holding at least one real column (stride ld may be wider)
    }
    Pack<real> own[VPL], sum[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        own[v] = pack_zero<real>();
        sum[v] = pack_zero<real>();
    }
    int cur = -1;

    // coalesced fetch of LPG triples (one per lane), software-pipelined one batch ahead
    int r = -1, c = 0;
    real y = real(0);
    if (beg + gl < end) {
        if (HINT) {
            r = ldg_stream(row + beg + gl, pol_stream);
            c = ldg_stream(col + beg + gl, pol_stream);
            y = ldg_stream(val + beg + gl, pol_stream);
        } else {
            r = __ldg(row + beg + gl);
            c = __ldg(col + beg + gl);
            y = __ldg(val + beg + gl);
        }
    }
    for (long long base = beg; base < end; base += LPG) {
        int rn = -1, cn = 0;
        real yn = real(0);
        const long long nidx = base + LPG + gl;
        if (nidx < end) {
            if (HINT) {
                rn = ldg_stream(row + nidx, pol_stream);
                cn = ldg_stream(col + nidx, pol_stream);
                yn = ldg_stream(val + nidx, pol_stream);
            } else {
                rn = __ldg(row + nidx);
                cn = __ldg(col + nidx);
                yn = __ldg(val + nidx);
            }
        }
#pragma unroll
        for (int t0 = 0; t0 < LPG; t0 += UNROLL) {
            if (base + t0 >= end) break;  // uniform inside the group
            Pack<real> g[UNROLL][VPL];
            int rr[UNROLL], ccs[UNROLL];
            real yy[UNROLL];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                const int cc = __shfl_sync(gmask, c, t0 + q, LPG);
                ccs[q] = cc;
                rr[q] = __shfl_sync(gmask, r, t0 + q, LPG);
                yy[q] = __shfl_sync(gmask, y, t0 + q, LPG);
                const real* src = xgat + (size_t)cc * ld;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if (HINT == 1)
                        g[q][v] = act[v] ? ldg_pack_hint(src + off[v], pol_keep) : pack_zero<real>();
                    else if (HINT == 3)
                        g[q][v] = act[v] ? ldg_pack_noalloc(src + off[v]) : pack_zero<real>();
                    else
                        g[q][v] = act[v] ? ldg_pack(src + off[v]) : pack_zero<real>();
                }
            }
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                if (rr[q] < 0) continue;  // past the end of the chunk (uniform inside the group)
                if (rr[q] != cur) {
                    if (cur >= 0) {
#pragma unroll
                        for (int v = 0; v < VPL; ++v)
                            if (act[v]) red_add_pack(acc + (size_t)cur * ld + off[v], sum[v]);
                    }
                    cur = rr[q];
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
                    for (int e = 0; e < EPV; ++e) s = fma(own[v].v[e], g[q][v].v[e], s);
                s = group_sum<LPG>(s, gmask);
                const real w = rdiv_fast(yy[q], s);
#pragma unroll
                for (int v = 0; v < VPL; ++v)
#pragma unroll
                    for (int e = 0; e < EPV; ++e) sum[v].v[e] = fma(w, g[q][v].v[e], sum[v].v[e]);
                if (FUSE) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (!act[v]) continue;
                        Pack<real> p;
#pragma unroll
                        for (int e = 0; e < EPV; ++e) p.v[e] = w * own[v].v[e];
                        red_add_pack(acc_minor + (size_t)ccs[q] * ld + off[v], p);
                    }
                }
            }
        }
        r = rn;
        c = cn;
        y = yn;
    }
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * ld + off[v], sum[v]);
    }
}

// =============================================================================================
// K2 (pipelined form) -- same contract as sweep_major_kernel without FUSE:
//       acc[r, :] += sum_{n in segment} (Y[n] / dot(xown[r], xgat[c_n])) * xgat[c_n, :]
// The probe of the bare access pattern (tools/gather_probe.cu, profiles/r01b_gather_probe_*.jsonl) moves
// 48M random 208-byte rows out of a 48 MB L2 window in 0.58-0.60 ms, while the kernel above needs
// 1.5 ms for the same gathers: every step serialises  shuffle -> gather -> dot -> butterfly -> divide ->
// FMA, so a warp waits one full L2 round trip PLUS the dependent arithmetic per nnz, and lane-group
// masked shuffles cost a MATCH/REDUX/VOTE/branch sequence each.  This form
//   * keeps control flow uniform across the WARP (every group runs the same number of steps; steps past
//     the end of a group's chunk are predicated off), so all shuffles use the full mask and compile to
//     bare SHFL;
//   * software-pipelines one step ahead: while step t is reduced and accumulated, the gathered row of
//     step t+1 -- and, when the major id changes at t+1, the group's own row -- are already in flight.
// =============================================================================================
template <typename real, int LPG, int VPL, int MINB, int HINT>
__global__ void __launch_bounds__(256, MINB)
sweep_major_v2_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                      long long nnz, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                      real* __restrict__ acc, int ld, int kw) {
    constexpr int EPV = Pack<real>::N;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPG;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // warp-uniform exit: the first group of this warp already starts past the end
    if (((tid - lane) / LPG) * (long long)chunk >= nnz) return;
    const long long group = tid / LPG;
    long long beg = group * (long long)chunk;
    if (beg > nnz) beg = nnz;  // later groups of the last warp stay alive with an empty range
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const int nbatch = (chunk + LPG - 1) / LPG;  // identical for every group of the grid

    uint64_t pol_keep = 0, pol_stream = 0;
    if (HINT) {
        pol_keep = l2_policy_keep();
        pol_stream = l2_policy_stream();
    }
    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);
    const char* gat_base = reinterpret_cast<const char*>(xgat);
    const char* own_base = reinterpret_cast<const char*>(xown);
    unsigned offb[VPL];  // byte offset of this lane's packs inside a row
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        offb[v] = (unsigned)((gl + LPG * v) * EPV) * (unsigned)sizeof(real);
        act[v] = (gl + LPG * v) * EPV < kw;
    }

    auto load_triple = [&](long long idx, int& r, int& c, real& y) {
        r = -1;
        c = 0;
        y = real(0);
        if (idx < end) {
            if (HINT) {
                r = ldg_stream(row + idx, pol_stream);
                c = ldg_stream(col + idx, pol_stream);
                y = ldg_stream(val + idx, pol_stream);
            } else {
                r = __ldg(row + idx);
                c = __ldg(col + idx);
                y = __ldg(val + idx);
            }
        }
    };
    auto gather = [&](int cc, bool valid, Pack<real>(&g)[VPL]) {
        const char* src = gat_base + (uint64_t)(unsigned)cc * row_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (valid && act[v]) {
                if (HINT == 1)
                    g[v] = ldg_pack_hint(reinterpret_cast<const real*>(src + offb[v]), pol_keep);
                else
                    g[v] = ldg_pack(reinterpret_cast<const real*>(src + offb[v]));
            } else {
                g[v] = pack_zero<real>();
            }
        }
    };

    Pack<real> own[VPL], own_nx[VPL], sum[VPL], g_nx[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        own[v] = pack_zero<real>();
        own_nx[v] = pack_zero<real>();
        sum[v] = pack_zero<real>();
    }
    int cur = -1;

    // batch 0 and the pipeline prologue (step 0 of batch 0)
    int r, c;
    real y;
    load_triple(beg + gl, r, c, y);
    int r_st = __shfl_sync(FULL, r, 0, LPG);
    real y_st = __shfl_sync(FULL, y, 0, LPG);
    {
        const int c0 = __shfl_sync(FULL, c, 0, LPG);
        gather(c0, r_st >= 0, g_nx);
    }
    bool chg_st = r_st >= 0;  // cur == -1: the first valid nnz always opens a row
    if (chg_st) {
        const char* src = own_base + (uint64_t)(unsigned)r_st * row_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) own_nx[v] = ldg_pack(reinterpret_cast<const real*>(src + offb[v]));
    }

    for (int b = 0; b < nbatch; ++b) {
        int rn, cn;
        real yn;
        load_triple((b + 1 < nbatch) ? beg + (long long)(b + 1) * LPG + gl : end, rn, cn, yn);
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            // ---- this step's operands were fetched one step ago
            Pack<real> g[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) g[v] = g_nx[v];
            const int rr = r_st;
            const real yy = y_st;
            const bool valid = rr >= 0;
            if (chg_st) {  // divergent between groups, no shuffles inside
                if (cur >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
                        if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
                }
                cur = rr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = own_nx[v];
                    sum[v] = pack_zero<real>();
                }
            }
            // ---- put the next step in flight (warp-uniform shuffles)
            int c2, r2;
            real y2;
            if (t + 1 < LPG) {
                c2 = __shfl_sync(FULL, c, t + 1, LPG);
                r2 = __shfl_sync(FULL, r, t + 1, LPG);
                y2 = __shfl_sync(FULL, y, t + 1, LPG);
            } else {
                c2 = __shfl_sync(FULL, cn, 0, LPG);
                r2 = __shfl_sync(FULL, rn, 0, LPG);
                y2 = __shfl_sync(FULL, yn, 0, LPG);
            }
            gather(c2, r2 >= 0, g_nx);
            chg_st = r2 >= 0 && r2 != cur;
            if (chg_st) {
                const char* src = own_base + (uint64_t)(unsigned)r2 * row_bytes;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (act[v]) own_nx[v] = ldg_pack(reinterpret_cast<const real*>(src + offb[v]));
            }
            r_st = r2;
            y_st = y2;
            // ---- reduce and accumulate the current step (pad / invalid lanes carry zeros)
            real s0 = real(0), s1 = real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                s0 = fma(own[v].v[0], g[v].v[0], s0);
                s1 = fma(own[v].v[1], g[v].v[1], s1);
                if (EPV == 4) {
                    s0 = fma(own[v].v[EPV - 2], g[v].v[EPV - 2], s0);
                    s1 = fma(own[v].v[EPV - 1], g[v].v[EPV - 1], s1);
                }
            }
            real s = s0 + s1;
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            const real w = valid ? rdiv_fast(yy, s) : real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) sum[v].v[e] = fma(w, g[v].v[e], sum[v].v[e]);
        }
        r = rn;
        c = cn;
        y = yn;
    }
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
    }
}

// =============================================================================================
// K2 (deep-pipeline form) -- same contract again:
//       acc[r, :] += sum_{n in segment} (Y[n] / dot(xown[r], xgat[c_n])) * xgat[c_n, :]
// Measured (profiles/r01b_tune_v2.jsonl): the one-step register pipeline above still needs 1.45 ms per
// pass where the bare gathers take 0.6 ms.  A warp advances one step per memory round trip, the round
// trip is the MAXIMUM over its lane groups' loads (gathers that miss L2, own rows and triples streamed
// from DRAM), and registers cap how many rows a warp can keep in flight.  Here the rows land in SHARED
// memory instead: every lane copies its own 16-byte packs with cp.async (LDGSTS, per-thread, no
// per-copy descriptor like the bulk/TMA path of hpf_sweep_tma.cuh) DEPTH-1 steps ahead of their use and
// reads back exactly the packs it copied, so no barrier or cross-lane hand-off is needed.  The own row
// of an upcoming major-id change is staged the same way in a second ring; triples are fetched two
// batches ahead.  Control flow is warp-uniform (full-mask shuffles).
//   shared memory per warp: 2 rings x DEPTH slots x VPL x 512 B.
// =============================================================================================
template <typename real, int LPG, int VPL, int MINB, int HINT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_major_v3_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                      long long nnz, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                      real* __restrict__ acc, int ld, int kw) {
    constexpr int EPV = Pack<real>::N;
    constexpr int DEPTH = 4;  // ring slots; LPG is a multiple of 4, so the slot of step t is t % 4 at compile time
    constexpr int LOOK = DEPTH - 1;
    static_assert(LPG % DEPTH == 0, "lane-group width must be a multiple of the ring depth");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t SLOT_BYTES = VPL * 32 * 16;            // one step of one warp: [v][lane] packs
    constexpr uint32_t WARP_BYTES = 2 * DEPTH * SLOT_BYTES;   // gather ring, then own-row ring
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((tid - lane) / LPG) * (long long)chunk >= nnz) return;  // warp-uniform
    const long long group = tid / LPG;
    long long beg = group * (long long)chunk;
    if (beg > nnz) beg = nnz;
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const int nbatch = (chunk + LPG - 1) / LPG;

    uint64_t pol_stream = 0;
    if (HINT) pol_stream = l2_policy_stream();
    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);
    const uint32_t ring_g = smem_u32(smem_raw) + (uint32_t)warp * WARP_BYTES + (uint32_t)lane * 16u;
    const uint32_t ring_o = ring_g + DEPTH * SLOT_BYTES;
    // every lane reads back exactly the cells it copies; cells of packs beyond the row's active width
    // are never copied, so zeroing them once makes every later read a plain LDS (no per-step predication)
#pragma unroll
    for (int q = 0; q < 2 * DEPTH * VPL; ++q) sts_pack<real>(ring_g + (uint32_t)q * 512u, pack_zero<real>());
    const unsigned off0 = (unsigned)(gl * EPV) * (unsigned)sizeof(real);  // byte offset of this lane's first pack
    const char* gat_lane = reinterpret_cast<const char*>(xgat) + off0;
    const char* own_lane = reinterpret_cast<const char*>(xown) + off0;
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) act[v] = (gl + LPG * v) * EPV < kw;

    auto load_triple = [&](long long idx, int& r, int& c, real& y) {
        r = -1;
        c = 0;
        y = real(0);
        if (idx < end) {
            if (HINT) {
                r = ldg_stream(row + idx, pol_stream);
                c = ldg_stream(col + idx, pol_stream);
                y = ldg_stream(val + idx, pol_stream);
            } else {
                r = __ldg(row + idx);
                c = __ldg(col + idx);
                y = __ldg(val + idx);
            }
        }
    };
    // stage one step: the gathered row always, the own row when the major id changes at that step
    auto stage = [&](int slot, int ra, int ca, int r_before) {
        if (ra >= 0) {
            const char* src = gat_lane + (uint64_t)(unsigned)ca * row_bytes;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v])
                    cp_async16(ring_g + (uint32_t)slot * SLOT_BYTES + (uint32_t)v * 512u, src + v * (LPG * 16));
            if (ra != r_before) {
                const char* so = own_lane + (uint64_t)(unsigned)ra * row_bytes;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (act[v])
                        cp_async16(ring_o + (uint32_t)slot * SLOT_BYTES + (uint32_t)v * 512u, so + v * (LPG * 16));
            }
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

    // triples: batch b in (r0,c0,y0), b+1 in (r1,c1,y1), b+2 loaded at the top of batch b
    int r0, c0, r1, c1, r2 = -1, c2 = 0;
    real y0, y1, y2 = real(0);
    load_triple(beg + gl, r0, c0, y0);
    load_triple((1 < nbatch) ? beg + LPG + gl : end, r1, c1, y1);
    // prologue: stage steps 0 .. DEPTH-2 (inside batch 0 since DEPTH-2 < LPG); rq[] = major ids of the
    // staged-but-not-consumed steps, oldest first
    int r_staged = -1;  // major id of the most recently staged valid step
    int rq[LOOK];
#pragma unroll
    for (int t = 0; t < LOOK; ++t) {
        const int ra = __shfl_sync(FULL, r0, t, LPG);
        const int ca = __shfl_sync(FULL, c0, t, LPG);
        stage(t, ra, ca, r_staged);
        if (ra >= 0) r_staged = ra;
        rq[t] = ra;
    }

    for (int b = 0; b < nbatch; ++b) {
        load_triple((b + 2 < nbatch) ? beg + (long long)(b + 2) * LPG + gl : end, r2, c2, y2);
#pragma unroll
        for (int t = 0; t < LPG; ++t) {
            // ---- stage step t + LOOK (this batch or the next one)
            int ra, ca;
            if (t + LOOK < LPG) {
                ra = __shfl_sync(FULL, r0, t + LOOK, LPG);
                ca = __shfl_sync(FULL, c0, t + LOOK, LPG);
            } else {
                ra = __shfl_sync(FULL, r1, t + LOOK - LPG, LPG);
                ca = __shfl_sync(FULL, c1, t + LOOK - LPG, LPG);
            }
            stage((t + LOOK) % DEPTH, ra, ca, r_staged);
            if (ra >= 0) r_staged = ra;
            cp_async_wait<LOOK>();  // everything but the newest LOOK groups has landed: step t is in
            // ---- consume step t
            const int rr = rq[0];
#pragma unroll
            for (int q = 0; q + 1 < LOOK; ++q) rq[q] = rq[q + 1];
            rq[LOOK - 1] = ra;
            const real yy = __shfl_sync(FULL, y0, t, LPG);
            const bool valid = rr >= 0;
            const uint32_t slot_off = (uint32_t)(t % DEPTH) * SLOT_BYTES;
            if (valid && rr != cur) {  // divergent between groups, no shuffles inside
                if (cur >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
                        if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
                }
                cur = rr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = lds_pack<real>(ring_o + slot_off + (uint32_t)v * 512u);
                    sum[v] = pack_zero<real>();
                }
            }
            Pack<real> g[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) g[v] = lds_pack<real>(ring_g + slot_off + (uint32_t)v * 512u);
            real s0 = real(0), s1 = real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                s0 = fma(own[v].v[0], g[v].v[0], s0);
                s1 = fma(own[v].v[1], g[v].v[1], s1);
                if (EPV == 4) {
                    s0 = fma(own[v].v[EPV - 2], g[v].v[EPV - 2], s0);
                    s1 = fma(own[v].v[EPV - 1], g[v].v[EPV - 1], s1);
                }
            }
            real s = s0 + s1;
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            // steps past the end of the chunk read a stale (finite) slot: their weight is forced to zero
            const real w = valid ? rdiv_rcp(yy, s) : real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) sum[v].v[e] = fma(w, g[v].v[e], sum[v].v[e]);
        }
        r0 = r1; c0 = c1; y0 = y1;
        r1 = r2; c1 = c2; y1 = y2;
    }
    cp_async_wait<0>();
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
    }
}

// =============================================================================================
// K2 (deep pipeline, shuffle-free triples).  ncu on the form above (profiles/r01b_ncu_full_v3_pipeline.csv):
// memory latency is hidden (long-scoreboard stalls 0.2 per issue) and the kernel is bound by the LSU pipe
// (59 %) / issue (66 %); of the ~10 LSU instructions per step, six are shuffles -- three of them only
// broadcast the step's (major id, minor id, count) from the lane that loaded it.  Here every lane
// reads the triples of FOUR consecutive steps itself with one 128-bit load per array (all lanes of a
// group read the same address, one sector), so the only shuffles left are the butterfly of the
// normaliser.  Batches are 4 steps (= the ring depth) for every lane-group width; needs chunk % 4 == 0
// and triple arrays padded by 4 entries (the host falls back to the form above otherwise).
// =============================================================================================
template <typename real, int LPG, int VPL, int MINB, int HINT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_major_v4_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
                      long long nnz, int chunk, const real* __restrict__ xown, const real* __restrict__ xgat,
                      real* __restrict__ acc, int ld, int kw) {
    constexpr int EPV = Pack<real>::N;
    constexpr int DEPTH = 4, LOOK = DEPTH - 1, B4 = 4;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t SLOT_BYTES = VPL * 32 * 16;
    constexpr uint32_t WARP_BYTES = 2 * DEPTH * SLOT_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPG;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((tid - lane) / LPG) * (long long)chunk >= nnz) return;  // warp-uniform
    const long long group = tid / LPG;
    long long beg = group * (long long)chunk;
    if (beg > nnz) beg = nnz;
    const long long end = (beg + chunk < nnz) ? beg + chunk : nnz;
    const int nbatch = (chunk + B4 - 1) / B4;

    const unsigned row_bytes = (unsigned)ld * (unsigned)sizeof(real);
    const uint32_t ring_g = smem_u32(smem_raw) + (uint32_t)warp * WARP_BYTES + (uint32_t)lane * 16u;
    const uint32_t ring_o = ring_g + DEPTH * SLOT_BYTES;
#pragma unroll
    for (int q = 0; q < 2 * DEPTH * VPL; ++q) sts_pack<real>(ring_g + (uint32_t)q * 512u, pack_zero<real>());
    const unsigned off0 = (unsigned)(gl * EPV) * (unsigned)sizeof(real);
    const char* gat_lane = reinterpret_cast<const char*>(xgat) + off0;
    const char* own_lane = reinterpret_cast<const char*>(xown) + off0;
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) act[v] = (gl + LPG * v) * EPV < kw;

    // one batch = 4 consecutive triples, identical in every lane of the group; entries past the end of
    // the chunk get major id -1
    auto load_batch = [&](long long idx, int (&r)[4], int (&c)[4], real (&y)[4]) {
        if (idx < end) {
            ldg4(row + idx, r);
            ldg4(col + idx, c);
            ldg4(val + idx, y);
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (idx + j >= end) r[j] = -1;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                r[j] = -1;
                c[j] = 0;
                y[j] = real(0);
            }
        }
    };
    auto stage = [&](int slot, int ra, int ca, int r_before) {
        if (ra >= 0) {
            const char* src = gat_lane + (uint64_t)(unsigned)ca * row_bytes;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (act[v])
                    cp_async16(ring_g + (uint32_t)slot * SLOT_BYTES + (uint32_t)v * 512u, src + v * (LPG * 16));
            if (ra != r_before) {
                const char* so = own_lane + (uint64_t)(unsigned)ra * row_bytes;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (act[v])
                        cp_async16(ring_o + (uint32_t)slot * SLOT_BYTES + (uint32_t)v * 512u, so + v * (LPG * 16));
            }
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

    int rA[4], cA[4], rB[4], cB[4], rC[4], cC[4];
    real yA[4], yB[4], yC[4];
    load_batch(beg, rA, cA, yA);
    load_batch(beg + B4, rB, cB, yB);
    int r_staged = -1;
    int rq[LOOK];
#pragma unroll
    for (int t = 0; t < LOOK; ++t) {
        stage(t, rA[t], cA[t], r_staged);
        if (rA[t] >= 0) r_staged = rA[t];
        rq[t] = rA[t];
    }

    for (int b = 0; b < nbatch; ++b) {
        load_batch(beg + (long long)(b + 2) * B4, rC, cC, yC);
#pragma unroll
        for (int t = 0; t < B4; ++t) {
            const int ra = (t + LOOK < B4) ? rA[(t + LOOK) % B4] : rB[(t + LOOK) % B4];
            const int ca = (t + LOOK < B4) ? cA[(t + LOOK) % B4] : cB[(t + LOOK) % B4];
            stage((t + LOOK) % DEPTH, ra, ca, r_staged);
            if (ra >= 0) r_staged = ra;
            cp_async_wait<LOOK>();
            const int rr = rq[0];
#pragma unroll
            for (int q = 0; q + 1 < LOOK; ++q) rq[q] = rq[q + 1];
            rq[LOOK - 1] = ra;
            const real yy = yA[t];
            const bool valid = rr >= 0;
            const uint32_t slot_off = (uint32_t)(t % DEPTH) * SLOT_BYTES;
            if (valid && rr != cur) {
                if (cur >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v)
                        if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
                }
                cur = rr;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    own[v] = lds_pack<real>(ring_o + slot_off + (uint32_t)v * 512u);
                    sum[v] = pack_zero<real>();
                }
            }
            Pack<real> g[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) g[v] = lds_pack<real>(ring_g + slot_off + (uint32_t)v * 512u);
            real s0 = real(0), s1 = real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                s0 = fma(own[v].v[0], g[v].v[0], s0);
                s1 = fma(own[v].v[1], g[v].v[1], s1);
                if (EPV == 4) {
                    s0 = fma(own[v].v[EPV - 2], g[v].v[EPV - 2], s0);
                    s1 = fma(own[v].v[EPV - 1], g[v].v[EPV - 1], s1);
                }
            }
            real s = s0 + s1;
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o, LPG);
            const real w = valid ? rdiv_rcp(yy, s) : real(0);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
#pragma unroll
                for (int e = 0; e < EPV; ++e) sum[v].v[e] = fma(w, g[v].v[e], sum[v].v[e]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rA[j] = rB[j]; cA[j] = cB[j]; yA[j] = yB[j];
            rB[j] = rC[j]; cB[j] = cC[j]; yB[j] = yC[j];
        }
    }
    cp_async_wait<0>();
    if (cur >= 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (act[v]) red_add_pack(acc + (size_t)cur * ld + (gl + LPG * v) * EPV, sum[v]);
    }
}

// =============================================================================================
// K2'  single-pass COO sweep with atomics on both sides: any nnz order, used for minibatches
//      (partial_fit pxi:438-459, SVI pxi:292-314) and as the cross-check of the two-pass sweep.
//      accU[u,:] += w_n * xi[i,:]    accI[i,:] += w_n * xu[u,:]     (optionally phi[n,:] written)
// =============================================================================================
template <typename real, int LPG, int VPL>
__global__ void __launch_bounds__(256)
sweep_coo_kernel(const int* __restrict__ iu, const int* __restrict__ ii, const real* __restrict__ val,
                 long long nnz, int chunk, const real* __restrict__ xu, const real* __restrict__ xi,
                 real* __restrict__ accU, real* __restrict__ accI, int ld,
                 real* __restrict__ phi, int k) {
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
            const real w = rdiv_fast(yy, s);
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
                if (phi != nullptr) {
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
