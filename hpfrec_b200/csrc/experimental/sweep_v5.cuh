// EXPERIMENTAL (not compiled into libhpf_b200.so, not reachable from the C ABI): the next revision of the
// deep-pipeline sweep kernel, kept here so that its SASS can be inspected offline
// (tools/sass_count.py) before it is wired into hpf_sweep_dispatch.inl and measured.
// Baseline = sweep_major_v3_kernel (hpf_sweep.cuh): ~60 issued instructions per 4-nnz step on the
// common path, issue-bound (66 %) / LSU-bound (59 %).
//
// Offline accounting so far (python tools/sass_count.py; static SASS per pipelined step, divergent
// own-row-staging and major-id-change blocks included):
//     v3  <float, 8 lanes, 2 packs>            105 / step of 4 nnz   (common path ~60 -> ~15 per nnz)
//     v5  <float, 8, 2, FULLROW>                89 / step of 4 nnz   (common path ~54 -> ~13.5 per nnz), 66 regs
//     v3  <float, 4 lanes, 4 packs, CTA 128>   158 / step of 8 nnz
//     v5  <float, 4, 4, CTA 128, FULLROW>      120 / step of 8 nnz   (common path ~80 -> ~10 per nnz), 88 regs
// FULLROW = copy and multiply all 16 packs of the zero-padded 256-byte row: the per-lane "pack is
// active" predicates disappear (compares, duplicated 64-bit address chains, constant-bank reloads) at
// the price of one more L2 sector per gathered row (8 instead of 7; L2 is at 50 %).  NOT MEASURED YET.
// To try it: include this file from hpf_sweep.cuh, add a dispatch entry ("kernel"=5) next to the v3
// one in hpf_sweep_dispatch.inl, add 5 to the kernel lists of tests/test_gpu_parity.py and to
// tools/tune_v2.py, run tools/gpu_session.sh.
#pragma once
#include "../hpf_device.cuh"

namespace hpf {

template <typename real, int LPG, int VPL, int MINB, int HINT, int BLOCK, bool FULLROW>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_major_v5_kernel(const int* __restrict__ row, const int* __restrict__ col, const real* __restrict__ val,
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
    // FULLROW: every pack of the (zero-padded) row stride is copied and multiplied.  Instructions are issued
    // per warp, so predicating the pad lanes off saves only L2 sectors (7 instead of 8 per 256-byte row),
    // while the per-lane predicates cost compares, duplicated address chains and predicate spills.
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) act[v] = FULLROW ? true : ((gl + LPG * v) * EPV < kw);

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

}  // namespace hpf
