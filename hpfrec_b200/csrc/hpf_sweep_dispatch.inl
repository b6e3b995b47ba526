// Launch wrappers and shape dispatch of the full-batch sweep kernel (included by hpf_engine.cu inside its
// anonymous namespace, after the measured-defaults block and struct hpf_engine).
// launch_sweep_major() resolves (row class, shape options) to one template instantiation of
// sweep_rows_kernel (hpf_sweep_rows.cuh).  The candidate shapes of the fp32 row classes stay compiled in
// and are selectable at run time ("lpg" / "block" / "minb" / "hint" / "fullrow" options), so the tuner
// (tools/tune.py), the test-suite and bench.py all exercise the library that ships.

// number of lane groups of a launch: whole chunks, whole warps (the triple arrays are padded, see build_order)
inline long long padded_groups(int64_t nnz, int chunk, int lpg) {
    const long long per_warp = 32 / lpg;
    long long groups = (nnz + chunk - 1) / chunk;
    return (groups + per_warp - 1) / per_warp * per_warp;
}

template <typename real, int LPG, int VPL, int D, int MINB, int BLOCK, int HINT, bool FULLROW, bool ROBUST, int PF = 0, int SG = 0>
int launch_sweep_rows(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                      const void* xgat, void* acc, const hpf::RescueArgs<real>& rescue) {
    auto kern = hpf::sweep_rows_kernel<real, LPG, VPL, D, MINB, BLOCK, HINT, FULLROW, ROBUST, PF, SG>;
    constexpr int smem = (BLOCK / 32) * (int)hpf::SweepSmem<VPL, SG ? D : 0>::WARP;
    static_assert(smem <= 227 * 1024, "shape does not fit shared memory");
    static thread_local bool configured[64] = {};  // the attribute is per device
    if (h->device >= 64 || !configured[h->device]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured[h->device] = true;
    }
    const long long groups = padded_groups(h->nnz, h->chunk, LPG);
    const long long warps = groups / (32 / LPG);
    const unsigned grid = (unsigned)((warps + BLOCK / 32 - 1) / (BLOCK / 32));
    kern<<<grid, BLOCK, smem, h->stream>>>(row, col, (const real*)val, groups, h->chunk, (const real*)xown,
                                           (const real*)xgat, (real*)acc, h->ld, h->kw, (float)h->keep_frac, rescue,
                                           (real*)nullptr);
    h->launches++;
    CKK();
    return HPF_OK;
}

// run-time flags -> template flags.  Two production forms per shape: the shared-memory ring with whole-stride
// copies (SG + FULLROW: the measured best, needs the row stride to be the whole class width) and the register
// form (any stride; also what ROBUST is built on).  The other combinations are measurement variants, compiled
// for the headline row class only (TUNE).
constexpr int kDefaultHint = 0;   // L2 policies measured neutral-to-harmful (profiles/r02_tune_*.jsonl, traffic probe)
template <typename real, int LPG, int VPL, int D, int MINB, int BLOCK, bool TUNE>
int launch_sweep_rows_flags(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                            const void* xgat, void* acc, const hpf::RescueArgs<real>& rescue, int hint, bool fullrow,
                            bool smem_gather, bool robust) {
    if (smem_gather && !robust) {
        if (fullrow && hint == 0) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, true, false, 0, 1>(h, row, col, val, xown, xgat, acc, rescue);
        if constexpr (TUNE) {
            if (fullrow) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 2, true, false, 0, 1>(h, row, col, val, xown, xgat, acc, rescue);
            return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, false, false, 0, 1>(h, row, col, val, xown, xgat, acc, rescue);
        } else if (h->strict) {
            return fail(HPF_EINVAL, "this smem_gather / fullrow / hint combination is only built for the k<=64 fp32 row class");
        }
        if (fullrow) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, true, false, 0, 1>(h, row, col, val, xown, xgat, acc, rescue);
    }
    if constexpr (D > 4) {
        return fail(HPF_EINVAL, "more than 4 rows in flight per lane group need the shared-memory ring (smem_gather=1, full row stride)");
    } else {
        if (robust) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, false, true>(h, row, col, val, xown, xgat, acc, rescue);
        if constexpr (TUNE) {
            if (h->v_prefetch == 1 && !fullrow) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, false, false, 1>(h, row, col, val, xown, xgat, acc, rescue);
            if (fullrow) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, true, false>(h, row, col, val, xown, xgat, acc, rescue);
            if (hint == 1) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 1, false, false>(h, row, col, val, xown, xgat, acc, rescue);
            if (hint == 2) return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 2, false, false>(h, row, col, val, xown, xgat, acc, rescue);
        } else if (h->strict && (fullrow || hint != 0)) {
            return fail(HPF_EINVAL, "hint / fullrow variants of the register form are only built for the k<=64 fp32 row class");
        }
        return launch_sweep_rows<real, LPG, VPL, D, MINB, BLOCK, 0, false, false>(h, row, col, val, xown, xgat, acc, rescue);
    }
}

struct RowsShape {
    int lpg, depth, block, minb;
};
// measured defaults per row class (16-byte packs per row: 8 = k<=32, 16 = k<=64, 32 = k<=128 in fp32)
inline RowsShape default_rows_shape(int packs, int real_bytes, bool smem_gather) {
    if (smem_gather) {
        if (real_bytes == 4 && packs == 8) return RowsShape{4, 4, 256, 3};
        if (real_bytes == 4 && packs == 16) return RowsShape{8, 4, 256, 3};
        if (real_bytes == 4 && packs == 32) return RowsShape{8, 4, 128, 3};
        if (real_bytes == 8 && packs == 32) return RowsShape{8, 4, 128, 3};
    } else {
        if (real_bytes == 4 && packs == 8) return RowsShape{8, 4, 256, 4};
        if (real_bytes == 4 && packs == 16) return RowsShape{8, 2, 256, 4};
        if (real_bytes == 4 && packs == 32) return RowsShape{16, 2, 256, 3};
        if (real_bytes == 8 && packs == 32) return RowsShape{8, 2, 128, 3};
    }
    return RowsShape{0, 0, 0, 0};
}

template <typename C>
int launch_sweep_major(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                       const void* xgat, void* acc, bool own_is_user) {
    using real = typename C::real;
    if (h->nnz == 0) return HPF_OK;
    constexpr int packs = C::lpg * C::vpl;  // capacity of the row class in 16-byte packs
    constexpr int EPV = 16 / (int)sizeof(real);
    hpf::RescueArgs<real> rescue{};
    const bool robust = h->robust_on;
    if (robust) {
        rescue.shp_own = (const real*)(own_is_user ? h->Gshp : h->Lshp);
        rescue.rte_own = (const real*)(own_is_user ? h->Grte : h->Lrte);
        rescue.shp_gat = (const real*)(own_is_user ? h->Lshp : h->Gshp);
        rescue.rte_gat = (const real*)(own_is_user ? h->Lrte : h->Grte);
        rescue.direct_own = (real*)(own_is_user ? h->dirU : h->dirI);
        rescue.k = h->k;
    }
    // FULLROW copies every pack of the row stride: only meaningful when the stride is the whole class width
    const bool full_ok = h->ld == packs * EPV;
    const bool smem_gather = h->v_smem_gather >= 0 ? h->v_smem_gather != 0 : (full_ok && !robust);
    const bool fullrow = full_ok && (h->v_fullrow >= 0 ? h->v_fullrow != 0 : smem_gather);
    const RowsShape def = default_rows_shape(packs, (int)sizeof(real), smem_gather);
    const int lpg = h->v_lpg ? h->v_lpg : def.lpg, depth = h->v_depth ? h->v_depth : def.depth;
    const int block = h->v_block ? h->v_block : def.block, minb = h->v_minb ? h->v_minb : def.minb;
    const int hint = h->v_hint >= 0 ? h->v_hint : kDefaultHint;
    if (h->chunk % 32 != 0 || h->chunk > kMaxChunk)
        return fail(HPF_EINVAL, "chunk must be a multiple of 32 and <= %d (got %d)", kMaxChunk, h->chunk);
#define HPF_S(L, D, B, M)                                                                                       \
    if (lpg == L && depth == D && block == B && minb == M)                                                      \
        return launch_sweep_rows_flags<real, L, packs / L, D, M, B, packs == 16 && sizeof(real) == 4>(          \
            h, row, col, val, xown, xgat, acc, rescue, hint, fullrow, smem_gather, robust);
    if constexpr (packs == 8 && sizeof(real) == 4) {
        HPF_S(4, 4, 256, 3) HPF_S(8, 4, 256, 4) HPF_S(8, 2, 256, 4) HPF_S(8, 4, 128, 6) HPF_S(4, 4, 256, 2) HPF_S(4, 4, 128, 6)
    }
    if constexpr (packs == 16 && sizeof(real) == 4) {
        HPF_S(8, 4, 256, 3) HPF_S(8, 2, 256, 4) HPF_S(8, 4, 128, 6) HPF_S(8, 4, 256, 2) HPF_S(8, 2, 256, 3)
        HPF_S(16, 4, 256, 4) HPF_S(4, 4, 128, 4) HPF_S(8, 8, 128, 4)
    }
    if constexpr (packs == 32 && sizeof(real) == 4) {
        HPF_S(8, 4, 128, 3) HPF_S(16, 4, 256, 3) HPF_S(16, 2, 256, 3) HPF_S(8, 2, 128, 4) HPF_S(8, 4, 128, 2)
    }
    if constexpr (packs == 32 && sizeof(real) == 8) {
        HPF_S(8, 4, 128, 3) HPF_S(16, 4, 128, 4) HPF_S(8, 2, 128, 3) HPF_S(16, 2, 128, 4)
    }
#undef HPF_S
    if (h->strict && (h->v_lpg || h->v_block || h->v_depth || h->v_minb))
        return fail(HPF_EINVAL, "no such sweep shape for this row class (lpg=%d depth=%d block=%d minb=%d)", lpg, depth, block, minb);
    // generic shape of the classes without a measured table entry: the class's lane-group width, 2 rows in flight
    return launch_sweep_rows_flags<real, C::lpg, C::vpl, 2, 2, 128, false>(h, row, col, val, xown, xgat, acc, rescue, hint,
                                                                          fullrow, smem_gather, robust);
}

// One-pass sweep of a device-assembled minibatch (triples grouped by the batched side, padded like an ordering):
// the batched side accumulates in registers per row, the other side takes one vector RED per pack per nnz.
template <typename C>
int launch_sweep_batch(hpf_engine* h, const int* major, const int* minor, const void* val, int64_t n, const void* xmajor,
                       const void* xminor, void* acc_major, void* acc_minor) {
    using real = typename C::real;
    if (n == 0) return HPF_OK;
    constexpr int BLOCK = 128, D = 2, MINB = 2;
    auto kern = hpf::sweep_rows_kernel<real, C::lpg, C::vpl, D, MINB, BLOCK, 0, false, false, 0, 0, 1>;
    constexpr int smem = (BLOCK / 32) * (int)hpf::SweepSmem<C::vpl, 0>::WARP;
    static thread_local bool configured[64] = {};
    if (h->device >= 64 || !configured[h->device]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured[h->device] = true;
    }
    const int chunk = 64;
    const long long groups = padded_groups(n, chunk, C::lpg);
    const long long warps = groups / (32 / C::lpg);
    const unsigned grid = (unsigned)((warps + BLOCK / 32 - 1) / (BLOCK / 32));
    kern<<<grid, BLOCK, smem, h->stream>>>(major, minor, (const real*)val, groups, chunk, (const real*)xmajor,
                                           (const real*)xminor, (real*)acc_major, h->ld, h->kw, 0.f, hpf::RescueArgs<real>{},
                                           (real*)acc_minor);
    h->launches++;
    CKK();
    return HPF_OK;
}

