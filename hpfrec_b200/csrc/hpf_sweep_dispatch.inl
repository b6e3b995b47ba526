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

template <typename real, int LPG, int VPL, int MINB, int BLOCK, int HINT, bool FULLROW, bool ROBUST>
int launch_sweep_rows(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                      const void* xgat, void* acc, const hpf::RescueArgs<real>& rescue) {
    auto kern = hpf::sweep_rows_kernel<real, LPG, VPL, MINB, BLOCK, HINT, FULLROW, ROBUST>;
    constexpr int smem = (BLOCK / 32) * (int)hpf::SweepSmem<VPL>::WARP;
    static thread_local bool configured[64] = {};  // the attribute is per device
    if (h->device >= 64 || !configured[h->device]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured[h->device] = true;
    }
    const long long groups = padded_groups(h->nnz, h->chunk, LPG);
    const long long warps = groups / (32 / LPG);
    const unsigned grid = (unsigned)((warps + BLOCK / 32 - 1) / (BLOCK / 32));
    kern<<<grid, BLOCK, smem, h->stream>>>(row, col, (const real*)val, groups, h->chunk, (const real*)xown,
                                           (const real*)xgat, (real*)acc, h->ld, h->kw, rescue);
    h->launches++;
    CKK();
    return HPF_OK;
}

// run-time flags -> template flags.  The hint-less and FULLROW forms are measurement variants, compiled for the
// headline row class only (TUNE); ROBUST is built with the default hints and without FULLROW.
template <typename real, int LPG, int VPL, int MINB, int BLOCK, bool TUNE>
int launch_sweep_rows_flags(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                            const void* xgat, void* acc, const hpf::RescueArgs<real>& rescue, int hint, bool fullrow,
                            bool robust) {
    if (robust) return launch_sweep_rows<real, LPG, VPL, MINB, BLOCK, 1, false, true>(h, row, col, val, xown, xgat, acc, rescue);
    if constexpr (TUNE) {
        if (fullrow) {
            if (hint) return launch_sweep_rows<real, LPG, VPL, MINB, BLOCK, 1, true, false>(h, row, col, val, xown, xgat, acc, rescue);
            return launch_sweep_rows<real, LPG, VPL, MINB, BLOCK, 0, true, false>(h, row, col, val, xown, xgat, acc, rescue);
        }
        if (!hint) return launch_sweep_rows<real, LPG, VPL, MINB, BLOCK, 0, false, false>(h, row, col, val, xown, xgat, acc, rescue);
    } else if (h->strict && (fullrow || !hint)) {
        return fail(HPF_EINVAL, "hint=0 / fullrow=1 are only built for the k<=64 fp32 row class");
    }
    return launch_sweep_rows<real, LPG, VPL, MINB, BLOCK, 1, false, false>(h, row, col, val, xown, xgat, acc, rescue);
}

// resident CTAs per SM that the shared-memory footprint of a shape allows (also its launch bound)
template <int VPL, int BLOCK>
constexpr int smem_ctas() {
    constexpr int per_cta = (BLOCK / 32) * (int)hpf::SweepSmem<VPL>::WARP + 1024;
    constexpr int fit = (227 * 1024) / per_cta;
    constexpr int by_threads = 2048 / BLOCK;
    return fit < 1 ? 1 : (fit > by_threads ? by_threads : (fit > 8 ? 8 : fit));
}

struct RowsShape {
    int lpg, block;
};
// measured defaults per fp32 row class (16-byte packs per row: 8 = k<=32, 16 = k<=64, 32 = k<=128)
inline RowsShape default_rows_shape(int packs, int real_bytes) {
    if (real_bytes == 4 && packs == 8) return RowsShape{4, 256};
    if (real_bytes == 4 && packs == 16) return RowsShape{4, 128};
    if (real_bytes == 4 && packs == 32) return RowsShape{8, 128};
    return RowsShape{0, 0};
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
    const RowsShape def = default_rows_shape(packs, (int)sizeof(real));
    const int lpg = h->v_lpg ? h->v_lpg : def.lpg;
    const int block = h->v_block ? h->v_block : def.block;
    const int hint = h->v_hint >= 0 ? (h->v_hint ? 1 : 0) : 1;
    // FULLROW copies every pack of the row stride: only meaningful when the stride is the whole class width
    const bool full_ok = h->ld == packs * EPV;
    const bool fullrow = full_ok && (h->v_fullrow >= 0 ? h->v_fullrow != 0 : false);
    if (h->chunk % 32 != 0 || h->chunk > kMaxChunk)
        return fail(HPF_EINVAL, "chunk must be a multiple of 32 and <= %d (got %d)", kMaxChunk, h->chunk);
#define HPF_S(L, B)                                                                                         \
    if (lpg == L && block == B)                                                                             \
        return launch_sweep_rows_flags<real, L, packs / L, smem_ctas<packs / L, B>(), B, packs == 16 && sizeof(real) == 4>( \
            h, row, col, val, xown, xgat, acc, rescue, hint, fullrow, robust);
    if constexpr (packs == 8 && sizeof(real) == 4) {
        HPF_S(4, 256) HPF_S(4, 128) HPF_S(8, 256)
    }
    if constexpr (packs == 16 && sizeof(real) == 4) {
        HPF_S(4, 128) HPF_S(8, 256) HPF_S(8, 128) HPF_S(4, 64)
    }
    if constexpr (packs == 32 && sizeof(real) == 4) {
        HPF_S(8, 128) HPF_S(16, 256) HPF_S(16, 128) HPF_S(8, 64)
    }
    if constexpr (packs == 32 && sizeof(real) == 8) {
        HPF_S(8, 128) HPF_S(16, 128)
    }
#undef HPF_S
    if (h->strict && (h->v_lpg || h->v_block))
        return fail(HPF_EINVAL, "no such sweep shape for this row class (lpg=%d block=%d)", lpg, block);
    // generic shape of the classes without a measured table entry (fp64, rows beyond 512 bytes)
    return launch_sweep_rows_flags<real, C::lpg, C::vpl, smem_ctas<C::vpl, 128>(), 128, false>(h, row, col, val, xown, xgat, acc,
                                                                                               rescue, hint, fullrow, robust);
}
