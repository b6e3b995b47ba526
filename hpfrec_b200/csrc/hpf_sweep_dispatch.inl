// Launch wrappers and shape dispatch of the CAVI sweep kernels (included by hpf_engine.cu inside its
// anonymous namespace, after the measured-defaults block and struct hpf_engine).  One wrapper per kernel
// generation; launch_sweep_major() resolves (sweep mode, kernel generation, row class, shape options)
// to one template instantiation.  Every candidate shape of the fp32 row classes is compiled in so that
// tuners, tests and bench.py run the library that ships.
// ---- kernel launch wrappers -----------------------------------------------------------------------
template <typename real, int LPG, int VPL, int UNROLL, int MINB, int HINT, int FUSE = 0>
int launch_sweep_variant(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                         const void* xgat, void* acc, void* acc_minor = nullptr) {
    const long long groups = (h->nnz + h->chunk - 1) / h->chunk;
    const long long threads = groups * LPG;
    hpf::sweep_major_kernel<real, LPG, VPL, UNROLL, MINB, HINT, FUSE><<<nblk(threads), 256, 0, h->stream>>>(
        row, col, (const real*)val, h->nnz, h->chunk, (const real*)xown, (const real*)xgat, (real*)acc,
        (real*)acc_minor, h->ld, h->kw);
    h->launches++;
    CKK();
    return HPF_OK;
}

template <typename real, int LPG, int VPL, int MINB, int HINT>
int launch_sweep_v2(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                    const void* xgat, void* acc) {
    const long long groups = (h->nnz + h->chunk - 1) / h->chunk;
    const long long threads = groups * LPG;
    hpf::sweep_major_v2_kernel<real, LPG, VPL, MINB, HINT><<<nblk(threads), 256, 0, h->stream>>>(
        row, col, (const real*)val, h->nnz, h->chunk, (const real*)xown, (const real*)xgat, (real*)acc, h->ld, h->kw);
    h->launches++;
    CKK();
    return HPF_OK;
}

template <typename real, int LPG, int VPL, int MINB, int HINT, int BLOCK>
int launch_sweep_v3(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                    const void* xgat, void* acc) {
    auto kern = hpf::sweep_major_v3_kernel<real, LPG, VPL, MINB, HINT, BLOCK>;
    constexpr int smem = (BLOCK / 32) * 2 * 4 * VPL * 512;  // warps x (gather ring + own ring) x 4 slots x VPL x 512 B
    // the attribute is per device: remember it per (instantiation, device)
    static thread_local bool configured[64] = {};
    if (h->device >= 64 || !configured[h->device]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured[h->device] = true;
    }
    const long long groups = (h->nnz + h->chunk - 1) / h->chunk;
    const long long threads = groups * LPG;
    kern<<<nblk(threads, BLOCK), BLOCK, smem, h->stream>>>(row, col, (const real*)val, h->nnz, h->chunk,
                                                           (const real*)xown, (const real*)xgat, (real*)acc, h->ld, h->kw);
    h->launches++;
    CKK();
    return HPF_OK;
}

template <typename real, int LPG, int VPL, int MINB, int HINT, int BLOCK>
int launch_sweep_v4(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                    const void* xgat, void* acc) {
    auto kern = hpf::sweep_major_v4_kernel<real, LPG, VPL, MINB, HINT, BLOCK>;
    constexpr int smem = (BLOCK / 32) * 2 * 4 * VPL * 512;
    static thread_local bool configured[64] = {};
    if (h->device >= 64 || !configured[h->device]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured[h->device] = true;
    }
    const long long groups = (h->nnz + h->chunk - 1) / h->chunk;
    const long long threads = groups * LPG;
    kern<<<nblk(threads, BLOCK), BLOCK, smem, h->stream>>>(row, col, (const real*)val, h->nnz, h->chunk,
                                                           (const real*)xown, (const real*)xgat, (real*)acc, h->ld, h->kw);
    h->launches++;
    CKK();
    return HPF_OK;
}

// staged-gather sweep (hpf_sweep_tma.cuh): 8 lanes per row, rows staged in shared memory by bulk copies
template <typename real, int VPL, int MINB>
int launch_sweep_tma_variant(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                             const void* xgat, void* acc) {
    auto kern = hpf::sweep_tma_kernel<real, VPL, MINB>;
    const size_t smem = 8 * (32 * (size_t)h->ld * sizeof(real) + 32 * 8);
    static thread_local size_t configured = 0;
    if (configured < smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long groups = (h->nnz + h->chunk - 1) / h->chunk;
    kern<<<nblk(groups * 8), 256, smem, h->stream>>>(row, col, (const real*)val, h->nnz, h->chunk, (const real*)xown,
                                                     (const real*)xgat, (real*)acc, h->ld);
    h->launches++;
    CKK();
    return HPF_OK;
}

template <typename C>
int launch_sweep_tma(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                     const void* xgat, void* acc) {
    using real = typename C::real;
    if (h->nnz == 0) return HPF_OK;
    constexpr int packs = C::lpg * C::vpl;
    if constexpr (packs > 32) {
        return fail(HPF_EINVAL, "staged-gather sweep supports rows up to 512 bytes");
    } else {
        constexpr int VPL = packs <= 8 ? 1 : (packs <= 16 ? 2 : 4);
        const int mb = h->v_minb ? h->v_minb : 3;
        if (mb == 2) return launch_sweep_tma_variant<real, VPL, 2>(h, row, col, val, xown, xgat, acc);
        if (mb == 4) return launch_sweep_tma_variant<real, VPL, 4>(h, row, col, val, xown, xgat, acc);
        return launch_sweep_tma_variant<real, VPL, 3>(h, row, col, val, xown, xgat, acc);
    }
}

// Sweep-kernel shape.  The default (lane-group width, min blocks/SM, load hints) per row-length class
// comes from measurements on B200 (profiles/).  The candidate shapes of the fp32 classes stay compiled
// in, and the "lpg" / "minb" / "hint" options select one at run time, so the tuners
// (tools/tune_r2.py), the test-suite and bench.py all exercise the library that ships.  A shape that
// does not exist for the row class at hand falls back to the default shape, unless option "strict" is
// set (the tuners set it so that a typo cannot be timed as a result).
template <typename C>
int launch_sweep_major(hpf_engine* h, const int* row, const int* col, const void* val, const void* xown,
                       const void* xgat, void* acc, void* acc_minor = nullptr) {
    using real = typename C::real;
    if (h->nnz == 0) return HPF_OK;
    constexpr int packs = C::lpg * C::vpl;  // capacity of the row class in 16-byte packs
    const bool fused = acc_minor != nullptr;
    // shape = measured default of the class, overridden field by field by the options
    const Shape def = default_shape(packs, (int)sizeof(real), fused);
    const int lpg = h->v_lpg ? h->v_lpg : def.lpg, mb = h->v_minb ? h->v_minb : def.minb;
    const int hint = h->v_hint >= 0 ? h->v_hint : def.hint;
    if (!fused && h->kernel_ver == 4 && h->chunk % 4 == 0) {  // deep pipeline with vector-loaded triples
        const Shape d4 = default_shape_v3(packs, (int)sizeof(real));
        const int l4 = h->v_lpg ? h->v_lpg : d4.lpg, m4 = h->v_minb ? h->v_minb : d4.minb;
        const int b4 = h->v_block ? h->v_block : default_block_v3(packs, (int)sizeof(real));
#define HPF_R(L, M, B)                          \
    if (l4 == L && m4 == M && b4 == B)          \
        return launch_sweep_v4<real, L, packs / L, M, 0, B>(h, row, col, val, xown, xgat, acc);
        if constexpr (packs == 16 && sizeof(real) == 4) {
            HPF_R(8, 2, 256) HPF_R(8, 3, 256) HPF_R(4, 2, 128) HPF_R(4, 3, 128) HPF_R(16, 4, 256) HPF_R(8, 4, 128)
        }
        if constexpr (packs == 8 && sizeof(real) == 4) {
            HPF_R(4, 2, 256) HPF_R(4, 3, 256) HPF_R(8, 4, 256)
        }
        if constexpr (packs == 32 && sizeof(real) == 4) {
            HPF_R(8, 3, 128) HPF_R(8, 2, 128) HPF_R(16, 2, 256) HPF_R(16, 3, 256)
        }
#undef HPF_R
        if (h->strict && (h->v_lpg || h->v_minb || h->v_block))
            return fail(HPF_EINVAL, "no such vector-triple sweep shape for this row class (lpg=%d minb=%d block=%d)", l4, m4, b4);
        constexpr int gm4 = C::vpl == 1 ? 4 : (C::vpl == 2 ? 3 : 1);
        return launch_sweep_v4<real, C::lpg, C::vpl, gm4, 0, 256>(h, row, col, val, xown, xgat, acc);
    }
    if (!fused && (h->kernel_ver == 3 || h->kernel_ver == 4)) {  // deep-pipeline two-pass kernel (sweep_major_v3_kernel, cp.async rings)
        const Shape d3 = default_shape_v3(packs, (int)sizeof(real));
        const int l3 = h->v_lpg ? h->v_lpg : d3.lpg, m3 = h->v_minb ? h->v_minb : d3.minb;
        const int h3 = h->v_hint >= 0 ? h->v_hint : d3.hint;
        const int b3 = h->v_block ? h->v_block : default_block_v3(packs, (int)sizeof(real));
#define HPF_Q(L, M, H, B)                               \
    if (l3 == L && m3 == M && h3 == H && b3 == B)       \
        return launch_sweep_v3<real, L, packs / L, M, H, B>(h, row, col, val, xown, xgat, acc);
        if constexpr (packs == 16 && sizeof(real) == 4) {
            HPF_Q(4, 2, 0, 128) HPF_Q(4, 3, 0, 128) HPF_Q(4, 3, 1, 128)
            HPF_Q(8, 2, 0, 256) HPF_Q(8, 3, 0, 256) HPF_Q(8, 4, 0, 128) HPF_Q(8, 6, 0, 128) HPF_Q(8, 2, 1, 256)
            HPF_Q(16, 4, 0, 256) HPF_Q(16, 6, 0, 256)
        }
        if constexpr (packs == 8 && sizeof(real) == 4) {
            HPF_Q(4, 2, 0, 256) HPF_Q(4, 3, 0, 256) HPF_Q(4, 6, 0, 128) HPF_Q(8, 4, 0, 256) HPF_Q(8, 6, 0, 256)
        }
        if constexpr (packs == 32 && sizeof(real) == 4) {
            HPF_Q(8, 2, 0, 128) HPF_Q(8, 3, 0, 128) HPF_Q(16, 2, 0, 256) HPF_Q(16, 3, 0, 256) HPF_Q(32, 4, 0, 256) HPF_Q(32, 6, 0, 256)
        }
#undef HPF_Q
        if (h->strict && (h->v_lpg || h->v_minb || h->v_hint >= 0 || h->v_block))
            return fail(HPF_EINVAL, "no such deep-pipeline sweep shape for this row class (lpg=%d minb=%d hint=%d block=%d)", l3, m3, h3, b3);
        // generic shape (fp64, rows beyond 512 bytes): resident CTAs follow the shared-memory footprint
        constexpr int gm = C::vpl == 1 ? 4 : (C::vpl == 2 ? 3 : 1);
        return launch_sweep_v3<real, C::lpg, C::vpl, gm, 0, 256>(h, row, col, val, xown, xgat, acc);
    }
    if (!fused && h->kernel_ver == 2) {  // pipelined two-pass kernel (sweep_major_v2_kernel)
        const Shape d2 = default_shape_v2(packs, (int)sizeof(real));
        const int l2 = h->v_lpg ? h->v_lpg : d2.lpg, m2 = h->v_minb ? h->v_minb : d2.minb;
        const int h2 = h->v_hint >= 0 ? h->v_hint : d2.hint;
#define HPF_P(L, M, H)                          \
    if (l2 == L && m2 == M && h2 == H)          \
        return launch_sweep_v2<real, L, packs / L, M, H>(h, row, col, val, xown, xgat, acc);
#define HPF_PL(L) HPF_P(L, 2, 0) HPF_P(L, 3, 0) HPF_P(L, 4, 0) HPF_P(L, 5, 0) HPF_P(L, 6, 0) \
                  HPF_P(L, 2, 1) HPF_P(L, 3, 1) HPF_P(L, 4, 1) HPF_P(L, 5, 1) HPF_P(L, 6, 1)
        if constexpr (packs == 16 && sizeof(real) == 4) {
            HPF_PL(4) HPF_PL(8) HPF_PL(16)
        }
        if constexpr (packs == 8 && sizeof(real) == 4) {
            HPF_P(4, 3, 0) HPF_P(4, 4, 0) HPF_P(4, 6, 0) HPF_P(8, 4, 0) HPF_P(8, 6, 0) HPF_P(8, 8, 0)
        }
        if constexpr (packs == 32 && sizeof(real) == 4) {
            HPF_P(8, 2, 0) HPF_P(8, 3, 0) HPF_P(16, 2, 0) HPF_P(16, 3, 0) HPF_P(16, 4, 0) HPF_P(32, 3, 0) HPF_P(32, 4, 0)
        }
#undef HPF_PL
#undef HPF_P
        if (h->strict && (h->v_lpg || h->v_minb || h->v_hint >= 0))
            return fail(HPF_EINVAL, "no such pipelined sweep shape for this row class (lpg=%d minb=%d hint=%d)", l2, m2, h2);
        // generic shape (fp64, rows beyond 512 bytes)
        if constexpr (packs <= 16) return launch_sweep_v2<real, 8, packs / 8, 3, 0>(h, row, col, val, xown, xgat, acc);
        else return launch_sweep_v2<real, C::lpg, C::vpl, 2, 0>(h, row, col, val, xown, xgat, acc);
    }
    if (fused) {
#define HPF_F(L, M, H)                          \
    if (lpg == L && mb == M && hint == H)       \
        return launch_sweep_variant<real, L, packs / L, 1, M, H, 1>(h, row, col, val, xown, xgat, acc, acc_minor);
#define HPF_FL(L) HPF_F(L, 2, 0) HPF_F(L, 3, 0) HPF_F(L, 4, 0) HPF_F(L, 3, 1) HPF_F(L, 4, 1) HPF_F(L, 3, 3) HPF_F(L, 4, 3)
        if constexpr (packs == 16 && sizeof(real) == 4) {
            HPF_FL(4) HPF_FL(8) HPF_FL(16) HPF_F(8, 5, 0) HPF_F(8, 6, 0) HPF_F(16, 5, 0) HPF_F(16, 6, 0)
        }
        if constexpr (packs == 8 && sizeof(real) == 4) {
            HPF_F(4, 2, 0) HPF_F(4, 3, 0) HPF_F(4, 4, 0) HPF_F(8, 3, 0) HPF_F(8, 4, 0) HPF_F(8, 6, 0)
        }
        if constexpr (packs == 32 && sizeof(real) == 4) {
            HPF_F(8, 2, 0) HPF_F(8, 3, 0) HPF_F(8, 4, 0) HPF_F(16, 2, 0) HPF_F(16, 3, 0) HPF_F(16, 4, 0) HPF_F(32, 2, 0) HPF_F(32, 4, 0)
        }
#undef HPF_FL
#undef HPF_F
    } else {
#define HPF_V(L, M, H)                          \
    if (lpg == L && mb == M && hint == H)       \
        return launch_sweep_variant<real, L, packs / L, 1, M, H>(h, row, col, val, xown, xgat, acc);
#define HPF_VL(L) HPF_V(L, 2, 0) HPF_V(L, 3, 0) HPF_V(L, 4, 0) HPF_V(L, 2, 1) HPF_V(L, 3, 1) HPF_V(L, 4, 1) \
                  HPF_V(L, 2, 3) HPF_V(L, 3, 3) HPF_V(L, 4, 3)
        if constexpr (packs == 16 && sizeof(real) == 4) {
            HPF_VL(4) HPF_VL(8) HPF_VL(16)
            HPF_V(8, 5, 0) HPF_V(8, 6, 0) HPF_V(8, 5, 1) HPF_V(8, 6, 1) HPF_V(8, 5, 3) HPF_V(8, 6, 3)
            HPF_V(16, 5, 0) HPF_V(16, 6, 0) HPF_V(16, 8, 0) HPF_V(16, 5, 3) HPF_V(16, 6, 3) HPF_V(16, 8, 3)
        }
        if constexpr (packs == 8 && sizeof(real) == 4) {
            HPF_VL(4) HPF_VL(8)
        }
        if constexpr (packs == 32 && sizeof(real) == 4) {
            HPF_VL(8) HPF_VL(16) HPF_VL(32)
        }
#undef HPF_VL
#undef HPF_V
    }
    if (h->strict && (h->v_lpg || h->v_minb || h->v_hint >= 0))
        return fail(HPF_EINVAL, "no such %s sweep shape for this row class (lpg=%d minb=%d hint=%d)",
                    fused ? "fused" : "two-pass", lpg, mb, hint);
    // generic shape of the classes without a measured table entry (fp64, rows beyond 512 bytes)
    if (fused) {
        if constexpr (packs <= 16) return launch_sweep_variant<real, 8, packs / 8, 1, 3, 0, 1>(h, row, col, val, xown, xgat, acc, acc_minor);
        else return launch_sweep_variant<real, C::lpg, C::vpl, 1, 2, 0, 1>(h, row, col, val, xown, xgat, acc, acc_minor);
    }
    if constexpr (packs <= 8) return launch_sweep_variant<real, 4, 2, 1, 2, 0>(h, row, col, val, xown, xgat, acc);
    else if constexpr (packs <= 16) return launch_sweep_variant<real, 4, 4, 1, 3, 1>(h, row, col, val, xown, xgat, acc);
    else if constexpr (packs <= 32) return launch_sweep_variant<real, 8, 4, 1, 4, 0>(h, row, col, val, xown, xgat, acc);
    else return launch_sweep_variant<real, C::lpg, C::vpl, 1, 2, 0>(h, row, col, val, xown, xgat, acc);
}

