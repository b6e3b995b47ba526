// tma_gather_probe -- can the TMA engine stage the sweep's row gathers faster than per-thread cp.async?
//
// The sweep kernel (hpfrec_b200/csrc/hpf_sweep.cuh) moves one 208..256-byte factor row per nnz from an
// L2-resident panel into shared memory with LDGSTS (cp.async, one 16-byte copy per lane).  sm_100 adds
// cp.async.bulk.tensor...tile::gather4: ONE TMA operation fetches FOUR arbitrary rows of a 2-D tensor map.
// This probe issues exactly the staging traffic of the sweep and nothing else, three ways:
//   ldgsts   every lane copies 2 x 16 bytes per step (4 rows x 256 B per warp step), cp.async groups
//   gather4  lane 0 of every warp issues one tile::gather4 per step (4 rows x 256 B), mbarrier completion
//   tile1    lane 0 issues four ordinary 2-D tile loads of one row each (the per-row TMA op rate)
// Every step's 1 KB is read back from shared memory by the warp (2 x LDS.128 per lane) so the data path is
// complete.  Output: one JSON line per (mode, window, depth, warps/CTA): ms for n rows and rows/s.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/tma_gather_probe.cu -o tools/bin/tma_gather_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)

constexpr int LD = 64;            // floats per row (256-byte stride, the k=50 fp32 layout)
constexpr int STEP_BYTES = 1024;  // 4 rows x 256 B

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void make_ids(int* ids, long long n, int rows, int windows, uint64_t seed) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long per = (n + windows - 1) / windows;
    const int w = (int)(p / per);
    const int rows_per = (rows + windows - 1) / windows;
    const int lo = w * rows_per;
    int span = rows - lo < rows_per ? rows - lo : rows_per;
    if (span < 1) span = 1;
    ids[p] = lo + (int)(mix64(seed + (uint64_t)p) % (uint64_t)span);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int4 r, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
            "r"(dst), "l"(tm), "r"(col), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tm), "r"(col), "r"(row), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// MODE 0 ldgsts, 1 gather4, 2 four single-row tile loads.  Every warp owns `steps` consecutive steps
// (4 ids each) of the id list.
template <int MODE, int DEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
stage_kernel(const int4* __restrict__ ids4, long long nsteps_total, int steps, const float* __restrict__ table,
             const __grid_constant__ CUtensorMap tmap, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * WARPS + warp;
    const long long s0 = w * steps;
    if (s0 >= nsteps_total) return;
    const int ns = (int)(nsteps_total - s0 < steps ? nsteps_total - s0 : steps);
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * (DEPTH * STEP_BYTES);
    const uint32_t bars = smem_u32(smem) + (uint32_t)WARPS * (DEPTH * STEP_BYTES) + (uint32_t)warp * (DEPTH * 8);
    if (MODE != 0) {
        if (lane < DEPTH) mbar_init(bars + lane * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue = [&](int t) {
        const int slot = t % DEPTH;
        const int4 c = __ldg(ids4 + s0 + t);
        if (MODE == 0) {
            const int g = lane >> 3, gl = lane & 7;
            const int r = g == 0 ? c.x : (g == 1 ? c.y : (g == 2 ? c.z : c.w));
            const char* src = reinterpret_cast<const char*>(table) + (size_t)r * (LD * 4) + gl * 16;
            const uint32_t dst = ring + slot * STEP_BYTES + g * 256 + gl * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 128), "l"(src + 128) : "memory");
        } else if (lane == 0) {
            const uint32_t bar = bars + slot * 8;
            mbar_expect_tx(bar, STEP_BYTES);
            if (MODE == 1) {
                tma_gather4(ring + slot * STEP_BYTES, &tmap, 0, c, bar);
            } else {
                tma_tile(ring + slot * STEP_BYTES, &tmap, 0, c.x, bar);
                tma_tile(ring + slot * STEP_BYTES + 256, &tmap, 0, c.y, bar);
                tma_tile(ring + slot * STEP_BYTES + 512, &tmap, 0, c.z, bar);
                tma_tile(ring + slot * STEP_BYTES + 768, &tmap, 0, c.w, bar);
            }
        }
        if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < DEPTH - 1; ++t) {
        if (t < ns) issue(t);
        else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int t = 0; t < ns; ++t) {
        const int slot = t % DEPTH;
        // stage step t + DEPTH - 1 into the slot consumed at step t - 1
        if (t + DEPTH - 1 < ns) issue(t + DEPTH - 1);
        else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        if (MODE == 0) {
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            __syncwarp();
        } else {
            mbar_wait(bars + slot * 8, (uint32_t)((t / DEPTH) & 1));
        }
        const float4 a = lds128(ring + slot * STEP_BYTES + lane * 16);
        const float4 b = lds128(ring + slot * STEP_BYTES + 512 + lane * 16);
        acc.x += a.x + b.x;
        acc.y += a.y + b.y;
        acc.z += a.z + b.z;
        acc.w += a.w + b.w;
        __syncwarp();  // every lane has read the slot before it is overwritten
    }
    if (MODE == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    const float s = acc.x + acc.y + acc.z + acc.w;
    if (s == 12345.678f) out[w] = s;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int DEPTH, int WARPS>
float run(const int4* ids4, long long nsteps, int steps, const float* table, const CUtensorMap& tm, float* out, cudaEvent_t e0,
          cudaEvent_t e1) {
    auto kern = stage_kernel<MODE, DEPTH, WARPS>;
    const int smem = WARPS * (DEPTH * STEP_BYTES) + WARPS * DEPTH * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long warps = (nsteps + steps - 1) / steps;
    const unsigned grid = (unsigned)((warps + WARPS - 1) / WARPS);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, WARPS * 32, smem>>>(ids4, nsteps, steps, table, tm, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    long long n = 48000000;
    int rows = 1000000;
    for (int a = 1; a < argc; ++a) {
        if (!strcmp(argv[a], "--n") && a + 1 < argc) n = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--rows") && a + 1 < argc) rows = atoi(argv[++a]);
    }
    n = n / 4 * 4;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "tma_gather_probe: no CUDA device\n");
        return 2;
    }
    int* ids = nullptr;
    float *table = nullptr, *out = nullptr;
    CK(cudaMalloc(&ids, sizeof(int) * (size_t)n));
    CK(cudaMalloc(&table, sizeof(float) * (size_t)rows * LD));
    CK(cudaMemset(table, 0, sizeof(float) * (size_t)rows * LD));
    CK(cudaMalloc(&out, sizeof(float) * (size_t)(n / 4 / 16 + 1)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) {
        fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
        return 2;
    }
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)LD, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)LD * 4};
    const cuuint32_t box[2] = {(cuuint32_t)LD, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)cr);
        return 2;
    }
    const double window_mb[] = {32, 64, 96, 1e9};
    const int steps = 64;  // 256 rows per warp
    const long long nsteps = n / 4;
    for (double wmb : window_mb) {
        const double table_mb = (double)rows * LD * 4 / 1048576.0;
        int windows = (int)((table_mb + wmb - 1e-9) / wmb);
        if (windows < 1) windows = 1;
        make_ids<<<(unsigned)((n + 255) / 256), 256>>>(ids, n, rows, windows, 0x5eedull + (uint64_t)windows);
        CK(cudaGetLastError());
        const int4* ids4 = reinterpret_cast<const int4*>(ids);
        struct Res { const char* mode; int depth, warps; float ms; };
        Res res[] = {
            {"ldgsts", 4, 8, run<0, 4, 8>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"ldgsts", 8, 8, run<0, 8, 8>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"ldgsts", 4, 16, run<0, 4, 16>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"gather4", 4, 8, run<1, 4, 8>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"gather4", 8, 8, run<1, 8, 8>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"gather4", 4, 16, run<1, 4, 16>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"gather4", 8, 16, run<1, 8, 16>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"gather4", 16, 4, run<1, 16, 4>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"tile1", 4, 8, run<2, 4, 8>(ids4, nsteps, steps, table, tm, out, e0, e1)},
            {"tile1", 8, 16, run<2, 8, 16>(ids4, nsteps, steps, table, tm, out, e0, e1)},
        };
        for (const Res& r : res)
            printf("{\"probe\": \"stage\", \"mode\": \"%s\", \"n_rows\": %lld, \"table_rows\": %d, \"row_bytes\": %d, "
                   "\"window_mb\": %.0f, \"windows\": %d, \"depth\": %d, \"warps_per_cta\": %d, \"ms\": %.4f, "
                   "\"grows_per_s\": %.3f, \"gbs\": %.1f}\n",
                   r.mode, n, rows, LD * 4, wmb > 1e8 ? table_mb : wmb, windows, r.depth, r.warps, r.ms,
                   (double)n / (r.ms * 1e-3) / 1e9, (double)n * LD * 4 / (r.ms * 1e-3) / 1e9);
        fflush(stdout);
    }
    return 0;
}
