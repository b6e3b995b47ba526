// gather_probe -- measured ceiling of the access pattern that bounds the CAVI sweep (sm_100a).
//
// The sweep's dominant cost is one k-wide factor-row gather per nnz per pass: 48M random
// rows of 200-224 bytes out of an L2-sized window of a (rows x ld) matrix.  This probe issues exactly
// that pattern with NOTHING else (no dot product, no softmax normaliser, no REDs): every lane group
// walks a chunk of pre-generated random row ids, loads the row with 128-bit loads and adds it into
// registers.  Its rows/s is the roofline of the gather path for a given (row stride, active bytes,
// window size, lane-group width, occupancy); the sweep kernel is reported against it.
//
// Build (hpfrec_b200/build.py does this):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
//                                          tools/gather_probe.cu -o tools/bin/gather_probe
// Run on the GPU box:  tools/bin/gather_probe [--n 48000000] [--rows 380000] [--quick]   (JSON lines)
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// ids[p] uniform inside the window that position p belongs to (positions are cut into `windows` equal
// runs, like the (panel, major)-sorted triples of the engine)
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

template <int NOALLOC>
__device__ __forceinline__ float4 load16(const float* p) {
    float4 r;
    if (NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "l"(p));
    else
        r = __ldg(reinterpret_cast<const float4*>(p));
    return r;
}

// L lanes per row, VPL 16-byte packs per lane; `packs` = active packs per row (<= L*VPL)
template <int L, int VPL, int MINB, int NOALLOC>
__global__ void __launch_bounds__(256, MINB)
gather_kernel(const int* __restrict__ ids, long long n, int chunk, const float* __restrict__ table, int ld,
              int packs, float* __restrict__ out) {
    const int gl = (threadIdx.x & 31) % L;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
    const long long beg = group * (long long)chunk;
    if (beg >= n) return;
    const long long end = beg + chunk < n ? beg + chunk : n;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = L == 32 ? 0xffffffffu : (((1u << L) - 1u) << (lane & ~(unsigned)(L - 1)));
    float4 acc[VPL];
    int off[VPL];
    bool act[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        off[v] = (gl + L * v) * 4;
        act[v] = (gl + L * v) < packs;
    }
    for (long long base = beg; base < end; base += L) {
        const int c = (base + gl < end) ? __ldg(ids + base + gl) : 0;
        const int cnt = end - base < L ? (int)(end - base) : L;
        for (int t = 0; t < cnt; ++t) {
            const int cc = __shfl_sync(gmask, c, t, L);
            const float* src = table + (size_t)cc * ld;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (!act[v]) continue;
                const float4 g = load16<NOALLOC>(src + off[v]);
                acc[v].x += g.x;
                acc[v].y += g.y;
                acc[v].z += g.z;
                acc[v].w += g.w;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) s += acc[v].x + acc[v].y + acc[v].z + acc[v].w;
    if (s == 12345.678f) out[group] = s;  // keeps the loads alive; never true for a zero-filled table
}

// the same walk issuing one vector RED (RED.E.ADD.F32x4) per active pack instead of a load: the ceiling
// of the one-pass sweep's scatter side
template <int L, int VPL, int MINB, int NOALLOC>
__global__ void __launch_bounds__(256, MINB)
red_kernel(const int* __restrict__ ids, long long n, int chunk, const float* __restrict__ table_c, int ld,
           int packs, float* __restrict__ out) {
    float* table = const_cast<float*>(table_c);
    const int gl = (threadIdx.x & 31) % L;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
    const long long beg = group * (long long)chunk;
    if (beg >= n) return;
    const long long end = beg + chunk < n ? beg + chunk : n;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = L == 32 ? 0xffffffffu : (((1u << L) - 1u) << (lane & ~(unsigned)(L - 1)));
    for (long long base = beg; base < end; base += L) {
        const int c = (base + gl < end) ? __ldg(ids + base + gl) : 0;
        const int cnt = end - base < L ? (int)(end - base) : L;
        for (int t = 0; t < cnt; ++t) {
            const int cc = __shfl_sync(gmask, c, t, L);
            float* dst = table + (size_t)cc * ld;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if ((gl + L * v) >= packs) continue;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + (gl + L * v) * 4), "f"(0.0f)
                             : "memory");
            }
        }
    }
    (void)out;
}

struct Variant {
    int L, minb, noalloc;
    void (*fn)(const int*, long long, int, const float*, int, int, float*);
};

#define V(L, M, N) {L, M, N, gather_kernel<L, 16 / L, M, N>}
#define R(L, M) {L, M, 0, red_kernel<L, 16 / L, M, 0>}
static const Variant kRedVariants[] = {R(4, 4), R(8, 4), R(8, 8), R(16, 8)};
static const Variant kVariants[] = {
    V(4, 2, 0), V(4, 3, 0), V(4, 4, 0), V(4, 3, 1), V(4, 4, 1),
    V(8, 2, 0), V(8, 3, 0), V(8, 4, 0), V(8, 6, 0), V(8, 8, 0), V(8, 4, 1), V(8, 6, 1), V(8, 8, 1),
    V(16, 4, 0), V(16, 6, 0), V(16, 8, 0), V(16, 6, 1), V(16, 8, 1),
};

int main(int argc, char** argv) {
    long long n = 48000000;
    int rows = 380000;
    bool quick = false;
    for (int a = 1; a < argc; ++a) {
        if (!strcmp(argv[a], "--n") && a + 1 < argc) n = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--rows") && a + 1 < argc) rows = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--quick")) quick = true;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "gather_probe: no CUDA device\n");
        return 2;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int* ids = nullptr;
    float *table = nullptr, *out = nullptr;
    const int chunk = 64;
    CK(cudaMalloc(&ids, sizeof(int) * (size_t)n));
    CK(cudaMalloc(&table, sizeof(float) * (size_t)rows * 64));
    CK(cudaMemset(table, 0, sizeof(float) * (size_t)rows * 64));
    CK(cudaMalloc(&out, sizeof(float) * (size_t)(n / chunk + 1)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    struct Layout { int ld, packs; };
    const Layout layouts[] = {{56, 14}, {56, 13}, {64, 13}, {64, 16}};
    const double window_mb[] = {16, 24, 48, 96, 1e9};
    for (const Layout& lay : layouts) {
        for (double wmb : window_mb) {
            const double table_mb = (double)rows * lay.ld * 4 / 1048576.0;
            int windows = (int)((table_mb + wmb - 1e-9) / wmb);
            if (windows < 1) windows = 1;
            if (wmb < 1e8 && wmb >= table_mb) continue;  // window larger than the table: same as "all"
            make_ids<<<(unsigned)((n + 255) / 256), 256>>>(ids, n, rows, windows, 0x5eedull + (uint64_t)windows);
            CK(cudaGetLastError());
            for (const Variant& v : kVariants) {
                if (quick && !(v.minb == 4 || v.minb == 8)) continue;
                const long long groups = (n + chunk - 1) / chunk;
                const unsigned grid = (unsigned)((groups * v.L + 255) / 256);
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    CK(cudaEventRecord(e0));
                    v.fn<<<grid, 256>>>(ids, n, chunk, table, lay.ld, lay.packs, out);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep > 0 && ms < best) best = ms;
                }
                CK(cudaGetLastError());
                const double rows_per_s = (double)n / (best * 1e-3);
                printf("{\"probe\": \"gather\", \"gpu\": \"%s\", \"n\": %lld, \"table_rows\": %d, \"ld_floats\": %d, "
                       "\"active_packs\": %d, \"window_mb\": %.0f, \"windows\": %d, \"lanes\": %d, \"minb\": %d, "
                       "\"no_allocate\": %d, \"ms\": %.4f, \"grows_per_s\": %.3f, \"useful_gbs\": %.1f, "
                       "\"sector_gbs\": %.1f}\n",
                       prop.name, n, rows, lay.ld, lay.packs, wmb > 1e8 ? table_mb : wmb, windows, v.L, v.minb,
                       v.noalloc, best, rows_per_s / 1e9, rows_per_s * lay.packs * 16 / 1e9,
                       rows_per_s * ((lay.packs * 16 + 31) / 32) * 32 / 1e9);
                fflush(stdout);
            }
            for (const Variant& v : kRedVariants) {
                if (lay.packs == 14) continue;
                const long long groups = (n + chunk - 1) / chunk;
                const unsigned grid = (unsigned)((groups * v.L + 255) / 256);
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaEventRecord(e0));
                    v.fn<<<grid, 256>>>(ids, n, chunk, table, lay.ld, lay.packs, out);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep > 0 && ms < best) best = ms;
                }
                CK(cudaGetLastError());
                printf("{\"probe\": \"red\", \"n\": %lld, \"table_rows\": %d, \"ld_floats\": %d, \"active_packs\": %d, "
                       "\"window_mb\": %.0f, \"windows\": %d, \"lanes\": %d, \"minb\": %d, \"ms\": %.4f, "
                       "\"grows_per_s\": %.3f}\n",
                       n, rows, lay.ld, lay.packs, wmb > 1e8 ? table_mb : wmb, windows, v.L, v.minb, best,
                       (double)n / (best * 1e-3) / 1e9);
                fflush(stdout);
            }
        }
    }
    cudaFree(ids);
    cudaFree(table);
    cudaFree(out);
    return 0;
}
