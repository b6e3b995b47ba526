// Device-side building blocks of the HPF engine (sm_100a): 16-byte packs, lane-group reductions,
// digamma.  No host code here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace hpf {

// ---------------------------------------------------------------------------------------------
// 16-byte pack of `real` (float4 / double2): every factor-row access in the engine is one 128-bit
// load or store, rows are padded to a multiple of 32 B so a row never shares a sector.
// ---------------------------------------------------------------------------------------------
template <typename real>
struct alignas(16) Pack {
    static constexpr int N = 16 / sizeof(real);
    real v[N];
};

template <typename real>
__device__ __forceinline__ Pack<real> pack_zero() {
    Pack<real> p;
#pragma unroll
    for (int e = 0; e < Pack<real>::N; ++e) p.v[e] = real(0);
    return p;
}

// read-only 128-bit load (LDG.E.128.CONSTANT); data is never written by the reading kernel
template <typename real>
__device__ __forceinline__ Pack<real> ldg_pack(const real* p) {
    Pack<real> r;
    *reinterpret_cast<int4*>(&r) = __ldg(reinterpret_cast<const int4*>(p));
    return r;
}
// plain 128-bit load of data this kernel may also write (no .nc)
template <typename real>
__device__ __forceinline__ Pack<real> ld_pack(const real* p) {
    Pack<real> r;
    *reinterpret_cast<int4*>(&r) = *reinterpret_cast<const int4*>(p);
    return r;
}
template <typename real>
__device__ __forceinline__ void st_pack(real* p, const Pack<real>& r) {
    *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(&r);
}

// L2 policies of the sweep's measurement variants (HINT template flag; measured neutral to harmful, off by default):
// streamed triples / own rows / REDs evict_first, gathered rows evict_last for a fraction of the lines.
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <typename real>
__device__ __forceinline__ Pack<real> ldg_pack_hint(const real* p, uint64_t pol) {
    Pack<real> r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int ldg_stream(const int* p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ldg_stream(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// ---------------------------------------------------------------------------------------------
// shared-memory address of a generic pointer, and a 128-bit shared-memory load
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
template <typename real>
__device__ __forceinline__ Pack<real> lds_pack(uint32_t addr) {
    Pack<real> r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
    return r;
}

// per-thread asynchronous 16-byte copy global -> shared (LDGSTS, L2-only caching) and its group fences:
// the landing zone of the gathered-row ring in sweep_rows_kernel
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// vector reduction into global memory: RED.E.ADD.F32x4 (sm_90+) for float, 2x RED.E.ADD.F64 for double
__device__ __forceinline__ void red_add_pack(float* p, const Pack<float>& r) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]),
                 "f"(r.v[2]), "f"(r.v[3])
                 : "memory");
}
__device__ __forceinline__ void red_add_pack(double* p, const Pack<double>& r) {
    atomicAdd(p, r.v[0]);
    atomicAdd(p + 1, r.v[1]);
}

// ---------------------------------------------------------------------------------------------
// lane groups: LPG consecutive lanes of a warp own one factor row; lane gl holds packs
// gl, gl+LPG, ... (VPL of them) so that one load instruction of the group covers LPG*16 contiguous bytes.
// ---------------------------------------------------------------------------------------------
template <int LPG>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (LPG == 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << LPG) - 1u) << (lane & ~(unsigned)(LPG - 1));
    }
}
template <int LPG, typename T>
__device__ __forceinline__ T group_sum(T x, unsigned mask) {
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o, LPG);
    return x;
}
template <int LPG, typename T>
__device__ __forceinline__ T group_max(T x, unsigned mask) {
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) {
        T y = __shfl_xor_sync(mask, x, o, LPG);
        x = x > y ? x : y;
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rlog(float x) { return logf(x); }
__device__ __forceinline__ double rlog(double x) { return log(x); }
__device__ __forceinline__ float rexp(float x) { return expf(x); }
__device__ __forceinline__ double rexp(double x) { return exp(x); }
// count / normaliser: 2-ulp MUFU division is ample for float (parity gate 1e-5), IEEE for double
__device__ __forceinline__ float rdiv_fast(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double rdiv_fast(double a, double b) { return a / b; }

// count / normaliser with a bare MUFU.RCP (1 ulp) and one multiply: the sweep spends its issue slots elsewhere,
// and the normaliser is a sum of products of numbers in (0, 1] with at least one term equal to the
// product of the two row maxima's neighbours -- far from the denormal range __fdividef guards against
__device__ __forceinline__ float rdiv_rcp(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}
__device__ __forceinline__ double rdiv_rcp(double a, double b) { return a / b; }

// digamma for x > 0.  The reference calls scipy.special.cython_special.psi (pxi:5, call sites
// pxi:570/588/685/717), i.e. cephes `psi`: upward recurrence psi(x) = psi(x+1) - 1/x until x >= 10,
// then the asymptotic series log(x) - 1/(2x) - sum_n B_2n / (2n x^2n).  Same scheme here (7 Bernoulli
// terms in double).  Validated against scipy on a dense grid in
// tests/test_gpu_parity.py::test_digamma_vs_scipy.
// The recurrence's sum of reciprocals is carried as ONE fraction (num/den += 1/x  <=>  num = num*x + den, den *= x):
// one fp64 division per call instead of up to ten (shapes near the prior need all ten steps; the fp64 row update
// was spending most of its instructions there).  3e-15 vs scipy on the test grid (the per-step divisions: 1.5e-15).
__device__ __forceinline__ void digamma_parts(double x, double& z, double& tail) {
    double num = 0.0, den = 1.0;
    while (x < 10.0) {
        num = fma(num, x, den);
        den *= x;
        x += 1.0;
    }
    const double r = 1.0 / x;
    const double r2 = r * r;
    double p = 8.33333333333333333333e-2;               //  1/12   (x^-14)
    p = fma(p, r2, -2.10927960927960927961e-2);         // -691/32760
    p = fma(p, r2, 7.57575757575757575758e-3);          //  1/132
    p = fma(p, r2, -4.16666666666666666667e-3);         // -1/240
    p = fma(p, r2, 3.96825396825396825397e-3);          //  1/252
    p = fma(p, r2, -8.33333333333333333333e-3);         // -1/120
    p = fma(p, r2, 8.33333333333333333333e-2);          //  1/12   (x^-2)
    z = x;
    tail = (-0.5 * r - r2 * p) - num / den;             // psi(x) = log(z) + tail
}
__device__ __forceinline__ double digamma(double x) {
    double z, tail;
    digamma_parts(x, z, tail);
    return log(z) + tail;
}
// float: psi(z) = log z - 1/(2z) - t*P3(t), t = 1/z^2, valid to 1 ulp for z >= 2 (P3 = degree-3
// least-squares fit of the asymptotic tail on t in (0, 1/4], fitted against scipy in double);
// x < 2 is shifted by two recurrence steps folded into one quotient:
// psi(x) = psi(x+2) - (2x+1)/(x(x+1)).  Max error vs scipy.special.psi on [1e-3, 1e7]: 4.2e-7
// (absolute where |psi| < 1, relative elsewhere) -- tests/test_gpu_parity.py::test_digamma_vs_scipy.
__device__ __forceinline__ void digamma_parts(float x, float& z, float& tail) {
    const bool small = x < 2.0f;
    z = small ? x + 2.0f : x;
    const float r = __fdividef(1.0f, z);
    const float t = r * r;
    float p = -2.1589174882362706e-3f;
    p = fmaf(p, t, 3.7109495907488447e-3f);
    p = fmaf(p, t, -8.320496848651517e-3f);
    p = fmaf(p, t, 8.333318236376897e-2f);
    const float corr = small ? __fdividef(fmaf(2.0f, x, 1.0f), x * (x + 1.0f)) : 0.0f;
    tail = fmaf(-0.5f, r, -t * p) - corr;  // psi(x) = log(z) + tail
}
__device__ __forceinline__ float digamma(float x) {
    float z, tail;
    digamma_parts(x, z, tail);
    return logf(z) + tail;
}

// E[log] of a Gamma(shape, rate) variable: psi(shape) - log(rate)  (the summand of pxi:570).
// float: the two logarithms are merged into one, log(z / rate).
__device__ __forceinline__ float elog(float shape, float rate) {
    float z, tail;
    digamma_parts(shape, z, tail);
    return logf(__fdividef(z, rate)) + tail;
}
// double: likewise one logarithm, log(z / rate) (z >= 10 after the recurrence; one more rounding of 1 ulp)
__device__ __forceinline__ double elog(double shape, double rate) {
    double z, tail;
    digamma_parts(shape, z, tail);
    return log(z / rate) + tail;
}

// expectation shape / rate: IEEE in double; 2-ulp MUFU quotient in float
__device__ __forceinline__ float rratio(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double rratio(double a, double b) { return a / b; }

}  // namespace hpf
