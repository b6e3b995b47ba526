// compile-only harness for tools/sass_count.py: instantiates the experimental kernel next to the shipped one
#include "../hpf_sweep.cuh"
#include "sweep_v5.cuh"
template __global__ void hpf::sweep_major_v3_kernel<float, 8, 2, 2, 0, 256>(const int*, const int*, const float*, long long, int,
                                                                           const float*, const float*, float*, int, int);
template __global__ void hpf::sweep_major_v5_kernel<float, 8, 2, 2, 0, 256, false>(const int*, const int*, const float*, long long,
                                                                                  int, const float*, const float*, float*, int, int);
template __global__ void hpf::sweep_major_v5_kernel<float, 8, 2, 2, 0, 256, true>(const int*, const int*, const float*, long long,
                                                                                 int, const float*, const float*, float*, int, int);
template __global__ void hpf::sweep_major_v5_kernel<float, 4, 4, 3, 0, 128, true>(const int*, const int*, const float*, long long,
                                                                                 int, const float*, const float*, float*, int, int);
template __global__ void hpf::sweep_major_v3_kernel<float, 4, 4, 3, 0, 128>(const int*, const int*, const float*, long long, int,
                                                                           const float*, const float*, float*, int, int);
