// Layout experiment for the state-update kernel: stream 21 doubles in / 8 doubles out per neuron, no arithmetic.
//   soa    : 21 separate arrays (current layout), one thread per neuron, grid covers all neurons
//   aosoa  : per 128-neuron tile one contiguous block [21][128] in, [8][128] out
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu && ./membench
#include <cstdio>
#include <cuda_runtime.h>
#define NF 21
#define NW 8
#define T 128
__global__ void k_soa(int N, double* const* in, double* const* out) {
  int i = blockIdx.x * T + threadIdx.x;
  if (i >= N) return;
  double v[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) v[f] = in[f][i];
  double s = 0;
#pragma unroll
  for (int f = NW; f < NF; ++f) s += v[f];
#pragma unroll
  for (int f = 0; f < NW; ++f) out[f][i] = v[f] + s * 1e-300;
}
struct Ptrs { double* in[NF]; double* out[NW]; };
__global__ void k_soa_p(int N, const __grid_constant__ Ptrs P) {
  int i = blockIdx.x * T + threadIdx.x;
  if (i >= N) return;
  double v[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) v[f] = P.in[f][i];
  double s = 0;
#pragma unroll
  for (int f = NW; f < NF; ++f) s += v[f];
#pragma unroll
  for (int f = 0; f < NW; ++f) P.out[f][i] = v[f] + s * 1e-300;
}
__global__ void k_aosoa(int ntiles, const double* in, double* out) {
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const double* b = in + (size_t)tile * NF * T;
    double v[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) v[f] = b[f * T + threadIdx.x];
    double s = 0;
#pragma unroll
    for (int f = NW; f < NF; ++f) s += v[f];
    double* o = out + (size_t)tile * NW * T;
#pragma unroll
    for (int f = 0; f < NW; ++f) o[f * T + threadIdx.x] = v[f] + s * 1e-300;
  }
}
// split layout: state tile [8][128] read+written in place, params tile [13][128] read only
__global__ void k_aosoa2(int ntiles, double* st, const double* pr) {
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    double* b = st + (size_t)tile * NW * T;
    const double* q = pr + (size_t)tile * (NF - NW) * T;
    double v[NF];
#pragma unroll
    for (int f = 0; f < NW; ++f) v[f] = b[f * T + threadIdx.x];
#pragma unroll
    for (int f = NW; f < NF; ++f) v[f] = q[(f - NW) * T + threadIdx.x];
    double s = 0;
#pragma unroll
    for (int f = NW; f < NF; ++f) s += v[f];
#pragma unroll
    for (int f = 0; f < NW; ++f) b[f * T + threadIdx.x] = v[f] + s * 1e-300;
  }
}
int main() {
  const int N = 1003000, ntiles = (N + T - 1) / T;
  const size_t NP = (size_t)ntiles * T;
  Ptrs P; double** din; double** dout;
  double* hin[NF]; double* hout[NW];
  for (int f = 0; f < NF; ++f) { cudaMalloc(&hin[f], NP * 8); cudaMemset(hin[f], 0, NP * 8); P.in[f] = hin[f]; }
  for (int f = 0; f < NW; ++f) { hout[f] = hin[f]; P.out[f] = hin[f]; }   // in place, like the engine
  cudaMalloc(&din, NF * 8); cudaMalloc(&dout, NW * 8);
  cudaMemcpy(din, hin, NF * 8, cudaMemcpyHostToDevice); cudaMemcpy(dout, hout, NW * 8, cudaMemcpyHostToDevice);
  double *a_in, *a_out, *st, *pr;
  cudaMalloc(&a_in, NP * NF * 8); cudaMalloc(&a_out, NP * NW * 8); cudaMalloc(&st, NP * NW * 8); cudaMalloc(&pr, NP * (NF - NW) * 8);
  cudaMemset(a_in, 0, NP * NF * 8); cudaMemset(st, 0, NP * NW * 8); cudaMemset(pr, 0, NP * (NF - NW) * 8);
  double* flush; cudaMalloc(&flush, 512ull << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)N * (NF + NW) * 8;
  auto run = [&](const char* name, auto launch) {
    float best = 1e9, sum = 0;
    for (int it = 0; it < 12; ++it) {
      cudaMemsetAsync(flush, it, 512ull << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    printf("%-28s best %.1f us (%.0f GB/s)  mean %.1f us   err=%s\n", name, best * 1e3, bytes / best / 1e6, sum / 10 * 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  run("soa ptr-array 7836 CTAs", [&] { k_soa<<<ntiles, T>>>(N, din, dout); });
  run("soa param 7836 CTAs", [&] { k_soa_p<<<ntiles, T>>>(N, P); });
  for (int c : {2, 4, 6, 8, 16}) {
    char nm[64];
    snprintf(nm, 64, "aosoa persistent %d CTA/SM", c);
    run(nm, [&] { k_aosoa<<<148 * c, T>>>(ntiles, a_in, a_out); });
    snprintf(nm, 64, "aosoa2 in-place %d CTA/SM", c);
    run(nm, [&] { k_aosoa2<<<148 * c, T>>>(ntiles, st, pr); });
  }
  run("aosoa one tile per CTA", [&] { k_aosoa<<<ntiles, T>>>(ntiles, a_in, a_out); });
  run("aosoa2 one tile per CTA", [&] { k_aosoa2<<<ntiles, T>>>(ntiles, st, pr); });
  return 0;
}
