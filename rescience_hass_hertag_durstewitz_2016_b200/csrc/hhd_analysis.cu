/*
 * Spike-train analyses on the device: the sufficient statistics of the reference's post-processing that is
 * quadratic in the number of neurons or linear in the number of spikes.
 *
 *   hhd_an_binned_trains   binary spike trains, one bit per bin
 *                          codes/AuxiliarEquations.py:537-548 (set_binary_vector), :422-437 (pop_rate),
 *                          codes_revision/AnalysesTools.py:176-196, 205-221 (get_binned_spiking_trains)
 *   hhd_an_lagged_counts   S_xy(i, j, L) = sum_t x_i[t] x_j[t + L] for every ordered pair and every lag, plus the window
 *                          sums — all a Pearson coefficient of two binary vectors needs
 *                          codes/AuxiliarEquations.py:550-595 (p_vector, p_joint_vector, Pearson), :622-690 (groups),
 *                          codes_revision/AnalysesTools.py:227-262 (ISIcorrelations: N^2 pandas Series.corr calls per lag,
 *                          the 81-minute step of Tests.py:67)
 *   hhd_an_isi_stats       per-neuron ISI mean and CV     codes_revision/AnalysesTools.py:421-429 (ISIstats)
 *
 * Integer results are exact, so the coefficients computed from them on the host (analysis.py) follow the reference's
 * formulas operand by operand.  Bit-packed trains: a pair and lag costs n_bins/64 AND + POPC; 1003 neurons x 15 000 bins
 * x 1 lag is 2.4e8 word operations — microseconds here, minutes in a pandas double loop.
 * No CPU fallback: every entry point fails with HHD_E_CUDA when no device is visible.
 */
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/hhd.h"

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <typename T> T* as() { return (T*)p; }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

#define AN_CU(call)                                                     \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      fprintf(stderr, "[hhd analysis] %s: %s\n", #call, cudaGetErrorString(e_)); \
      return HHD_E_CUDA;                                                \
    }                                                                   \
  } while (0)

/* numpy's floor division of two doubles (npy_divmod): the quotient of the exact remainder, not floor(a / b) */
__device__ __forceinline__ double np_floor_divide(double a, double b) {
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
  if (div != 0.0) {
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
  }
  return copysign(0.0, a / b);
}

/* one thread per spike: row of its neuron (or none), bin by the reference's rule, set the bit */
__global__ void k_an_bin(long long n_spikes, const int* neuron, const double* t_ms, const int* row_of, int n_map,
                         double t0, double t1, double bin_ms, int rule, long long n_bins, int words, unsigned long long* bits) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_spikes) return;
  const int j = neuron[q];
  if (j < 0 || j >= n_map) return;
  const int row = row_of[j];
  if (row < 0) return;
  const double t = t_ms[q];                                      /* SpikeMonitor.t / ms, the caller's floats */
  if (!(t >= t0 && t < t1)) return;
  double b;
  if (rule == HHD_BIN_TRUNC) b = trunc((t - t0) / bin_ms);      /* ((spike_times - t0) / dt).astype(int) */
  else b = np_floor_divide(t - t0, bin_ms);                      /* ((train - t0) // delta_t).astype(int) */
  const long long bin = (long long)b;
  if (bin < 0 || bin >= n_bins) return;                          /* numpy would raise; cannot happen for t < t1 */
  atomicOr(&bits[(size_t)row * words + (bin >> 6)], 1ull << (bin & 63));
}

__global__ void k_an_unpack(int n_sel, long long n_bins, int words, const unsigned long long* bits, unsigned char* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n_sel * n_bins) return;
  const long long row = q / n_bins, bin = q % n_bins;
  out[q] = (unsigned char)((bits[(size_t)row * words + (bin >> 6)] >> (bin & 63)) & 1ull);
}

/* population count per bin (pop_rate's column sum) */
__global__ void k_an_bin_counts(int n_sel, long long n_bins, int words, const unsigned long long* bits, int* out) {
  const long long bin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= n_bins) return;
  int c = 0;
  for (int r = 0; r < n_sel; ++r) c += (int)((bits[(size_t)r * words + (bin >> 6)] >> (bin & 63)) & 1ull);
  out[bin] = c;
}

/* bits [64 w + L, 64 w + L + 64) of a row, zero beyond n_bins (rows are zero-padded to whole words + one spare word) */
__device__ __forceinline__ unsigned long long shifted_word(const unsigned long long* row, int w, int L, int words) {
  const int ws = w + (L >> 6), sh = L & 63;
  const unsigned long long lo = ws < words ? row[ws] : 0ull;
  if (!sh) return lo;
  const unsigned long long hi = ws + 1 < words ? row[ws + 1] : 0ull;
  return (lo >> sh) | (hi << (64 - sh));
}

#define AN_TILE 32
#define AN_CHUNK 64   /* words of a row staged per pass */
/* S_xy(i, j, L) = sum_t x_i[t] x_j[t + L]: a 32 x 32 tile of pairs per CTA and lag; the i rows are staged as they are, the j
 * rows already shifted by L, chunk by chunk; thread (ti, tj) of the 1024 accumulates its pair. */
__global__ void __launch_bounds__(AN_TILE * AN_TILE) k_an_sxy(int n_sel, int words, const unsigned long long* bits, int n_lags,
                                                               const int* lags, int* out) {
  __shared__ unsigned long long xi[AN_TILE][AN_CHUNK + 1], xj[AN_TILE][AN_CHUNK + 1];
  const int lag_idx = blockIdx.z, L = lags[lag_idx];
  const int i0 = blockIdx.y * AN_TILE, j0 = blockIdx.x * AN_TILE;
  const int ti = threadIdx.x / AN_TILE, tj = threadIdx.x % AN_TILE;
  int acc = 0;
  for (int w0 = 0; w0 < words; w0 += AN_CHUNK) {
    for (int q = threadIdx.x; q < AN_TILE * AN_CHUNK; q += blockDim.x) {
      const int r = q / AN_CHUNK, w = w0 + q % AN_CHUNK;
      unsigned long long a = 0ull, b = 0ull;
      if (w < words) {
        if (i0 + r < n_sel) a = bits[(size_t)(i0 + r) * words + w];
        if (j0 + r < n_sel) b = shifted_word(bits + (size_t)(j0 + r) * words, w, L, words);
      }
      xi[r][q % AN_CHUNK] = a;
      xj[r][q % AN_CHUNK] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int w = 0; w < AN_CHUNK; ++w) acc += __popcll(xi[ti][w] & xj[tj][w]);
    __syncthreads();
  }
  if (i0 + ti < n_sel && j0 + tj < n_sel) out[((size_t)lag_idx * n_sel + (i0 + ti)) * n_sel + (j0 + tj)] = acc;
}

/* head(i, L) = sum_{t < T - L} x_i[t],  tail(i, L) = sum_{t >= L} x_i[t] */
__global__ void k_an_windows(int n_sel, long long n_bins, int words, const unsigned long long* bits, int n_lags, const int* lags,
                             int* head, int* tail) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_sel * n_lags) return;
  const int li = q / n_sel, r = q % n_sel, L = lags[li];
  const unsigned long long* row = bits + (size_t)r * words;
  int h = 0, t = 0;
  const long long keep = n_bins - L;                                   /* head: bins [0, keep) */
  for (int w = 0; w < words; ++w) {
    const long long b0 = (long long)w * 64;
    unsigned long long x = row[w];
    unsigned long long hm = ~0ull, tm = ~0ull;
    if (keep <= b0) hm = 0ull; else if (keep < b0 + 64) hm = (1ull << (keep - b0)) - 1ull;
    if ((long long)L >= b0 + 64) tm = 0ull; else if ((long long)L > b0) tm = ~((1ull << (L - b0)) - 1ull);
    h += __popcll(x & hm);
    t += __popcll(x & tm);
  }
  head[q] = h;
  tail[q] = t;
}

/* per selected neuron: ISI mean and CV (np.std, ddof 0: sqrt(mean((x - mean)^2))) of its spikes in [t0, t1).  The spikes
 * were sorted by neuron (stable, so each neuron's spikes are in time order); the thread finds its neuron's segment by
 * binary search and walks it twice (mean, then variance — the two passes numpy makes). */
__global__ void k_an_isi(long long n_spikes, const int* sorted_neuron, const double* sorted_t, int n_sel,
                         const int* sel, int has_window, double t0, double t1, double* out_mean, double* out_cv, int* out_count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_sel) return;
  const int j = sel[q];
  long long lo = 0, hi = n_spikes;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (sorted_neuron[mid] < j) lo = mid + 1; else hi = mid; }
  const long long s0 = lo;
  hi = n_spikes;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (sorted_neuron[mid] <= j) lo = mid + 1; else hi = mid; }
  const long long s1 = lo;
  int n = 0;
  double sum = 0.0, prev = 0.0;
  bool have = false;
  for (long long r = s0; r < s1; ++r) {
    const double t = sorted_t[r];
    if (has_window && !(t >= t0 && t < t1)) continue;
    if (have) { sum += t - prev; ++n; }
    prev = t;
    have = true;
  }
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  out_count[q] = n;
  if (n == 0) { out_mean[q] = nan; out_cv[q] = nan; return; }       /* np.mean([]) */
  const double mean = sum / (double)n;
  double ss = 0.0;
  have = false;
  for (long long r = s0; r < s1; ++r) {
    const double t = sorted_t[r];
    if (has_window && !(t >= t0 && t < t1)) continue;
    if (have) { const double d = (t - prev) - mean; ss += d * d; }
    prev = t;
    have = true;
  }
  out_mean[q] = mean;
  out_cv[q] = sqrt(ss / (double)n) / mean;
}

}  // namespace

static int an_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HHD_E_CUDA;
  return cudaSetDevice(device) == cudaSuccess ? HHD_OK : HHD_E_CUDA;
}

extern "C" int hhd_an_words(int64_t n_bins) { return (int)((n_bins + 63) / 64 + 1); }

extern "C" int hhd_an_binned_trains(int32_t device, int64_t n_spikes, const int32_t* neuron, const double* t_ms,
                                    int32_t n_sel, const int32_t* sel, double t0_ms, double t1_ms, double bin_ms, int32_t rule,
                                    int64_t n_bins, uint64_t* out_bits, uint8_t* out_bins, int32_t* out_bin_counts) {
  if (n_spikes < 0 || n_sel <= 0 || !sel || n_bins <= 0 || !(bin_ms > 0.0) || (n_spikes && (!neuron || !t_ms))) return HHD_E_ARG;
  if (rule != HHD_BIN_TRUNC && rule != HHD_BIN_FLOORDIV) return HHD_E_ARG;
  int rc = an_device(device);
  if (rc) return rc;
  const int words = hhd_an_words(n_bins);
  int n_map = 0;
  for (int q = 0; q < n_sel; ++q) { if (sel[q] < 0) return HHD_E_ARG; if (sel[q] + 1 > n_map) n_map = sel[q] + 1; }
  std::vector<int> row_of((size_t)n_map, -1);
  for (int q = 0; q < n_sel; ++q) {
    if (row_of[(size_t)sel[q]] >= 0) return HHD_E_ARG;               /* a neuron may be selected once */
    row_of[(size_t)sel[q]] = q;
  }
  DevBuf d_n, d_s, d_map, d_bits;
  AN_CU(d_n.alloc((size_t)n_spikes * 4));
  AN_CU(d_s.alloc((size_t)n_spikes * 8));
  AN_CU(d_map.alloc((size_t)n_map * 4));
  AN_CU(d_bits.alloc((size_t)n_sel * words * 8));
  if (n_spikes) {
    AN_CU(cudaMemcpy(d_n.p, neuron, (size_t)n_spikes * 4, cudaMemcpyHostToDevice));
    AN_CU(cudaMemcpy(d_s.p, t_ms, (size_t)n_spikes * 8, cudaMemcpyHostToDevice));
  }
  AN_CU(cudaMemcpy(d_map.p, row_of.data(), (size_t)n_map * 4, cudaMemcpyHostToDevice));
  AN_CU(cudaMemset(d_bits.p, 0, (size_t)n_sel * words * 8));
  if (n_spikes)
    k_an_bin<<<(unsigned)((n_spikes + 255) / 256), 256>>>(n_spikes, d_n.as<int>(), d_s.as<double>(), d_map.as<int>(), n_map,
                                                            t0_ms, t1_ms, bin_ms, rule, n_bins, words, d_bits.as<unsigned long long>());
  AN_CU(cudaGetLastError());
  if (out_bits) AN_CU(cudaMemcpy(out_bits, d_bits.p, (size_t)n_sel * words * 8, cudaMemcpyDeviceToHost));
  if (out_bins) {
    DevBuf d_out;
    const long long tot = (long long)n_sel * n_bins;
    AN_CU(d_out.alloc((size_t)tot));
    k_an_unpack<<<(unsigned)((tot + 255) / 256), 256>>>(n_sel, n_bins, words, d_bits.as<unsigned long long>(), d_out.as<unsigned char>());
    AN_CU(cudaGetLastError());
    AN_CU(cudaMemcpy(out_bins, d_out.p, (size_t)tot, cudaMemcpyDeviceToHost));
  }
  if (out_bin_counts) {
    DevBuf d_cnt;
    AN_CU(d_cnt.alloc((size_t)n_bins * 4));
    k_an_bin_counts<<<(unsigned)((n_bins + 255) / 256), 256>>>(n_sel, n_bins, words, d_bits.as<unsigned long long>(), d_cnt.as<int>());
    AN_CU(cudaGetLastError());
    AN_CU(cudaMemcpy(out_bin_counts, d_cnt.p, (size_t)n_bins * 4, cudaMemcpyDeviceToHost));
  }
  AN_CU(cudaDeviceSynchronize());
  return HHD_OK;
}

extern "C" int hhd_an_lagged_counts(int32_t device, int32_t n_sel, int64_t n_bins, const uint64_t* bits, int32_t n_lags,
                                    const int32_t* lags, int32_t* out_sxy, int32_t* out_head, int32_t* out_tail) {
  if (n_sel <= 0 || n_bins <= 0 || !bits || n_lags <= 0 || !lags || !out_sxy || !out_head || !out_tail) return HHD_E_ARG;
  for (int q = 0; q < n_lags; ++q) if (lags[q] < 0 || lags[q] >= n_bins) return HHD_E_ARG;
  int rc = an_device(device);
  if (rc) return rc;
  const int words = hhd_an_words(n_bins);
  DevBuf d_bits, d_lags, d_sxy, d_head, d_tail;
  AN_CU(d_bits.alloc((size_t)n_sel * words * 8));
  AN_CU(d_lags.alloc((size_t)n_lags * 4));
  AN_CU(d_sxy.alloc((size_t)n_lags * n_sel * n_sel * 4));
  AN_CU(d_head.alloc((size_t)n_lags * n_sel * 4));
  AN_CU(d_tail.alloc((size_t)n_lags * n_sel * 4));
  AN_CU(cudaMemcpy(d_bits.p, bits, (size_t)n_sel * words * 8, cudaMemcpyHostToDevice));
  AN_CU(cudaMemcpy(d_lags.p, lags, (size_t)n_lags * 4, cudaMemcpyHostToDevice));
  const int tiles = (n_sel + AN_TILE - 1) / AN_TILE;
  for (int l0 = 0; l0 < n_lags; l0 += 32768) {                      /* gridDim.z limit */
    const int nl = n_lags - l0 < 32768 ? n_lags - l0 : 32768;
    k_an_sxy<<<dim3((unsigned)tiles, (unsigned)tiles, (unsigned)nl), AN_TILE * AN_TILE>>>(
        n_sel, words, d_bits.as<unsigned long long>(), nl, d_lags.as<int>() + l0, d_sxy.as<int>() + (size_t)l0 * n_sel * n_sel);
  }
  AN_CU(cudaGetLastError());
  k_an_windows<<<(unsigned)((n_sel * n_lags + 255) / 256), 256>>>(n_sel, n_bins, words, d_bits.as<unsigned long long>(), n_lags,
                                                                   d_lags.as<int>(), d_head.as<int>(), d_tail.as<int>());
  AN_CU(cudaGetLastError());
  AN_CU(cudaMemcpy(out_sxy, d_sxy.p, (size_t)n_lags * n_sel * n_sel * 4, cudaMemcpyDeviceToHost));
  AN_CU(cudaMemcpy(out_head, d_head.p, (size_t)n_lags * n_sel * 4, cudaMemcpyDeviceToHost));
  AN_CU(cudaMemcpy(out_tail, d_tail.p, (size_t)n_lags * n_sel * 4, cudaMemcpyDeviceToHost));
  return HHD_OK;
}

extern "C" int hhd_an_isi_stats(int32_t device, int64_t n_spikes, const int32_t* neuron, const double* t_ms,
                                int32_t n_sel, const int32_t* sel, int32_t has_window, double t0_ms, double t1_ms,
                                double* out_mean, double* out_cv, int32_t* out_count) {
  if (n_spikes < 0 || n_sel <= 0 || !sel || !out_mean || !out_cv || !out_count || (n_spikes && (!neuron || !t_ms))) return HHD_E_ARG;
  int rc = an_device(device);
  if (rc) return rc;
  DevBuf d_n, d_s, d_n2, d_s2, d_sel, d_mean, d_cv, d_cnt, d_tmp;
  AN_CU(d_n.alloc((size_t)n_spikes * 4));
  AN_CU(d_s.alloc((size_t)n_spikes * 8));
  AN_CU(d_n2.alloc((size_t)n_spikes * 4));
  AN_CU(d_s2.alloc((size_t)n_spikes * 8));
  AN_CU(d_sel.alloc((size_t)n_sel * 4));
  AN_CU(d_mean.alloc((size_t)n_sel * 8));
  AN_CU(d_cv.alloc((size_t)n_sel * 8));
  AN_CU(d_cnt.alloc((size_t)n_sel * 4));
  if (n_spikes) {
    AN_CU(cudaMemcpy(d_n.p, neuron, (size_t)n_spikes * 4, cudaMemcpyHostToDevice));
    AN_CU(cudaMemcpy(d_s.p, t_ms, (size_t)n_spikes * 8, cudaMemcpyHostToDevice));
    /* stable LSD radix sort by neuron: the input is in time order, so every neuron's spikes stay in time order */
    size_t tmp_bytes = 0;
    AN_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_n.as<int>(), d_n2.as<int>(), d_s.as<double>(), d_s2.as<double>(), (long long)n_spikes));
    AN_CU(d_tmp.alloc(tmp_bytes));
    AN_CU(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_n.as<int>(), d_n2.as<int>(), d_s.as<double>(), d_s2.as<double>(), (long long)n_spikes));
  }
  AN_CU(cudaMemcpy(d_sel.p, sel, (size_t)n_sel * 4, cudaMemcpyHostToDevice));
  k_an_isi<<<(unsigned)((n_sel + 127) / 128), 128>>>(n_spikes, d_n2.as<int>(), d_s2.as<double>(), n_sel, d_sel.as<int>(), has_window,
                                                      t0_ms, t1_ms, d_mean.as<double>(), d_cv.as<double>(), d_cnt.as<int>());
  AN_CU(cudaGetLastError());
  AN_CU(cudaMemcpy(out_mean, d_mean.p, (size_t)n_sel * 8, cudaMemcpyDeviceToHost));
  AN_CU(cudaMemcpy(out_cv, d_cv.p, (size_t)n_sel * 8, cudaMemcpyDeviceToHost));
  AN_CU(cudaMemcpy(out_count, d_cnt.p, (size_t)n_sel * 4, cudaMemcpyDeviceToHost));
  return HHD_OK;
}
