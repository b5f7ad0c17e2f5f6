// cc_summ.cu - posterior summaries of the recorded draws ON THE DEVICE: what the consumers of the fit object
// compute in R on the host (R/summarize_plot.R:46-58 coda::gelman.diag, :68-153 min / quantiles / mean / max / coda::effectiveSize /
// coda::geweke.diag per parameter, :162-264 the overall IFR), here next to the draws in HBM - at the cfg4 scale the gathered
// cold-rung draws are 4096 chains x iterations x 48 parameters.
//
//   summ_transpose_kernel       draws[chain][iteration][param] (any chain / iteration stride) -> series[param][chain][iteration]
//   summ_series_kernel          one CTA per series: min, max, mean, variance, type-7 quantiles by radix select (no sort),
//                               spectral density at zero by coda's AR fit (Yule-Walker, order by AIC) -> ESS, Geweke z
//   summ_moments_kernel         one warp per (chain, param): mean and unbiased variance (two passes)
//   summ_gelman_kernel          one CTA per parameter over the per-chain moments: drjacoby's R-hat and coda's PSRF point estimate
//                               with every ingredient of its upper limit except the F quantile
//   summ_overall_ifr_kernel     one thread per draw: sum_a ifr_a w_a with population or attack-rate x population weights
//
// The host formulas these restate are in covidcurve_b200/summaries.py (the R-side behaviour restated in numpy; coda itself is not
// available in this image, so parity of ESS / Geweke / gelman.diag with coda is pinned only through that restatement).
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/covidcurve_b200.h"
#include "cc_internal.h"

namespace {

constexpr int kSummThreads = 256;
constexpr int kMaxOrder = 96;  // 10 log10(n) <= 96 for n < 3.9e9

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to NV values per thread; result valid in every thread
template <int NV>
__device__ inline void block_sum_n(double (&v)[NV], double* red /* [NV * 8] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; q++) v[q] = warp_sum_d(v[q]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NV; q++) red[q * 8 + w] = v[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double s = 0.0;
    for (int j = 0; j < nw; j++) s += red[q * 8 + j];
    v[q] = s;
  }
}

// ---------------------------------------------------------------------------------------------- transpose
// in : element (c, i, p) at draws[c * sc + (first + i) * si + p],  c < C, i < n, p < P
// out: series[p][c * n + i]
__global__ void summ_transpose_kernel(const double* __restrict__ draws, long long C, long long n, int P, long long sc, long long si,
                                      long long first, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const long long R = C * n;
  const long long r0 = (long long)blockIdx.x * 32;
  const int p0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const long long r = r0 + j;
    const int p = p0 + tx;
    if (r < R && p < P) {
      const long long c = r / n, i = r - c * n;
      tile[j][tx] = draws[c * sc + (first + i) * si + p];
    }
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j;
    const long long r = r0 + tx;
    if (r < R && p < P) out[(long long)p * R + r] = tile[tx][j];
  }
}

// ---------------------------------------------------------------------------------------------- spectrum at zero
// coda::spectrum0.ar of x[0..n): AR(p) by Yule-Walker (Levinson-Durbin), order <= floor(10 log10 n) chosen by AIC, innovations
// variance scaled by n / (n - (p + 1)) like R's ar.yw, spectrum = v / (1 - sum phi)^2.  All threads call; result in all threads.
struct SummSmem {
  double red[8 * 8];
  double acov[kMaxOrder + 1];
  double out[4];
  unsigned int hist[6][256];
  unsigned long long prefix[6];
  long long rank[6];
};

__device__ double spectrum0_ar_block(const double* __restrict__ x, long long n, SummSmem& sm, double* mean_out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double s[1] = {0.0};
  for (long long i = tid; i < n; i += nt) s[0] += x[i];
  block_sum_n<1>(s, sm.red);
  const double mean = (n > 0) ? s[0] / (double)n : 0.0;
  if (mean_out) *mean_out = mean;
  if (n < 3) return 0.0;
  int order_max = (int)floor(10.0 * log10((double)n));
  if ((long long)order_max > n - 1) order_max = (int)(n - 1);
  if (order_max > kMaxOrder) order_max = kMaxOrder;
  // autocovariances, eight lags per pass
  for (int k0 = 0; k0 <= order_max; k0 += 8) {
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = tid; i < n; i += nt) {
      const double xi = x[i] - mean;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const long long j = i + k0 + q;
        if (j < n) a[q] = fma(xi, x[j] - mean, a[q]);
      }
    }
    block_sum_n<8>(a, sm.red);
    if (tid == 0)
      for (int q = 0; q < 8; q++)
        if (k0 + q <= order_max) sm.acov[k0 + q] = a[q] / (double)n;
  }
  __syncthreads();
  if (tid == 0) {
    double res = 0.0;
    const double* acov = sm.acov;
    if (acov[0] > 0.0) {  // np.var(x) == 0 -> 0
      double phi[kMaxOrder], best_phi_sum = 0.0, tmp[kMaxOrder];
      double v = acov[0];
      double best_aic = (double)n * log(v), best_v = v;
      int best_k = 0;
      for (int k = 1; k <= order_max; k++) {
        if (!(v > 0.0)) break;
        double dot = 0.0;
        for (int j = 0; j < k - 1; j++) dot += phi[j] * acov[k - 1 - j];
        const double kappa = (acov[k] - dot) / v;
        for (int j = 0; j < k - 1; j++) tmp[j] = phi[j] - kappa * phi[k - 2 - j];
        for (int j = 0; j < k - 1; j++) phi[j] = tmp[j];
        phi[k - 1] = kappa;
        v = v * (1.0 - kappa * kappa);
        if (!(v > 0.0)) break;
        const double aic = (double)n * log(v) + 2.0 * k;
        if (aic < best_aic) {
          best_aic = aic;
          best_v = v;
          best_k = k;
          double ps = 0.0;
          for (int j = 0; j < k; j++) ps += phi[j];
          best_phi_sum = ps;
        }
      }
      const double vv = best_v * (double)n / (double)(n - (best_k + 1));
      res = vv / ((1.0 - best_phi_sum) * (1.0 - best_phi_sum));
    }
    sm.out[0] = res;
  }
  __syncthreads();
  const double r = sm.out[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ unsigned long long sort_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// order statistics x_(rank[q]), q < 6, of x[0..n) by most-significant-digit radix select: eight passes of 8 bits, six
// histograms per pass (one per requested rank), nothing is moved.  Results in sm.prefix[q] as sortable keys.
__device__ void radix_select6(const double* __restrict__ x, long long n, SummSmem& sm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 6) sm.prefix[tid] = 0ull;
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    for (int j = tid; j < 6 * 256; j += nt) (&sm.hist[0][0])[j] = 0u;
    __syncthreads();
    unsigned long long pre[6];
#pragma unroll
    for (int q = 0; q < 6; q++) pre[q] = sm.prefix[q];
    const unsigned long long himask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
    for (long long i = tid; i < n; i += nt) {
      const unsigned long long k = sort_key(x[i]);
      const unsigned int dgt = (unsigned int)((k >> shift) & 0xffu);
#pragma unroll
      for (int q = 0; q < 6; q++)
        if ((k & himask) == pre[q]) atomicAdd(&sm.hist[q][dgt], 1u);
    }
    __syncthreads();
    if (tid < 6) {
      long long r = sm.rank[tid];
      int dsel = 255;
      for (int b = 0; b < 256; b++) {
        const long long c = (long long)sm.hist[tid][b];
        if (r < c) { dsel = b; break; }
        r -= c;
      }
      sm.rank[tid] = r;
      sm.prefix[tid] |= ((unsigned long long)dsel) << shift;
    }
    __syncthreads();
  }
}

// one CTA per series; out[series][CC_SUMM_FIELDS]
__global__ void __launch_bounds__(kSummThreads) summ_series_kernel(const double* __restrict__ series, long long len, long long n_series,
                                                                   double* __restrict__ out) {
  __shared__ SummSmem sm;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (long long sidx = blockIdx.x; sidx < n_series; sidx += gridDim.x) {
    const double* x = series + sidx * len;
    __syncthreads();
    // min / max
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (long long i = tid; i < len; i += nt) {
      const double v = x[i];
      mn = fmin(mn, v);
      mx = fmax(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) {
      sm.red[tid >> 5] = mn;
      sm.red[8 + (tid >> 5)] = mx;
    }
    __syncthreads();
    for (int j = 0; j < (nt >> 5); j++) {
      mn = fmin(mn, sm.red[j]);
      mx = fmax(mx, sm.red[8 + j]);
    }
    __syncthreads();
    // mean, variance, spectrum at zero of the whole series
    double mean;
    const double s0 = spectrum0_ar_block(x, len, sm, &mean);
    double ss[1] = {0.0};
    for (long long i = tid; i < len; i += nt) {
      const double dv = x[i] - mean;
      ss[0] = fma(dv, dv, ss[0]);
    }
    block_sum_n<1>(ss, sm.red);
    const double var = (len > 1) ? ss[0] / (double)(len - 1) : 0.0;
    const double ess = (s0 == 0.0) ? 0.0 : (double)len * var / s0;
    // Geweke: windows 1 .. ceil(1 + 0.1 (n - 1)) and floor(n - 0.5 (n - 1)) .. n (1-based, inclusive), as coda takes them
    long long end1 = (long long)ceil(1.0 + 0.1 * (double)(len - 1));
    long long start2 = (long long)floor((double)len - 0.5 * (double)(len - 1));
    if (end1 < 1) end1 = 1;
    if (start2 < 1) start2 = 1;
    if (end1 > len) end1 = len;
    double ma, mb;
    const double sa = spectrum0_ar_block(x, end1, sm, &ma);
    const double sb = spectrum0_ar_block(x + (start2 - 1), len - (start2 - 1), sm, &mb);
    const double z = (ma - mb) / sqrt(sa / (double)end1 + sb / (double)(len - (start2 - 1)));
    // type-7 quantiles 2.5 %, 50 %, 97.5 %
    const double qs[3] = {0.025, 0.5, 0.975};
    double hfrac[3];
    if (tid == 0) {
      for (int q = 0; q < 3; q++) {
        const double h = (double)(len - 1) * qs[q];
        long long lo = (long long)floor(h);
        if (lo > len - 1) lo = len - 1;
        const long long hi = (lo + 1 < len) ? lo + 1 : len - 1;
        sm.rank[2 * q] = lo;
        sm.rank[2 * q + 1] = hi;
      }
    }
    for (int q = 0; q < 3; q++) {
      const double h = (double)(len - 1) * qs[q];
      hfrac[q] = h - floor(h);
    }
    __syncthreads();
    radix_select6(x, len, sm);
    if (tid == 0) {
      double qv[3];
      for (int q = 0; q < 3; q++) {
        const double a = key_value(sm.prefix[2 * q]), b = key_value(sm.prefix[2 * q + 1]);
        const double t = hfrac[q], diff = b - a;
        qv[q] = (t >= 0.5) ? (b - diff * (1.0 - t)) : (a + diff * t);  // numpy's lerp; R's type 7 differs by an ulp at most
      }
      double* o = out + sidx * CC_SUMM_FIELDS;
      o[0] = mn; o[1] = qv[0]; o[2] = qv[1]; o[3] = mean; o[4] = qv[2]; o[5] = mx; o[6] = ess; o[7] = z;
      o[8] = 0.3989422804014326779399460599343819 * exp(-0.5 * z * z);  // GewekeP = dnorm(z), summarize_plot.R:151
      o[9] = var; o[10] = s0; o[11] = (double)len;
    }
  }
}

// ---------------------------------------------------------------------------------------------- per-chain moments
// series[p][c * n + i] -> mean[c * P + p], var[c * P + p]; one warp per (p, c)
__global__ void summ_moments_kernel(const double* __restrict__ series, long long C, long long n, int P, double* __restrict__ mean,
                                    double* __restrict__ var) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= C * P) return;
  const long long p = wid / C, c = wid - p * C;
  const double* x = series + (p * C + c) * n;
  double s = 0.0;
  for (long long i = lane; i < n; i += 32) s += x[i];
  const double mu = warp_sum_d(s) / (double)n;
  double q = 0.0;
  for (long long i = lane; i < n; i += 32) {
    const double dv = x[i] - mu;
    q = fma(dv, dv, q);
  }
  q = warp_sum_d(q);
  if (lane == 0) {
    mean[c * P + p] = mu;
    var[c * P + p] = (n > 1) ? q / (double)(n - 1) : 0.0;
  }
}

// one CTA per parameter: out[p][CC_GELMAN_FIELDS] from the per-chain means and variances of n draws each
__global__ void __launch_bounds__(kSummThreads) summ_gelman_kernel(const double* __restrict__ mean, const double* __restrict__ var,
                                                                   long long C, long long n, int P, double* __restrict__ out) {
  __shared__ double red[8 * 8];
  const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  double a[2] = {0.0, 0.0};
  for (long long c = tid; c < C; c += nt) {
    a[0] += mean[c * P + p];
    a[1] += var[c * P + p];
  }
  block_sum_n<2>(a, red);
  const double muhat = a[0] / (double)C, W = a[1] / (double)C;
  // xbar^2 mean
  double b0[1] = {0.0};
  for (long long c = tid; c < C; c += nt) {
    const double xb = mean[c * P + p];
    b0[0] += xb * xb;
  }
  block_sum_n<1>(b0, red);
  const double m2 = b0[0] / (double)C;
  double b[4] = {0.0, 0.0, 0.0, 0.0};  // sum (xbar-muhat)^2, sum (s2-W)^2, sum (s2-W)(xbar^2-m2), sum (s2-W)(xbar-muhat)
  for (long long c = tid; c < C; c += nt) {
    const double xb = mean[c * P + p], s2 = var[c * P + p];
    const double dx = xb - muhat, ds = s2 - W;
    b[0] = fma(dx, dx, b[0]);
    b[1] = fma(ds, ds, b[1]);
    b[2] = fma(ds, xb * xb - m2, b[2]);
    b[3] = fma(ds, dx, b[3]);
  }
  block_sum_n<4>(b, red);
  if (tid == 0) {
    const double Cd = (double)C, nd = (double)n;
    const double var_xbar = b[0] / (Cd - 1.0);
    const double B = nd * var_xbar;
    const double var_w = (b[1] / (Cd - 1.0)) / Cd;
    const double var_b = 2.0 * B * B / (Cd - 1.0);
    const double cov_wb = (nd / Cd) * (b[2] / (Cd - 1.0) - 2.0 * muhat * (b[3] / (Cd - 1.0)));
    const double V = (nd - 1.0) * W / nd + (1.0 + 1.0 / Cd) * B / nd;
    const double var_V = ((nd - 1.0) * (nd - 1.0) * var_w + (1.0 + 1.0 / Cd) * (1.0 + 1.0 / Cd) * var_b +
                          2.0 * (nd - 1.0) * (1.0 + 1.0 / Cd) * cov_wb) / (nd * nd);
    const double df_V = 2.0 * V * V / var_V;
    const double df_adj = (df_V + 3.0) / (df_V + 1.0);
    const double W_df = 2.0 * W * W / var_w;
    const double R2_fixed = (nd - 1.0) / nd;
    const double R2_random = (1.0 + 1.0 / Cd) * (1.0 / nd) * (B / W);
    const double Vj = (1.0 - 1.0 / nd) * W + B / nd;  // drjacoby's gelman_rubin
    double* o = out + (long long)p * CC_GELMAN_FIELDS;
    o[0] = sqrt(Vj / W);
    o[1] = sqrt(df_adj * (R2_fixed + R2_random));
    o[2] = W; o[3] = B; o[4] = W_df; o[5] = df_adj; o[6] = R2_fixed; o[7] = R2_random;
  }
}

// est[c * n + i] = sum_a ifr_a w_a; w = popN / sum popN ("pop") or noise_a popN_a / sum_b noise_b popN_b ("arpop")
__global__ void summ_overall_ifr_kernel(const double* __restrict__ draws, long long C, long long n, long long sc, long long si,
                                        const int* __restrict__ idx_ifr, const int* __restrict__ idx_noise,
                                        const double* __restrict__ pop, int A, double* __restrict__ est) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= C * n) return;
  const long long c = r / n, i = r - c * n;
  const double* row = draws + c * sc + i * si;
  double denom = 0.0;
  for (int a = 0; a < A; a++) denom += (idx_noise ? row[idx_noise[a]] : 1.0) * pop[a];
  double s = 0.0;
  for (int a = 0; a < A; a++) {
    const double wi = ((idx_noise ? row[idx_noise[a]] : 1.0) * pop[a]) / denom;
    s += row[idx_ifr[a]] * wi;
  }
  est[r] = s;
}

// ---------------------------------------------------------------------------------------------- host side
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

#define SUMM_CUDA(expr)                                                                                   \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) return cc_set_error(CC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// draws on the device: either the caller's pointer (CC_MEM_DEVICE) or an upload of the caller's host array (dense copy)
struct DrawsView {
  const double* d = nullptr;
  long long sc = 0, si = 0;
  DevBuf own;
};
int view_draws(const double* draws, long long C, long long n, int P, long long sc, long long si, int mem, cudaStream_t st, DrawsView& v) {
  if (mem == CC_MEM_DEVICE) {
    v.d = draws; v.sc = sc; v.si = si;
    return CC_OK;
  }
  SUMM_CUDA(v.own.alloc((size_t)C * n * P * sizeof(double)));
  // rows of P doubles, one 2-D copy per chain (iteration stride si on the host side)
  for (long long c = 0; c < C; c++)
    SUMM_CUDA(cudaMemcpy2DAsync(v.own.as<double>() + c * n * P, (size_t)P * sizeof(double), draws + c * sc, (size_t)si * sizeof(double),
                                (size_t)P * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, st));
  v.d = v.own.as<double>(); v.sc = n * P; v.si = P;
  return CC_OK;
}

int check_shape(const char* who, const void* draws, long long C, long long n, int P, long long sc, long long si) {
  if (!draws) return cc_set_error(CC_E_INVALID, std::string(who) + ": null draws");
  if (C < 1 || n < 1 || P < 1) return cc_set_error(CC_E_INVALID, std::string(who) + ": need n_chains, n_iter, n_params >= 1");
  if (si < P || sc < 1) return cc_set_error(CC_E_INVALID, std::string(who) + ": strides must cover one row of n_params doubles");
  return CC_OK;
}

}  // namespace

extern "C" int cc_summarize_draws(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params,
                                  int64_t stride_chain, int64_t stride_iter, int32_t by_chain, int mem, void* stream, double* out) {
  if (int rc = check_shape("cc_summarize_draws", draws, n_chains, n_iter, n_params, stride_chain, stride_iter)) return rc;
  if (!out) return cc_set_error(CC_E_INVALID, "cc_summarize_draws: null output");
  SUMM_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long C = n_chains, n = n_iter, R = C * n;
  const int P = n_params;
  DrawsView v;
  if (int rc = view_draws(draws, C, n, P, stride_chain, stride_iter, mem, st, v)) return rc;
  DevBuf ser, res;
  SUMM_CUDA(ser.alloc((size_t)R * P * sizeof(double)));
  const long long G = by_chain ? C : 1, len = by_chain ? n : R, n_series = G * P;
  SUMM_CUDA(res.alloc((size_t)n_series * CC_SUMM_FIELDS * sizeof(double)));
  dim3 tg((unsigned)((R + 31) / 32), (unsigned)((P + 31) / 32)), tb(32, 8);
  summ_transpose_kernel<<<tg, tb, 0, st>>>(v.d, C, n, P, v.sc, v.si, 0, ser.as<double>());
  summ_series_kernel<<<(unsigned)std::min<long long>(n_series, 148LL * 64), kSummThreads, 0, st>>>(ser.as<double>(), len, n_series,
                                                                                                 res.as<double>());
  SUMM_CUDA(cudaGetLastError());
  // the series are ordered [param][group]; the caller gets [group][param][field]
  std::vector<double> h((size_t)n_series * CC_SUMM_FIELDS);
  SUMM_CUDA(cudaMemcpyAsync(h.data(), res.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  SUMM_CUDA(cudaStreamSynchronize(st));
  for (long long p = 0; p < P; p++)
    for (long long g = 0; g < G; g++)
      for (int f = 0; f < CC_SUMM_FIELDS; f++) out[(g * P + p) * CC_SUMM_FIELDS + f] = h[(size_t)((p * G + g) * CC_SUMM_FIELDS + f)];
  return CC_OK;
}

extern "C" int cc_chain_moments(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params,
                                int64_t stride_chain, int64_t stride_iter, int64_t first_iter, int mem, void* stream, double* mean,
                                double* var) {
  if (int rc = check_shape("cc_chain_moments", draws, n_chains, n_iter, n_params, stride_chain, stride_iter)) return rc;
  if (!mean || !var) return cc_set_error(CC_E_INVALID, "cc_chain_moments: null output");
  if (first_iter < 0 || first_iter >= n_iter) return cc_set_error(CC_E_INVALID, "cc_chain_moments: need 0 <= first_iter < n_iter");
  SUMM_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long C = n_chains, n = n_iter - first_iter;
  const int P = n_params;
  DrawsView v;
  if (int rc = view_draws(draws, C, n_iter, P, stride_chain, stride_iter, mem, st, v)) return rc;
  DevBuf ser, dm, dv;
  SUMM_CUDA(ser.alloc((size_t)C * n * P * sizeof(double)));
  double *pm = mean, *pv = var;
  if (mem != CC_MEM_DEVICE) {
    SUMM_CUDA(dm.alloc((size_t)C * P * sizeof(double)));
    SUMM_CUDA(dv.alloc((size_t)C * P * sizeof(double)));
    pm = dm.as<double>(); pv = dv.as<double>();
  }
  dim3 tg((unsigned)((C * n + 31) / 32), (unsigned)((P + 31) / 32)), tb(32, 8);
  summ_transpose_kernel<<<tg, tb, 0, st>>>(v.d, C, n, P, v.sc, v.si, first_iter, ser.as<double>());
  const long long warps = C * P;
  summ_moments_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(ser.as<double>(), C, n, P, pm, pv);
  SUMM_CUDA(cudaGetLastError());
  if (mem != CC_MEM_DEVICE) {
    SUMM_CUDA(cudaMemcpyAsync(mean, pm, (size_t)C * P * sizeof(double), cudaMemcpyDeviceToHost, st));
    SUMM_CUDA(cudaMemcpyAsync(var, pv, (size_t)C * P * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  SUMM_CUDA(cudaStreamSynchronize(st));  // the scratch is freed on return
  return CC_OK;
}

extern "C" int cc_gelman_from_moments(int device, const double* mean, const double* var, int64_t n_chains, int64_t n_iter_used,
                                      int32_t n_params, int mem, void* stream, double* out) {
  if (!mean || !var || !out) return cc_set_error(CC_E_INVALID, "cc_gelman_from_moments: null argument");
  if (n_chains < 2) return cc_set_error(CC_E_INVALID, "cc_gelman_from_moments: Gelman-Rubin requires at least two chains");
  if (n_iter_used < 2 || n_params < 1) return cc_set_error(CC_E_INVALID, "cc_gelman_from_moments: need n_iter_used >= 2, n_params >= 1");
  SUMM_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = (size_t)n_chains * n_params * sizeof(double);
  DevBuf dm, dv, res;
  const double *pm = mean, *pv = var;
  if (mem != CC_MEM_DEVICE) {
    SUMM_CUDA(dm.alloc(nb));
    SUMM_CUDA(dv.alloc(nb));
    SUMM_CUDA(cudaMemcpyAsync(dm.p, mean, nb, cudaMemcpyHostToDevice, st));
    SUMM_CUDA(cudaMemcpyAsync(dv.p, var, nb, cudaMemcpyHostToDevice, st));
    pm = dm.as<double>(); pv = dv.as<double>();
  }
  SUMM_CUDA(res.alloc((size_t)n_params * CC_GELMAN_FIELDS * sizeof(double)));
  summ_gelman_kernel<<<(unsigned)n_params, kSummThreads, 0, st>>>(pm, pv, n_chains, n_iter_used, n_params, res.as<double>());
  SUMM_CUDA(cudaGetLastError());
  SUMM_CUDA(cudaMemcpyAsync(out, res.p, (size_t)n_params * CC_GELMAN_FIELDS * sizeof(double), cudaMemcpyDeviceToHost, st));
  SUMM_CUDA(cudaStreamSynchronize(st));
  return CC_OK;
}

extern "C" int cc_overall_ifr_draws(int device, const double* draws, int64_t n_chains, int64_t n_iter, int32_t n_params,
                                    int64_t stride_chain, int64_t stride_iter, const int32_t* idx_ifr, const int32_t* idx_noise,
                                    const double* popN, int32_t n_strata, int mem, void* stream, double* est) {
  if (int rc = check_shape("cc_overall_ifr_draws", draws, n_chains, n_iter, n_params, stride_chain, stride_iter)) return rc;
  if (!idx_ifr || !popN || !est || n_strata < 1) return cc_set_error(CC_E_INVALID, "cc_overall_ifr_draws: null argument");
  for (int a = 0; a < n_strata; a++)
    if (idx_ifr[a] < 0 || idx_ifr[a] >= n_params || (idx_noise && (idx_noise[a] < 0 || idx_noise[a] >= n_params)))
      return cc_set_error(CC_E_INVALID, "cc_overall_ifr_draws: parameter index out of range");
  SUMM_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long C = n_chains, n = n_iter, R = C * n;
  DrawsView v;
  if (int rc = view_draws(draws, C, n, n_params, stride_chain, stride_iter, mem, st, v)) return rc;
  DevBuf di, dn, dp, de;
  SUMM_CUDA(di.alloc(n_strata * sizeof(int)));
  SUMM_CUDA(dp.alloc(n_strata * sizeof(double)));
  SUMM_CUDA(cudaMemcpyAsync(di.p, idx_ifr, n_strata * sizeof(int), cudaMemcpyHostToDevice, st));
  SUMM_CUDA(cudaMemcpyAsync(dp.p, popN, n_strata * sizeof(double), cudaMemcpyHostToDevice, st));
  if (idx_noise) {
    SUMM_CUDA(dn.alloc(n_strata * sizeof(int)));
    SUMM_CUDA(cudaMemcpyAsync(dn.p, idx_noise, n_strata * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  double* pe = est;
  if (mem != CC_MEM_DEVICE) {
    SUMM_CUDA(de.alloc((size_t)R * sizeof(double)));
    pe = de.as<double>();
  }
  summ_overall_ifr_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(v.d, C, n, v.sc, v.si, di.as<int>(), idx_noise ? dn.as<int>() : nullptr,
                                                                      dp.as<double>(), n_strata, pe);
  SUMM_CUDA(cudaGetLastError());
  if (mem != CC_MEM_DEVICE) SUMM_CUDA(cudaMemcpyAsync(est, pe, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost, st));
  SUMM_CUDA(cudaStreamSynchronize(st));
  return CC_OK;
}
