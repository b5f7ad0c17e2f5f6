// covidcurve_b200.cu - kernels + C ABI (include/covidcurve_b200.h) of the B200-native COVIDCurve hot path.
//
// Built for sm_100a only (see __graft_entry__.build()).  There is no CPU fallback anywhere in this file: every
// compute entry point needs a CUDA device and returns CC_E_CUDA otherwise.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/covidcurve_b200.h"
#include "cc_internal.h"
#include "cc_eval.cuh"
#include "cc_eval_warp.cuh"
#include "cc_eval_tile.cuh"
#include "cc_mcmc_dev.cuh"
#include "cc_model.cuh"
#include "cc_sim.cuh"

using namespace ccdev;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CC_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      return fail(CC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
  } while (0)

int cc_set_error(int code, const std::string& msg) { return fail(code, msg); }

extern "C" const char* cc_last_error(void) { return g_err.c_str(); }
extern "C" int cc_abi_version(void) { return CC_ABI_VERSION; }
extern "C" int cc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ kernels
constexpr int kEvalThreads = 128;

#ifndef CC_P2_SYNC
#define CC_P2_SYNC 1   // 1: the warps of a CTA start every evaluation together; 2: every stage (developer experiments); 0: free-running
#endif
constexpr int kWarpsPerCta = 4;
// Shape of pass C of the batched likelihood: few LARGE CTAs whose warps start every evaluation together.  The per-day code
// is ~100 kB of instructions per evaluation, far beyond the SM's instruction cache: free-running warps each stream it from
// L2 (ncu: `no_instruction` was the top stall of the box workload, 3.4 per issue); warps of one CTA that run the same stage at
// the same time share one stream.  Full lock-step (a barrier per stage) loses again because the FP64-heavy stages then
// collide on the pipe.  Measured on B200, box workload, M evals/s (value / e2e): 5 CTAs x 4 warps free-running 73.8 / 63.5;
// the same with the barrier 77.9 / 68; 2 x 8: 79.5 / 74.3; 2 x 10: 83.1 / 74.7; 1 x 20: 80.6 / 77.1; 2 x 10 with a barrier
// per stage 78.9 / 74.9 (profiles/r1_summary.md).
// The shape is chosen per model at creation (cc_model_create) among a few candidates, the one with the most resident warps per SM
// that fits the model's shared memory: p2_warps(nt) is the LARGEST CTA a kernel instantiation is compiled for (launch bounds), the
// kernel itself reads its warp count from blockDim.
struct P2Shape { int warps, ctas; };
#ifdef CC_P2_WARPS
__host__ __device__ constexpr int p2_warps(int) { return CC_P2_WARPS; }
#else
__host__ __device__ constexpr int p2_warps(int nt) { return nt <= 4 ? 12 : nt <= 5 ? 10 : nt <= 7 ? 15 : 6; }
#endif
#ifdef CC_P2_CTAS
__host__ __device__ constexpr int p2_ctas(int) { return CC_P2_CTAS; }
#else
__host__ __device__ constexpr int p2_ctas(int nt) { return nt <= 5 ? 2 : 1; }
#endif
// candidates in order of preference (measured on B200, M evals/s box / posterior-like: T = 200 2 x 12 warps 92.9 / 93.8 against 2 x 10
// 90.5 / 92.9; T = 300 2 x 10 (2 x 11, 2 x 8, 1 x 20 slower); T = 400 1 x 15 62.4 / 59.0 against 2 x 7 60.1 / 58.0 and 3 x 5 59.6 / 59.1)
static inline int p2_candidates(int nt, P2Shape* out) {
#if defined(CC_P2_WARPS) || defined(CC_P2_CTAS)
  out[0] = {p2_warps(nt), p2_ctas(nt)};
  return 1;
#else
  int n = 0;
  if (nt <= 4) { out[n++] = {12, 2}; out[n++] = {10, 2}; }
  else if (nt <= 5) { out[n++] = {10, 2}; }
  else if (nt <= 7) { out[n++] = {15, 1}; out[n++] = {7, 2}; }
  else { out[n++] = {6, 1}; }
  for (int w = p2_warps(nt) - 1; w >= 1 && n < 24; w--) out[n++] = {w, 1};  // models with many strata / knots: whatever still fits
  return n;
#endif
}
#ifndef CC_MIN_CTAS
#define CC_MIN_CTAS 4  // resident CTAs per SM the evaluation kernels are compiled for (caps registers at 128 per thread)
#endif
// The cached Metropolis sweep runs as ONE 12-warp CTA per SM (168 registers, 15.5 kB of shared memory per warp at cfg3)
// whose warps start every single-site update together: same parameter, same stages, one instruction stream
// (measured, cfg3 64 x 50: 3 CTAs x 4 free-running warps 1.52 M iterations/s, 1 x 12 free-running 1.51, 1 x 12 with the
// barrier 1.63, 1 x 14 with the barrier 1.62).
#ifndef CC_SWEEP_CTAS
#define CC_SWEEP_CTAS 1
#endif
#ifndef CC_SWEEP_WARPS
#define CC_SWEEP_WARPS 12
#endif
#ifndef CC_SWEEP_SYNC
#define CC_SWEEP_SYNC 1
#endif
constexpr int kSweepWarps = CC_SWEEP_WARPS;

// ---- batched likelihood as a four-pass pipeline (five launches) -------------------------------------------------------
// pass A (one THREAD per vector): phase 1 -> record rec[f * cap + v] in global memory, a regime key, a histogram of the keys
// pass B (tiny, two launches): bucket offsets, then a scatter that groups the vector indices by key (order within a bucket is
//         irrelevant: every vector is evaluated independently, so the values do not depend on the permutation)
// pass C (one WARP per vector, CTA tiles of the permutation drawn from a ticket): phase 2, the per-day work.
// pass D (one THREAD per vector): phase 3.
// Grouping by regime keeps the warps of a CTA on the same code path (small gamma table / large table / general algorithm /
// no per-day work at all) and lets the launch start with the expensive vectors (history: profiles/r1_summary.md).
constexpr int kNumKeys = 8;       // histogram slots: key 0 = no per-day work, 1 .. 4 = size of the gamma coefficient table, 7 = general
                                  // incomplete-gamma algorithm; the phase-2 tile ticket lives behind the two arrays.  (Ordering the
                                  // vectors by an estimated cost of the per-day pass in 126 buckets instead - so that the warps of a
                                  // CTA also FINISH every evaluation together - measured slower: 88.4 vs 92.0 M evals/s, round 2.)
constexpr int kLoglikeLaunches = 5;  // kernels per cc_loglike_batch pipeline (launch_loglike)
#ifndef CC_P2_TILE
#define CC_P2_TILE 2
#endif
constexpr int kPhase2Tile = CC_P2_TILE;    // vectors per warp of a dynamically fetched CTA tile of pass C

template <int NT>
__global__ void __launch_bounds__(128) loglike_phase1_kernel(const KModel m, const double* __restrict__ theta, long long n, long long ld,
                                                              double* __restrict__ rec, long long cap, int* __restrict__ key,
                                                              unsigned int* __restrict__ counts) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const RecLayout L = rec_layout(m.K, m.A, m.S);
  int k = -1;
  if (v < n) {
    auto par = [&](int i) { return theta[(long long)i * ld + v]; };
    auto put = [&](int f, double x) { rec[(long long)f * cap + v] = x; };
    put(L.o_status, 0.0);
    p1_weights(m, L, par, put);
    k = 0;  // knots not increasing: no per-day work
    if (p1_spline(m, L, par, put)) {
      p1_gamma<NT>(m, L, par, put);
      p1_sero(m, L, par, put);
      const double fast = rec[(long long)(L.gam + 8) * cap + v];
      const double h = rec[(long long)(L.gam + 2) * cap + v];
      k = (fast == 0.0) ? (kNumKeys - 1) : (h <= 3.6) ? 1 : (h <= 16.0) ? 2 : (h <= 32.0) ? 3 : 4;
    }
    key[v] = k;
  }
  // warp-aggregated histogram
  const unsigned peers = __match_any_sync(0xffffffffu, k);
  if (k >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&counts[k], (unsigned)__popc(peers));
}

__global__ void bucket_offsets_kernel(const unsigned int* __restrict__ counts, unsigned int* __restrict__ cursor) {
  // heavy vectors first, so that the long-running tiles start early and the tail of the launch is made of cheap ones
  unsigned int run = 0;
  for (int q = kNumKeys - 1; q >= 0; q--) {
    cursor[q] = run;
    run += counts[q];
  }
}

__global__ void bucket_scatter_kernel(const int* __restrict__ key, long long n, unsigned int* __restrict__ cursor, int* __restrict__ perm) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = (v < n) ? key[v] : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, k);
  if (k < 0) return;
  const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(&cursor[k], (unsigned)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  perm[base + __popc(peers & ((1u << lane) - 1))] = (int)v;
}

// tile_phase2 with a CTA barrier in front of every stage: all warps of the CTA execute the same stage code at the same
// time, so the SM fetches one instruction stream per CTA instead of one per warp.  `live` = this warp has a vector.
template <int NT, typename Par>
__device__ __forceinline__ void cta_phase2_impl(const KModel& m, const RecLayout& L, const TileSmem& w, const TileTables& tb, double* __restrict__ scr,
                                                long long stride, bool live, Par par) {
  const int lane = threadIdx.x & 31;
  __syncthreads();
  if (live) {
    for (int f = lane; f < L.n2; f += 32) w.rec[f] = scr[f * stride];
    __syncwarp();
    live = w.rec[L.flags] != 0.0;  // knots not increasing: pass D returns the failure value
  }
  if (live) tile_sero_prefix<NT>(m, L, w, tb);
  __syncthreads();
  if (live) tile_gamma<NT>(m, L, w, tb);
  __syncthreads();
  if (live) {
    const double tot = stage_spline<NT>(m, L, w.rec, w.xs);
    const bool ok = stage_popn(m, L, w.rec, tot);
    if (lane == 0) scr[L.o_status * stride] = ok ? 0.0 : 1.0;
    live = ok;
    __syncwarp();
  }
  __syncthreads();
  if (live) stage_surveys(m, w.xs, w.sk, [&](int sv, double v) { scr[(L.o_base + sv) * stride] = v; });
  __syncthreads();
  if (live) stage_conv<NT>(m, w.xs, w.gs, w.auc, conv_lag_blocks(m, L, w.rec));
  __syncthreads();
  if (live) {
    double l1, s_all, s_unc;
    stage_l1_direct(m, tb, w.auc, w.rec[L.snm], w.rec + L.ne, par, l1, s_all, s_unc, w.xs, w.gs);
    if (lane == 0) {
      scr[L.o_l1 * stride] = l1 - m.l1_const;
      scr[L.o_texpd * stride] = s_all;
      scr[L.o_sunc * stride] = s_unc;
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(32 * p2_warps(NT), p2_ctas(NT)) loglike_phase2_kernel(const KModel m, long long n, double* __restrict__ rec,
                                                                                          long long cap, const int* __restrict__ perm,
                                                                                          unsigned int* __restrict__ ticket,
                                                                                          const double* __restrict__ theta, long long ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int s_ticket[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RecLayout L = rec_layout(m.K, m.A, m.S);
  const TileTables tb = tile_tables_load(m, nullptr);
  const TileSmem w = carve_tile(reinterpret_cast<double*>(smem_raw) + (size_t)warp * tile_smem_doubles(m.K, m.A, m.S, NT), L, NT);
  tile_smem_init(w);
  // dynamic queue of CTA tiles (kPhase2Tile vectors per warp) in bucket order (heavy regimes first): a CTA that drew cheap
  // vectors comes back for more, so the launch ends within one short tile of the ideal instead of one long tile of the
  // worst regime.  The ticket of the NEXT tile is drawn at the start of the current one and published at its end (two slots),
  // so the round trip of the atomic is hidden behind the evaluations instead of stalling the whole CTA once per tile.
  if (threadIdx.x == 0) s_ticket[0] = atomicAdd(ticket, 1u);
  __syncthreads();
  for (unsigned int it = 0;; it++) {
    const long long c0 = (long long)s_ticket[it & 1] * (kPhase2Tile * (int)(blockDim.x >> 5));
    if (c0 >= n) break;
    unsigned int next_ticket = 0;
    if (threadIdx.x == 0) next_ticket = atomicAdd(ticket, 1u);
    const long long t0 = c0 + (long long)warp * kPhase2Tile;
    const long long pos = t0 + lane;
    const int vl = (lane < kPhase2Tile && pos < n) ? perm[pos] : 0;
    const int cnt = (int)max((long long)0, min((long long)kPhase2Tile, n - t0));
    for (int e = 0; e < kPhase2Tile; e++) {
#if CC_P2_SYNC == 2
      const long long v = __shfl_sync(0xffffffffu, vl, min(e, max(cnt - 1, 0)));
      cta_phase2_impl<NT>(m, L, w, tb, rec + v, cap, e < cnt, [&](int i) { return theta[(long long)i * ld + v]; });
#else
#if CC_P2_SYNC == 1
      if (e) __syncthreads();  // the warps of the CTA start every evaluation together (the first of a tile: behind the ticket barrier)
#endif
      if (e < cnt) {
        const long long v = __shfl_sync(0xffffffffu, vl, e);
        tile_phase2<NT>(m, L, w, tb, rec + v, (int)0, [&](int i) { return theta[(long long)i * ld + v]; }, cap);
      }
#endif
    }
    if (threadIdx.x == 0) s_ticket[(it + 1) & 1] = next_ticket;
    __syncthreads();
  }
}

// pass D (one THREAD per vector, original order -> coalesced): chained binomial, serology term, total + clamp
__global__ void __launch_bounds__(128) loglike_phase3_kernel(const KModel m, long long n, const double* __restrict__ rec, long long cap,
                                                              double* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const RecLayout L = rec_layout(m.K, m.A, m.S);
  // tile_phase3 with one lane per vector and no cross-lane reduction (E = 32)
  out[v] = tile_phase3(m, L, rec + v, 0, 0, 1, 32, cap);
}

// decode lane -> (vector e of the tile, sub-lane) for tiles of E = 2^le vectors
struct TileLane { int e, sub, G, E; };
__device__ __forceinline__ TileLane tile_lane(int le) {
  TileLane t;
  const int lane = threadIdx.x & 31;
  t.E = 1 << le;
  t.G = 32 >> le;
  t.e = lane & (t.E - 1);
  t.sub = lane >> le;
  return t;
}

// initial state of every particle: phi, loglike, logprior
template <int NT>
__global__ void __launch_bounds__(32 * kWarpsPerCta, CC_MIN_CTAS) mcmc_init_tile_kernel(const KModel m, const McmcState s, int le, double* __restrict__ scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const RecLayout L = rec_layout(m.K, m.A, m.S);
  const TileTables tb = tile_tables_load(m, reinterpret_cast<double*>(smem_raw));
  const TileSmem w = carve_tile(reinterpret_cast<double*>(smem_raw) + tile_tables_doubles(m.T) + (size_t)warp * tile_smem_doubles(m.K, m.A, m.S, NT), L, NT);
  tile_smem_init(w);
  const TileLane tl = tile_lane(le);
  const int gwarp = blockIdx.x * kWarpsPerCta + warp, nwarps = gridDim.x * kWarpsPerCta;
  double* scr = scratch + (size_t)gwarp * L.F * 32;
  const int P = s.n_chains * s.rungs, d = m.d;
  for (int p0 = gwarp * tl.E; p0 < P; p0 += nwarps * tl.E) {
    const int p = p0 + tl.e;
    const bool valid = p < P;
    const double* th = s.theta + (size_t)(valid ? p : 0) * d;
    auto par = [&](int j) { return th[j]; };
    if (valid)
      for (int i = tl.sub; i < d; i += tl.G) {
        const double lo = m.pmin[i], hi = m.pmax[i];
        s.phi[(size_t)p * d + i] = theta_to_phi(trans_type(lo, hi), th[i], lo, hi);
      }
    if (valid && tl.sub == 0) tile_phase1<NT>(m, L, scr, tl.e, par);
    __syncwarp();
    const int cnt = min(tl.E, P - p0);
    for (int q = 0; q < cnt; q++) {
      const double* thq = s.theta + (size_t)(p0 + q) * d;
      tile_phase2<NT>(m, L, w, tb, scr, q, [&](int j) { return thq[j]; });
    }
    __syncwarp();
    const double ll = tile_phase3(m, L, scr, tl.e, tl.sub, tl.G, tl.E);
    const double lp = tile_logprior(m, tl.sub, tl.G, tl.E, par);
    if (valid && tl.sub == 0) {
      s.ll[p] = ll;
      s.lp[p] = lp;
    }
    __syncwarp();
  }
}

// ---- the cached Metropolis sweep.
// The cached state of a particle - its record and every per-day array of its CURRENT parameter vector - lives in global
// memory (s.cache, ~13 kB per particle at the cfg3 shape: L2-resident for thousands of particles) and persists from one
// iteration to the next.  Shared memory holds only the working set of the proposal in flight: the two records and ONE
// buffer per array kind, filled either by recomputing the stage the proposed parameter touches or by loading the cached
// array a dependent stage needs.  An accepted proposal writes the arrays it changed back; a rejected one costs nothing.
// (The first version double-buffered all arrays in shared memory - 28 kB per warp, 8 warps per SM, and one full
// re-evaluation per particle and sweep; this layout needs 14.6 kB per warp: 12 warps per SM.)
struct SweepSmem {
  double *recA, *recB, *xs, *gs, *auc, *sk, *pcur, *pterm;
  double *q_thp, *q_php, *q_adj, *q_ptn, *q_logu;  // the sweep's d proposals, drawn up front (q_logu[i] becomes the accept flag)
};
struct SweepCacheLayout {
  int rec, xs, gs, auc, sk, scal, size;  // offsets (doubles) into a particle's cache block; scal: tot, s1, sa_obs, s_all_auc, s_unc, irregular, valid, l2, l3
  int nxs, ngs, nauc, nsk, nrec;
};
__host__ __device__ inline SweepCacheLayout sweep_cache_layout(int K, int A, int S, int NT) {
  const RecLayout L = rec_layout(K, A, S);
  SweepCacheLayout c;
  c.nrec = (L.F + 3) & ~3; c.nxs = 12 * (8 * NT + 7); c.ngs = 64 * NT + 16; c.nauc = 64 * NT; c.nsk = 64 * NT + 4;
  c.rec = 0; c.xs = c.rec + c.nrec; c.gs = c.xs + c.nxs; c.auc = c.gs + c.ngs; c.sk = c.auc + c.nauc; c.scal = c.sk + c.nsk;
  c.size = c.scal + 12;
  return c;
}
__host__ __device__ inline size_t sweep_smem_doubles(int d, int K, int A, int S, int NT) {
  const SweepCacheLayout c = sweep_cache_layout(K, A, S, NT);
  const size_t dp = (size_t)((d + 3) & ~3);
  return 2 * (size_t)c.nrec + c.nxs + c.ngs + c.nauc + c.nsk + dp + dp + 4 + 5 * dp;
}
__device__ inline SweepSmem carve_sweep(double* q, int d, const SweepCacheLayout& c) {
  SweepSmem w;
  const int dp = (d + 3) & ~3;
  w.recA = q; q += c.nrec; w.recB = q; q += c.nrec;
  w.xs = q; q += c.nxs;
  w.gs = q; q += c.ngs;
  w.auc = q; q += c.nauc;
  w.sk = q; q += c.nsk;
  w.pcur = q; q += dp;
  w.pterm = q; q += dp + 4;
  w.q_thp = q; q += dp; w.q_php = q; q += dp; w.q_adj = q; q += dp; w.q_ptn = q; q += dp; w.q_logu = q;
  return w;
}

enum : int { kDepWeights = 1, kDepSpline = 2, kDepGamma = 4, kDepSero = 8,
             kKeepL3 = 16,   // the proposal cannot change the serology term (an IFR, the onset-to-death delay)
             kKeepL2 = 32,   // the proposal cannot change the age split (test sensitivity / specificity, seroconversion, seroreversion)
             kKeepNe = 64 }; // the proposal cannot change the attack-rate weights ne[a] (every parameter but the noise scalars)

template <typename T>
__device__ __forceinline__ void swap_ptr(T*& a, T*& b) { T* t = a; a = b; b = t; }

// warp-wide copy of n doubles (both sides 8-byte aligned; global side coalesced)
// (no __restrict__ / const on purpose: the cache block is read back by the thread that wrote it within one kernel, so the
// loads must stay on the coherent path)
__device__ __forceinline__ void warp_copy(double* dst, double* src, int n) {
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int j = lane; j < n; j += 32) dst[j] = src[j];
}

// One sweep of d single-site Metropolis-Hastings updates for every (chain, rung); one warp per particle, walking particles.
// A proposal for parameter i redoes only the stages that depend on it (m.dep[i]): the IFR / attack-rate / test parameters
// touch no per-day stage at all, the knots skip the gamma table, mod/sod skip the spline, the seroconversion rate skips
// both.  Every cached quantity is a pure function of the current parameter vector, so the values equal a from-scratch
// evaluation.  `iteration` is the 1-based drjacoby iteration this sweep produces (2 .. burnin+samples).
template <int NT>
__global__ void __launch_bounds__(32 * kSweepWarps, CC_SWEEP_CTAS) mcmc_sweep_cached_kernel(const KModel m, const McmcState s, int iteration) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const RecLayout L = rec_layout(m.K, m.A, m.S);
  const SweepCacheLayout CL = sweep_cache_layout(m.K, m.A, m.S, NT);
  const int d = m.d;
  SweepSmem w = carve_sweep(reinterpret_cast<double*>(smem_raw) + (size_t)warp * sweep_smem_doubles(d, m.K, m.A, m.S, NT), d, CL);
  const TileTables tb = tile_tables_load(m, nullptr);
  for (int i = lane; i < kXsLead; i += 32) w.xs[i] = 0.0;  // the zeros in front of the series and of the delay kernel are
  if (lane < 7) w.gs[lane] = 0.0;                          // never overwritten (full-array copies carry them along)
  __syncwarp();
  const int P = s.n_chains * s.rungs;
  const bool adapt = s.adapt_in_sampling || iteration <= s.burnin;
  // tile_gamma keeps its coefficient table in the (not yet written) convolution output buffer and its reciprocal table
  // behind the zero blocks of the series buffer
  auto view = [&](double* rec) {
    TileSmem v;
    v.rec = rec; v.xs = w.xs; v.gs = w.gs; v.auc = w.auc; v.sk = w.sk; v.ctab = w.auc;
    return v;
  };
  // Poisson term from the cached sums (or day by day when they do not apply), written into `rec`
  auto l1_needs_days = [&](const double* rec, const L1Sums& q) {
    const double snm = rec[L.snm];
    return q.irregular || !(snm > 0.0) || !isfinite(snm);
  };
  auto finish_l1 = [&](double* rec, const L1Sums& q, double* aucs, auto parf) {
    const double snm = rec[L.snm];
    double l1, texpd;
    if (!l1_needs_days(rec, q)) {
      l1 = (q.s1 + m.l1_sumx * log(snm)) - snm * q.sa_obs;
      texpd = snm * q.s_all_auc;
    } else {
      double su;
      stage_l1_direct(m, tb, aucs, snm, rec + L.ne, parf, l1, texpd, su);
    }
    __syncwarp();
    if (lane == 0) {
      rec[L.o_l1] = l1 - m.l1_const;
      rec[L.o_texpd] = texpd;
      rec[L.o_sunc] = q.s_unc;
    }
    __syncwarp();
  };

  // named barrier shared by the warps of the CTA; arrivals are counted per warp, wherever in the code the warp is
  auto cta_bar = [&]() {
#if CC_SWEEP_SYNC
    asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
#endif
  };
  // Particles are handed out dynamically, one per warp of the CTA and ticket, in rung-major order starting with the hottest
  // rung: the warps of a CTA then hold neighbouring chains at the SAME temperature (similar acceptance rates and validity,
  // hence similar work between the per-update barriers), the expensive particles go first and CTAs that drew cheap ones
  // come back for more - the sweep ends within one particle of the ideal instead of one static round.
  __shared__ int s_q0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_q0 = atomicAdd(s.ticket, wpc);
    __syncthreads();
    const int q = s_q0 + warp;
    if (s_q0 >= P) break;
    if (q >= P) {  // no particle left for this warp in the last chunk: keep the CTA's barrier count
      for (int i = 0; i < d; i++) cta_bar();
      continue;
    }
    const int ro = q / s.n_chains, c = q - ro * s.n_chains;
    const int r = (s.cold_pos == 0) ? (s.rungs - 1 - ro) : ro;  // hot end of the ladder first
    const int b = c * s.rungs + r;
    const int pid = s.at_rung[b];  // particle (within chain) sitting at rung position r
    const int p = c * s.rungs + pid;
    const int slot = s.bw_per_particle ? p : b;
    const double beta = s.beta[r];
    double* th = s.theta + (size_t)p * d;
    double* ph = s.phi + (size_t)p * d;
    double* bw = s.bw + (size_t)slot * d;
    int* bwi = s.bw_index + (size_t)slot * d;
    unsigned int* macc = s.move_acc + (size_t)b * d;
    double* cache = s.cache + (size_t)p * CL.size;
    double ll = s.ll[p], lp = s.lp[p];

    __syncwarp();
    for (int j = lane; j < d; j += 32) w.pcur[j] = th[j];
    __syncwarp();
    for (int j = lane; j < d; j += 32) w.pterm[j] = prior_term_shared(m, j, w.pcur[j]);
    if (lane < 3) w.pterm[d + lane] = (m.jac_rows[lane] >= 0) ? m.jac_counts[lane] * log(w.pcur[max(m.jac_rows[lane], 0)]) : 0.0;
    // All d proposals of this sweep at once, one parameter per lane: parameter i's random numbers are keyed by (chain,
    // particle, iteration, i), and phi[i], bw[i] and theta[i] are touched by update i alone, so drawing them ahead of the
    // sequential accept/reject loop changes no value - it only takes the Philox rounds, the transforms, the proposed
    // prior term and log(u) off the per-update dependency chain.
    for (int j = lane; j < d; j += 32) {
      const ccrng::Draw dr = ccrng::draw(s.seed, ccrng::kStreamMove, (uint32_t)(s.chain_offset + c), (uint32_t)pid,
                                         (uint32_t)iteration, (uint32_t)j);
      const double lo = m.pmin[j], hi = m.pmax[j];
      const int tt = trans_type(lo, hi);
      const double php = ph[j] + bw[j] * dr.z;
      const double thp = phi_to_theta(tt, php, lo, hi);
      w.q_php[j] = php;
      w.q_thp[j] = thp;
      w.q_adj[j] = mh_adjust(tt, w.pcur[j], thp, lo, hi);
      w.q_ptn[j] = prior_term_shared(m, j, thp);
      w.q_logu[j] = log(dr.u);
    }
    __syncwarp();
    double tot_cur = 0.0;
    L1Sums sums_cur = {0.0, 0.0, 0.0, 0.0, 1};
    double l2_cur = 0.0, l3_cur = 0.0;  // age-split and serology sums of the current state
    bool unbuilt = cache[CL.scal + 6] == 0.0;
    if (!unbuilt) {
      // caches of the current state, left by the previous sweep
      warp_copy(w.recA, cache + CL.rec, CL.nrec);
      tot_cur = cache[CL.scal + 0];
      sums_cur.s1 = cache[CL.scal + 1]; sums_cur.sa_obs = cache[CL.scal + 2]; sums_cur.s_all_auc = cache[CL.scal + 3];
      sums_cur.s_unc = cache[CL.scal + 4]; sums_cur.irregular = (int)cache[CL.scal + 5];
      l2_cur = cache[CL.scal + 7]; l3_cur = cache[CL.scal + 8];
    } else {
      // First sweep of a particle: nothing is cached yet.  An all-zero record reads as "no usable caches" (flags == 0), so
      // every proposal is evaluated from scratch until the first one is accepted and becomes the cached state.
      for (int f = lane; f < CL.nrec; f += 32) w.recA[f] = 0.0;
    }
    __syncwarp();

    for (int i = 0; i < d; i++) {
      cta_bar();  // the warps of the CTA work on the same parameter (same stages, same code) at the same time
      const int dep = m.dep[i];
      const double thp = w.q_thp[i], php = w.q_php[i], adj = w.q_adj[i];
      auto par = [&](int j) { return j == i ? thp : w.pcur[j]; };
      for (int f = lane; f < L.F; f += 32) w.recB[f] = w.recA[f];
      __syncwarp();
      // (a proposal that cannot move the attack-rate weights keeps them: no normalising sum, no divisions)
      if ((dep & kDepWeights) || unbuilt) p1_weights_warp(m, L, par, w.recB, (dep & kKeepNe) && !unbuilt);
      if (lane == 0) {
        auto putB = [&](int f, double v) { w.recB[f] = v; };
        bool ok = w.recB[L.flags] != 0.0;
        if ((dep & kDepSpline) || unbuilt) ok = p1_spline(m, L, par, putB);
        // a vector whose knots were disordered has no gamma / sero constants yet: build them when the knots come back
        const bool need_all = ok && (w.recA[L.flags] == 0.0);
        if (ok && ((dep & kDepGamma) || need_all)) p1_gamma<NT>(m, L, par, putB);
        if (ok && ((dep & kDepSero) || need_all)) p1_sero(m, L, par, putB);
        w.recB[L.o_status] = 0.0;
      }
      __syncwarp();
      const bool valid = w.recB[L.flags] != 0.0;
      // What the cache holds for the current state.  Knots out of order (or nothing built yet): no per-day array at all.
      // Population check failed (binomial.cpp:170-184): the hazard prefix sums, the delay kernel and the series WERE built
      // and committed - they precede the check - only what follows it (convolution, Poisson sums, survey sums, age split,
      // serology term) is missing.  The hot rung (beta = 0) accepts such states all the time; redoing every stage from them
      // made its particles five times as expensive as the others and, with one warp per particle, the critical path of
      // the whole sweep.
      const bool cur_noarrays = (w.recA[L.flags] == 0.0);
      const bool from_invalid = cur_noarrays || (w.recA[L.o_status] != 0.0);  // nothing downstream of the check is cached
      const bool do_sero = valid && ((dep & kDepSero) || cur_noarrays);
      const bool do_gamma = valid && ((dep & kDepGamma) || cur_noarrays);
      const bool do_spline = valid && ((dep & kDepSpline) || cur_noarrays);
      const bool need_conv = do_spline || do_gamma || from_invalid, need_surv = do_spline || do_sero || from_invalid;
      bool new_auc = false;
      double tot_p = tot_cur;
      L1Sums sums_p = sums_cur;
      if (valid) {
        // stage order: the hazard convolution of the seroreversion model borrows the series and kernel buffers, and the
        // gamma table uses the series buffer as scratch, so both run before the series is produced or fetched
        if (do_sero) tile_sero_prefix<NT>(m, L, view(w.recB), tb);
        if (do_gamma) tile_gamma<NT>(m, L, view(w.recB), tb);
        if (do_spline) tot_p = stage_spline<NT>(m, L, w.recB, w.xs);
        else if (need_conv || need_surv) warp_copy(w.xs + kXsLead, cache + CL.xs + kXsLead, CL.nxs - kXsLead);
        // population check (binomial.cpp:170-184): ne[a] * sum(infxn) against the stratum sizes - with the weights and the curve
        // untouched the outcome is the one the current state recorded
        const bool same_popn = (dep & kKeepNe) && !do_spline && !cur_noarrays && !unbuilt;
        const bool ok = same_popn ? (w.recA[L.o_status] == 0.0) : stage_popn(m, L, w.recB, tot_p);
        if (lane == 0) w.recB[L.o_status] = ok ? 0.0 : 1.0;
        __syncwarp();
        if (ok) {
          if (need_conv) {
            if (!do_gamma) {
              warp_copy(w.gs, cache + CL.gs, CL.ngs);
              __syncwarp();
            }
            stage_conv<NT>(m, w.xs, w.gs, w.auc, conv_lag_blocks(m, L, w.recB));
            new_auc = true;
            sums_p = stage_l1_sums(m, tb, w.auc);
            if (sums_p.irregular) {  // some expected count is not a positive normal number: the reference's subnormal arithmetic
              conv_fixup_tiny(m.T, w.xs, w.gs, w.auc);
              sums_p = stage_l1_sums(m, tb, w.auc);
            }
          }
          if (need_surv) {
            if (!do_sero) {
              warp_copy(w.sk, cache + CL.sk, CL.nsk);
              __syncwarp();
            }
            stage_surveys(m, w.xs, w.sk, [&](int sv, double v) { w.recB[L.o_base + sv] = v; });
          }
          if (!new_auc && l1_needs_days(w.recB, sums_p)) {  // rare: the day-by-day Poisson term needs the cached convolution
            warp_copy(w.auc, cache + CL.auc, CL.nauc);
            __syncwarp();
          }
          finish_l1(w.recB, sums_p, w.auc, par);
        }
      }
      __syncwarp();
      // age split and serology term: recomputed only if the proposed parameter can change them
      double l2_p = l2_cur, l3_p = l3_cur;
      const bool p3ok = sweep_phase3(m, L, w.recB, from_invalid || !(dep & kKeepL2), from_invalid || !(dep & kKeepL3), l2_p, l3_p);
      double llp = p3ok ? w.recB[L.o_l1] + (l2_p + l3_p) : -CC_OVERFLO;
      if (!isfinite(llp)) llp = -CC_OVERFLO;
      // prior: only term i (and a Jacobian term if i is a reference parameter) changes
      double pt_new = 0.0, jac_new[3] = {0.0, 0.0, 0.0};
      {
        pt_new = w.q_ptn[i];
#pragma unroll
        for (int j = 0; j < 3; j++) jac_new[j] = (m.jac_rows[j] == i) ? m.jac_counts[j] * log(thp) : w.pterm[d + j];
        double v = 0.0;
        for (int j = lane; j < d; j += 32)
          if (m.prior_family[j] != CC_PRIOR_NONE) v += (j == i) ? pt_new : w.pterm[j];
        if (lane < 3 && m.jac_rows[lane] >= 0) v += (lane == 0) ? jac_new[0] : (lane == 1) ? jac_new[1] : jac_new[2];
        v = warp_sum(v);
        if (!isfinite(v)) v = -CC_OVERFLO;
        const double lpp = v;
        const double mh = beta * (llp - ll) + (lpp - lp) + adj;
        const bool acc = w.q_logu[i] < mh;  // warp-uniform
        __syncwarp();
        if (acc) {
          ll = llp;
          lp = lpp;
          // commit: the proposal's record and the arrays it recomputed become the cached state
          warp_copy(cache + CL.rec, w.recB, CL.nrec);
          if (do_spline) warp_copy(cache + CL.xs + kXsLead, w.xs + kXsLead, CL.nxs - kXsLead);
          if (do_gamma) warp_copy(cache + CL.gs, w.gs, CL.ngs);
          if (new_auc) warp_copy(cache + CL.auc, w.auc, CL.nauc);
          if (do_sero) warp_copy(cache + CL.sk, w.sk, CL.nsk);
          swap_ptr(w.recA, w.recB);
          unbuilt = false;
          tot_cur = tot_p;
          sums_cur = sums_p;
          l2_cur = l2_p;
          l3_cur = l3_p;
          if (lane == 0) {
            cache[CL.scal + 7] = l2_cur; cache[CL.scal + 8] = l3_cur;
            cache[CL.scal + 0] = tot_cur;
            cache[CL.scal + 1] = sums_cur.s1; cache[CL.scal + 2] = sums_cur.sa_obs; cache[CL.scal + 3] = sums_cur.s_all_auc;
            cache[CL.scal + 4] = sums_cur.s_unc; cache[CL.scal + 5] = (double)sums_cur.irregular;
            cache[CL.scal + 6] = 1.0;
            w.pcur[i] = thp;
            w.pterm[i] = pt_new;
            for (int j = 0; j < 3; j++) w.pterm[d + j] = jac_new[j];
            th[i] = thp;
            ph[i] = php;
            macc[i] += 1u;
          }
        }
        __syncwarp();
        if (lane == 0) w.q_logu[i] = acc ? 1.0 : 0.0;  // remembered for the bandwidth update below
        __syncwarp();
      }
    }
    if (adapt) {
      // Robbins-Monro on log(bw) towards a 0.234 acceptance rate, one parameter per lane
      __syncwarp();
      for (int j = lane; j < d; j += 32) {
        const double step = (w.q_logu[j] != 0.0) ? (1.0 - 0.234) : -0.234;
        bw[j] = bw[j] * exp(step / sqrt((double)bwi[j]));  // = exp(log bw + step / sqrt(n))
        bwi[j] += 1;
      }
    }
    if (lane == 0) {
      s.ll[p] = ll;
      s.lp[p] = lp;
    }
  }
}

// One thread per parameter vector (the prior is d cheap terms).
__global__ void logprior_batch_kernel(const KModel m, const double* __restrict__ theta, long long n, long long ld,
                                      double* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  out[v] = thread_logprior(m, theta + v, (int)ld);
}

// Intermediates for the post-processing re-entry points (R/summarize_plot.R:276-884).
__global__ void __launch_bounds__(kEvalThreads) curves_kernel(const KModel m, const double* __restrict__ theta, long long n,
                                                              long long ld, double* __restrict__ o_infxn,
                                                              double* __restrict__ o_ne, double* __restrict__ o_auc,
                                                              double* __restrict__ o_sero, double* __restrict__ o_deaths) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // curves need the hazard over all T days (summarize_plot.R:733-771), so carve with M = T
  KModel mc = m;
  mc.M = m.T;
  mc.S = 0;  // no survey sums needed
  EvalSmem sm = carve_smem(smem_raw, m.d, m.T, m.T);
  const int T = m.T, A = m.A;
  for (long long v = blockIdx.x; v < n; v += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < m.d; i += blockDim.x) sm.p[i] = theta[(long long)i * ld + v];
    __syncthreads();
    (void)block_loglike(mc, sm, false);
    __syncthreads();
    // an invalid knot order leaves no curve: report NaN like an aborted evaluation would
    const bool have = sm.sc->nodex_pass != 0;
    if (o_ne)
      for (int a = threadIdx.x; a < A; a += blockDim.x) o_ne[v * A + a] = sm.sc->ne[a];
    if (o_infxn)
      for (int i = threadIdx.x; i < T; i += blockDim.x) o_infxn[v * T + i] = have ? sm.infxn[i] : d_nan();
    const bool conv_ok = have && sm.sc->popn_pass;
    if (o_auc)
      for (int i = threadIdx.x; i < T; i += blockDim.x) o_auc[v * T + i] = conv_ok ? sm.auc[i] : d_nan();
    if (o_sero) {
      // sero_full[j] = sum_{i<=j} infxn[i] * haz[j-i]; computed whether or not the population check passed (:729)
      if (have) {
        __syncthreads();
        causal_conv(sm.infxn, sm.haz, T, sm.tmp1);
        __syncthreads();
      }
      for (int i = threadIdx.x; i < T; i += blockDim.x) o_sero[v * T + i] = have ? sm.tmp1[i] : d_nan();
    }
    if (o_deaths) {
      // posterior_check_infxns_to_death (R/summarize_plot.R:605-641): gamma DENSITY lookup, then per (day, stratum) the sum over
      // infection days in the reference's order, every product rounded like its `a * b * c` (no contraction)
      __syncthreads();
      const double sod = sm.p[m.idx_sod], mod = sm.p[m.idx_mod];
      const double shape = 1 / (sod * sod), scale = mod * (sod * sod);
      for (int k = threadIdx.x; k <= T; k += blockDim.x) sm.pg[k] = dgamma_density((double)k, shape, scale);
      __syncthreads();
      auto par = [&](int r) { return sm.p[r]; };
      for (int o = threadIdx.x; o < T * A; o += blockDim.x) {
        const int j = o / A, a = o - j * A;
        const double nea = sm.sc->ne[a], ifr = ifr_of(m, par, a);
        double acc = 0.0;
        for (int i = 0; i <= j; i++) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(nea, sm.infxn[i]), ifr), sm.pg[j - i]));
        o_deaths[(v * T + j) * A + a] = have ? acc : d_nan();
      }
    }
  }
}

// Metropolis coupling (sequential over adjacent rung pairs, like drjacoby) + recording of this iteration.
// One CTA per chain.  `row` < 0: do not record.  `do_swap` = 0 for the initial state.
__global__ void mcmc_swap_record_kernel(const KModel m, const McmcState s, int iteration, long long row, int do_swap) {
  const int c = blockIdx.x;
  const int R = s.rungs, d = m.d;
  int* at = s.at_rung + (size_t)c * R;
  if (threadIdx.x == 0 && do_swap && s.coupling_on && R > 1) {
    const bool sampling = iteration > s.burnin;
    for (int k = 0; k < R - 1; k++) {
      const int p1 = c * R + at[k], p2 = c * R + at[k + 1];
      const double l1 = s.ll[p1], l2 = s.ll[p2];
      const double b1 = s.beta[k], b2 = s.beta[k + 1];
      const double a = (l2 - l1) * (b1 - b2);
      const ccrng::Draw dr = ccrng::draw(s.seed, ccrng::kStreamSwap, (uint32_t)(s.chain_offset + c), (uint32_t)k,
                                         (uint32_t)iteration, 0u);
      const bool acc = log(dr.u) < a;
      if (sampling) s.swap_try[(size_t)c * (R - 1) + k] += 1u;
      if (acc) {
        const int t = at[k];
        at[k] = at[k + 1];
        at[k + 1] = t;
        if (sampling) s.swap_acc[(size_t)c * (R - 1) + k] += 1u;
      }
    }
  }
  __syncthreads();
  if (row < 0) return;
  for (int q = 0; q < s.n_rec_rungs; q++) {
    const int r = s.record_all ? q : s.cold_pos;
    const int p = c * R + at[r];
    const size_t o = ((size_t)c * s.n_rec_rungs + q) * (size_t)s.rows_cap + (size_t)row;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s.rec_theta[o * d + i] = s.theta[(size_t)p * d + i];
    if (threadIdx.x == 0) {
      s.rec_ll[o] = s.ll[p];
      s.rec_lp[o] = s.lp[p];
    }
  }
}

// ---- FP64 FMA peak probe: 8 independent register-resident DFMA chains per thread
__global__ void fp64_probe_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double sum = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (sum == 123.456) out[0] = sum;  // keep the chains alive without a store in the common case
}

// ---- FP64 tensor-core (DMMA m8n8k4) probe, optionally interleaved with independent DFMA chains, to learn whether
// the two share a pipe on this part (decides where the Toeplitz contraction should run).
template <int NFMA>
__global__ void dmma_probe_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int u = 0; u < 8; u++) c[u][0] = c[u][1] = threadIdx.x * 1e-9 + u;
  double x[8];
#pragma unroll
  for (int u = 0; u < 8; u++) x[u] = threadIdx.x * 1e-9 + u;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      dmma884(c[u][0], c[u][1], a, b);
#pragma unroll
      for (int f = 0; f < NFMA; f++) x[(u + f) & 7] = fma(x[(u + f) & 7], a, b);
    }
  }
  double sum = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) sum += c[u][0] + c[u][1] + x[u];
  if (sum == 123.456) out[0] = sum;
}

// ------------------------------------------------------------------------------------------------ warp-kernel dispatch
// NT = number of 64-day output tiles per evaluation; instantiated for the shapes in use (T <= 64 NT)
#ifdef CC_NT_ONLY  // developer builds: instantiate a single tile count (faster compile)
#define CC_FOR_EACH_NT(X) X(CC_NT_ONLY)
#else
#define CC_FOR_EACH_NT(X) X(4) X(5) X(7) X(16)
#endif
static int pick_nt(int T) {
  const int need = warp_tiles(T);
#define CC_PICK(NT_) if (need <= NT_) return NT_;
  CC_FOR_EACH_NT(CC_PICK)
#undef CC_PICK
  return -1;
}
static size_t warp_cta_smem(const KModel& k, int nt) {
  return (tile_tables_doubles(k.T) + (size_t)kWarpsPerCta * tile_smem_doubles(k.K, k.A, k.S, nt)) * sizeof(double);
}

static cudaError_t warp_kernels_opt_in(const KModel& k, int nt) {
  const int bytes = (int)warp_cta_smem(k, nt);
  cudaError_t e = cudaSuccess;
#define CC_OPT(NT_)                                                                                                   \
  if (nt == NT_) {                                                                                                    \
    {                                                                                                                 \
      cudaFuncAttributes fp;                                                                                          \
      e = cudaFuncGetAttributes(&fp, loglike_phase2_kernel<NT_>);                                                     \
      if (e == cudaSuccess)                                                                                           \
        e = cudaFuncSetAttribute(loglike_phase2_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                 227 * 1024 - (int)fp.sharedSizeBytes);                                               \
    }                                                                                                                 \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mcmc_init_tile_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);  \
    if (e == cudaSuccess) {                                                                                           \
      cudaFuncAttributes fa;                                                                                          \
      e = cudaFuncGetAttributes(&fa, mcmc_sweep_cached_kernel<NT_>);                                                  \
      if (e == cudaSuccess)                                                                                           \
        e = cudaFuncSetAttribute(mcmc_sweep_cached_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                 227 * 1024 - (int)fa.sharedSizeBytes);                                               \
    }                                                                                                                 \
  }
  CC_FOR_EACH_NT(CC_OPT)
#undef CC_OPT
  return e;
}

// ------------------------------------------------------------------------------------------------ model
static double host_stirlerr(double n) {
  const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778, S2 = 0.00079365079365079365079365,
               S3 = 0.000595238095238095238095238, S4 = 0.0008417508417508417508417508;
  if (n <= 15.0) {
    double nn = n + n;
    if (nn == (double)(int)nn) return h_sferr_halves[(int)nn];
    return std::lgamma(n + 1.) - (n + 0.5) * std::log(n) + n - CC_LN_SQRT_2PI;
  }
  double nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

template <typename T>
static int upload(cc_model* m, const std::vector<T>& h, const T** dptr) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  CC_CUDA(cudaMalloc(&p, bytes));
  m->dev_allocs.push_back(p);
  if (!h.empty()) CC_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = static_cast<const T*>(p);
  return CC_OK;
}

static bool idx_ok(int i, int d) { return i >= 0 && i < d; }

extern "C" int cc_model_create(const cc_model_desc* de, int device, cc_model** out) {
  if (!de || !out) return fail(CC_E_INVALID, "cc_model_create: null argument");
  *out = nullptr;
  if (de->abi_version != CC_ABI_VERSION) return fail(CC_E_INVALID, "cc_model_create: abi_version mismatch");
  const int d = de->n_params, A = de->n_strata, K = de->n_knots, T = de->days_obsd, S = de->n_sero_obs, M = de->max_seroday_obsd;
  if (d < 1 || d > CC_MAX_PARAMS) return fail(CC_E_INVALID, "n_params out of range [1, CC_MAX_PARAMS]");
  if (A < 1 || A > CC_MAX_STRATA) return fail(CC_E_INVALID, "n_strata out of range [1, CC_MAX_STRATA]");
  if (K < 3 || K > CC_MAX_KNOTS) return fail(CC_E_INVALID, "n_knots out of range [3, CC_MAX_KNOTS]");
  if (T < 1 || T > CC_MAX_DAYS) return fail(CC_E_INVALID, "days_obsd out of range [1, CC_MAX_DAYS]");
  if (S < 1 || S > CC_MAX_SEROSURVEYS) return fail(CC_E_INVALID, "n_sero_obs out of range [1, CC_MAX_SEROSURVEYS]");
  if (M < 1 || M > T) return fail(CC_E_INVALID, "max_seroday_obsd must be in [1, days_obsd] (the reference reads infxn_spline[0..M-1])");
  if (!de->demog || !de->obs_deaths || !de->prop_strata_obs_deaths || !de->sero_a || !de->sero_b || !de->sero_survey_start ||
      !de->sero_survey_end || !de->idx_knots || !de->idx_infxn || !de->idx_ifr || !de->pmin || !de->pmax || !de->prior_family ||
      !de->prior_p1 || !de->prior_p2)
    return fail(CC_E_INVALID, "cc_model_create: null array in descriptor");
  if (A > 1 && !de->idx_noise) return fail(CC_E_INVALID, "idx_noise is required when n_strata > 1");
  for (int s = 0; s < S; s++)
    if (de->sero_survey_start[s] < 1 || de->sero_survey_end[s] < de->sero_survey_start[s] || de->sero_survey_end[s] > M)
      return fail(CC_E_INVALID, "serosurvey windows must satisfy 1 <= start <= end <= max_seroday_obsd");
  bool ok = idx_ok(de->idx_sens, d) && idx_ok(de->idx_spec, d) && idx_ok(de->idx_sero_con_rate, d) && idx_ok(de->idx_mod, d) &&
            idx_ok(de->idx_sod, d);
  if (de->account_serorev) ok = ok && idx_ok(de->idx_sero_rev_shape, d) && idx_ok(de->idx_sero_rev_scale, d);
  for (int i = 0; i < K - 1; i++) ok = ok && idx_ok(de->idx_knots[i], d);
  for (int i = 0; i < K; i++) ok = ok && idx_ok(de->idx_infxn[i], d);
  for (int a = 0; a < A; a++) ok = ok && idx_ok(de->idx_ifr[a], d);
  if (A > 1)
    for (int a = 0; a < A; a++) ok = ok && idx_ok(de->idx_noise[a], d);
  if (de->reparam_knots) ok = ok && idx_ok(de->idx_rel_knot, d);
  if (de->reparam_infxn) ok = ok && idx_ok(de->idx_rel_infxn, d);
  if (de->reparam_ifr) ok = ok && idx_ok(de->idx_max_ma, d);
  for (int j = 0; j < 3; j++) ok = ok && (de->jac_rows[j] < d);
  if (!ok) return fail(CC_E_INVALID, "role map index out of range");
  for (int i = 0; i < d; i++)
    if (de->prior_family[i] < CC_PRIOR_NONE || de->prior_family[i] > CC_PRIOR_DBETA) return fail(CC_E_INVALID, "unknown prior family");

  int ndev = cc_device_count();
  if (ndev <= 0) return fail(CC_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(CC_E_INVALID, "device index out of range");
  CC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CC_CUDA(cudaGetDeviceProperties(&prop, device));

  cc_model* m = new (std::nothrow) cc_model();
  if (!m) return fail(CC_E_NOMEM, "out of host memory");
  m->device = device;
  m->sm_count = prop.multiProcessorCount;
  KModel& k = m->k;
  k.d = d; k.A = A; k.K = K; k.T = T; k.S = S; k.M = M;
  k.rcensor_day = de->rcensor_day;
  {
    const char* gm = std::getenv("CC_GAMMA_MODE");  // developer switch: 1 = always use the general incomplete-gamma algorithm
    k.gamma_mode = (gm && gm[0] == '1') ? 1 : 0;
    const char* ds = std::getenv("CC_DEBUG_STOP");
    k.debug_stop = ds ? std::atoi(ds) : 0;
    k.debug_dump = nullptr;
  }
  k.serorev = de->account_serorev ? 1 : 0;
  k.binomial = de->binomial_likelihood ? 1 : 0;
  k.reparam_ifr = de->reparam_ifr ? 1 : 0;
  k.reparam_knots = de->reparam_knots ? 1 : 0;
  k.reparam_infxn = de->reparam_infxn ? 1 : 0;
  k.idx_sens = de->idx_sens; k.idx_spec = de->idx_spec; k.idx_rate = de->idx_sero_con_rate;
  // unused role rows are 0, not -1: the flags decide whether they are read, and a compiler-speculated load stays in bounds
  k.idx_shape = de->account_serorev ? de->idx_sero_rev_shape : 0;
  k.idx_scale = de->account_serorev ? de->idx_sero_rev_scale : 0;
  k.idx_mod = de->idx_mod; k.idx_sod = de->idx_sod;
  k.idx_rel_knot = de->reparam_knots ? de->idx_rel_knot : 0;
  k.idx_rel_infxn = de->reparam_infxn ? de->idx_rel_infxn : 0;
  k.idx_max_ma = de->reparam_ifr ? de->idx_max_ma : 0;
  for (int i = 0; i < K - 1; i++) k.idx_knots[i] = (short)de->idx_knots[i];
  for (int i = 0; i < K; i++) k.idx_infxn[i] = (short)de->idx_infxn[i];
  for (int a = 0; a < A; a++) {
    k.idx_ifr[a] = (short)de->idx_ifr[a];
    k.idx_noise[a] = (short)(A > 1 ? de->idx_noise[a] : 0);
    k.demog[a] = (int)de->demog[a];  // binomial.cpp:10 (Rcpp::as<std::vector<int>>)
  }
  // L2 probabilities are data-only: paobsd[a] / running paobsd_sum (binomial.cpp:238-250)
  {
    double psum = 0.0;
    for (int a = 0; a < A; a++) psum += de->prop_strata_obs_deaths[a];
    for (int a = 0; a < A; a++) {
      k.l2_prob[a] = de->prop_strata_obs_deaths[a] / psum;
      psum -= de->prop_strata_obs_deaths[a];
    }
  }
  for (int s = 0; s < S; s++) {
    k.sero_start[s] = de->sero_survey_start[s];
    k.sero_end[s] = de->sero_survey_end[s];
  }
  for (int j = 0; j < 3; j++) {
    k.jac_rows[j] = de->jac_rows[j];
    k.jac_counts[j] = de->jac_counts[j];
  }
  if (!de->reparam_ifr) k.jac_rows[0] = -1;
  if (!de->reparam_knots) k.jac_rows[1] = -1;
  if (!de->reparam_infxn) k.jac_rows[2] = -1;

  std::vector<int> obsd(T);
  std::vector<double> pois_c(T), lgx1(T);
  for (int i = 0; i < T; i++) {
    obsd[i] = (int)de->obs_deaths[i];  // binomial.cpp:204
    const double x = (double)obsd[i];
    pois_c[i] = (x > 0) ? (-0.5 * std::log(CC_2PI * x) - host_stirlerr(x)) : 0.0;
    lgx1[i] = (x >= 0) ? std::lgamma(x + 1.0) : 0.0;
  }
  std::vector<double> logk(T + 1, 0.0);
  for (int i = 1; i <= T; i++) logk[i] = std::log((double)i);
  std::vector<LogTab> logtab(128);
  for (int i = 0; i < 128; i++) {
    const float c = (float)(1.0 / (1.0 + ((double)i + 0.5) / 128.0));  // 24-bit reciprocal of the bin centre
    logtab[i].c = (double)c;
    logtab[i].mlogc = (double)(-logl((long double)c));
  }
  std::vector<double> rk(std::max(T, 640) + 2, 0.0);
  for (size_t i = 1; i < rk.size(); i++) rk[i] = 1.0 / (double)i;
  {
    long double c = 0.0L;
    for (int i = 0; i < T; i++)
      if ((i + 1) < k.rcensor_day && obsd[i] >= 0) c += lgammal((long double)obsd[i] + 1.0L);
    k.l1_const = (double)c;
    double sx = 0.0;
    for (int i = 0; i < T; i++)
      if ((i + 1) < k.rcensor_day && obsd[i] > 0) sx += (double)obsd[i];
    k.l1_sumx = sx;
  }
  std::vector<int> dep(d, 0);
  {
    auto mark = [&](int row, int bit) { if (row >= 0 && row < d) dep[row] |= bit; };
    for (int a = 0; a < A; a++) {
      mark(de->idx_ifr[a], kDepWeights | kKeepL3);
      if (A > 1) mark(de->idx_noise[a], kDepWeights);
    }
    mark(de->idx_sens, kDepWeights | kKeepL2);
    mark(de->idx_spec, kDepWeights | kKeepL2);
    if (de->reparam_ifr) mark(de->idx_max_ma, kDepWeights | kKeepL3);
    for (int i = 0; i < K - 1; i++) mark(de->idx_knots[i], kDepSpline);
    for (int i = 0; i < K; i++) mark(de->idx_infxn[i], kDepSpline);
    if (de->reparam_knots) mark(de->idx_rel_knot, kDepSpline);
    if (de->reparam_infxn) mark(de->idx_rel_infxn, kDepSpline);
    mark(de->idx_mod, kDepGamma | kKeepL3);
    mark(de->idx_sod, kDepGamma | kKeepL3);
    mark(de->idx_sero_con_rate, kDepSero | kKeepL2);
    if (de->account_serorev) {
      mark(de->idx_sero_rev_shape, kDepSero | kKeepL2);
      mark(de->idx_sero_rev_scale, kDepSero | kKeepL2);
    }
    // a row that plays two roles (not possible through the R interface, but the descriptor allows it) keeps nothing
    for (int i = 0; i < d; i++) {
      const int roles = ((dep[i] & kDepWeights) ? 1 : 0) + ((dep[i] & kDepSpline) ? 1 : 0) + ((dep[i] & kDepGamma) ? 1 : 0) + ((dep[i] & kDepSero) ? 1 : 0);
      if (roles != 1 || ((dep[i] & kKeepL2) && (dep[i] & kKeepL3))) dep[i] &= ~(kKeepL2 | kKeepL3);
      bool noise = false;
      for (int a = 0; a < A && A > 1; a++) noise = noise || (de->idx_noise[a] == i);
      if (!noise) dep[i] |= kKeepNe;
    }
  }
  std::vector<double> sa(de->sero_a, de->sero_a + (size_t)S * A), sb(de->sero_b, de->sero_b + (size_t)S * A);
  std::vector<int> l3_kind((size_t)S * A, 0);
  std::vector<double> l3_lch((size_t)S * A, 0.0);
  for (size_t q = 0; q < (size_t)S * A; q++) {
    const double x = sa[q], nn = sb[q];
    if (x == -1 && nn == -1) l3_kind[q] = 0;  // binomial.cpp:334
    else if (std::isfinite(x) && std::isfinite(nn) && x == std::floor(x) && nn == std::floor(nn) && x >= 0 && x <= nn && nn < 9e15) {
      l3_kind[q] = 1;
      l3_lch[q] = (double)(lgammal((long double)nn + 1.0L) - lgammal((long double)x + 1.0L) - lgammal((long double)(nn - x) + 1.0L));
    } else l3_kind[q] = 2;
  }
  std::vector<double> pmin(de->pmin, de->pmin + d), pmax(de->pmax, de->pmax + d);
  std::vector<int> fam(de->prior_family, de->prior_family + d);
  std::vector<double> p1(de->prior_p1, de->prior_p1 + d), p2(de->prior_p2, de->prior_p2 + d);
  int rc = CC_OK;
  if ((rc = upload(m, obsd, &k.obsd)) || (rc = upload(m, pois_c, &k.pois_c)) || (rc = upload(m, lgx1, &k.lgx1)) ||
      (rc = upload(m, logk, &k.logk)) || (rc = upload(m, rk, &k.rk)) || (rc = upload(m, dep, &k.dep)) || (rc = upload(m, logtab, &k.logtab)) || (rc = upload(m, l3_kind, &k.l3_kind)) || (rc = upload(m, l3_lch, &k.l3_lchoose)) ||
      (rc = upload(m, sa, &k.sero_a)) || (rc = upload(m, sb, &k.sero_b)) || (rc = upload(m, pmin, &k.pmin)) ||
      (rc = upload(m, pmax, &k.pmax)) || (rc = upload(m, fam, &k.prior_family)) || (rc = upload(m, p1, &k.prior_p1)) ||
      (rc = upload(m, p2, &k.prior_p2))) {
    cc_model_destroy(m);
    return rc;
  }
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream3, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev1);
  for (int q = 0; q < 3 && e == cudaSuccess; q++) e = cudaEventCreateWithFlags(&m->ev_stage[q], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_slot0, cudaEventDisableTiming);
  // opt in to large dynamic shared memory once
  m->nt = pick_nt(T);
  if (e == cudaSuccess && m->nt < 0) {
    cc_model_destroy(m);
    return fail(CC_E_INVALID, "days_obsd too large for the instantiated kernels");
  }
  if (e == cudaSuccess) e = warp_kernels_opt_in(k, m->nt);
  if (e == cudaSuccess) {
    // shape of the per-day pass: the candidate with the most resident warps per SM that fits this model's shared memory
    const size_t per_warp = warp_cta_smem(k, m->nt) / kWarpsPerCta;
    P2Shape cand[24];
    const int nc = p2_candidates(m->nt, cand);
    int best_w = 0, best_bps = 0;
    for (int c = 0; c < nc && e == cudaSuccess; c++) {
      const size_t bytes = per_warp * cand[c].warps;
      if (bytes > 227 * 1024 - 1040) continue;
      int bps = 0;
      cudaError_t oe = cudaSuccess;
      switch (m->nt) {
#define CC_OCC(NT_) case NT_: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, loglike_phase2_kernel<NT_>, 32 * cand[c].warps, bytes); break;
        CC_FOR_EACH_NT(CC_OCC)
#undef CC_OCC
        default: break;
      }
      if (oe != cudaSuccess) { cudaGetLastError(); continue; }
      bps = std::min(bps, cand[c].ctas);
      if (bps * cand[c].warps > best_bps * best_w) { best_w = cand[c].warps; best_bps = bps; }
    }
    if (e == cudaSuccess && best_w < 1) {
      cc_model_destroy(m);
      return fail(CC_E_INVALID, "model too large for one CTA's shared memory");
    }
    m->p2_warps = best_w;
    m->grid_p2 = std::max(1, best_bps) * m->sm_count;
    m->grid_max = std::max(1, best_bps * best_w / kWarpsPerCta) * m->sm_count;  // in 4-warp CTAs (initial-state kernel of the sweep)
    if (std::getenv("CC_VERBOSE"))
      std::fprintf(stderr, "covidcurve_b200: per-day pass %d CTAs/SM x %d warps, %zu B smem/CTA, persistent grid %d\n", best_bps, best_w,
                   per_warp * best_w, m->grid_p2);
  }
  const size_t smem = eval_smem_bytes(d, T, T);
  if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(curves_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cc_model_destroy(m);
    return fail(CC_E_CUDA, std::string("cc_model_create: ") + cudaGetErrorString(e));
  }
  if (std::getenv("CC_DEBUG_DUMP")) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)4 * T * sizeof(double)) == cudaSuccess) {
      cudaMemset(p, 0, (size_t)4 * T * sizeof(double));
      m->dev_allocs.push_back(p);
      m->k.debug_dump = static_cast<double*>(p);
    }
  }
  *out = m;
  return CC_OK;
}

// developer aid (not part of the public header): intermediates of vector 0 of the last batch; needs CC_DEBUG_DUMP=1 at model creation
extern "C" int cc_debug_fetch(cc_model* m, double* out4T) {
  if (!m || !m->k.debug_dump || !out4T) return CC_E_INVALID;
  cudaDeviceSynchronize();
  return cudaMemcpy(out4T, m->k.debug_dump, (size_t)4 * m->k.T * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess ? CC_OK : CC_E_CUDA;
}

extern "C" void cc_model_destroy(cc_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (void* p : m->dev_allocs) cudaFree(p);
  for (int q = 0; q < 4; q++) {
    if (m->work[q].rec) cudaFree(m->work[q].rec);
    if (m->work[q].key) cudaFree(m->work[q].key);
    if (m->work[q].perm) cudaFree(m->work[q].perm);
    if (m->work[q].counts) cudaFree(m->work[q].counts);
  }
  if (m->d_stage) cudaFree(m->d_stage);
  if (m->d_out) cudaFree(m->d_out);
  if (m->h_pin) cudaFreeHost(m->h_pin);
  for (int q = 0; q < 3; q++)
    if (m->ev_stage[q]) cudaEventDestroy(m->ev_stage[q]);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->ev_slot0) cudaEventDestroy(m->ev_slot0);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->stream2) cudaStreamDestroy(m->stream2);
  if (m->stream3) cudaStreamDestroy(m->stream3);
  delete m;
}

// ------------------------------------------------------------------------------------------------ batch calls
static int ensure_work(cc_model* m, int slot, long long n) {
  cc_model::Work& wk = m->work[slot];
  if (!wk.counts) CC_CUDA(cudaMalloc((void**)&wk.counts, (2 * kNumKeys + 8) * sizeof(unsigned int)));
  if (wk.cap >= n) return CC_OK;
  if (wk.rec) cudaFree(wk.rec);
  if (wk.key) cudaFree(wk.key);
  if (wk.perm) cudaFree(wk.perm);
  wk.rec = nullptr; wk.key = nullptr; wk.perm = nullptr; wk.cap = 0;
  const long long cap = (n + 1023) & ~1023LL;
  const RecLayout L = rec_layout(m->k.K, m->k.A, m->k.S);
  CC_CUDA(cudaMalloc((void**)&wk.rec, (size_t)L.F * cap * sizeof(double)));
  CC_CUDA(cudaMalloc((void**)&wk.key, (size_t)cap * sizeof(int)));
  CC_CUDA(cudaMalloc((void**)&wk.perm, (size_t)cap * sizeof(int)));
  wk.cap = cap;
  return CC_OK;
}

// the batched likelihood: pass A (thread per vector) -> regime buckets -> pass C (warp per vector) -> pass D; 5 launches on one stream
static int launch_loglike(cc_model* m, const double* d_theta, long long n, long long ld, double* d_out, cudaStream_t st, int slot) {
  const KModel& k = m->k;
  if (n > 0x7fffffffLL) return fail(CC_E_INVALID, "batch too large (2^31 vectors per call)");
  int rc = ensure_work(m, slot, n);
  if (rc) return rc;
  cc_model::Work& wk = m->work[slot];
  unsigned int* counts = wk.counts;
  unsigned int* cursor = wk.counts + kNumKeys;
  CC_CUDA(cudaMemsetAsync(counts, 0, (2 * kNumKeys + 8) * sizeof(unsigned int), st));
  const unsigned gridA = (unsigned)((n + 127) / 128);
  const size_t smem = warp_cta_smem(k, m->nt);
  const long long tiles = (n + kPhase2Tile - 1) / kPhase2Tile;
  const int p2w = m->p2_warps;
  const unsigned gridC = (unsigned)std::min<long long>((tiles + p2w - 1) / p2w, (long long)m->grid_p2);
  switch (m->nt) {
#define CC_LAUNCH(NT_)                                                                                                        \
  case NT_:                                                                                                                   \
    loglike_phase1_kernel<NT_><<<gridA, 128, 0, st>>>(k, d_theta, n, ld, wk.rec, wk.cap, wk.key, counts);                     \
    bucket_offsets_kernel<<<1, 1, 0, st>>>(counts, cursor);                                                                   \
    bucket_scatter_kernel<<<gridA, 128, 0, st>>>(wk.key, n, cursor, wk.perm);                                                 \
    loglike_phase2_kernel<NT_><<<gridC, 32 * p2w, smem / kWarpsPerCta * p2w, st>>>(k, n, wk.rec, wk.cap, wk.perm, counts + 2 * kNumKeys, d_theta, ld);                      \
    loglike_phase3_kernel<<<gridA, 128, 0, st>>>(k, n, wk.rec, wk.cap, d_out);                                                \
    break;
    CC_FOR_EACH_NT(CC_LAUNCH)
#undef CC_LAUNCH
    default: return fail(CC_E_INVALID, "unsupported days_obsd for the instantiated kernels");
  }
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

static int launch_logprior(cc_model* m, const double* d_theta, long long n, long long ld, double* d_out, cudaStream_t st, int /*slot*/) {
  const int bs = 128;
  const long long grid = (n + bs - 1) / bs;
  logprior_batch_kernel<<<(unsigned)grid, bs, 0, st>>>(m->k, d_theta, n, ld, d_out);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

static int ensure_stage(cc_model* m, size_t theta_bytes, size_t out_bytes) {
  if (m->d_stage_bytes < theta_bytes) {
    if (m->d_stage) cudaFree(m->d_stage);
    m->d_stage = nullptr;
    m->d_stage_bytes = 0;
    CC_CUDA(cudaMalloc((void**)&m->d_stage, theta_bytes));
    m->d_stage_bytes = theta_bytes;
  }
  if (m->d_out_bytes < out_bytes) {
    if (m->d_out) cudaFree(m->d_out);
    m->d_out = nullptr;
    m->d_out_bytes = 0;
    CC_CUDA(cudaMalloc((void**)&m->d_out, out_bytes));
    m->d_out_bytes = out_bytes;
  }
  return CC_OK;
}

typedef int (*launch_fn)(cc_model*, const double*, long long, long long, double*, cudaStream_t, int);

// true when `p` is page-locked (or managed) host memory the copy engine can read directly
static bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// A few long-lived helper threads for packing pageable caller memory into the pinned ring (one thread moves ~10 GB/s, the
// GPU consumes 272-byte vectors at 90 M evaluations/s = 24 GB/s; spawning threads per chunk cost ~10 % of a 128k-vector chunk).
class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool p;
    return p;
  }
  int size() const { return (int)th_.size() + 1; }
  // run job(t) for t = 0 .. n-1 (n <= size()); job 0 on the calling thread
  void run(int n, const std::function<void(int)>& job) {
    if (n <= 1 || getpid() != pid_) {  // (a forked child has no helper threads: the caller drains the work queue alone)
      job(0);
      return;
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      job_ = &job;
      njobs_ = n;
      pending_ = n - 1;
      gen_++;
    }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> g(mu_);
    done_.wait(g, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  CopyPool() : pid_(getpid()) {
    unsigned hc = std::thread::hardware_concurrency();
    int n = (int)std::min(8u, std::max(1u, hc / 2));
    if (const char* e = std::getenv("CC_COPY_THREADS")) n = std::max(1, std::min(32, std::atoi(e)));
    for (int t = 1; t < n; t++) th_.emplace_back([this, t] { loop(t); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void loop(int t) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (t < njobs_) job = job_;
      }
      if (job) {
        (*job)(t);
        std::lock_guard<std::mutex> g(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int njobs_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  const pid_t pid_;
};

// row-wise host copy dst[r * dld + 0..n) <- src[r * sld + 0..n), r < rows, split into (row, column-block) pieces over the pool
static void host_copy_rows(double* dst, size_t dld, const double* src, size_t sld, size_t n, int rows) {
  const size_t bytes = n * sizeof(double) * (size_t)rows;
  CopyPool& pool = CopyPool::get();
  const int nth = (int)std::min<size_t>((size_t)pool.size(), std::max<size_t>(1, bytes >> 21));
  const size_t blk = 1 << 15;                                   // doubles per piece (256 kB)
  const size_t nblk = (n + blk - 1) / blk, pieces = nblk * (size_t)rows;
  std::atomic<size_t> next{0};
  pool.run(nth, [&](int) {
    for (size_t p = next.fetch_add(1); p < pieces; p = next.fetch_add(1)) {
      const size_t r = p / nblk, c0 = (p - r * nblk) * blk, cn = std::min(blk, n - c0);
      std::memcpy(dst + r * dld + c0, src + r * sld + c0, cn * sizeof(double));
    }
  });
}

// Host-buffer path: column chunks are copied H2D (d rows of `chunk` doubles), evaluated and copied back, triple-buffered
// over three streams so that the copy engine never waits for a staging slot: while chunk k computes, chunk k+1 is uploaded
// and chunk k+2 can already start uploading (with two slots the upload of k+2 had to wait for the kernels of k, which cost
// 25 % when four ranks share the host's PCIe bandwidth).
//   * page-locked caller memory (cudaHostAlloc / cudaHostRegister, torch pin_memory) is read and written by the copy engine
//     in place (2-D copies);
//   * PAGEABLE caller memory (malloc, an R vector's REAL()) is staged through the library's own pinned ring: the calling
//     thread packs chunk k+1 into a pinned slot while the GPU works on chunk k, and unpacks finished outputs.
static int run_host_batch(cc_model* m, launch_fn fn, const double* theta, long long n, long long ld, double* out) {
  const int d = m->k.d;
  constexpr int kSlots = 3;
  const bool pinned_in = host_ptr_is_pinned(theta), pinned_out = host_ptr_is_pinned(out);
  // Chunk sizes grow geometrically (64k, 128k, ... up to 512k vectors): only the first chunk's upload and the last
  // chunk's download are exposed, so they are kept small, while the bulk runs in large launches (measured on B200 at
  // 2^20 vectors: fixed 128k chunks 56.8 M evals/s, 256k 58.6, 512k 60.1; profiles/r1_summary.md).  The pinned ring of the
  // pageable path is capped at 128k vectors per slot (d x 1 MB of page-locked memory each).
  long long cmax = pinned_in ? (1 << 19) : (1 << 17);
  if (const char* e = std::getenv("CC_HOST_CHUNK")) cmax = std::max<long long>(1024, (std::atoll(e) + 1023) & ~1023LL);  // developer tuning
  cmax = std::min<long long>(cmax, (n + 1023) & ~1023LL);
  int rc = ensure_stage(m, (size_t)kSlots * d * cmax * sizeof(double), (size_t)kSlots * cmax * sizeof(double));
  if (rc) return rc;
  const size_t pin_slot = (size_t)(d + 1) * cmax;  // doubles: [d][cmax] parameters, then [cmax] outputs
  if (!pinned_in || !pinned_out) {
    const size_t need = (size_t)kSlots * pin_slot * sizeof(double);
    if (m->h_pin_bytes < need) {
      if (m->h_pin) cudaFreeHost(m->h_pin);
      m->h_pin = nullptr;
      m->h_pin_bytes = 0;
      CC_CUDA(cudaHostAlloc((void**)&m->h_pin, need, cudaHostAllocDefault));
      m->h_pin_bytes = need;
    }
  }
  if (fn == launch_loglike)  // size the work slots for the largest chunk now: growing them mid-pipeline would synchronise the device
    for (int slot = 1; slot <= kSlots; slot++)
      if ((rc = ensure_work(m, slot, std::min(cmax, n)))) return rc;
  cudaStream_t st[kSlots] = {m->stream, m->stream2, m->stream3};
  // on any failure: let the copies already queued into / out of caller memory finish before the caller sees the error
  auto bail = [&](int code) {
    for (int j = 0; j < kSlots; j++) cudaStreamSynchronize(st[j]);
    return code;
  };
#define CC_CUDA_BAIL(expr)                                                                                 \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) return bail(fail(CC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)
  CC_CUDA_BAIL(cudaEventRecord(m->ev0, st[0]));
  m->last_launches = 0;
  struct Pending { long long c0 = 0, cn = 0; bool busy = false; } pend[kSlots];
  auto drain = [&](int q) -> int {  // outputs of the chunk that last used slot q: pinned slot -> caller memory
    if (!pend[q].busy) return CC_OK;
    CC_CUDA_BAIL(cudaEventSynchronize(m->ev_stage[q]));
    if (!pinned_out) std::memcpy(out + pend[q].c0, m->h_pin + (size_t)q * pin_slot + (size_t)d * cmax, (size_t)pend[q].cn * sizeof(double));
    pend[q].busy = false;
    return CC_OK;
  };
  int q = 0;
  long long chunk = std::min<long long>(cmax, 1 << 16);
  for (long long c0 = 0; c0 < n; q = (q + 1) % kSlots) {
    long long cn = std::min(chunk, n - c0);
    if (n - c0 - cn < (1 << 15)) cn = std::min(n - c0, cmax);  // do not leave a sliver behind
    double* dth = m->d_stage + (size_t)q * d * cmax;
    double* dout = m->d_out + (size_t)q * cmax;
    double* hp = m->h_pin ? m->h_pin + (size_t)q * pin_slot : nullptr;
    if ((!pinned_in || !pinned_out) && (rc = drain(q))) return rc;  // the slot's previous chunk has left the pinned ring
    if (pinned_in) {
      CC_CUDA_BAIL(cudaMemcpy2DAsync(dth, (size_t)cmax * sizeof(double), theta + c0, (size_t)ld * sizeof(double),
                                     (size_t)cn * sizeof(double), (size_t)d, cudaMemcpyHostToDevice, st[q]));
    } else {
      host_copy_rows(hp, (size_t)cmax, theta + c0, (size_t)ld, (size_t)cn, d);
      CC_CUDA_BAIL(cudaMemcpy2DAsync(dth, (size_t)cmax * sizeof(double), hp, (size_t)cmax * sizeof(double), (size_t)cn * sizeof(double),
                                     (size_t)d, cudaMemcpyHostToDevice, st[q]));
    }
    rc = fn(m, dth, cn, cmax, dout, st[q], 1 + q);
    if (rc) return bail(rc);
    m->last_launches += (fn == launch_loglike) ? kLoglikeLaunches : 1;
    CC_CUDA_BAIL(cudaMemcpyAsync(pinned_out ? out + c0 : hp + (size_t)d * cmax, dout, (size_t)cn * sizeof(double), cudaMemcpyDeviceToHost, st[q]));
    CC_CUDA_BAIL(cudaEventRecord(m->ev_stage[q], st[q]));
    pend[q].c0 = c0; pend[q].cn = cn; pend[q].busy = true;
    c0 += cn;
    chunk = std::min<long long>(2 * chunk, cmax);
  }
  for (int j = 1; j < kSlots; j++)
    if (pend[j].busy) CC_CUDA_BAIL(cudaStreamWaitEvent(st[0], m->ev_stage[j], 0));
  CC_CUDA_BAIL(cudaEventRecord(m->ev1, st[0]));
  for (int j = 0; j < kSlots; j++)
    if ((rc = drain(j))) return rc;
  CC_CUDA_BAIL(cudaStreamSynchronize(st[0]));
#undef CC_CUDA_BAIL
  float ms = 0.f;
  CC_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
  m->last_ms = ms;
  return CC_OK;
}

static int batch_call(cc_model* m, launch_fn fn, const double* theta, int64_t n, int64_t ld, double* out, int mem, void* stream,
                      const char* who) {
  if (!m || !theta || !out) return fail(CC_E_INVALID, std::string(who) + ": null argument");
  if (n < 0 || ld < n) return fail(CC_E_INVALID, std::string(who) + ": need 0 <= n <= ld");
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(m->device));
  if (mem == CC_MEM_DEVICE) {
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    // device-pointer calls share work slot 0 (records, regime buckets, tile ticket): a call on another stream than the
    // previous one first waits for it, so calls on one model never overlap whatever streams the caller uses
    if (m->slot0_used && m->slot0_stream != st) CC_CUDA(cudaStreamWaitEvent(st, m->ev_slot0, 0));
    CC_CUDA(cudaEventRecord(m->ev0, st));
    int rc = fn(m, theta, n, ld, out, st, 0);
    if (rc) return rc;
    CC_CUDA(cudaEventRecord(m->ev1, st));
    CC_CUDA(cudaEventRecord(m->ev_slot0, st));
    m->slot0_used = true;
    m->slot0_stream = st;
    m->last_launches = (fn == launch_loglike) ? kLoglikeLaunches : 1;
    m->last_ms = -1.0;  // resolved lazily by cc_last_kernel_ms
    return CC_OK;
  }
  if (mem != CC_MEM_HOST) return fail(CC_E_INVALID, std::string(who) + ": mem must be CC_MEM_HOST or CC_MEM_DEVICE");
  return run_host_batch(m, fn, theta, n, ld, out);
}

extern "C" int cc_loglike_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* out, int mem, void* stream) {
  return batch_call(m, launch_loglike, theta, n, ld, out, mem, stream, "cc_loglike_batch");
}

extern "C" int cc_logprior_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* out, int mem, void* stream) {
  return batch_call(m, launch_logprior, theta, n, ld, out, mem, stream, "cc_logprior_batch");
}

extern "C" int cc_loglike(cc_model* m, const double* params, int /*param_i: ignored like binomial.cpp:7*/, double* out) {
  return batch_call(m, launch_loglike, params, 1, 1, out, CC_MEM_HOST, nullptr, "cc_loglike");
}

extern "C" int cc_logprior(cc_model* m, const double* params, int /*param_i*/, double* out) {
  return batch_call(m, launch_logprior, params, 1, 1, out, CC_MEM_HOST, nullptr, "cc_logprior");
}

// ---- the reference's 4-argument call shape (header: cc_call_args): role map from the names, model handles cached by content
namespace {
struct CallCacheEntry { uint64_t key = 0; int device = -1; cc_model* model = nullptr; uint64_t stamp = 0; };
std::mutex g_call_mu;
CallCacheEntry g_call_cache[4];
uint64_t g_call_stamp = 0;

uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
  return h;
}
int find_name(const cc_call_args* a, const std::string& nm) {
  for (int i = 0; i < a->n_params; i++)
    if (a->names[i] && nm == a->names[i]) return i;
  return -1;
}
}  // namespace

extern "C" int cc_loglike_call(const cc_call_args* a, const double* params, int param_i, int device, double* out) {
  if (!a || !params || !out) return fail(CC_E_INVALID, "cc_loglike_call: null argument");
  if (!a->names || !a->obs_deaths || !a->prop_strata_obs_deaths || !a->sero_a || !a->sero_b || !a->demog || !a->sero_survey_start ||
      !a->sero_survey_end)
    return fail(CC_E_INVALID, "cc_loglike_call: null field in cc_call_args");
  const int d = a->n_params, A = a->n_strata, K = a->n_knots, T = a->days_obsd, S = a->n_sero_obs;
  if (d < 1 || d > CC_MAX_PARAMS || A < 1 || A > CC_MAX_STRATA || K < 3 || K > CC_MAX_KNOTS || T < 1 || T > CC_MAX_DAYS || S < 1 ||
      S > CC_MAX_SEROSURVEYS)
    return fail(CC_E_INVALID, "cc_loglike_call: sizes out of range");
  uint64_t h = 1469598103934665603ULL;
  for (int i = 0; i < d; i++) {
    if (!a->names[i]) return fail(CC_E_INVALID, "cc_loglike_call: params must be a fully named vector");
    h = fnv1a(h, a->names[i], std::strlen(a->names[i]) + 1);
  }
  const int32_t scal[9] = {d, A, K, a->rcensor_day, T, S, a->max_seroday_obsd, a->account_serorev, a->binomial_likelihood};
  h = fnv1a(h, scal, sizeof scal);
  h = fnv1a(h, a->obs_deaths, sizeof(double) * T);
  h = fnv1a(h, a->prop_strata_obs_deaths, sizeof(double) * A);
  h = fnv1a(h, a->sero_a, sizeof(double) * S * A);
  h = fnv1a(h, a->sero_b, sizeof(double) * S * A);
  h = fnv1a(h, a->demog, sizeof(double) * A);
  h = fnv1a(h, a->sero_survey_start, sizeof(int32_t) * S);
  h = fnv1a(h, a->sero_survey_end, sizeof(int32_t) * S);

  std::lock_guard<std::mutex> lock(g_call_mu);
  cc_model* m = nullptr;
  for (auto& e : g_call_cache)
    if (e.model && e.key == h && e.device == device) { m = e.model; e.stamp = ++g_call_stamp; }
  if (!m) {
    // role map by name (binomial.cpp:21-55; generated form Agecpp2strings.R:255-368)
    auto need = [&](const std::string& nm, int* row) {
      *row = find_name(a, nm);
      return *row >= 0;
    };
    cc_model_desc de;
    std::memset(&de, 0, sizeof de);
    de.abi_version = CC_ABI_VERSION;
    de.n_params = d; de.n_strata = A; de.n_knots = K; de.days_obsd = T; de.n_sero_obs = S;
    de.max_seroday_obsd = a->max_seroday_obsd; de.rcensor_day = a->rcensor_day; de.account_serorev = a->account_serorev ? 1 : 0;
    de.binomial_likelihood = a->binomial_likelihood ? 1 : 0;
    de.demog = a->demog; de.obs_deaths = a->obs_deaths; de.prop_strata_obs_deaths = a->prop_strata_obs_deaths;
    de.sero_a = a->sero_a; de.sero_b = a->sero_b; de.sero_survey_start = a->sero_survey_start; de.sero_survey_end = a->sero_survey_end;
    std::string missing;
    auto want = [&](const std::string& nm, int32_t* row) { int r; if (!need(nm, &r)) { if (missing.empty()) missing = nm; r = 0; } *row = r; };
    want("sens", &de.idx_sens); want("spec", &de.idx_spec); want("sero_con_rate", &de.idx_sero_con_rate);
    want("mod", &de.idx_mod); want("sod", &de.idx_sod);
    de.idx_sero_rev_shape = de.idx_sero_rev_scale = -1;
    if (a->account_serorev) { want("sero_rev_shape", &de.idx_sero_rev_shape); want("sero_rev_scale", &de.idx_sero_rev_scale); }
    std::vector<int32_t> knots(K - 1), infxn(K), ifr(A), noise(A);
    for (int i = 0; i < K - 1; i++) want("x" + std::to_string(i + 1), &knots[i]);
    for (int i = 0; i < K; i++) want("y" + std::to_string(i + 1), &infxn[i]);
    for (int i = 0; i < A; i++) want("ma" + std::to_string(i + 1), &ifr[i]);
    for (int i = 0; i < A; i++) want("ne" + std::to_string(i + 1), &noise[i]);
    if (!missing.empty()) return fail(CC_E_INVALID, "cc_loglike_call: params has no element named '" + missing + "'");
    de.idx_knots = knots.data(); de.idx_infxn = infxn.data(); de.idx_ifr = ifr.data(); de.idx_noise = noise.data();
    de.idx_rel_knot = de.idx_rel_infxn = de.idx_max_ma = -1;
    std::vector<double> lo(d, -INFINITY), hi(d, INFINITY), zero(d, 0.0);
    std::vector<int32_t> fam(d, CC_PRIOR_NONE);
    de.pmin = lo.data(); de.pmax = hi.data(); de.prior_family = fam.data(); de.prior_p1 = zero.data(); de.prior_p2 = zero.data();
    de.jac_rows[0] = de.jac_rows[1] = de.jac_rows[2] = -1;
    int rc = cc_model_create(&de, device, &m);
    if (rc) return rc;
    CallCacheEntry* victim = &g_call_cache[0];
    for (auto& e : g_call_cache)
      if (!e.model) { victim = &e; break; } else if (e.stamp < victim->stamp) victim = &e;
    if (victim->model) cc_model_destroy(victim->model);
    victim->key = h; victim->device = device; victim->model = m; victim->stamp = ++g_call_stamp;
  }
  return cc_loglike(m, params, param_i, out);
}

extern "C" int cc_last_kernel_ms(cc_model* m, double* ms, int32_t* launches) {
  if (!m) return fail(CC_E_INVALID, "cc_last_kernel_ms: null model");
  if (m->last_ms < 0) {
    CC_CUDA(cudaEventSynchronize(m->ev1));
    float f = 0.f;
    CC_CUDA(cudaEventElapsedTime(&f, m->ev0, m->ev1));
    m->last_ms = f;
  }
  if (ms) *ms = m->last_ms;
  if (launches) *launches = m->last_launches;
  return CC_OK;
}

extern "C" int cc_curves_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* infxn, double* ne, double* auc,
                               double* sero_full) {
  if (!m || !theta) return fail(CC_E_INVALID, "cc_curves_batch: null argument");
  if (n < 0 || ld < n) return fail(CC_E_INVALID, "cc_curves_batch: need 0 <= n <= ld");
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(m->device));
  const KModel& k = m->k;
  const int d = k.d, T = k.T, A = k.A;
  double *dth = nullptr, *di = nullptr, *dn = nullptr, *da = nullptr, *ds = nullptr;
  auto cleanup = [&]() {
    cudaFree(dth); cudaFree(di); cudaFree(dn); cudaFree(da); cudaFree(ds);
  };
  cudaError_t e = cudaMalloc((void**)&dth, (size_t)d * n * sizeof(double));
  if (e == cudaSuccess && infxn) e = cudaMalloc((void**)&di, (size_t)n * T * sizeof(double));
  if (e == cudaSuccess && ne) e = cudaMalloc((void**)&dn, (size_t)n * A * sizeof(double));
  if (e == cudaSuccess && auc) e = cudaMalloc((void**)&da, (size_t)n * T * sizeof(double));
  if (e == cudaSuccess && sero_full) e = cudaMalloc((void**)&ds, (size_t)n * T * sizeof(double));
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(dth, (size_t)n * sizeof(double), theta, (size_t)ld * sizeof(double), (size_t)n * sizeof(double), (size_t)d,
                          cudaMemcpyHostToDevice, m->stream);
  if (e == cudaSuccess) {
    const size_t smem = eval_smem_bytes(d, T, T);
    curves_kernel<<<(unsigned)std::min<long long>(n, 65535LL * 16), kEvalThreads, smem, m->stream>>>(k, dth, n, n, di, dn, da, ds, nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && infxn) e = cudaMemcpyAsync(infxn, di, (size_t)n * T * sizeof(double), cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess && ne) e = cudaMemcpyAsync(ne, dn, (size_t)n * A * sizeof(double), cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess && auc) e = cudaMemcpyAsync(auc, da, (size_t)n * T * sizeof(double), cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess && sero_full) e = cudaMemcpyAsync(sero_full, ds, (size_t)n * T * sizeof(double), cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
  cleanup();
  if (e != cudaSuccess) return fail(CC_E_CUDA, std::string("cc_curves_batch: ") + cudaGetErrorString(e));
  return CC_OK;
}

extern "C" int cc_post_deaths_batch(cc_model* m, const double* theta, int64_t n, int64_t ld, double* deaths) {
  if (!m || !theta || !deaths) return fail(CC_E_INVALID, "cc_post_deaths_batch: null argument");
  if (n < 0 || ld < n) return fail(CC_E_INVALID, "cc_post_deaths_batch: need 0 <= n <= ld");
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(m->device));
  const KModel& k = m->k;
  const int d = k.d, T = k.T, A = k.A;
  double *dth = nullptr, *dd = nullptr;
  const size_t nout = (size_t)n * T * A * sizeof(double);
  cudaError_t e = cudaMalloc((void**)&dth, (size_t)d * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&dd, nout);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(dth, (size_t)n * sizeof(double), theta, (size_t)ld * sizeof(double), (size_t)n * sizeof(double), (size_t)d,
                          cudaMemcpyHostToDevice, m->stream);
  if (e == cudaSuccess) {
    const size_t smem = eval_smem_bytes(d, T, T);
    curves_kernel<<<(unsigned)std::min<long long>(n, 65535LL * 16), kEvalThreads, smem, m->stream>>>(k, dth, n, n, nullptr, nullptr, nullptr,
                                                                                                   nullptr, dd);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(deaths, dd, nout, cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
  else cudaStreamSynchronize(m->stream);
  cudaFree(dth);
  cudaFree(dd);
  if (e != cudaSuccess) return fail(CC_E_CUDA, std::string("cc_post_deaths_batch: ") + cudaGetErrorString(e));
  return CC_OK;
}

// ------------------------------------------------------------------------------------------------ MCMC host side
struct cc_mcmc {
  cc_model* m = nullptr;
  McmcState s{};
  std::vector<void*> allocs;
  std::vector<double> beta;
  std::vector<int32_t> rec_iter;  // iteration number of every recorded row
  long long iteration = 0;        // iterations done (1 after create)
  long long rows = 0;
  long long evals = 0;
  size_t smem = 0;
  int le = 0;          // log2 of the particles per warp tile (initial-state kernel)
  int sweep_wpc = 4;   // warps per CTA, grid and shared memory of the cached sweep kernel
  unsigned sweep_grid = 1;
  size_t sweep_smem = 0;
  unsigned grid = 1;
  double* scratch = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // bracket the launches of the last cc_mcmc_advance on the library's stream
  double last_ms = 0.0;
};

template <typename T>
static int dalloc(cc_mcmc* s, T** p, size_t n, bool zero = true) {
  void* q = nullptr;
  CC_CUDA(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
  s->allocs.push_back(q);
  if (zero) CC_CUDA(cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T)));
  *p = static_cast<T*>(q);
  return CC_OK;
}

extern "C" void cc_mcmc_destroy(cc_mcmc* s) {
  if (!s) return;
  cudaSetDevice(s->m->device);
  for (void* p : s->allocs) cudaFree(p);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
}

static int record_and_swap(cc_mcmc* r, int iteration, int do_swap) {
  McmcState& s = r->s;
  const int thin = s.thin <= 1 ? 1 : s.thin;
  long long row = -1;
  if (iteration % thin == 0 && r->rows < s.rows_cap) row = r->rows;
  if (!(do_swap && s.coupling_on && s.rungs > 1) && row < 0) return CC_OK;
  const int bs = std::min(128, std::max(32, ((s.d + 31) / 32) * 32));
  mcmc_swap_record_kernel<<<s.n_chains, bs, 0, r->m->stream>>>(r->m->k, s, iteration, row, do_swap);
  CC_CUDA(cudaGetLastError());
  if (row >= 0) {
    r->rec_iter.push_back(iteration);
    r->rows++;
  }
  return CC_OK;
}

extern "C" int cc_mcmc_create(cc_model* m, const cc_mcmc_opts* o, cc_mcmc** out) {
  if (!m || !o || !out) return fail(CC_E_INVALID, "cc_mcmc_create: null argument");
  *out = nullptr;
  if (o->burnin < 1 || o->samples < 0 || o->rungs < 1 || o->n_chains < 1 || o->chain_offset < 0)
    return fail(CC_E_INVALID, "cc_mcmc_create: need burnin >= 1, samples >= 0, rungs >= 1, n_chains >= 1");
  if (!o->init && !o->init_default) return fail(CC_E_INVALID, "cc_mcmc_create: init or init_default is required");
  CC_CUDA(cudaSetDevice(m->device));
  const int d = m->k.d, R = o->rungs, C = o->n_chains;
  const size_t P = (size_t)R * C;

  // thermodynamic powers: equally spaced on [0, 1] raised to GTI_pow; a single rung is the cold chain (power 1)
  std::vector<double> beta(R);
  if (o->beta_manual) {
    for (int r = 0; r < R; r++) beta[r] = o->beta_manual[r];
  } else if (R == 1) {
    beta[0] = 1.0;
  } else {
    for (int r = 0; r < R; r++) beta[r] = std::pow((double)r / (double)(R - 1), o->GTI_pow);
  }
  const bool cold_last = o->beta_manual ? true : (o->cold_rung_last != 0);
  if (!o->beta_manual && !cold_last) std::reverse(beta.begin(), beta.end());
  int cold_pos = 0;
  for (int r = 0; r < R; r++)
    if (beta[r] >= beta[cold_pos]) cold_pos = r;
  if (!cold_last)
    for (int r = R - 1; r >= 0; r--)
      if (beta[r] >= beta[cold_pos]) cold_pos = r;

  cc_mcmc* r = new (std::nothrow) cc_mcmc();
  if (!r) return fail(CC_E_NOMEM, "out of host memory");
  r->m = m;
  r->beta = beta;
  if (cudaEventCreate(&r->ev0) != cudaSuccess || cudaEventCreate(&r->ev1) != cudaSuccess) {
    cc_mcmc_destroy(r);
    return fail(CC_E_CUDA, "cc_mcmc_create: cudaEventCreate failed");
  }
  McmcState& s = r->s;
  s.d = d; s.rungs = R; s.n_chains = C; s.chain_offset = o->chain_offset;
  s.burnin = o->burnin; s.samples = o->samples; s.coupling_on = o->coupling_on ? 1 : 0;
  s.record_all = o->record_rungs ? 1 : 0;
  s.thin = o->thin <= 1 ? 1 : o->thin;
  s.bw_per_particle = o->bw_per_particle ? 1 : 0;
  s.adapt_in_sampling = o->no_adapt_in_sampling ? 0 : 1;
  s.seed = o->seed;
  s.n_rec_rungs = s.record_all ? R : 1;
  s.cold_pos = cold_pos;
  s.rows_cap = ((long long)o->burnin + o->samples) / s.thin;
  if (s.rows_cap < 1) s.rows_cap = 1;
  int rc = CC_OK;
  const size_t rec = (size_t)C * s.n_rec_rungs * (size_t)s.rows_cap;
  if ((rc = dalloc(r, &s.theta, P * d)) || (rc = dalloc(r, &s.phi, P * d)) || (rc = dalloc(r, &s.ll, P)) ||
      (rc = dalloc(r, &s.lp, P)) || (rc = dalloc(r, &s.bw, P * d)) || (rc = dalloc(r, &s.bw_index, P * d)) ||
      (rc = dalloc(r, &s.move_acc, P * d)) || (rc = dalloc(r, &s.at_rung, P)) || (rc = dalloc(r, &s.beta, (size_t)R)) ||
      (rc = dalloc(r, &s.swap_acc, (size_t)C * std::max(R - 1, 1))) || (rc = dalloc(r, &s.swap_try, (size_t)C * std::max(R - 1, 1))) ||
      (rc = dalloc(r, &s.rec_theta, rec * d, false)) || (rc = dalloc(r, &s.rec_ll, rec, false)) ||
      (rc = dalloc(r, &s.rec_lp, rec, false)) ||
      (rc = dalloc(r, &s.cache, P * (size_t)sweep_cache_layout(m->k.K, m->k.A, m->k.S, m->nt).size)) || (rc = dalloc(r, &s.ticket, 1))) {
    cc_mcmc_destroy(r);
    return rc;
  }
  {
    std::vector<double> th(P * d), bw(P * d, 1.0);
    std::vector<int> bwi(P * d, 1), at(P);
    for (int c = 0; c < C; c++)
      for (int q = 0; q < R; q++) {
        const double* src = o->init ? o->init + (size_t)c * d : o->init_default;
        std::copy(src, src + d, th.begin() + ((size_t)c * R + q) * d);
        at[(size_t)c * R + q] = q;
      }
    cudaError_t e = cudaMemcpy(s.theta, th.data(), th.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s.bw, bw.data(), bw.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s.bw_index, bwi.data(), bwi.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s.at_rung, at.data(), at.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s.beta, beta.data(), beta.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cc_mcmc_destroy(r);
      return fail(CC_E_CUDA, std::string("cc_mcmc_create: ") + cudaGetErrorString(e));
    }
  }
  r->smem = warp_cta_smem(m->k, m->nt);
  {
    // particles per warp tile: as many as keep every warp slot of the GPU busy (E = 1 for small runs)
    const long long slots = (long long)m->grid_max * kWarpsPerCta;
    int le = 5;
    while (le > 0 && (long long)((P + (1u << le) - 1) >> le) < slots) le--;
    r->le = le;
    const long long tiles = ((long long)P + (1 << le) - 1) >> le;
    r->grid = (unsigned)std::min<long long>((tiles + kWarpsPerCta - 1) / kWarpsPerCta, (long long)m->grid_max);
    {
      // warps per CTA and CTAs per SM of the sweep: whatever keeps the most warps resident (shared memory and registers)
      const size_t per_warp = sweep_smem_doubles(d, m->k.K, m->k.A, m->k.S, m->nt) * sizeof(double);
      int best_wpc = 0, best_warps = 0, best_ctas = 0;
      for (int wpc = kSweepWarps; wpc >= 1; wpc = (wpc > 4) ? wpc - 2 : wpc >> 1) {
        int ctas = 0;
        cudaError_t oe = cudaSuccess;
        switch (m->nt) {
#define CC_OCC(NT_) case NT_: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, mcmc_sweep_cached_kernel<NT_>, 32 * wpc, per_warp * wpc); break;
          CC_FOR_EACH_NT(CC_OCC)
#undef CC_OCC
          default: break;
        }
        if (oe != cudaSuccess) ctas = 0;
        if (ctas * wpc > best_warps) { best_warps = ctas * wpc; best_wpc = wpc; best_ctas = ctas; }
      }
      if (best_wpc == 0) {
        cudaGetLastError();
        cc_mcmc_destroy(r);
        return fail(CC_E_INVALID, "model too large for the shared memory of the Metropolis sweep kernel");
      }
      // fewer particles than resident warp slots: thinner CTAs, so that every SM gets its share
      const long long per_sm = ((long long)P + (long long)best_ctas * m->sm_count - 1) / ((long long)best_ctas * m->sm_count);
      if (per_sm < best_wpc) best_wpc = (int)std::max<long long>(1, per_sm);
      r->sweep_wpc = best_wpc;
      r->sweep_smem = per_warp * best_wpc;
      r->sweep_grid = (unsigned)std::min<long long>(((long long)P + best_wpc - 1) / best_wpc, (long long)best_ctas * m->sm_count);
      if (std::getenv("CC_VERBOSE"))
        std::fprintf(stderr, "covidcurve_b200: sweep kernel %d warps/CTA x %d CTAs/SM, %zu B smem/CTA, grid %u, cache %.1f MB\n", best_wpc,
                     best_ctas, r->sweep_smem, r->sweep_grid, (double)P * sweep_cache_layout(m->k.K, m->k.A, m->k.S, m->nt).size * 8e-6);
    }
    const RecLayout L = rec_layout(m->k.K, m->k.A, m->k.S);
    if ((rc = dalloc(r, &r->scratch, (size_t)r->grid * kWarpsPerCta * L.F * 32, false))) {
      cc_mcmc_destroy(r);
      return rc;
    }
    // the memsets / copies above ran on the legacy default stream; m->stream is non-blocking and does not order behind it
    if (cudaError_t es = cudaDeviceSynchronize(); es != cudaSuccess) {
      cc_mcmc_destroy(r);
      return fail(CC_E_CUDA, std::string("cc_mcmc_create: ") + cudaGetErrorString(es));
    }
    switch (m->nt) {
#define CC_LAUNCH(NT_) case NT_: mcmc_init_tile_kernel<NT_><<<r->grid, 32 * kWarpsPerCta, r->smem, m->stream>>>(m->k, s, r->le, r->scratch); break;
      CC_FOR_EACH_NT(CC_LAUNCH)
#undef CC_LAUNCH
      default: break;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cc_mcmc_destroy(r);
    return fail(CC_E_CUDA, std::string("mcmc_init_kernel: ") + cudaGetErrorString(e));
  }
  r->evals += (long long)P;
  r->iteration = 1;
  rc = record_and_swap(r, 1, 0);
  if (rc == CC_OK) {
    e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) rc = fail(CC_E_CUDA, std::string("cc_mcmc_create: ") + cudaGetErrorString(e));
  }
  if (rc) {
    cc_mcmc_destroy(r);
    return rc;
  }
  *out = r;
  return CC_OK;
}

extern "C" int cc_mcmc_advance(cc_mcmc* r, int32_t n_iter) {
  if (!r) return fail(CC_E_INVALID, "cc_mcmc_advance: null state");
  if (n_iter < 0) return fail(CC_E_INVALID, "cc_mcmc_advance: n_iter < 0");
  cc_model* m = r->m;
  CC_CUDA(cudaSetDevice(m->device));
  McmcState& s = r->s;
  const long long total = (long long)s.burnin + s.samples;
  const unsigned P = (unsigned)(s.n_chains * s.rungs);
  CC_CUDA(cudaEventRecord(r->ev0, m->stream));
  for (int k = 0; k < n_iter; k++) {
    if (r->iteration >= total) return fail(CC_E_STATE, "cc_mcmc_advance: burnin + samples iterations already done");
    const int it = (int)(r->iteration + 1);
    CC_CUDA(cudaMemsetAsync(s.ticket, 0, sizeof(int), m->stream));
    switch (m->nt) {
#define CC_LAUNCH(NT_) case NT_: mcmc_sweep_cached_kernel<NT_><<<r->sweep_grid, 32 * r->sweep_wpc, r->sweep_smem, m->stream>>>(m->k, s, it); break;
      CC_FOR_EACH_NT(CC_LAUNCH)
#undef CC_LAUNCH
      default: break;
    }
    CC_CUDA(cudaGetLastError());
    r->evals += (long long)P * s.d;
    int rc = record_and_swap(r, it, 1);
    if (rc) return rc;
    r->iteration = it;
  }
  CC_CUDA(cudaEventRecord(r->ev1, m->stream));
  CC_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0.f;
  CC_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  r->last_ms = ms;
  return CC_OK;
}

extern "C" int cc_mcmc_last_advance_ms(cc_mcmc* r, double* ms) {
  if (!r || !ms) return fail(CC_E_INVALID, "cc_mcmc_last_advance_ms: null argument");
  *ms = r->last_ms;
  return CC_OK;
}

extern "C" int64_t cc_mcmc_iterations(const cc_mcmc* s) { return s ? s->iteration : 0; }
extern "C" int64_t cc_mcmc_rows(const cc_mcmc* s) { return s ? s->rows : 0; }
extern "C" int64_t cc_mcmc_loglike_evals(const cc_mcmc* s) { return s ? s->evals : 0; }

// recorded draws are stored with a row capacity; fetch compacts to [chains][rec rungs][rows][..]
extern "C" int cc_mcmc_fetch(cc_mcmc* r, double* theta, double* loglike, double* logprior, int32_t* iteration) {
  if (!r) return fail(CC_E_INVALID, "cc_mcmc_fetch: null state");
  CC_CUDA(cudaSetDevice(r->m->device));
  const McmcState& s = r->s;
  const size_t groups = (size_t)s.n_chains * s.n_rec_rungs, rows = (size_t)r->rows, cap = (size_t)s.rows_cap, d = (size_t)s.d;
  CC_CUDA(cudaStreamSynchronize(r->m->stream));
  if (theta)
    CC_CUDA(cudaMemcpy2D(theta, rows * d * sizeof(double), s.rec_theta, cap * d * sizeof(double), rows * d * sizeof(double), groups,
                         cudaMemcpyDeviceToHost));
  if (loglike)
    CC_CUDA(cudaMemcpy2D(loglike, rows * sizeof(double), s.rec_ll, cap * sizeof(double), rows * sizeof(double), groups,
                         cudaMemcpyDeviceToHost));
  if (logprior)
    CC_CUDA(cudaMemcpy2D(logprior, rows * sizeof(double), s.rec_lp, cap * sizeof(double), rows * sizeof(double), groups,
                         cudaMemcpyDeviceToHost));
  if (iteration) std::copy(r->rec_iter.begin(), r->rec_iter.end(), iteration);
  return CC_OK;
}

extern "C" int cc_mcmc_device_draws(cc_mcmc* r, double** theta, double** loglike, double** logprior, int64_t* rows_cap) {
  if (!r) return fail(CC_E_INVALID, "cc_mcmc_device_draws: null state");
  if (theta) *theta = r->s.rec_theta;
  if (loglike) *loglike = r->s.rec_ll;
  if (logprior) *logprior = r->s.rec_lp;
  if (rows_cap) *rows_cap = r->s.rows_cap;
  return CC_OK;
}

extern "C" int cc_mcmc_diagnostics(cc_mcmc* r, double* beta, double* mc_accept, double* move_accept, double* bw) {
  if (!r) return fail(CC_E_INVALID, "cc_mcmc_diagnostics: null state");
  CC_CUDA(cudaSetDevice(r->m->device));
  const McmcState& s = r->s;
  CC_CUDA(cudaStreamSynchronize(r->m->stream));
  const size_t P = (size_t)s.n_chains * s.rungs, d = (size_t)s.d;
  if (beta) std::copy(r->beta.begin(), r->beta.end(), beta);
  if (mc_accept && s.rungs > 1) {
    const size_t n = (size_t)s.n_chains * (s.rungs - 1);
    std::vector<unsigned> a(n), t(n);
    CC_CUDA(cudaMemcpy(a.data(), s.swap_acc, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CC_CUDA(cudaMemcpy(t.data(), s.swap_try, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) mc_accept[i] = t[i] ? (double)a[i] / (double)t[i] : NAN;
  }
  if (move_accept) {
    std::vector<unsigned> a(P * d);
    CC_CUDA(cudaMemcpy(a.data(), s.move_acc, P * d * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < P * d; i++) move_accept[i] = (double)a[i];
  }
  if (bw) CC_CUDA(cudaMemcpy(bw, s.bw, P * d * sizeof(double), cudaMemcpyDeviceToHost));
  return CC_OK;
}

// ------------------------------------------------------------------------------------------------ nmath test hook
__global__ void nmath_kernel(int fn, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                             long long n, double* __restrict__ out, const LogTab* __restrict__ tab) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b[i], z = c[i];
  double r;
  switch (fn) {
    case CC_FN_PGAMMA: {  // pgamma(x, shape = y, scale = z), lower tail
      if (isnan(x) || isnan(y) || isnan(z)) r = x + y + z;
      else if (y < 0 || z <= 0) r = d_nan();
      else r = pgamma_lower(gamma_shape(y), x / z);
      break;
    }
    case CC_FN_DPOIS: r = dpois_log_int(x, y); break;
    case CC_FN_DBINOM: r = dbinom_log(x, y, z); break;
    case CC_FN_DNORM: r = dnorm_log(x, y, z); break;
    case CC_FN_DBETA: r = dbeta_log(x, y, z); break;
    case CC_FN_DUNIF: r = dunif_log(x, y, z); break;
    case CC_FN_PWEIBULL: r = pweibull_lower(x, y, z); break;
    case CC_FN_STIRLERR: r = stirlerr(x); break;
    case CC_FN_BD0: r = bd0(x, y); break;
    case CC_FN_FASTLOG: r = fast_log(x, tab); break;
    case CC_FN_RPOIS: {
      SimRng g{(unsigned long long)z, (uint32_t)y, (uint32_t)(i & 0xffffffffLL), 7u, 0u, {0, 0, 0, 0}, 0};
      r = rpois_dev(x, g);
      break;
    }
    default: r = d_nan();
  }
  out[i] = r;
}

extern "C" int cc_nmath_batch(int device, int fn, const double* a, const double* b, const double* c, int64_t n, double* out) {
  if (!a || !b || !c || !out || n < 0) return fail(CC_E_INVALID, "cc_nmath_batch: bad argument");
  if (fn < CC_FN_PGAMMA || fn > CC_FN_RPOIS) return fail(CC_E_INVALID, "cc_nmath_batch: unknown function id");
  if (cc_device_count() <= 0) return fail(CC_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (n == 0) return CC_OK;
  CC_CUDA(cudaSetDevice(device));
  double* d = nullptr;
  CC_CUDA(cudaMalloc((void**)&d, (size_t)4 * n * sizeof(double) + 128 * sizeof(LogTab)));
  LogTab* dtab = reinterpret_cast<LogTab*>(d + 4 * n);
  {
    LogTab ht[128];
    for (int i = 0; i < 128; i++) {
      const float cf = (float)(1.0 / (1.0 + ((double)i + 0.5) / 128.0));
      ht[i].c = (double)cf;
      ht[i].mlogc = (double)(-logl((long double)cf));
    }
    CC_CUDA(cudaMemcpy(dtab, ht, sizeof ht, cudaMemcpyHostToDevice));
  }
  cudaError_t e = cudaMemcpy(d, a, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + n, b, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + 2 * n, c, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    nmath_kernel<<<(unsigned)((n + 127) / 128), 128>>>(fn, d, d + n, d + 2 * n, n, d + 3 * n, dtab);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(CC_E_CUDA, std::string("cc_nmath_batch: ") + cudaGetErrorString(e));
  return CC_OK;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int cc_fp64_peak_probe(int device, int iters, double* tflops, double* ms_out) {
  int ndev = cc_device_count();
  if (ndev <= 0) return fail(CC_E_CUDA, "no CUDA device available");
  if (device < 0 || device >= ndev) return fail(CC_E_INVALID, "device index out of range");
  if (iters < 1) iters = 4096;
  CC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CC_CUDA(cudaGetDeviceProperties(&prop, device));
  double* dummy = nullptr;
  CC_CUDA(cudaMalloc((void**)&dummy, sizeof(double)));
  cudaEvent_t e0, e1;
  CC_CUDA(cudaEventCreate(&e0));
  CC_CUDA(cudaEventCreate(&e1));
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 8; rep++) {
    CC_CUDA(cudaEventRecord(e0, 0));
    fp64_probe_kernel<<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7);
    CC_CUDA(cudaEventRecord(e1, 0));
    CC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dummy);
  const double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
  if (tflops) *tflops = flops / ((double)best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return CC_OK;
}

extern "C" int cc_dmma_probe(int device, int iters, int fma_per_mma, double* mma_tflops, double* fma_tflops, double* ms_out) {
  int ndev = cc_device_count();
  if (ndev <= 0) return fail(CC_E_CUDA, "no CUDA device available");
  if (device < 0 || device >= ndev) return fail(CC_E_INVALID, "device index out of range");
  if (iters < 1) iters = 4096;
  CC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CC_CUDA(cudaGetDeviceProperties(&prop, device));
  double* dummy = nullptr;
  CC_CUDA(cudaMalloc((void**)&dummy, sizeof(double)));
  cudaEvent_t e0, e1;
  CC_CUDA(cudaEventCreate(&e0));
  CC_CUDA(cudaEventCreate(&e1));
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CC_CUDA(cudaEventRecord(e0, 0));
    switch (fma_per_mma) {
      case 0: dmma_probe_kernel<0><<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7); break;
      case 1: dmma_probe_kernel<1><<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7); break;
      case 2: dmma_probe_kernel<2><<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7); break;
      case 4: dmma_probe_kernel<4><<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7); break;
      case 8: dmma_probe_kernel<8><<<blocks, threads>>>(dummy, iters, 0.999999, 1e-7); break;
      default: cudaFree(dummy); return fail(CC_E_INVALID, "fma_per_mma must be 0, 1, 2, 4 or 8");
    }
    CC_CUDA(cudaEventRecord(e1, 0));
    CC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dummy);
  const double warps = (double)threads / 32 * blocks;
  const double mma_flops = 2.0 * 256.0 * 8.0 * (double)iters * warps;                       // 8x8x4 FMAs per warp-level mma
  const double fma_flops = 2.0 * 8.0 * fma_per_mma * (double)iters * (double)threads * blocks;
  if (mma_tflops) *mma_tflops = mma_flops / ((double)best * 1e-3) / 1e12;
  if (fma_tflops) *fma_tflops = fma_flops / ((double)best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return CC_OK;
}

// ------------------------------------------------------------------------------------------------ forward simulator
extern "C" int cc_simulate_infxn_2_death(int device, const double* infections, int32_t T, const double* ifr, const double* rho,
                                         const double* popN, int32_t A, double m_od, double s_od, double sero_delay_rate,
                                         double smplfrac, double sero_rev_shape, double sero_rev_scale, uint64_t seed, int64_t n_rep,
                                         double* deaths, double* serocon, double* serorev, double* mean_deaths, double* mean_serocon) {
  if (!infections || !ifr || !rho || !popN) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: null argument");
  if (T < 1 || A < 1 || n_rep < 0) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: need T >= 1, A >= 1, n_rep >= 0");
  if (n_rep > 0 && (!deaths || !serocon)) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: null output");
  if (!(m_od > 0.0) || !(s_od > 0.0) || !(sero_delay_rate > 0.0) || !(smplfrac > 0.0) || !(smplfrac <= 1.0))
    return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: m_od, s_od, sero_delay_rate must be positive and smplfrac in (0, 1]");
  const bool rev = (sero_rev_shape != 0.0) || (sero_rev_scale != 0.0);
  if (rev && (!(sero_rev_shape > 0.0) || !(sero_rev_scale > 0.0) || (n_rep > 0 && !serorev)))
    return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: seroreversion needs shape > 0, scale > 0 and the serorev output");
  if (rev && T > 2048) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: seroreversion supports up to 2048 days");
  if ((long long)n_rep * T * A > (1LL << 40)) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: too many cells");
  if (cc_device_count() <= 0) return fail(CC_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
  double wsum = 0.0;
  for (int a = 0; a < A; a++) {
    if (!(popN[a] >= 0.0) || !(rho[a] >= 0.0) || !(ifr[a] >= 0.0) || !(ifr[a] <= 1.0))
      return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: popN, rho must be >= 0 and ifr in [0, 1]");
    wsum += popN[a] * rho[a];
  }
  if (!(wsum > 0.0)) return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: popN * rho sums to zero");
  for (int t = 0; t < T; t++)
    if (!(infections[t] >= 0.0) || !std::isfinite(infections[t]))
      return fail(CC_E_INVALID, "cc_simulate_infxn_2_death: infections must be finite and >= 0");
  std::vector<double> wd(A), ws(A);
  for (int a = 0; a < A; a++) {
    const double pa = popN[a] * rho[a] / wsum;  // simulators.R:172
    wd[a] = pa * ifr[a];
    ws[a] = pa * smplfrac;
  }
  CC_CUDA(cudaSetDevice(device));
  const size_t cells = (size_t)T * A, tot = (size_t)n_rep * cells, tot1 = std::max<size_t>(tot, 1);
  const size_t tt = rev ? (size_t)T * T : 0;
  double* d = nullptr;
  CC_CUDA(cudaMalloc((void**)&d, ((size_t)5 * T + 2 * A + (rev ? 3 : 2) * tot1 + 2 * tt) * sizeof(double)));
  double *d_inf = d, *d_dG = d + T, *d_dH = d + 2 * T, *d_mud = d + 3 * T, *d_mus = d + 4 * T, *d_wd = d + 5 * T, *d_ws = d_wd + A,
         *d_deaths = d_ws + A, *d_sero = d_deaths + tot1, *d_rev = d_sero + tot1, *d_pi = d_rev + (rev ? tot1 : 0), *d_lam = d_pi + tt;
  cudaError_t e = cudaMemcpy(d_inf, infections, (size_t)T * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_wd, wd.data(), (size_t)A * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_ws, ws.data(), (size_t)A * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    sim_means_kernel<<<1, 256>>>(d_inf, T, m_od, s_od, sero_delay_rate, d_dG, d_dH, d_mud, d_mus);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && tot > 0) {
    sim_draw_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(d_mud, d_mus, d_wd, d_ws, T, A, (long long)n_rep, seed, d_deaths, d_sero);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && rev) {
    sim_serorev_pi_kernel<<<(unsigned)((tt + 255) / 256), 256>>>(T, sero_delay_rate, sero_rev_shape, sero_rev_scale, d_pi);
    sim_serorev_means_kernel<<<(unsigned)((tt + 255) / 256), 256>>>(d_inf, d_pi, T, d_lam);
    e = cudaGetLastError();
    if (e == cudaSuccess && tot > 0) {
      e = cudaMemset(d_rev, 0, tot * sizeof(double));
      if (e == cudaSuccess) {
        sim_serorev_draw_kernel<<<(unsigned)((tot + 127) / 128), 128>>>(d_lam, d_mus, d_ws, T, A, (long long)n_rep, seed, d_sero, d_rev);
        e = cudaGetLastError();
      }
    }
  }
  if (e == cudaSuccess && tot > 0) e = cudaMemcpy(deaths, d_deaths, tot * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && tot > 0) e = cudaMemcpy(serocon, d_sero, tot * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && tot > 0 && rev) e = cudaMemcpy(serorev, d_rev, tot * sizeof(double), cudaMemcpyDeviceToHost);
  std::vector<double> mud, mus;
  if (e == cudaSuccess && (mean_deaths || mean_serocon)) {
    mud.resize(T); mus.resize(T);
    e = cudaMemcpy(mud.data(), d_mud, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(mus.data(), d_mus, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  if (e != cudaSuccess) return fail(CC_E_CUDA, std::string("cc_simulate_infxn_2_death: ") + cudaGetErrorString(e));
  for (int j = 0; j < T && (mean_deaths || mean_serocon); j++)
    for (int a = 0; a < A; a++) {
      if (mean_deaths) mean_deaths[(size_t)j * A + a] = mud[j] * wd[a];
      if (mean_serocon) mean_serocon[(size_t)j * A + a] = mus[j] * ws[a];
    }
  return CC_OK;
}
