// cc_rng.cuh - counter-based device RNG for the GPU-resident Metropolis step.
//
// drjacoby (the reference's MCMC engine, call site R/main.R:179-193) draws from R's Mersenne-Twister through
// Rcpp (rnorm1 / runif_0_1); that stream is inherently sequential per chain.  Here every draw is a pure
// function of (seed, stream tag, global chain id, particle-or-pair id, iteration, parameter) through
// Philox4x32-10 (Salmon et al., SC'11), so results do not depend on how chains are sharded over GPUs, and a
// CPU restatement with the same keys reproduces the trajectory.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define CC_HD __host__ __device__ __forceinline__
#else
#define CC_HD inline
#endif

namespace ccrng {

struct u4 { uint32_t x, y, z, w; };

CC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

// Philox4x32 with 10 rounds. ctr = 128-bit counter, (k0, k1) = 64-bit key.
CC_HD u4 philox4x32_10(u4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(M0, c.x, hi0, lo0);
    mulhilo(M1, c.z, hi1, lo1);
    u4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// 53-bit uniform in the open interval (0, 1)
CC_HD double u01(uint32_t a, uint32_t b) {
  uint64_t v = (((uint64_t)a << 32) | (uint64_t)b) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

enum Stream : uint32_t { kStreamMove = 0x4d4f5645u /* "MOVE" */, kStreamSwap = 0x53574150u /* "SWAP" */ };

struct Draw { double z; double u; };  // one standard normal + one uniform(0,1)

// The draws of one single-site proposal: z ~ N(0,1) for the proposal, u ~ U(0,1) for the accept test.
// id = particle id within the chain (move stream) or pair index (swap stream).
CC_HD Draw draw(uint64_t seed, uint32_t stream, uint32_t chain, uint32_t id, uint32_t iteration, uint32_t param) {
  u4 c;
  c.x = iteration;
  c.y = param;
  c.z = id;
  c.w = chain;
  u4 r = philox4x32_10(c, (uint32_t)seed ^ stream, (uint32_t)(seed >> 32));
  // second block for the uniform of the accept test (keeps the normal's two words independent of it)
  c.y = param | 0x80000000u;
  u4 q = philox4x32_10(c, (uint32_t)seed ^ stream, (uint32_t)(seed >> 32));
  const double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
  Draw d;
  d.z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);  // Box-Muller
  d.u = u01(q.x, q.y);
  return d;
}

}  // namespace ccrng
