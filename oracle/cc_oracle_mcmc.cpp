/*
 * oracle/cc_oracle_mcmc.cpp - TEST INFRASTRUCTURE ONLY (see cc_oracle.h).
 *
 * CPU restatement of the MCMC engine the reference delegates to: drjacoby::run_mcmc (third-party, NOT under
 * /root/reference; pinned mrc-ide/drjacoby@e83884e0b2c48309ebd7eccdcbfc4495f0f80cac in DESCRIPTION:33-34; call site
 * R/main.R:179-193).  PARITY UNPINNED against drjacoby itself (its source and R's Mersenne-Twister stream are not
 * available here); what is restated is its published algorithm:
 *   - per-parameter transform to an unbounded phi from the finiteness of (min, max),
 *   - single-site Metropolis-Hastings in phi with proposal N(phi_i, bw_i), acceptance
 *         log U < beta * (ll' - ll) + (lp' - lp) + adj_i,
 *   - Robbins-Monro: log bw_i += (1 - 0.234)/sqrt(n_i) on accept, -= 0.234/sqrt(n_i) on reject, n_i++,
 *   - Metropolis coupling: sequentially for adjacent rungs k, k+1, swap w.p. exp((ll_{k+1} - ll_k)(beta_k - beta_{k+1})),
 *   - iteration 1 is the initial state; every rung is recorded every iteration.
 * The random numbers are those of the product (Philox4x32-10 keyed by seed / chain / particle / iteration / parameter,
 * restated here from the published algorithm, Salmon et al. SC'11) so that the GPU run can be checked draw by draw.
 * The loop is sequential per chain, chains farmed over threads like drjacoby's `cluster=`.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "cc_oracle.h"

namespace {

struct U4 { uint32_t v[4]; };

U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.v[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.v[2];
    U4 n;
    n.v[0] = (uint32_t)(p1 >> 32) ^ c.v[1] ^ k0;
    n.v[1] = (uint32_t)p1;
    n.v[2] = (uint32_t)(p0 >> 32) ^ c.v[3] ^ k1;
    n.v[3] = (uint32_t)p0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

double u01(uint32_t a, uint32_t b) {
  const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
  return ((double)v + 0.5) / 9007199254740992.0;
}

const uint32_t STREAM_MOVE = 0x4d4f5645u, STREAM_SWAP = 0x53574150u;

void draw(uint64_t seed, uint32_t stream, uint32_t chain, uint32_t id, uint32_t iteration, uint32_t param, double* z, double* u) {
  U4 c = {{iteration, param, id, chain}};
  const uint32_t k0 = (uint32_t)seed ^ stream, k1 = (uint32_t)(seed >> 32);
  const U4 r = philox4x32_10(c, k0, k1);
  c.v[1] = param | 0x80000000u;
  const U4 q = philox4x32_10(c, k0, k1);
  const double u1 = u01(r.v[0], r.v[1]), u2 = u01(r.v[2], r.v[3]);
  *z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
  *u = u01(q.v[0], q.v[1]);
}

int trans_type(double lo, double hi) {
  const bool flo = std::isfinite(lo), fhi = std::isfinite(hi);
  return flo ? (fhi ? 3 : 2) : (fhi ? 1 : 0);
}
double to_phi(int tt, double th, double lo, double hi) {
  switch (tt) {
    case 1: return std::log(hi - th);
    case 2: return std::log(th - lo);
    case 3: return std::log(th - lo) - std::log(hi - th);
    default: return th;
  }
}
double to_theta(int tt, double ph, double lo, double hi) {
  switch (tt) {
    case 1: return hi - std::exp(ph);
    case 2: return lo + std::exp(ph);
    case 3: { const double e = std::exp(ph); return (hi * e + lo) / (1 + e); }
    default: return ph;
  }
}
double adjust(int tt, double th, double thp, double lo, double hi) {
  switch (tt) {
    case 1: return std::log(hi - thp) - std::log(hi - th);
    case 2: return std::log(thp - lo) - std::log(th - lo);
    case 3: return std::log(hi - thp) + std::log(thp - lo) - std::log(hi - th) - std::log(th - lo);
    default: return 0.0;
  }
}

struct Particle {
  std::vector<double> theta, phi;
  double ll, lp;
};

void run_chain(const cco_desc* m, const double* init, const cco_mcmc_opts* o, const std::vector<double>& beta, int chain,
               double* out_theta, double* out_ll, double* out_lp, double* out_accept) {
  const int d = m->d, R = o->rungs, n_it = o->burnin + o->samples;
  std::vector<Particle> part(R);
  std::vector<int> at(R);
  // proposal bandwidths and Robbins-Monro counters belong to the rung position (temperature)
  std::vector<std::vector<double>> bw(R, std::vector<double>(d, 1.0));
  std::vector<std::vector<int>> bwi(R, std::vector<int>(d, 1));
  std::vector<double> acc(std::max(R - 1, 1), 0.0), tried(std::max(R - 1, 1), 0.0);
  for (int r = 0; r < R; r++) {
    at[r] = r;
    part[r].theta.assign(init, init + d);
    part[r].phi.resize(d);
    for (int i = 0; i < d; i++) part[r].phi[i] = to_phi(trans_type(m->pmin[i], m->pmax[i]), init[i], m->pmin[i], m->pmax[i]);
    part[r].ll = cco_loglike(m, init, 1, nullptr);
    part[r].lp = cco_logprior(m, init, 1);
  }
  auto record = [&](int it) {
    for (int r = 0; r < R; r++) {
      const Particle& p = part[at[r]];
      const size_t o_ = ((size_t)chain * R + r) * n_it + (size_t)(it - 1);
      std::copy(p.theta.begin(), p.theta.end(), out_theta + o_ * d);
      out_ll[o_] = p.ll;
      out_lp[o_] = p.lp;
    }
  };
  record(1);
  std::vector<double> prop(d);
  for (int it = 2; it <= n_it; it++) {
    for (int r = 0; r < R; r++) {
      Particle& p = part[at[r]];
      prop = p.theta;
      for (int i = 0; i < d; i++) {
        double z, u;
        draw(o->seed, STREAM_MOVE, (uint32_t)(chain + o->chain_offset), (uint32_t)at[r], (uint32_t)it, (uint32_t)i, &z, &u);
        const double lo = m->pmin[i], hi = m->pmax[i];
        const int tt = trans_type(lo, hi);
        const double php = p.phi[i] + bw[r][i] * z;
        const double thp = to_theta(tt, php, lo, hi);
        const double adj = adjust(tt, p.theta[i], thp, lo, hi);
        prop[i] = thp;
        const double llp = cco_loglike(m, prop.data(), i + 1, nullptr);
        const double lpp = cco_logprior(m, prop.data(), i + 1);
        const double mh = beta[r] * (llp - p.ll) + (lpp - p.lp) + adj;
        const bool accept = std::log(u) < mh;
        if (accept) {
          p.theta[i] = thp;
          p.phi[i] = php;
          p.ll = llp;
          p.lp = lpp;
        } else {
          prop[i] = p.theta[i];
        }
        const double step = accept ? (1.0 - 0.234) : -0.234;
        bw[r][i] = std::exp(std::log(bw[r][i]) + step / std::sqrt((double)bwi[r][i]));
        bwi[r][i] += 1;
      }
    }
    if (o->coupling_on && R > 1) {
      for (int k = 0; k < R - 1; k++) {
        const double l1 = part[at[k]].ll, l2 = part[at[k + 1]].ll;
        const double a = (l2 - l1) * (beta[k] - beta[k + 1]);
        double z, u;
        draw(o->seed, STREAM_SWAP, (uint32_t)(chain + o->chain_offset), (uint32_t)k, (uint32_t)it, 0u, &z, &u);
        const bool accept = std::log(u) < a;
        if (it > o->burnin) tried[k] += 1;
        if (accept) {
          std::swap(at[k], at[k + 1]);
          if (it > o->burnin) acc[k] += 1;
        }
      }
    }
    record(it);
  }
  if (out_accept)
    for (int k = 0; k < std::max(R - 1, 1); k++) out_accept[(size_t)chain * std::max(R - 1, 1) + k] = tried[k] > 0 ? acc[k] / tried[k] : NAN;
}

}  // namespace

extern "C" int cco_philox_kat(const unsigned ctr[4], const unsigned key[2], unsigned out[4]) {
  U4 c = {{ctr[0], ctr[1], ctr[2], ctr[3]}};
  const U4 r = philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; i++) out[i] = r.v[i];
  return 0;
}

extern "C" void cco_draw(unsigned long long seed, unsigned stream, unsigned chain, unsigned id, unsigned iteration, unsigned param,
                         double* z, double* u) {
  draw(seed, stream, chain, id, iteration, param, z, u);
}

extern "C" int cco_mcmc(const cco_desc* m, const double* init, const cco_mcmc_opts* o, double* out_theta, double* out_ll,
                        double* out_lp, double* out_accept, double* beta_out) {
  if (!m || !init || !o || !out_theta || !out_ll || !out_lp) return -1;
  if (o->burnin < 1 || o->samples < 0 || o->rungs < 1 || o->chains < 1) return -2;
  const int R = o->rungs;
  std::vector<double> beta(R);
  if (o->beta_manual) {
    for (int r = 0; r < R; r++) beta[r] = o->beta_manual[r];
  } else if (R == 1) {
    beta[0] = 1.0;
  } else {
    for (int r = 0; r < R; r++) beta[r] = std::pow((double)r / (double)(R - 1), o->GTI_pow);
    if (!o->cold_rung_last) std::reverse(beta.begin(), beta.end());
  }
  if (beta_out) std::copy(beta.begin(), beta.end(), beta_out);
  const int nth = std::max(1, std::min(o->nthreads, o->chains));
  std::vector<std::thread> th;
  for (int t = 0; t < nth; t++)
    th.emplace_back([&, t]() {
      for (int c = t; c < o->chains; c += nth)
        run_chain(m, init + (o->init_per_chain ? (size_t)c * m->d : 0), o, beta, c, out_theta, out_ll, out_lp, out_accept);
    });
  for (auto& t : th) t.join();
  return 0;
}
