// ESWOA fitness (src/baselines/WOA.py:87-105, `ESWOA.calc`) for a whole batch of whale positions in one launch.
//
// The reference evaluates one composition at a time in numpy float64 (popSize x MAX_Iter x instances calls):
//   conValues[i] = np.cumprod(q_{.,2+i})[-1]                  sequential float64 products, i in {0,1}
//   violate      = #{ i : conValues[i] < lo_i  or  conValues[i] > hi_i }
//   serviceNum   = #{ k : q_{k,0} > 0 }
//   objFunc      = (np.sum(q_{.,0}) / serviceNum + 1 - np.min(q_{.,1})) / 2
//   fitness      = violate + objFunc                          (WOA.py:59,78,119,151)
// One thread = one position; every float64 operation is issued in numpy's order (np.sum's 8-accumulator pairwise
// scheme included) so fitness values -- and therefore every comparison `bestFitness > fitness` of the search -- are
// bit-identical to the reference's.
#include "common.cuh"

namespace gnnpn {
namespace {

// numpy's pairwise summation for contiguous float64 (numpy/core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE):
// < 8 elements sequential, <= 128 elements eight running accumulators, above that a recursive split at
// n/2 rounded down to a multiple of 8.  The recursion is unrolled at compile time (depth 3 covers n <= 512: the
// pieces are <= 263, <= 135, <= 71 elements), so the kernel needs no device call stack.
__device__ __forceinline__ double numpy_block_sum_f64(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i;
  for (i = 8; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}
template <int DEPTH>
__device__ __forceinline__ double numpy_sum_f64(const double* a, int n) {
  if (n <= 128) return numpy_block_sum_f64(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(numpy_sum_f64<DEPTH - 1>(a, n2), numpy_sum_f64<DEPTH - 1>(a + n2, n - n2));
}
template <>
__device__ __forceinline__ double numpy_sum_f64<0>(const double* a, int n) { return numpy_block_sum_f64(a, n); }

constexpr int kMaxTasksWoa = 512;

// MEAN_ALL: `np.average(qos[0])` over all K picks (src/ML2PN.py:6-12, `ML2PN.calc`) instead of sum / #{q0 > 0}
template <bool MEAN_ALL>
__global__ void woa_fitness_kernel(const double* __restrict__ qos, const int32_t* __restrict__ idx, int64_t idx_ld,
                                   const int32_t* __restrict__ klen, const double* __restrict__ bounds, int64_t P,
                                   int Kmax, int32_t* __restrict__ viol_out, double* __restrict__ obj_out,
                                   double* __restrict__ fit_out) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int K = klen ? klen[p] : Kmax;
  double q0s[kMaxTasksWoa];
  double prod[2] = {1.0, 1.0};
  double min_q1 = INFINITY;
  int used = 0;
  for (int k = 0; k < K; ++k) {
    const double* row = qos + (int64_t)idx[p * idx_ld + k] * 4;
    const double q0 = row[0], q1 = row[1];
    q0s[k] = q0;
    used += q0 > 0.0;
    min_q1 = fmin(min_q1, q1);
    prod[0] = k == 0 ? row[2] : __dmul_rn(prod[0], row[2]);
    prod[1] = k == 0 ? row[3] : __dmul_rn(prod[1], row[3]);
  }
  const double* b = bounds + p * 4;
  int viol = 0;
  for (int i = 0; i < 2; ++i) viol += (prod[i] < b[2 * i] || prod[i] > b[2 * i + 1]) ? 1 : 0;
  double obj = __ddiv_rn(numpy_sum_f64<3>(q0s, K), (double)(MEAN_ALL ? K : used));
  obj = __dadd_rn(obj, 1.0);
  obj = __dsub_rn(obj, min_q1);
  obj = __ddiv_rn(obj, 2.0);
  if (viol_out) viol_out[p] = viol;
  if (obj_out) obj_out[p] = obj;
  if (fit_out) fit_out[p] = __dadd_rn((double)viol, obj);
}


// ------------------------------------------------------------------------------------------------------------------
// Device-resident ESWOA search (WOA.py:107-162): one CTA per composition instance, one thread per whale, all MAX_Iter
// iterations in one launch.  Same sequential semantics as the reference's loop (and as gnnpn_sc_b200/WOA.py's host
// engine): global phase = in-place mutations, best-so-far replayed in whale order; local phase = moves computed
// speculatively against the current best, committed up to the first improvement, the rest recomputed; `bestPops`
// aliasing kept (best_ref).  Random numbers are counter-based (Philox4x32-10, key = instance seed, counter =
// (slot, whale, phase, iteration)), so the host engine reproduces the run draw for draw with the same generator.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t slot, uint32_t whale, uint32_t phase, uint32_t t) {
  uint32_t c[4] = {slot, whale, phase, t};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  return __dmul_rn(__dadd_rn(__dmul_rn((double)(c[0] >> 5), 67108864.0), (double)(c[1] >> 6)), 1.0 / 9007199254740992.0);
}

constexpr int kWoaMaxK = 128;       // tasks per instance in the device search
constexpr int kWoaThreads = 128;    // whales per instance (popSize <= 128)

struct WoaSearchArgs {
  const double* qos;        // [T, 4]
  const int32_t* base;      // [I, Kmax] first row of every task's candidate list
  const int32_t* size;      // [I, Kmax]
  const int32_t* klen;      // [I]
  const double* bounds;     // [I, 4]
  int32_t* pops;            // [I, P, Kmax] in/out
  double* best_fit;         // [I] in/out
  int32_t* best_ref;        // [I] in/out: whale whose row IS the best position, or -1
  int32_t* best_vec;        // [I, Kmax] in/out
  const uint64_t* seeds;    // [I]
  double* traj;             // [I, iters] best fitness after every iteration
  int P, Kmax, iters;
};

// fitness of one position given as local indices (Python semantics for negatives) into the instance's candidate lists
__device__ double woa_fitness_local(const double* __restrict__ qos, const int32_t* pos, const int32_t* base, const int32_t* size,
                                    int K, const double* b) {
  double q0s[kWoaMaxK];
  double prod[2] = {1.0, 1.0};
  double min_q1 = INFINITY;
  int used = 0;
  for (int k = 0; k < K; ++k) {
    int v = pos[k] % size[k];
    if (v < 0) v += size[k];
    const double* row = qos + (int64_t)(base[k] + v) * 4;
    q0s[k] = row[0];
    used += row[0] > 0.0;
    min_q1 = fmin(min_q1, row[1]);
    prod[0] = k == 0 ? row[2] : __dmul_rn(prod[0], row[2]);
    prod[1] = k == 0 ? row[3] : __dmul_rn(prod[1], row[3]);
  }
  int viol = 0;
  for (int i = 0; i < 2; ++i) viol += (prod[i] < b[2 * i] || prod[i] > b[2 * i + 1]) ? 1 : 0;
  double obj = __ddiv_rn(numpy_sum_f64<0>(q0s, K), (double)used);
  obj = __dadd_rn(obj, 1.0);
  obj = __dsub_rn(obj, min_q1);
  obj = __ddiv_rn(obj, 2.0);
  return __dadd_rn((double)viol, obj);
}

__global__ void __launch_bounds__(kWoaThreads) woa_search_kernel(const WoaSearchArgs a) {
  extern __shared__ double sm_d[];
  const int I = blockIdx.x, i = threadIdx.x;
  const int P = a.P, KM = a.Kmax, K = a.klen[I];
  double* fit = sm_d;                        // [P]
  int32_t* pos = reinterpret_cast<int32_t*>(sm_d + P);   // [P][KM]
  int32_t* newp = pos + P * KM;              // [P][KM] speculative moves
  int32_t* bestv = newp + P * KM;            // [KM]
  int32_t* sz = bestv + KM;
  int32_t* bs = sz + KM;
  int32_t* flag = bs + KM;                   // [P] has_move
  __shared__ double s_best;
  __shared__ int s_ref, s_nxt, s_star, s_skip;
  const uint64_t seed = a.seeds[I];
  const double* bnd = a.bounds + (int64_t)I * 4;
  for (int e = i; e < P * KM; e += blockDim.x) pos[e] = a.pops[(int64_t)I * P * KM + e];
  for (int e = i; e < KM; e += blockDim.x) {
    bestv[e] = a.best_vec[(int64_t)I * KM + e];
    sz[e] = e < K ? a.size[(int64_t)I * KM + e] : 1;
    bs[e] = e < K ? a.base[(int64_t)I * KM + e] : 0;
  }
  if (i == 0) { s_best = a.best_fit[I]; s_ref = a.best_ref[I]; }
  __syncthreads();
  const bool whale = i < P;
  for (int t = 0; t < a.iters; ++t) {
    // ---- global phase (WOA.py:110-122)
    const double prob = __dmul_rn(0.2, __dsub_rn(1.0, __ddiv_rn((double)t, (double)a.iters)));
    double f = INFINITY;
    if (whale && philox_uniform(seed, 0, i, 0, t) < prob) {
      const int rand = (int)(philox_uniform(seed, 1, i, 0, t) * K);
      const int randi = (int)(philox_uniform(seed, 2, i, 0, t) * sz[rand]);
      pos[i * KM + rand] = randi;
      if (s_ref == i) bestv[rand] = randi;                 // the best IS this whale's row: it moves with it
      f = woa_fitness_local(a.qos, pos + i * KM, bs, sz, K, bnd);
    }
    if (whale) fit[i] = f;
    __syncthreads();
    if (i == 0) {
      int ref = s_ref; double best = s_best; int changed = 0;
      for (int w = 0; w < P; ++w)
        if (best > fit[w]) { best = fit[w]; ref = w; changed = 1; }
      s_best = best; s_ref = ref; s_star = changed;
      s_skip = 0.2 > philox_uniform(seed, 0, 0, 1, t);        // exploration skip (WOA.py:124-128)
    }
    __syncthreads();
    if (s_star) {
      for (int e = i; e < K; e += blockDim.x) bestv[e] = pos[s_ref * KM + e];
      __syncthreads();
    }
    if (!s_skip) {
      // ---- local phase (WOA.py:130-155)
      double A = 0, C = 0, e1 = 0, e2 = 0, pp = 1.0;
      if (whale) {
        const double aa = __dsub_rn(2.0, __ddiv_rn(__dmul_rn(2.0, (double)t), (double)a.iters));
        const double r = philox_uniform(seed, 0, i, 2, t), l = philox_uniform(seed, 1, i, 2, t);
        pp = philox_uniform(seed, 2, i, 2, t);
        A = __dsub_rn(__dmul_rn(__dmul_rn(2.0, aa), r), aa);
        C = __dmul_rn(2.0, r);
        e1 = exp(l);
        e2 = cos(__dmul_rn(__dmul_rn(2.0, 3.141592653589793), l));
      }
      const int mode = !whale ? 0 : (pp < 0.5 ? (fabs(A) < 1.0 ? 1 : 0) : 2);
      if (i == 0) s_nxt = 0;
      __syncthreads();
      while (true) {
        const int nxt = s_nxt;
        f = INFINITY;
        if (whale && i >= nxt) {
          flag[i] = mode != 0;
          if (mode) {
            for (int k = 0; k < K; ++k) {
              const double b = (double)bestv[k], x = (double)pos[i * KM + k];
              double v;
              if (mode == 1) v = __dsub_rn(b, __dmul_rn(A, __dsub_rn(__dmul_rn(C, b), x)));
              else v = __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(x, b), e1), e2), b);
              int q = (int)rint(v);
              if (abs(q) >= sz[k]) { q %= sz[k]; if (q < 0) q += sz[k]; }
              newp[i * KM + k] = q;
            }
            f = woa_fitness_local(a.qos, newp + i * KM, bs, sz, K, bnd);
          }
          fit[i] = f;
        }
        __syncthreads();
        if (i == 0) {
          int star = P;
          for (int w = nxt; w < P; ++w)
            if (flag[w] && s_best > fit[w]) { star = w; break; }
          s_star = star;
        }
        __syncthreads();
        const int star = s_star;
        const int old_ref = s_ref;
        __syncthreads();
        if (whale && i >= nxt && i <= star && flag[i]) {
          for (int k = 0; k < K; ++k) pos[i * KM + k] = newp[i * KM + k];
          if (old_ref == i && i != star) s_ref = -1;        // pops[i] was rebound: the old list stays the best (WOA.py:150)
        }
        __syncthreads();
        if (star < P) {
          for (int e = i; e < K; e += blockDim.x) bestv[e] = pos[star * KM + e];
          if (i == 0) { s_best = fit[star]; s_ref = star; s_nxt = star + 1; }
        }
        __syncthreads();
        if (star >= P || star + 1 >= P) break;
      }
    }
    if (i == 0) a.traj[(int64_t)I * a.iters + t] = s_best;
    __syncthreads();
  }
  for (int e = i; e < P * KM; e += blockDim.x) a.pops[(int64_t)I * P * KM + e] = pos[e];
  for (int e = i; e < KM; e += blockDim.x) a.best_vec[(int64_t)I * KM + e] = bestv[e];
  if (i == 0) { a.best_fit[I] = s_best; a.best_ref[I] = s_ref; }
}

}  // namespace
}  // namespace gnnpn

using namespace gnnpn;

extern "C" int gnnpn_woa_fitness_f64(const double* qos, int64_t n_services, const int32_t* idx, int64_t idx_ld,
                                     const int32_t* klen, const double* bounds, int64_t P, int Kmax,
                                     int32_t* viol_out, double* obj_out, double* fit_out, void* stream) {
  GNNPN_REQUIRE(qos && idx && bounds, GNNPN_ENULL);
  GNNPN_REQUIRE(Kmax >= 1 && Kmax <= kMaxTasksWoa && idx_ld >= Kmax && n_services >= 1 && P >= 0, GNNPN_ESHAPE);
  if (P == 0) return GNNPN_OK;
  woa_fitness_kernel<false><<<(unsigned)ceil_div(P, 64), 64, 0, (cudaStream_t)stream>>>(qos, idx, idx_ld, klen, bounds, P,
                                                                                       Kmax, viol_out, obj_out, fit_out);
  return after_launch();
}

extern "C" int gnnpn_ml2pn_score_f64(const double* qos, int64_t n_rows, const int32_t* idx, int64_t idx_ld,
                                     const int32_t* klen, const double* bounds, int64_t P, int Kmax,
                                     int32_t* viol_out, double* obj_out, double* score_out, void* stream) {
  GNNPN_REQUIRE(qos && idx && bounds, GNNPN_ENULL);
  GNNPN_REQUIRE(Kmax >= 1 && Kmax <= kMaxTasksWoa && idx_ld >= Kmax && n_rows >= 1 && P >= 0, GNNPN_ESHAPE);
  if (P == 0) return GNNPN_OK;
  woa_fitness_kernel<true><<<(unsigned)ceil_div(P, 64), 64, 0, (cudaStream_t)stream>>>(qos, idx, idx_ld, klen, bounds, P,
                                                                                      Kmax, viol_out, obj_out, score_out);
  return after_launch();
}

extern "C" int gnnpn_woa_search_f64(const double* qos, int64_t n_services, const int32_t* base, const int32_t* size,
                                    const int32_t* klen, const double* bounds, int32_t* pops, double* best_fit,
                                    int32_t* best_ref, int32_t* best_vec, const uint64_t* seeds, double* traj,
                                    int64_t n_instances, int popSize, int Kmax, int iters, void* stream) {
  GNNPN_REQUIRE(qos && base && size && klen && bounds && pops && best_fit && best_ref && best_vec && seeds && traj, GNNPN_ENULL);
  GNNPN_REQUIRE(popSize >= 1 && popSize <= kWoaThreads && Kmax >= 1 && Kmax <= kWoaMaxK && iters >= 0 && n_services >= 1 &&
                    n_instances >= 0, GNNPN_ESHAPE);
  if (n_instances == 0 || iters == 0) return GNNPN_OK;
  WoaSearchArgs a{qos, base, size, klen, bounds, pops, best_fit, best_ref, best_vec, seeds, traj, popSize, Kmax, iters};
  const size_t smem = (size_t)popSize * 8 + (size_t)(2 * popSize * Kmax + 3 * Kmax + popSize) * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(woa_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  woa_search_kernel<<<(unsigned)n_instances, kWoaThreads, smem, (cudaStream_t)stream>>>(a);
  return after_launch();
}
