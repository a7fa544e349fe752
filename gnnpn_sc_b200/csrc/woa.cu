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
  double obj = __ddiv_rn(numpy_sum_f64<3>(q0s, K), (double)used);
  obj = __dadd_rn(obj, 1.0);
  obj = __dsub_rn(obj, min_q1);
  obj = __ddiv_rn(obj, 2.0);
  if (viol_out) viol_out[p] = viol;
  if (obj_out) obj_out[p] = obj;
  if (fit_out) fit_out[p] = __dadd_rn((double)viol, obj);
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
  woa_fitness_kernel<<<(unsigned)ceil_div(P, 64), 64, 0, (cudaStream_t)stream>>>(qos, idx, idx_ld, klen, bounds, P, Kmax,
                                                                                viol_out, obj_out, fit_out);
  return after_launch();
}
