// Train-mode BatchNorm1d (+ optional ReLU) forward and backward for the ML stage's training path
// (src/models/modelML.py:75-93,98-104 use torch.nn.BatchNorm1d in training mode under TrainML.train, trainML.py:34-47):
//   forward   mean_c, var_c (biased) over the M rows of y [M, C];  out = act((y - mean) * rstd * gamma + beta);
//             running_mean / running_var updated with momentum (unbiased variance), as torch does
//   backward  dgamma = sum dy * xhat, dbeta = sum dy, dx = gamma * rstd * (dy - dbeta / M - xhat * dgamma / M)
//             (dy is first masked by the ReLU: out > 0)
// One CTA per 32 channels, 32 x 8 threads; every thread walks rows ty, ty + 8, ... for its channel (coalesced along the
// channels), partial sums are combined in a fixed order -> deterministic.  Two-pass variance (mean first).
#include "common.cuh"

namespace gnnpn {
namespace {

// column sums are accumulated in float64 (a few thousand rows per channel: cheap, and the statistics / dgamma / dbeta then
// carry no summation-order error at fp32 level)
__device__ __forceinline__ double block_col_sum(double v, double (*red)[33], int tx, int ty) {
  red[ty][tx] = v;
  __syncthreads();
  double s = 0.0;
  if (ty == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    red[0][tx] = s;
  }
  __syncthreads();
  s = red[0][tx];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(256) bn_train_fwd_kernel(const float* __restrict__ y, int64_t ldy, int64_t M, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float eps, float momentum, int relu, float* __restrict__ out,
                                                           int64_t ldo, float* __restrict__ save_mean,
                                                           float* __restrict__ save_rstd, float* __restrict__ running_mean,
                                                           float* __restrict__ running_var) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  double s = 0.0;
  if (ok) for (int64_t m = ty; m < M; m += 8) s += (double)y[m * ldy + c];
  const double mean_d = block_col_sum(s, red, tx, ty) / (double)M;
  double q = 0.0;
  if (ok) for (int64_t m = ty; m < M; m += 8) { const double d = (double)y[m * ldy + c] - mean_d; q += d * d; }
  const double var_d = block_col_sum(q, red, tx, ty) / (double)M;
  const float mean = (float)mean_d, var = (float)var_d;
  const float rstd = (float)(1.0 / sqrt(var_d + (double)eps));
  if (!ok) return;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  for (int64_t m = ty; m < M; m += 8) {
    float v = (y[m * ldy + c] - mean) * rstd * g + b;
    if (relu) v = fmaxf(v, 0.f);
    out[m * ldo + c] = v;
  }
  if (ty == 0) {
    save_mean[c] = mean;
    save_rstd[c] = rstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * ((float)M / (float)(M > 1 ? M - 1 : 1));
  }
}

__global__ void __launch_bounds__(256) bn_train_bwd_kernel(const float* __restrict__ y, int64_t ldy, const float* __restrict__ out,
                                                           int64_t ldo, const float* __restrict__ dout, int64_t ldd, int64_t M,
                                                           int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ save_mean,
                                                           const float* __restrict__ save_rstd, int relu,
                                                           float* __restrict__ dx, int64_t ldx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  const float mean = ok ? save_mean[c] : 0.f, rstd = ok ? save_rstd[c] : 0.f;
  double sb = 0.0, sg = 0.0;
  if (ok)
    for (int64_t m = ty; m < M; m += 8) {
      float d = dout[m * ldd + c];
      if (relu && !(out[m * ldo + c] > 0.f)) d = 0.f;
      sb += (double)d;
      sg += (double)d * (double)((y[m * ldy + c] - mean) * rstd);
    }
  const float db = (float)block_col_sum(sb, red, tx, ty);
  const float dg = (float)block_col_sum(sg, red, tx, ty);
  if (!ok) return;
  const float g = gamma ? gamma[c] : 1.f;
  const float inv_m = 1.0f / (float)M;
  for (int64_t m = ty; m < M; m += 8) {
    float d = dout[m * ldd + c];
    if (relu && !(out[m * ldo + c] > 0.f)) d = 0.f;
    const float xh = (y[m * ldy + c] - mean) * rstd;
    dx[m * ldx + c] = g * rstd * (d - db * inv_m - xh * dg * inv_m);
  }
  if (ty == 0) { dgamma[c] = dg; dbeta[c] = db; }
}

}  // namespace
}  // namespace gnnpn

using namespace gnnpn;

extern "C" {

int gnnpn_bn_train_forward_f32(const float* y, int64_t ldy, int64_t M, int C, const float* gamma, const float* beta,
                               float eps, float momentum, int relu, float* out, int64_t ldo, float* save_mean,
                               float* save_rstd, float* running_mean, float* running_var, void* stream) {
  GNNPN_REQUIRE(y && out && save_mean && save_rstd, GNNPN_ENULL);
  GNNPN_REQUIRE(M >= 1 && C >= 1 && ldy >= C && ldo >= C, GNNPN_ESHAPE);
  bn_train_fwd_kernel<<<(unsigned)ceil_div(C, 32), 256, 0, (cudaStream_t)stream>>>(y, ldy, M, C, gamma, beta, eps, momentum, relu,
                                                                                  out, ldo, save_mean, save_rstd,
                                                                                  running_mean, running_var);
  return after_launch();
}

int gnnpn_bn_train_backward_f32(const float* y, int64_t ldy, const float* out, int64_t ldo, const float* dout, int64_t ldd,
                                int64_t M, int C, const float* gamma, const float* save_mean, const float* save_rstd,
                                int relu, float* dx, int64_t ldx, float* dgamma, float* dbeta, void* stream) {
  GNNPN_REQUIRE(y && out && dout && save_mean && save_rstd && dx && dgamma && dbeta, GNNPN_ENULL);
  GNNPN_REQUIRE(M >= 1 && C >= 1 && ldy >= C && ldo >= C && ldd >= C && ldx >= C, GNNPN_ESHAPE);
  bn_train_bwd_kernel<<<(unsigned)ceil_div(C, 32), 256, 0, (cudaStream_t)stream>>>(y, ldy, out, ldo, dout, ldd, M, C, gamma,
                                                                                  save_mean, save_rstd, relu, dx, ldx, dgamma,
                                                                                  dbeta);
  return after_launch();
}

}  // extern "C"
