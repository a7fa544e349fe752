// FFMA implementation of the batched LSTM step (strict fp32, the parity anchor).
// Tile: 128 instances x 128 gate columns (= 32 hidden units x 4 gates), k-tiles of 16,
// 256 threads, 8x8 accumulators per thread, register-prefetched double-buffered smem.
#include "lstm_step.cuh"

namespace gnnpn {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TPB = 256;
constexpr int AS_LD = BM + 4;   // +4 floats: keeps float4 alignment, spreads k-rows over banks

__global__ void __launch_bounds__(TPB, 2) lstm_step_ffma_kernel(const LstmStepArgs a) {
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // global -> register staging roles
  const int a_row = tid & (BM - 1);          // instance row inside the tile
  const int a_half = tid >> 7;               // which 8 of the 16 k-columns
  const int b_row = tid >> 5;                // 0..7 (+8)
  const int b_col4 = tid & 31;               // float4 column

  const int gm = m0 + a_row;
  const bool row_ok = gm < a.M;
  const float* h_row = (row_ok && !a.first) ? a.h_in + (int64_t)gm * a.h_in_ld : nullptr;
  const float* x_row = nullptr;
  if (row_ok && a.use_x) {
    const int r = a.x_row >= 0 ? a.x_row : a.gather[gm];
    x_row = a.x + (int64_t)gm * a.x_inst_ld + (int64_t)r * a.F;
  }

  const int kt_begin = a.first ? kH / BK : 0;
  const int kt_end = kH / BK + (a.use_x ? (a.F + BK - 1) / BK : 0);   // one or two k-tiles of raw input columns

  float4 ra[2], rb[2];
  auto load_regs = [&](int kt) {
    if (kt < kH / BK) {
      if (h_row) {
        const float4* p = reinterpret_cast<const float4*>(h_row + kt * BK + a_half * 8);
        ra[0] = __ldg(p);
        ra[1] = __ldg(p + 1);
      } else {
        ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = (kt - kH / BK) * BK + a_half * 8 + i;
        v[i] = (x_row && f < a.F) ? __ldg(x_row + f) : 0.f;
      }
      ra[0] = make_float4(v[0], v[1], v[2], v[3]);
      ra[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    const float4* w = reinterpret_cast<const float4*>(a.P + (int64_t)(kt * BK + b_row) * kG + n0) + b_col4;
    rb[0] = __ldg(w);
    rb[1] = __ldg(w + 8 * (kG / 4));
  };
  auto store_smem = [&](int buf) {
    float* as = &As[buf][a_half * 8][a_row];
    as[0 * AS_LD] = ra[0].x; as[1 * AS_LD] = ra[0].y; as[2 * AS_LD] = ra[0].z; as[3 * AS_LD] = ra[0].w;
    as[4 * AS_LD] = ra[1].x; as[5 * AS_LD] = ra[1].y; as[6 * AS_LD] = ra[1].z; as[7 * AS_LD] = ra[1].w;
    *reinterpret_cast<float4*>(&Bs[buf][b_row][b_col4 * 4]) = rb[0];
    *reinterpret_cast<float4*>(&Bs[buf][b_row + 8][b_col4 * 4]) = rb[1];
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (kt_begin < kt_end) {
    load_regs(kt_begin);
    store_smem(0);
  }
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    __syncthreads();
    const bool more = kt + 1 < kt_end;
    if (more) load_regs(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    buf ^= 1;
  }

  // ---- fused cell update: this thread holds gates i,f,g,o of hidden units jA and jB for 8 instances
  const int jA = (n0 >> 2) + tx, jB = jA + 16;
  const float4 biasA = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + tx * 4));
  const float4 biasB = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 64 + tx * 4));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= a.M) continue;
    float* crow = a.c + (int64_t)m * kH;
    float* hrow = a.h_out + (int64_t)m * a.h_out_ld;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = u ? jB : jA;
      const float4 bb = u ? biasB : biasA;
      const float gi = acc[i][u * 4 + 0] + bb.x;
      const float gf = acc[i][u * 4 + 1] + bb.y;
      const float gg = acc[i][u * 4 + 2] + bb.z;
      const float go = acc[i][u * 4 + 3] + bb.w;
      const float c_old = a.first ? 0.f : crow[j];
      const float c_new = sigmoid_accurate(gf) * c_old + sigmoid_accurate(gi) * tanhf(gg);
      crow[j] = c_new;
      hrow[j] = sigmoid_accurate(go) * tanhf(c_new);
    }
  }
}

}  // namespace

int launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return GNNPN_OK;
  dim3 grid((unsigned)ceil_div(a.M, BM), kG / BN);
  lstm_step_ffma_kernel<<<grid, TPB, 0, stream>>>(a);
  return after_launch();
}

}  // namespace gnnpn
