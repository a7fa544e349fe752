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
    const float* cin = (a.c_in ? a.c_in : a.c) + (int64_t)m * kH;
    float* hrow = a.h_out + (int64_t)m * a.h_out_ld;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = u ? jB : jA;
      const float4 bb = u ? biasB : biasA;
      const float gi = acc[i][u * 4 + 0] + bb.x;
      const float gf = acc[i][u * 4 + 1] + bb.y;
      const float gg = acc[i][u * 4 + 2] + bb.z;
      const float go = acc[i][u * 4 + 3] + bb.w;
      const float c_old = a.first ? 0.f : cin[j];        // training replay: the previous step's saved slot (c_in), not this step's
      const float si = sigmoid_accurate(gi), sf = sigmoid_accurate(gf), tg = tanhf(gg), so = sigmoid_accurate(go);
      const float c_new = sf * c_old + si * tg;
      crow[j] = c_new;
      hrow[j] = so * tanhf(c_new);
      if (a.gates_out) *reinterpret_cast<float4*>(a.gates_out + (int64_t)m * kG + 4 * j) = make_float4(si, sf, tg, so);
    }
  }
}

// ---- small-batch variant (M <= 1024, e.g. the reference's training batch of 128): the 128 x 128 tile above leaves a
// batch of 128 on 8 CTAs (~18 us per step); here a CTA owns 32 instances x 32 gate columns (8 hidden units) with the whole
// K = H + F panel in shared memory -> 128 CTAs for a batch of 128, ~2 us per step.  Same fmaf order over k (h part, then
// the raw input columns) as the big tile: bit-identical results.
constexpr int SM_ROWS = 32, SM_COLS = 32, SM_TPB = 128, SM_K = kH + 8, SM_ALD = SM_K + 1;
constexpr int SM_SMEM = (SM_ROWS * SM_ALD + SM_K * SM_COLS) * 4;

__global__ void __launch_bounds__(SM_TPB) lstm_step_small_kernel(const LstmStepArgs a) {
  extern __shared__ __align__(16) float smem_f[];
  float* As = smem_f;                          // [32 rows][SM_ALD]
  float* Bs = smem_f + SM_ROWS * SM_ALD;       // [K][32 cols]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * SM_COLS;
  const int m0 = blockIdx.y * SM_ROWS;
  const int Kx = a.use_x ? a.F : 0;            // raw input columns (<= 8)
  const int k_begin = a.first ? kH : 0;
  const int k_end = kH + Kx;
  // A panel: [h | x] rows of this tile
  if (!a.first) {
    for (int i = tid; i < SM_ROWS * (kH / 4); i += SM_TPB) {
      const int r = i / (kH / 4), c4 = i % (kH / 4);
      const int m = m0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) v = __ldg(reinterpret_cast<const float4*>(a.h_in + (int64_t)m * a.h_in_ld) + c4);
      float* d = As + r * SM_ALD + c4 * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  }
  for (int i = tid; i < SM_ROWS * Kx; i += SM_TPB) {
    const int r = i / Kx, f = i % Kx;
    const int m = m0 + r;
    float v = 0.f;
    if (m < a.M) {
      const int row = a.x_row >= 0 ? a.x_row : a.gather[m];
      v = __ldg(a.x + (int64_t)m * a.x_inst_ld + (int64_t)row * a.F + f);
    }
    As[r * SM_ALD + kH + f] = v;
  }
  // B panel: packed weights, k-major rows of kG gate columns
  for (int i = tid; i < (k_end - k_begin) * (SM_COLS / 4); i += SM_TPB) {
    const int k = k_begin + i / (SM_COLS / 4), c4 = i % (SM_COLS / 4);
    reinterpret_cast<float4*>(Bs + k * SM_COLS)[c4] = __ldg(reinterpret_cast<const float4*>(a.P + (int64_t)k * kG + n0) + c4);
  }
  __syncthreads();
  // thread = 1 instance row x 2 hidden units (8 gate columns)
  const int r = tid >> 2, ug = tid & 3;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* arow = As + r * SM_ALD;
  for (int k = k_begin; k < k_end; ++k) {
    const float av = arow[k];
    const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * SM_COLS + ug * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * SM_COLS + ug * 8 + 4);
    acc[0] = fmaf(av, b0.x, acc[0]); acc[1] = fmaf(av, b0.y, acc[1]); acc[2] = fmaf(av, b0.z, acc[2]); acc[3] = fmaf(av, b0.w, acc[3]);
    acc[4] = fmaf(av, b1.x, acc[4]); acc[5] = fmaf(av, b1.y, acc[5]); acc[6] = fmaf(av, b1.z, acc[6]); acc[7] = fmaf(av, b1.w, acc[7]);
  }
  const int m = m0 + r;
  if (m >= a.M) return;
  float* crow = a.c + (int64_t)m * kH;
  const float* cin = (a.c_in ? a.c_in : a.c) + (int64_t)m * kH;
  float* hrow = a.h_out + (int64_t)m * a.h_out_ld;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = (n0 >> 2) + ug * 2 + u;
    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + 4 * j));
    const float gi = acc[u * 4 + 0] + bb.x, gf = acc[u * 4 + 1] + bb.y, gg = acc[u * 4 + 2] + bb.z, go = acc[u * 4 + 3] + bb.w;
    const float c_old = a.first ? 0.f : cin[j];
    const float si = sigmoid_accurate(gi), sf = sigmoid_accurate(gf), tg = tanhf(gg), so = sigmoid_accurate(go);
    const float c_new = sf * c_old + si * tg;
    crow[j] = c_new;
    hrow[j] = so * tanhf(c_new);
    if (a.gates_out) *reinterpret_cast<float4*>(a.gates_out + (int64_t)m * kG + 4 * j) = make_float4(si, sf, tg, so);
  }
}

}  // namespace

int launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return GNNPN_OK;
  if (a.M <= 1024 && a.F <= 8) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(lstm_step_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    dim3 grid(kG / SM_COLS, (unsigned)ceil_div(a.M, SM_ROWS));
    lstm_step_small_kernel<<<grid, SM_TPB, SM_SMEM, stream>>>(a);
    return after_launch();
  }
  dim3 grid((unsigned)ceil_div(a.M, BM), kG / BN);
  lstm_step_ffma_kernel<<<grid, TPB, 0, stream>>>(a);
  return after_launch();
}

}  // namespace gnnpn
