// Pointer-network entry points: weight packing, encoder scan, fused greedy decode,
// interface-faithful logits materialisation, composition objective / reward.
#include <math.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "lstm_step.cuh"
#include "tc_lstm.cuh"
#include "tc_seq.cuh"
#include "pointer.cuh"
#include "options.cuh"

namespace gnnpn {
namespace {

// ---------------------------------------------------------------------------
// weight packing (fp64 accumulate, once per model)
// ---------------------------------------------------------------------------
__device__ float folded_weight(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                               const float* __restrict__ w_e, int H, int F, int r, int k) {
  // column k of the "[h | x]" operand for torch gate row r:  k < H -> W_hh ; H <= k < H+F -> (W_ih . W_e)[:, k-H]
  if (k < H) return w_hh[(int64_t)r * H + k];
  const int f = k - H;
  if (f >= F) return 0.f;
  double s = 0.0;
  for (int i = 0; i < H; ++i) s += (double)w_ih[(int64_t)r * H + i] * (double)w_e[(int64_t)i * F + f];
  return (float)s;
}

__global__ void pack_lstm_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                 const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                 const float* __restrict__ w_e, const float* __restrict__ b_e,
                                 const float* __restrict__ start, int H, int F, int Fpad, int Kp,
                                 float* __restrict__ packed) {
  const int G = 4 * H;
  const int rows = H + Fpad + 2;  // + bias row + start row
  const int64_t ffma_total = (int64_t)rows * G;
  const int64_t tc_total = (int64_t)G * Kp;
  const int64_t tc16_total = (int64_t)G * kKp16;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ffma_total + tc_total + tc16_total;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (e >= ffma_total + tc_total) {              // fp16-split tcgen05 operand: [4H][kKp16] halfs, weights x 2^4
      const int64_t t = e - ffma_total - tc_total;
      const int nn = (int)(t / kKp16), k = (int)(t % kKp16);
      const int r = (nn & 3) * H + (nn >> 2);
      const float w = kW16Scale * folded_weight(w_ih, w_hh, w_e, H, F, r, k);
      const __half hi = __float2half_rn(w);
      __half* base = reinterpret_cast<__half*>(packed + ffma_total + 2 * tc_total);
      base[t] = hi;
      base[tc16_total + t] = __float2half_rn(w - __half2float(hi));
      continue;
    }
    if (e >= ffma_total) {                         // tcgen05 operand: [4H gate columns][Kp], K contiguous
      const int64_t t = e - ffma_total;
      const int nn = (int)(t / Kp), k = (int)(t % Kp);
      const int r = (nn & 3) * H + (nn >> 2);
      const float w = folded_weight(w_ih, w_hh, w_e, H, F, r, k);
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(w));
      const float hi = __uint_as_float(hb);
      packed[ffma_total + t] = hi;
      packed[ffma_total + tc_total + t] = w - hi;
      continue;
    }
    const int k = (int)(e / G), nn = (int)(e % G);
    const int j = nn >> 2, g = nn & 3;
    const int r = g * H + j;                       // torch row: gate-major
    float out = 0.f;
    if (k < H + Fpad) {
      out = folded_weight(w_ih, w_hh, w_e, H, F, r, k);
    } else if (k == H + Fpad) {                    // bias = b_ih + b_hh + W_ih . b_e
      double s = (double)b_ih[r] + (double)b_hh[r];
      for (int i = 0; i < H; ++i) s += (double)w_ih[(int64_t)r * H + i] * (double)b_e[i];
      out = (float)s;
    } else {                                       // start = b_ih + b_hh + W_ih . start_input
      if (start) {
        double s = (double)b_ih[r] + (double)b_hh[r];
        for (int i = 0; i < H; ++i) s += (double)w_ih[(int64_t)r * H + i] * (double)start[i];
        out = (float)s;
      }
    }
    packed[e] = out;
  }
}

// ---------------------------------------------------------------------------
// pointer step (Dot attention): one warp per instance
//   u_j   = <enc_out[b, kN+j, :], q[b,:]>            j in [0,N)
//   l_j   = use_tanh ? C*tanh(u_j) : u_j            -> win_logits[b, kN+j]
//   w_j   = l_j + alpha*latent[b, kN+j]
//   p     = softmax_j(w)  (exp(w-max)/sum, fp32)    -> win_probs[b, kN+j]
//   pick  = first j with maximal p                  -> idx_out[b] = kN + j
// Positions outside the window carry -inf after modelPN.py:220-222 and contribute exp(-inf)=0.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pointer_step_dot_kernel(
    const PointerStepArgs pa, int k, const float* __restrict__ q, int64_t q_ld, const float* __restrict__ inputs, int F,
    void* __restrict__ a_hi_next, void* __restrict__ a_lo_next, int64_t a_ld, int a_f16) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= pa.n) return;
  const float4* qp = reinterpret_cast<const float4*>(q + b * q_ld);
  const float4 q0 = __ldg(qp + 2 * lane), q1 = __ldg(qp + 2 * lane + 1);     // the lane's floats [8*lane, 8*lane+8)
  const int fed = pointer_step_warp(pa, k, b, q0, q1, lane);
  if (a_hi_next && lane < F) {
    // tensor-core path: the chosen candidate's raw row becomes columns [H, H+F) of the next step's A operand
    const float v = __ldg(inputs + (b * pa.L + fed) * (int64_t)F + lane);
    if (a_f16) {
      const __half hi = __float2half_rn(v);
      reinterpret_cast<__half*>(a_hi_next)[b * a_ld + kH + lane] = hi;
      reinterpret_cast<__half*>(a_lo_next)[b * a_ld + kH + lane] = __float2half_rn(v - __half2float(hi));
    } else {
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
      reinterpret_cast<float*>(a_hi_next)[b * a_ld + kH + lane] = __uint_as_float(hb);
      reinterpret_cast<float*>(a_lo_next)[b * a_ld + kH + lane] = v - __uint_as_float(hb);
    }
  }
}

// ---------------------------------------------------------------------------
// full logits: one CTA per (instance, 32-row slab of L); queries of the instance in smem
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) full_logits_dot_kernel(
    const float* __restrict__ enc_out, const float* __restrict__ dec_h, const int32_t* __restrict__ idx,
    int use_tanh, float C, int64_t n, int L, int K, float* __restrict__ out) {
  extern __shared__ float sq[];                // [K][kH]
  const int64_t b = blockIdx.x;
  const int l0 = blockIdx.y * 32;
  const float4* qsrc = reinterpret_cast<const float4*>(dec_h + b * (int64_t)K * kH);
  for (int i = threadIdx.x; i < K * kH / 4; i += blockDim.x) reinterpret_cast<float4*>(sq)[i] = __ldg(qsrc + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int l = l0 + warp; l < min(L, l0 + 32); l += 8) {
    const float4* rp = reinterpret_cast<const float4*>(enc_out + (b * L + l) * (int64_t)kH);
    const float4 r0 = __ldg(rp + 2 * lane), r1 = __ldg(rp + 2 * lane + 1);   // canonical dot: lane owns floats [8*lane, +8)
    for (int k = 0; k < K; ++k) {
      const float4 a0 = reinterpret_cast<const float4*>(sq + k * kH)[2 * lane];
      const float4 a1 = reinterpret_cast<const float4*>(sq + k * kH)[2 * lane + 1];
      const float d = dot_reduce(dot8(r0, r1, a0, a1), lane);
      if (lane == 0) out[((int64_t)k * n + b) * L + l] = use_tanh ? C * tanhf(d) : d;
    }
  }
  // cumulative visited mask: step k sees -inf at the picks of steps 0..k-1 (modelPN.py:165-173)
  __syncthreads();
  for (int pair = threadIdx.x; pair < K * K; pair += blockDim.x) {
    const int k = pair / K, jprev = pair % K;
    if (jprev >= k) continue;
    const int pos = idx[(int64_t)jprev * n + b];
    if (pos >= l0 && pos < l0 + 32 && pos < L) out[((int64_t)k * n + b) * L + pos] = -INFINITY;
  }
}

// ---------------------------------------------------------------------------
// composition objective (calc / reward): one thread per instance, fp32 sequential products
// and numpy's float32 pairwise-sum order for sum(q0) so results are bit-identical to the CPU path
// ---------------------------------------------------------------------------
__device__ float numpy_pairwise_sum_f32(const float* v, int n) {
  // numpy pairwise_sum for n < 128 (PW_BLOCKSIZE): 8 accumulators, then the tail in order.
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, v[i]);
    return r;
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], v[i + j]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, v[i]);
  return res;
}

// numpy splits n > 128 as (n/2 rounded down to a multiple of 8) + rest, recursively; unrolled by
// template depth (3 levels cover n <= 1024) because device recursion has no static stack bound.
template <int DEPTH>
__device__ float numpy_sum_f32(const float* v, int n) {
  if (n <= 128) return numpy_pairwise_sum_f32(v, n);
  if constexpr (DEPTH == 0) {
    return numpy_pairwise_sum_f32(v, n);   // unreachable for n <= kMaxTasks
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(numpy_sum_f32<DEPTH - 1>(v, n2), numpy_sum_f32<DEPTH - 1>(v + n2, n - n2));
  }
}

constexpr int kMaxTasks = 512;

__global__ void reward_kernel(const float* __restrict__ inputs, const int32_t* __restrict__ idx, int64_t n,
                              int L, int F, int K, int tag, int32_t* __restrict__ viol_out,
                              float* __restrict__ obj_out, float* __restrict__ rew_out) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= n) return;
  float q0s[kMaxTasks];
  float prod[2] = {1.f, 1.f};
  float min_q1 = INFINITY;
  int used = 0;
  float lo[2] = {0, 0}, hi[2] = {0, 0};
  for (int k = 0; k < K; ++k) {
    const float* row = inputs + (b * L + idx[(int64_t)k * n + b]) * (int64_t)F + tag;
    const float q0 = row[0], q1 = row[1];
    q0s[k] = q0;
    used += q0 > 0.f;
    min_q1 = fminf(min_q1, q1);
    prod[0] = k == 0 ? row[2] : __fmul_rn(prod[0], row[2]);
    prod[1] = k == 0 ? row[3] : __fmul_rn(prod[1], row[3]);
    if (k == 0) { lo[0] = row[4]; hi[0] = row[5]; lo[1] = row[6]; hi[1] = row[7]; }
  }
  int viol = 0;
  for (int i = 0; i < 2; ++i) viol += (prod[i] < lo[i] || prod[i] > hi[i]) ? 1 : 0;
  const float s = numpy_sum_f32<2>(q0s, K);
  float obj = __fdiv_rn(s, (float)used);
  obj = __fadd_rn(obj, 1.0f);
  obj = __fsub_rn(obj, min_q1);
  obj = __fdiv_rn(obj, 2.0f);
  if (viol_out) viol_out[b] = viol;
  if (obj_out) obj_out[b] = obj;
  if (rew_out) {
    const double v = (double)viol + (double)obj;          // python: round(violate + objFunc, 5)
    rew_out[b] = (float)(rint(v * 1e5) / 1e5);
  }
}

// option "persistent": bit 0 = persistent encoder scan, bit 1 = persistent fused decode (default 3); 0 selects the
// one-launch-per-step kernels (kept for inputs wider than 8 columns and as the A/B reference)
int seq_mode() { return options().persistent.load(std::memory_order_relaxed); }

// raw input range check: *flag |= 1 if any value is NaN / inf or has |x| >= limit
__global__ void check_inputs_kernel(const float* __restrict__ x, int64_t count, float limit, int32_t* __restrict__ flag) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    bad |= !(fabsf(__ldg(x + i)) < limit);
  if (bad) atomicOr(flag, 1);
}

// blocked encodings -> row-major [n, L, kH]; one thread per float4 of the output (coalesced writes)
__global__ void enc_unblock_kernel(const float* __restrict__ blk, int64_t n, int L, float* __restrict__ out) {
  const int64_t total = n * L * (kH / 4);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(e % (kH / 4));
    const int64_t bl = e / (kH / 4);
    const int l = (int)(bl % L);
    const int64_t b = bl / L;
    const int u = c4 * 4, nt = u >> 5, g = (u >> 3) & 3;
    const int64_t src = ((((b >> 7) * L + l) * 8 + nt) * 4 + g) * 1024 + (b & 127) * 8 + (u & 7);
    reinterpret_cast<float4*>(out)[e] = __ldg(reinterpret_cast<const float4*>(blk + src));
  }
}

}  // namespace
}  // namespace gnnpn

using namespace gnnpn;

extern "C" {

size_t gnnpn_pn_packed_lstm_floats(int hidden, int in_features) {
  (void)in_features;
  return hidden == kH ? kPackedFloats : 0;
}

size_t gnnpn_pn_workspace_bytes(int64_t n, int hidden) {
  return hidden == kH && n >= 0 ? tc_lstm_workspace_bytes(n) : 0;
}

int gnnpn_pn_pack_lstm_f32(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                           const float* w_embed, const float* b_embed, const float* start_input,
                           int hidden, int in_features, float* packed, void* stream) {
  GNNPN_REQUIRE(w_ih && w_hh && b_ih && b_hh && w_embed && b_embed && packed, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= kXPad, GNNPN_ESHAPE);
  pack_lstm_kernel<<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, w_embed, b_embed,
                                                                  start_input, hidden, in_features, kXPad, kKp,
                                                                  packed);
  return after_launch();
}

float gnnpn_pn_input_limit(void) { return 65504.0f; }

int gnnpn_pn_check_inputs_f32(const float* inputs, int64_t count, int32_t* flag, void* stream) {
  GNNPN_REQUIRE(inputs && flag, GNNPN_ENULL);
  GNNPN_REQUIRE(count >= 0, GNNPN_ESHAPE);
  if (count == 0) return GNNPN_OK;
  const int64_t blocks = ceil_div(count, 256 * 8);
  check_inputs_kernel<<<(unsigned)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(
      inputs, count, gnnpn_pn_input_limit(), flag);
  return after_launch();
}

int gnnpn_pn_enc_layout(int64_t n, int L, int in_features, int K, int N, int has_workspace) {
  (void)L; (void)K;
  if (!has_workspace || seq_mode() != 3 || in_features < 1 || in_features > 8) return GNNPN_ENC_ROWMAJOR;
  if (!tc_seq_fused_decode_supported(N)) return GNNPN_ENC_ROWMAJOR;
  if (tc_colsplit_wanted(n) || tc_colsplit_wanted_encode(n)) return GNNPN_ENC_ROWMAJOR;   // small batch: cluster scan
  return GNNPN_ENC_BLOCKED128;
}

size_t gnnpn_pn_enc_out_floats(int64_t n, int L, int hidden, int layout) {
  if (n < 0 || L < 1 || hidden != kH) return 0;
  const int64_t rows = layout == GNNPN_ENC_BLOCKED128 ? (n + 127) / 128 * 128 : n;
  return (size_t)rows * L * kH;
}

int gnnpn_pn_enc_to_rowmajor_f32(const float* enc_blocked, int64_t n, int L, int hidden, float* enc_out, void* stream) {
  GNNPN_REQUIRE(enc_blocked && enc_out, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && L >= 1 && n >= 0, GNNPN_ESHAPE);
  GNNPN_REQUIRE(aligned16(enc_blocked) && aligned16(enc_out), GNNPN_EALIGN);
  if (n == 0) return GNNPN_OK;
  enc_unblock_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(enc_blocked, n, L, enc_out);
  return after_launch();
}

int gnnpn_lstm_encode_f32(const float* inputs, int64_t n, int L, int in_features, int hidden,
                          const float* packed, float* enc_out, float* c_state, void* workspace,
                          size_t workspace_bytes, int enc_layout, void* stream) {
  GNNPN_REQUIRE(enc_layout == GNNPN_ENC_ROWMAJOR || enc_layout == GNNPN_ENC_BLOCKED128, GNNPN_ESHAPE);
  GNNPN_REQUIRE(inputs && packed && enc_out && c_state, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= kXPad && L >= 1 && n >= 0, GNNPN_ESHAPE);
  GNNPN_REQUIRE(n < (1ll << 31), GNNPN_ERANGE);
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(c_state) && aligned16(packed), GNNPN_EALIGN);
  if (n == 0) return GNNPN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace && (seq_mode() & 1) && in_features <= 8) {
    // ---- persistent tcgen05 scan: one launch for all L steps, h resident in shared memory
    GNNPN_REQUIRE((reinterpret_cast<uintptr_t>(enc_out) & 31u) == 0 && (reinterpret_cast<uintptr_t>(c_state) & 31u) == 0,
                  GNNPN_EALIGN);
    GNNPN_REQUIRE(workspace_bytes >= tc_lstm_workspace_bytes(n), GNNPN_EWORKSPACE);
    float* scr = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    SeqEncodeArgs sa{inputs, n, L, in_features, packed, enc_out, c_state, scr, enc_layout};
    if (enc_layout == GNNPN_ENC_BLOCKED128) return tc_seq_encode(sa, st);         // the CTA-pair scan owns this layout
    if (tc_colsplit_wanted_encode(n)) return tc_colsplit_encode(sa, scr, st);     // small batch: column-split cluster scan
    return tc_seq_encode(sa, st);
  }
  GNNPN_REQUIRE(enc_layout == GNNPN_ENC_ROWMAJOR, GNNPN_EUNSUPPORTED);
  if (workspace) {
    // ---- tcgen05 recurrence, one launch per step: [h|x] kept as hi/lo pairs in two ping-pong buffers
    TcLstmPlan plan;
    int rc = tc_lstm_plan(&plan, workspace, workspace_bytes, n, packed);
    if (rc) return rc;
    if ((rc = tc_lstm_reset(plan, inputs, (int64_t)L * in_features, 0, in_features, st))) return rc;
    TcLstmStep s{};
    s.use_x = 1; s.bias = packed + kOffBias; s.c = c_state; s.h_out_ld = (int64_t)L * kH;
    s.x_next = inputs; s.x_inst_ld = (int64_t)L * in_features; s.F = in_features;
    for (int t = 0; t < L; ++t) {
      s.cur = t & 1; s.first = t == 0;
      s.h_out = enc_out + (int64_t)t * kH;
      s.x_row_next = t + 1 < L ? t + 1 : -1;
      if ((rc = tc_lstm_step(plan, s, st))) return rc;
    }
    return GNNPN_OK;
  }
  LstmStepArgs a{};
  a.x = inputs; a.x_inst_ld = (int64_t)L * in_features; a.F = in_features; a.use_x = 1; a.gather = nullptr;
  a.P = packed; a.bias = packed + kOffBias;
  a.c = c_state; a.M = (int)n;
  a.h_in_ld = a.h_out_ld = (int64_t)L * kH;
  for (int t = 0; t < L; ++t) {
    a.first = t == 0;
    a.h_in = t == 0 ? nullptr : enc_out + (int64_t)(t - 1) * kH;
    a.h_out = enc_out + (int64_t)t * kH;
    a.x_row = t;
    int rc = launch_lstm_step(a, st);
    if (rc) return rc;
  }
  return GNNPN_OK;
}

int gnnpn_pn_decode_greedy_f32(const float* inputs, const float* enc_out, float* c_state,
                               const float* latent_win, float alpha, const float* packed,
                               int attention, const float* att_params, int use_tanh, float C,
                               int64_t n, int L, int in_features, int hidden, int K, int N,
                               float* dec_h, int32_t* idx_out, float* win_logits, float* win_probs,
                               const int32_t* forced_idx, const float* sample_uniform, void* workspace,
                               size_t workspace_bytes, int enc_layout, void* stream) {
  GNNPN_REQUIRE(enc_layout == GNNPN_ENC_ROWMAJOR || enc_layout == GNNPN_ENC_BLOCKED128, GNNPN_ESHAPE);
  GNNPN_REQUIRE(inputs && enc_out && c_state && packed && idx_out && win_logits && win_probs, GNNPN_ENULL);
  GNNPN_REQUIRE(dec_h || enc_layout == GNNPN_ENC_BLOCKED128, GNNPN_ENULL);   // optional only in the fused decoder
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= kXPad, GNNPN_ESHAPE);
  GNNPN_REQUIRE(K >= 1 && N >= 1 && N <= kMaxWindow && (int64_t)K * N == L, GNNPN_ESHAPE);
  GNNPN_REQUIRE(n >= 0 && n < (1ll << 31), GNNPN_ERANGE);
  GNNPN_REQUIRE(attention == GNNPN_ATT_DOT, GNNPN_EUNSUPPORTED);
  (void)att_params;
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(dec_h) && aligned16(c_state) && aligned16(packed), GNNPN_EALIGN);   // NULL dec_h passes
  if (n == 0) return GNNPN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float* bias = packed + kOffBias;
  const float* start = packed + kOffStart;
  const unsigned att_blocks = (unsigned)ceil_div(n, 8);
  const bool use_tc = workspace != nullptr;
  if (use_tc && (seq_mode() & 2) && in_features <= 8) {
    GNNPN_REQUIRE((reinterpret_cast<uintptr_t>(enc_out) & 31u) == 0 && (reinterpret_cast<uintptr_t>(c_state) & 31u) == 0 &&
                      (reinterpret_cast<uintptr_t>(dec_h) & 31u) == 0, GNNPN_EALIGN);
    SeqDecodeArgs sa{inputs, enc_out, c_state, latent_win, alpha, packed, use_tanh, C, n, L, in_features, K, N,
                     dec_h, idx_out, win_logits, win_probs, forced_idx, sample_uniform,
                     reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023)),
                     enc_layout};
    GNNPN_REQUIRE(workspace_bytes >= tc_lstm_workspace_bytes(n), GNNPN_EWORKSPACE);
    if (enc_layout == GNNPN_ENC_BLOCKED128) return tc_seq_decode(sa, st);          // fused pointer dots (CTA-pair scan)
    if (tc_colsplit_wanted(n)) return tc_colsplit_decode(sa, sa.c_scratch, st);    // small batch: column-split cluster scan
    return tc_seq_decode(sa, st);
  }
  GNNPN_REQUIRE(enc_layout == GNNPN_ENC_ROWMAJOR, GNNPN_EUNSUPPORTED);
  TcLstmPlan plan;
  TcLstmStep ts{};
  LstmStepArgs a{};
  int rc;
  if (use_tc) {
    if ((rc = tc_lstm_plan(&plan, workspace, workspace_bytes, n, packed))) return rc;
    // decoder state starts from the encoder's last hidden state (modelPN.py:191,205)
    if ((rc = tc_lstm_load_h(plan, 0, enc_out + (int64_t)(L - 1) * kH, (int64_t)L * kH, st))) return rc;
    if ((rc = tc_lstm_zero(plan, 1, st))) return rc;
    ts.c = c_state; ts.h_out_ld = (int64_t)K * kH; ts.x_next = nullptr; ts.x_row_next = -1;
    ts.F = in_features; ts.first = 0;
  } else {
    a.x = inputs; a.x_inst_ld = (int64_t)L * in_features; a.F = in_features; a.x_row = -1;
    a.P = packed; a.c = c_state; a.M = (int)n; a.first = 0;
    a.h_out_ld = (int64_t)K * kH;
  }
  for (int k = 0; k < K; ++k) {
    const int32_t* fed_prev = k == 0 ? nullptr : (forced_idx ? forced_idx : idx_out) + (int64_t)(k - 1) * n;
    if (use_tc) {
      ts.cur = k & 1; ts.use_x = k != 0; ts.bias = k == 0 ? start : bias;
      ts.h_out = dec_h + (int64_t)k * kH;
      if ((rc = tc_lstm_step(plan, ts, st))) return rc;
    } else {
      if (k == 0) {
        a.h_in = enc_out + (int64_t)(L - 1) * kH; a.h_in_ld = (int64_t)L * kH;
        a.use_x = 0; a.bias = start; a.gather = nullptr;
      } else {
        a.h_in = dec_h + (int64_t)(k - 1) * kH; a.h_in_ld = (int64_t)K * kH;
        a.use_x = 1; a.bias = bias; a.gather = fed_prev;
      }
      a.h_out = dec_h + (int64_t)k * kH;
      if ((rc = launch_lstm_step(a, st))) return rc;
    }
    const int nxt = (k + 1) & 1;
    PointerStepArgs pa;
    pa.enc_out = enc_out; pa.enc_inst_ld = (int64_t)L * kH; pa.latent_win = latent_win; pa.alpha = alpha;
    pa.use_tanh = use_tanh; pa.C = C; pa.n = n; pa.L = L; pa.N = N;
    pa.idx_out = idx_out; pa.win_logits = win_logits; pa.win_probs = win_probs;
    pa.forced = forced_idx; pa.uniform = sample_uniform;
    pointer_step_dot_kernel<<<att_blocks, 256, 0, st>>>(
        pa, k, dec_h + (int64_t)k * kH, (int64_t)K * kH, inputs, in_features,
        use_tc ? plan.hi[nxt] : nullptr, use_tc ? plan.lo[nxt] : nullptr, use_tc ? (int64_t)plan.ld : 0,
        use_tc ? plan.f16 : 0);
    if ((rc = after_launch())) return rc;
  }
  return GNNPN_OK;
}

// Differentiable replay, forward half (see pn_train.cu): strict-fp32 FFMA LSTM steps + pointer steps, teacher-forced on
// the picks `idx`, saving the post-activation gates and the cell state of every step for the backward pass.
int gnnpn_pn_train_forward_f32(const float* inputs, const float* packed_enc, const float* packed_dec, const int32_t* idx,
                               const float* latent_win, float alpha, int use_tanh, float C, int64_t n, int L,
                               int in_features, int hidden, int K, int N, float* enc_out, float* gates_e, float* c_e,
                               float* dec_h, float* gates_d, float* c_d, float* win_logits, float* win_probs,
                               int32_t* idx_free, void* stream) {
  GNNPN_REQUIRE(inputs && packed_enc && packed_dec && idx && enc_out && gates_e && c_e && dec_h && gates_d && c_d &&
                    win_logits && win_probs && idx_free, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= 8 && K >= 1 && N >= 1 && N <= kMaxWindow &&
                    (int64_t)K * N == L && n >= 0 && n < (1ll << 31), GNNPN_ESHAPE);
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(dec_h) && aligned16(gates_e) && aligned16(gates_d), GNNPN_EALIGN);
  if (n == 0) return GNNPN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  LstmStepArgs a{};
  a.x = inputs; a.x_inst_ld = (int64_t)L * in_features; a.F = in_features; a.use_x = 1; a.gather = nullptr;
  a.P = packed_enc; a.bias = packed_enc + kOffBias; a.M = (int)n;
  a.h_in_ld = a.h_out_ld = (int64_t)L * kH;
  for (int t = 0; t < L; ++t) {
    a.first = t == 0;
    a.h_in = t == 0 ? nullptr : enc_out + (int64_t)(t - 1) * kH;
    a.h_out = enc_out + (int64_t)t * kH;
    a.x_row = t;
    a.gates_out = gates_e + (size_t)t * n * kG;
    a.c = c_e + (size_t)t * n * kH;                                  // this step's slot of the saved cell states
    a.c_in = t == 0 ? nullptr : c_e + (size_t)(t - 1) * n * kH;
    if ((rc = launch_lstm_step(a, st))) return rc;
  }
  a.P = packed_dec; a.x_row = -1; a.first = 0; a.h_out_ld = (int64_t)K * kH;
  for (int k = 0; k < K; ++k) {
    if (k == 0) {
      a.h_in = enc_out + (int64_t)(L - 1) * kH; a.h_in_ld = (int64_t)L * kH;
      a.use_x = 0; a.bias = packed_dec + kOffStart; a.gather = nullptr;
    } else {
      a.h_in = dec_h + (int64_t)(k - 1) * kH; a.h_in_ld = (int64_t)K * kH;
      a.use_x = 1; a.bias = packed_dec + kOffBias; a.gather = idx + (int64_t)(k - 1) * n;
    }
    a.h_out = dec_h + (int64_t)k * kH;
    a.gates_out = gates_d + (size_t)k * n * kG;
    a.c = c_d + (size_t)k * n * kH;
    a.c_in = k == 0 ? c_e + (size_t)(L - 1) * n * kH : c_d + (size_t)(k - 1) * n * kH;
    if ((rc = launch_lstm_step(a, st))) return rc;
    PointerStepArgs pa;
    pa.enc_out = enc_out; pa.enc_inst_ld = (int64_t)L * kH; pa.latent_win = latent_win; pa.alpha = alpha;
    pa.use_tanh = use_tanh; pa.C = C; pa.n = n; pa.L = L; pa.N = N;
    pa.idx_out = idx_free; pa.win_logits = win_logits; pa.win_probs = win_probs;
    pa.forced = idx; pa.uniform = nullptr;
    pointer_step_dot_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(pa, k, dec_h + (int64_t)k * kH, (int64_t)K * kH,
                                                                     inputs, in_features, nullptr, nullptr, 0, 0);
    if ((rc = after_launch())) return rc;
  }
  return GNNPN_OK;
}

// The attention of the decode loop alone (modelPN.py:213-228 restricted to the windows): for given decoder states
// dec_h [n, K, H] every step's window logits / probabilities / first-max picks.  One launch of the stand-alone pointer
// kernel per step; reads every encoding row exactly once.  Used to re-derive window logits for saved states and as the
// attention-only roofline point of bench.py (no LSTM step in the launches).
int gnnpn_pn_attention_windows_f32(const float* enc_out, const float* dec_h, const float* latent_win, float alpha,
                                   int use_tanh, float C, int64_t n, int L, int hidden, int K, int N, int32_t* idx_out,
                                   float* win_logits, float* win_probs, void* stream) {
  GNNPN_REQUIRE(enc_out && dec_h && idx_out && win_logits && win_probs, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && K >= 1 && N >= 1 && N <= kMaxWindow && (int64_t)K * N == L && n >= 0, GNNPN_ESHAPE);
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(dec_h), GNNPN_EALIGN);
  if (n == 0) return GNNPN_OK;
  PointerStepArgs pa;
  pa.enc_out = enc_out; pa.enc_inst_ld = (int64_t)L * kH; pa.latent_win = latent_win; pa.alpha = alpha;
  pa.use_tanh = use_tanh; pa.C = C; pa.n = n; pa.L = L; pa.N = N;
  pa.idx_out = idx_out; pa.win_logits = win_logits; pa.win_probs = win_probs;
  pa.forced = nullptr; pa.uniform = nullptr;
  for (int k = 0; k < K; ++k) {
    pointer_step_dot_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(
        pa, k, dec_h + (int64_t)k * kH, (int64_t)K * kH, nullptr, 0, nullptr, nullptr, 0, 0);
    const int rc = after_launch();
    if (rc) return rc;
  }
  return GNNPN_OK;
}

int gnnpn_pn_train_forward_tc_f32(const float* inputs, const float* packed_enc, const float* packed_dec,
                                  const int32_t* forced_idx, const float* sample_uniform, const float* latent_win,
                                  float alpha, int use_tanh, float C, int64_t n, int L, int in_features, int hidden, int K,
                                  int N, float* enc_out, float* gates_e, float* c_e, float* dec_h, float* gates_d,
                                  float* c_d, float* win_logits, float* win_probs, int32_t* idx_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  GNNPN_REQUIRE(inputs && packed_enc && packed_dec && enc_out && gates_e && c_e && dec_h && gates_d && c_d && win_logits &&
                    win_probs && idx_out && workspace, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= 8 && K >= 1 && N >= 1 && N <= kMaxWindow &&
                    (int64_t)K * N == L && n >= 0 && n < (1ll << 31), GNNPN_ESHAPE);
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(dec_h) && aligned16(gates_e) && aligned16(gates_d) && aligned16(c_e) &&
                    aligned16(c_d), GNNPN_EALIGN);
  GNNPN_REQUIRE(workspace_bytes >= tc_lstm_workspace_bytes(n), GNNPN_EWORKSPACE);
  // the column-split cluster scan only (training batches: the reference's is 128); larger batches use the FFMA replay
  GNNPN_REQUIRE(tc_colsplit_wanted(n) && tc_colsplit_wanted_encode(n), GNNPN_EUNSUPPORTED);
  if (n == 0) return GNNPN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* scr = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  const size_t nh = (size_t)n * kH;
  float* c_last_e = c_e + (size_t)(L - 1) * nh;          // the encoder's final cell state = its last save
  float* c_last_d = c_d + (size_t)(K - 1) * nh;          // decoder: in (copy of c_last_e) / out (= its last save)
  SeqEncodeArgs ea{inputs, n, L, in_features, packed_enc, enc_out, c_last_e, scr, GNNPN_ENC_ROWMAJOR};
  ea.save_gates = gates_e; ea.save_c = c_e;
  int rc = tc_colsplit_encode(ea, scr, st);
  if (rc) return rc;
  cudaError_t ce = cudaMemcpyAsync(c_last_d, c_last_e, nh * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (ce != cudaSuccess) return (int)ce;
  SeqDecodeArgs da{inputs, enc_out, c_last_d, latent_win, alpha, packed_dec, use_tanh, C, n, L, in_features, K, N, dec_h,
                   idx_out, win_logits, win_probs, forced_idx, sample_uniform, scr, GNNPN_ENC_ROWMAJOR};
  da.save_gates = gates_d; da.save_c = c_d;
  return tc_colsplit_decode(da, scr, st);
}

int gnnpn_pn_full_logits_f32(const float* enc_out, const float* dec_h, const int32_t* idx,
                             int attention, const float* att_params, int use_tanh, float C,
                             int64_t n, int L, int hidden, int K, float* logits_full, void* stream) {
  GNNPN_REQUIRE(enc_out && dec_h && idx && logits_full, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && K >= 1 && L >= 1, GNNPN_ESHAPE);
  GNNPN_REQUIRE(attention == GNNPN_ATT_DOT, GNNPN_EUNSUPPORTED);
  (void)att_params;
  const size_t smem = (size_t)K * kH * sizeof(float);
  GNNPN_REQUIRE(smem <= 200 * 1024, GNNPN_ESHAPE);
  GNNPN_REQUIRE(n < 65536ll * 32768ll, GNNPN_ERANGE);
  if (n == 0) return GNNPN_OK;
  cudaError_t e = cudaFuncSetAttribute(full_logits_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)n, (unsigned)ceil_div(L, 32));
  full_logits_dot_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(enc_out, dec_h, idx, use_tanh, C, n, L, K,
                                                                  logits_full);
  return after_launch();
}

int gnnpn_pn_reward_f32(const float* inputs, const int32_t* idx, int64_t n, int L, int in_features,
                        int K, int tag, int32_t* viol_out, float* obj_out, float* reward_high_out,
                        void* stream) {
  GNNPN_REQUIRE(inputs && idx, GNNPN_ENULL);
  GNNPN_REQUIRE(K >= 1 && K <= kMaxTasks && (tag == 0 || tag == 1) && in_features >= tag + 8, GNNPN_ESHAPE);
  if (n == 0) return GNNPN_OK;
  reward_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      inputs, idx, n, L, in_features, K, tag, viol_out, obj_out, reward_high_out);
  return after_launch();
}

}  // extern "C"
