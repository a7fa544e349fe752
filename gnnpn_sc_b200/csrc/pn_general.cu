// General pointer-network decode: every variant of PointerNet.forward's decode loop (modelPN.py:204-239) that the
// fused fast paths (pn.cu / tc_seq.cu: Dot attention, no glimpses, window <= 32, <= 8 raw columns) do not cover:
//   * attention "Bahdanau" (modelPN.py:83-91,103-109):  u_l = V . tanh(W_query q + W_ref ref_l),
//     with E = W_ref ref + b_ref computed ONCE per forward (the reference recomputes the 1x1 conv every step);
//   * n_glimpses > 0 (modelPN.py:208-211): q <- sum_l softmax_l(glimpse logits, visited -> -inf) * ref'_l over ALL L
//     positions (ref' = enc_out for Dot, the glimpse's own W_ref conv output for Bahdanau);
//   * any window width N (scale-up: N = 1000) and up to 32 raw input columns (embedding_size = 20).
// One launch per piece and per step (LSTM cell, [query GEMM, glimpse] x n_glimpses, query GEMM, pointer); one CTA
// per composition instance in the attention kernels.  For Dot / no glimpse / N <= 32 the pointer arithmetic is the
// same as pointer_step_warp's (canonical dot, fmaf latent, sequential softmax sum) -> bit-identical picks / logits.
#include <math.h>
#include <cuda_fp16.h>
#include "lstm_step.cuh"
#include "tc_lstm.cuh"
#include "pointer.cuh"

namespace gnnpn {
int launch_gemm_ffma(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                     const float* scale, const float* shift, int act, float* C, int64_t ldc, int64_t M, int N,
                     int K, cudaStream_t st);
namespace {

constexpr int kAttThreads = 256, kAttWarps = kAttThreads / 32;
constexpr size_t kAttBlockFloats = 2 * (size_t)kH * kH + 3 * kH;   // [W_query | b_query | W_ref | b_ref | V]
constexpr size_t kOffWq = 0, kOffBq = (size_t)kH * kH, kOffWr = kOffBq + kH, kOffBr = kOffWr + (size_t)kH * kH,
                 kOffV = kOffBr + kH;

// sum over the lane's 8 elements of V * tanh(qw + r): explicit fma chain (fixed rounding order)
__device__ __forceinline__ float bahdanau8(const float4 r0, const float4 r1, const float4 w0, const float4 w1,
                                           const float4 v0, const float4 v1) {
  float s = v0.x * tanhf(w0.x + r0.x);
  s = fmaf(v0.y, tanhf(w0.y + r0.y), s); s = fmaf(v0.z, tanhf(w0.z + r0.z), s); s = fmaf(v0.w, tanhf(w0.w + r0.w), s);
  s = fmaf(v1.x, tanhf(w1.x + r1.x), s); s = fmaf(v1.y, tanhf(w1.y + r1.y), s); s = fmaf(v1.z, tanhf(w1.z + r1.z), s);
  s = fmaf(v1.w, tanhf(w1.w + r1.w), s);
  return s;
}

struct GeneralArgs {
  const float* rows;         // [n, L, kH]  enc_out (Dot) or E = W_ref enc_out + b_ref (Bahdanau)
  const float* q;            // [n] rows q_ld apart: the query (Dot) or W_query q + b_query (Bahdanau)
  int64_t q_ld;
  const float* V;            // [kH] (Bahdanau)
  const int32_t* fed;        // [K, n] picks the visited mask follows (forced picks when teacher forcing) -- steps < k
  int64_t n;
  int L, k;
};

// ---------------------------------------------------------------------------------------------------------
// glimpse (modelPN.py:208-211): one CTA per instance, online softmax over all L rows, 8 warps stride the rows
// ---------------------------------------------------------------------------------------------------------
template <bool BAHD>
__global__ void __launch_bounds__(kAttThreads) glimpse_step_kernel(const GeneralArgs a, float* __restrict__ q_out,
                                                                   int64_t q_out_ld) {
  extern __shared__ uint32_t smem_u[];
  uint32_t* visited = smem_u;                              // bitmap of L bits
  const int words = (a.L + 31) / 32;
  const int words4 = (words + 3) & ~3;                     // keep the float4 accesses below 16-byte aligned
  float* red_ms = reinterpret_cast<float*>(smem_u + words4);           // [8 warps][2] running max, sum
  float* red_acc = red_ms + 4 * kAttWarps;                               // [8 warps][kH] weighted row sums
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < words; i += kAttThreads) visited[i] = 0u;
  __syncthreads();
  for (int j = threadIdx.x; j < a.k; j += kAttThreads) {
    const int pos = a.fed[(int64_t)j * a.n + b];
    if (pos >= 0 && pos < a.L) atomicOr(&visited[pos >> 5], 1u << (pos & 31));
  }
  __syncthreads();
  const float4* qp = reinterpret_cast<const float4*>(a.q + b * a.q_ld);
  const float4 q0 = qp[lane], q1 = qp[32 + lane];
  float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
  if (BAHD) { v0 = __ldg(reinterpret_cast<const float4*>(a.V) + lane); v1 = __ldg(reinterpret_cast<const float4*>(a.V) + 32 + lane); }
  float m = -INFINITY, s = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float* base = a.rows + b * (int64_t)a.L * kH;
  for (int l = warp; l < a.L; l += kAttWarps) {
    if (visited[l >> 5] & (1u << (l & 31))) continue;      // exp(-inf) = 0: contributes nothing
    const float4* rp = reinterpret_cast<const float4*>(base + (int64_t)l * kH);
    const float4 r0 = __ldg(rp + lane), r1 = __ldg(rp + 32 + lane);
    const float u = warp_sum(BAHD ? bahdanau8(r0, r1, q0, q1, v0, v1) : dot8(r0, r1, q0, q1));   // glimpse: no C*tanh
    const float nm = fmaxf(m, u);
    const float sc = expf(m - nm), p = expf(u - nm);       // m = -inf on the first row: sc = 0
    s = fmaf(s, sc, p);
    acc[0] = fmaf(acc[0], sc, p * r0.x); acc[1] = fmaf(acc[1], sc, p * r0.y);
    acc[2] = fmaf(acc[2], sc, p * r0.z); acc[3] = fmaf(acc[3], sc, p * r0.w);
    acc[4] = fmaf(acc[4], sc, p * r1.x); acc[5] = fmaf(acc[5], sc, p * r1.y);
    acc[6] = fmaf(acc[6], sc, p * r1.z); acc[7] = fmaf(acc[7], sc, p * r1.w);
    m = nm;
  }
  if (lane == 0) { red_ms[2 * warp] = m; red_ms[2 * warp + 1] = s; }
  // lane owns hidden elements 4*lane..+3 and 128+4*lane..+3 (float4 index lane and 32+lane)
  reinterpret_cast<float4*>(red_acc + warp * kH)[lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  reinterpret_cast<float4*>(red_acc + warp * kH)[32 + lane] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  float M = -INFINITY;
  for (int w = 0; w < kAttWarps; ++w) M = fmaxf(M, red_ms[2 * w]);
  float S = 0.f;
  for (int w = 0; w < kAttWarps; ++w) {
    const float mw = red_ms[2 * w];
    if (mw > -INFINITY) S += red_ms[2 * w + 1] * expf(mw - M);
  }
  for (int h = threadIdx.x; h < kH; h += kAttThreads) {
    float o = 0.f;
    for (int w = 0; w < kAttWarps; ++w) {
      const float mw = red_ms[2 * w];
      if (mw > -INFINITY) o = fmaf(red_acc[w * kH + h], expf(mw - M), o);
    }
    q_out[b * q_out_ld + h] = o / S;
  }
}

// ---------------------------------------------------------------------------------------------------------
// pointer on window k, any N: one CTA per instance
// ---------------------------------------------------------------------------------------------------------
struct PointerGeneralArgs {
  GeneralArgs g;
  const float* latent_win;   // [n, L] or nullptr
  float alpha;
  int use_tanh;
  float C;
  int N;
  int32_t* idx_out;          // [K, n]
  float* win_logits;         // [n, L]
  float* win_probs;          // [n, L]
  const int32_t* forced;     // [K, n] or nullptr
  const float* uniform;      // [K, n] or nullptr
  const float* inputs;       // [n, L, F]
  int F;
  void* a_hi_next;           // tensor-core LSTM path: A operand of the next step (nullptr: FFMA path gathers by idx)
  void* a_lo_next;
  int64_t a_ld;
  int a_f16;
};

template <bool BAHD>
__global__ void __launch_bounds__(kAttThreads) pointer_general_kernel(const PointerGeneralArgs p) {
  extern __shared__ float smem_f[];
  const int N = p.N, k = p.g.k;
  float* sw = smem_f;                  // [N] masked working logits
  float* se = smem_f + N;              // [N] exp(w - max)
  __shared__ float red_v[kAttWarps];
  __shared__ int red_i[kAttWarps];
  __shared__ float s_bcast[2];
  __shared__ int s_pick;
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Dot: canonical pointer dot (pointer.cuh), lane owns floats [8*lane, 8*lane+8); Bahdanau: float4 lane / 32+lane
  const float4* qp = reinterpret_cast<const float4*>(p.g.q + b * p.g.q_ld);
  const int i0 = BAHD ? lane : 2 * lane, i1 = BAHD ? 32 + lane : 2 * lane + 1;
  const float4 q0 = qp[i0], q1 = qp[i1];
  float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
  if (BAHD) { v0 = __ldg(reinterpret_cast<const float4*>(p.g.V) + lane); v1 = __ldg(reinterpret_cast<const float4*>(p.g.V) + 32 + lane); }
  const float* base = p.g.rows + (b * (int64_t)p.g.L + (int64_t)k * N) * kH;
  const int64_t wbase = b * (int64_t)p.g.L + (int64_t)k * N;
  for (int j = warp; j < N; j += kAttWarps) {
    const float4* rp = reinterpret_cast<const float4*>(base + (int64_t)j * kH);
    const float4 r0 = ldg_stream(rp + i0), r1 = ldg_stream(rp + i1);
    const float d = BAHD ? warp_sum(bahdanau8(r0, r1, q0, q1, v0, v1)) : dot_reduce(dot8(r0, r1, q0, q1), lane);
    if (lane == 0) {
      const float l = p.use_tanh ? p.C * tanhf(d) : d;
      p.win_logits[wbase + j] = l;
      sw[j] = p.latent_win ? fmaf(p.alpha, __ldg(p.latent_win + wbase + j), l) : l;
    }
  }
  __syncthreads();
  // max (exact, order-free)
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += kAttThreads) mx = fmaxf(mx, sw[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red_v[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m2 = red_v[0];
    for (int w = 1; w < kAttWarps; ++w) m2 = fmaxf(m2, red_v[w]);
    s_bcast[0] = m2;
  }
  __syncthreads();
  mx = s_bcast[0];
  for (int j = threadIdx.x; j < N; j += kAttThreads) se[j] = expf(sw[j] - mx);
  __syncthreads();
  if (threadIdx.x == 0) {                    // sequential sum in candidate order (same order as the fused kernels)
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += se[j];
    s_bcast[1] = s;
  }
  __syncthreads();
  const float s = s_bcast[1];
  // probabilities + first maximal probability (torch.max tie rule, modelPN.py:225-226)
  float best = -1.f;
  int best_j = 0x7fffffff;
  for (int j = threadIdx.x; j < N; j += kAttThreads) {
    const float pj = se[j] / s;
    p.win_probs[wbase + j] = pj;
    se[j] = pj;
    if (pj > best) { best = pj; best_j = j; }            // ascending j per thread: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
  }
  if (lane == 0) { red_v[warp] = best; red_i[warp] = best_j; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = red_v[0];
    int bj = red_i[0];
    for (int w = 1; w < kAttWarps; ++w)
      if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bj)) { bv = red_v[w]; bj = red_i[w]; }
    if (p.uniform) {                       // sample="sample": inverse-CDF draw (see pointer_finish_warp)
      const float u = __ldg(p.uniform + (int64_t)k * p.g.n + b);
      float cum = 0.f;
      int pick = -1, last_pos = 0;
      for (int j = 0; j < N; ++j) {
        const float pj = se[j];
        cum += pj;
        if (pj > 0.f) last_pos = j;
        if (pick < 0 && u < cum) pick = j;
      }
      bj = pick < 0 ? last_pos : pick;
    }
    p.idx_out[(int64_t)k * p.g.n + b] = k * N + bj;
    s_pick = p.forced ? p.forced[(int64_t)k * p.g.n + b] : k * N + bj;
  }
  __syncthreads();
  if (p.a_hi_next && threadIdx.x < p.F) {
    const int f = threadIdx.x;
    const float v = __ldg(p.inputs + (b * p.g.L + s_pick) * (int64_t)p.F + f);
    if (p.a_f16) {
      const __half hi = __float2half_rn(v);
      reinterpret_cast<__half*>(p.a_hi_next)[b * p.a_ld + kH + f] = hi;
      reinterpret_cast<__half*>(p.a_lo_next)[b * p.a_ld + kH + f] = __float2half_rn(v - __half2float(hi));
    } else {
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
      reinterpret_cast<float*>(p.a_hi_next)[b * p.a_ld + kH + f] = __uint_as_float(hb);
      reinterpret_cast<float*>(p.a_lo_next)[b * p.a_ld + kH + f] = v - __uint_as_float(hb);
    }
  }
}

// dense logits for Bahdanau: out[k, b, l] = C*tanh( V . tanh(qw[b,k,:] + E[b,l,:]) ), -inf at the picks of steps < k
__global__ void __launch_bounds__(256) full_logits_bahdanau_kernel(
    const float* __restrict__ E, const float* __restrict__ qw, const float* __restrict__ V,
    const int32_t* __restrict__ idx, int use_tanh, float C, int64_t n, int L, int K, float* __restrict__ out) {
  extern __shared__ float sq[];                // [K][kH]
  const int64_t b = blockIdx.x;
  const int l0 = blockIdx.y * 32;
  const float4* qsrc = reinterpret_cast<const float4*>(qw + b * (int64_t)K * kH);
  for (int i = threadIdx.x; i < K * kH / 4; i += blockDim.x) reinterpret_cast<float4*>(sq)[i] = __ldg(qsrc + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 v0 = __ldg(reinterpret_cast<const float4*>(V) + lane), v1 = __ldg(reinterpret_cast<const float4*>(V) + 32 + lane);
  for (int l = l0 + warp; l < min(L, l0 + 32); l += 8) {
    const float4* rp = reinterpret_cast<const float4*>(E + (b * L + l) * (int64_t)kH);
    const float4 r0 = __ldg(rp + lane), r1 = __ldg(rp + 32 + lane);
    for (int k = 0; k < K; ++k) {
      const float4 a0 = reinterpret_cast<const float4*>(sq + k * kH)[lane];
      const float4 a1 = reinterpret_cast<const float4*>(sq + k * kH)[32 + lane];
      const float d = warp_sum(bahdanau8(r0, r1, a0, a1, v0, v1));
      if (lane == 0) out[((int64_t)k * n + b) * L + l] = use_tanh ? C * tanhf(d) : d;
    }
  }
  __syncthreads();
  for (int pair = threadIdx.x; pair < K * K; pair += blockDim.x) {
    const int k = pair / K, jprev = pair % K;
    if (jprev >= k) continue;
    const int pos = idx[(int64_t)jprev * n + b];
    if (pos >= l0 && pos < l0 + 32 && pos < L) out[((int64_t)k * n + b) * L + pos] = -INFINITY;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct GeneralWs {
  size_t tc, e_p, e_g, qw, qw_all, total;
};
GeneralWs general_ws(int64_t n, int L, int K, int attention, int n_glimpses, int use_tc) {
  GeneralWs w{};
  size_t off = 0;
  w.tc = off; off += use_tc ? align_up(tc_lstm_workspace_bytes(n), 1024) : 0;
  const size_t big = align_up((size_t)n * L * kH * sizeof(float), 1024);
  w.e_p = off; off += attention == GNNPN_ATT_BAHDANAU ? big : 0;
  w.e_g = off; off += (attention == GNNPN_ATT_BAHDANAU && n_glimpses > 0) ? big : 0;
  w.qw = off; off += align_up((size_t)n * kH * sizeof(float), 1024);
  w.qw_all = off; off += attention == GNNPN_ATT_BAHDANAU ? align_up((size_t)n * K * kH * sizeof(float), 1024) : 0;
  w.total = off + 1024;
  return w;
}

}  // namespace
}  // namespace gnnpn

using namespace gnnpn;

extern "C" {

size_t gnnpn_pn_att_block_floats(int hidden) { return hidden == kH ? kAttBlockFloats : 0; }

size_t gnnpn_pn_decode_general_workspace_bytes(int64_t n, int L, int K, int hidden, int attention, int n_glimpses,
                                               int use_tc) {
  if (hidden != kH || n < 0 || L < 1 || K < 1) return 0;
  return general_ws(n, L, K, attention, n_glimpses, use_tc).total;
}

int gnnpn_pn_decode_general_f32(const float* inputs, const float* enc_out, float* c_state, const float* latent_win,
                                float alpha, const float* packed, int attention, const float* att_params,
                                int n_glimpses, int use_tanh, float C, int64_t n, int L, int in_features, int hidden,
                                int K, int N, float* dec_h, float* dec_q, float* qw_pointer, int32_t* idx_out,
                                float* win_logits, float* win_probs, const int32_t* forced_idx,
                                const float* sample_uniform, int use_tc, void* workspace, size_t workspace_bytes,
                                void* stream) {
  GNNPN_REQUIRE(inputs && enc_out && c_state && packed && dec_h && dec_q && idx_out && win_logits && win_probs &&
                    workspace, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && in_features >= 1 && in_features <= kXPad, GNNPN_ESHAPE);
  GNNPN_REQUIRE(K >= 1 && N >= 1 && (int64_t)K * N == L && n_glimpses >= 0, GNNPN_ESHAPE);
  GNNPN_REQUIRE((size_t)N * 2 * sizeof(float) <= 160 * 1024 && (size_t)L / 8 <= 160 * 1024, GNNPN_ESHAPE);
  GNNPN_REQUIRE(n >= 0 && n < (1ll << 31), GNNPN_ERANGE);
  GNNPN_REQUIRE(attention == GNNPN_ATT_DOT || attention == GNNPN_ATT_BAHDANAU, GNNPN_EUNSUPPORTED);
  const bool bahd = attention == GNNPN_ATT_BAHDANAU;
  GNNPN_REQUIRE(!bahd || (att_params && qw_pointer), GNNPN_ENULL);
  GNNPN_REQUIRE(n_glimpses == 0 || dec_q != dec_h, GNNPN_ESHAPE);
  GNNPN_REQUIRE(aligned16(enc_out) && aligned16(dec_h) && aligned16(dec_q) && aligned16(c_state) && aligned16(packed) &&
                    (!att_params || aligned16(att_params)), GNNPN_EALIGN);
  if (n == 0) return GNNPN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const GeneralWs wsl = general_ws(n, L, K, attention, n_glimpses, use_tc);
  GNNPN_REQUIRE(workspace_bytes >= wsl.total, GNNPN_EWORKSPACE);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* e_p = reinterpret_cast<float*>(ws + wsl.e_p);
  float* e_g = reinterpret_cast<float*>(ws + wsl.e_g);
  float* qw = reinterpret_cast<float*>(ws + wsl.qw);
  const float* att_p = att_params;                                   // pointer block
  const float* att_g = att_params ? att_params + kAttBlockFloats : nullptr;   // glimpse block (follows)
  int rc;
  if (bahd) {
    // E = W_ref . enc_out + b_ref for all positions, once (Conv1d(H,H,1) == a row-wise Linear, modelPN.py:87,105)
    if ((rc = launch_gemm_ffma(enc_out, kH, att_p + kOffWr, kH, att_p + kOffBr, nullptr, nullptr, GNNPN_ACT_NONE, e_p,
                               kH, n * (int64_t)L, kH, kH, st))) return rc;
    if (n_glimpses > 0 &&
        (rc = launch_gemm_ffma(enc_out, kH, att_g + kOffWr, kH, att_g + kOffBr, nullptr, nullptr, GNNPN_ACT_NONE, e_g,
                               kH, n * (int64_t)L, kH, kH, st))) return rc;
  }
  const float* bias = packed + kOffBias;
  const float* start = packed + kOffStart;
  TcLstmPlan plan;
  TcLstmStep ts{};
  LstmStepArgs a{};
  if (use_tc) {
    if ((rc = tc_lstm_plan(&plan, ws + wsl.tc, tc_lstm_workspace_bytes(n), n, packed))) return rc;
    if ((rc = tc_lstm_load_h(plan, 0, enc_out + (int64_t)(L - 1) * kH, (int64_t)L * kH, st))) return rc;
    if ((rc = tc_lstm_zero(plan, 1, st))) return rc;
    ts.c = c_state; ts.h_out_ld = (int64_t)K * kH; ts.x_next = nullptr; ts.x_row_next = -1;
    ts.F = in_features; ts.first = 0;
  } else {
    a.x = inputs; a.x_inst_ld = (int64_t)L * in_features; a.F = in_features; a.x_row = -1;
    a.P = packed; a.c = c_state; a.M = (int)n; a.first = 0;
    a.h_out_ld = (int64_t)K * kH;
  }
  const int32_t* fed_all = forced_idx ? forced_idx : idx_out;
  const size_t glimpse_smem = (size_t)((((L + 31) / 32) + 3) & ~3) * 4 + (size_t)kAttWarps * (4 + kH) * 4;
  const size_t ptr_smem = (size_t)N * 2 * sizeof(float);
  if (glimpse_smem > 48 * 1024) {
    cudaFuncSetAttribute(glimpse_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)glimpse_smem);
    cudaFuncSetAttribute(glimpse_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)glimpse_smem);
  }
  if (ptr_smem > 48 * 1024) {
    cudaFuncSetAttribute(pointer_general_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ptr_smem);
    cudaFuncSetAttribute(pointer_general_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ptr_smem);
  }
  for (int k = 0; k < K; ++k) {
    // ---- decoder LSTM cell (modelPN.py:205)
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
        a.use_x = 1; a.bias = bias; a.gather = fed_all + (int64_t)(k - 1) * n;
      }
      a.h_out = dec_h + (int64_t)k * kH;
      if ((rc = launch_lstm_step(a, st))) return rc;
    }
    // ---- glimpses (modelPN.py:208-211)
    const float* q_cur = dec_h + (int64_t)k * kH;
    for (int g = 0; g < n_glimpses; ++g) {
      GeneralArgs ga{};
      ga.fed = fed_all; ga.n = n; ga.L = L; ga.k = k;
      if (bahd) {
        if ((rc = launch_gemm_ffma(q_cur, (int64_t)K * kH, att_g + kOffWq, kH, att_g + kOffBq, nullptr, nullptr,
                                   GNNPN_ACT_NONE, qw, kH, n, kH, kH, st))) return rc;
        ga.rows = e_g; ga.q = qw; ga.q_ld = kH; ga.V = att_g + kOffV;
        glimpse_step_kernel<true><<<(unsigned)n, kAttThreads, glimpse_smem, st>>>(ga, dec_q + (int64_t)k * kH,
                                                                                  (int64_t)K * kH);
      } else {
        ga.rows = enc_out; ga.q = q_cur; ga.q_ld = (int64_t)K * kH; ga.V = nullptr;
        glimpse_step_kernel<false><<<(unsigned)n, kAttThreads, glimpse_smem, st>>>(ga, dec_q + (int64_t)k * kH,
                                                                                   (int64_t)K * kH);
      }
      if ((rc = after_launch())) return rc;
      q_cur = dec_q + (int64_t)k * kH;
    }
    if (n_glimpses == 0 && dec_q != dec_h) {
      if (cudaMemcpy2DAsync(dec_q + (int64_t)k * kH, (size_t)K * kH * 4, dec_h + (int64_t)k * kH, (size_t)K * kH * 4,
                            (size_t)kH * 4, (size_t)n, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return (int)cudaGetLastError();
    }
    // ---- pointer on window k (modelPN.py:213-228)
    PointerGeneralArgs pa{};
    pa.g.fed = fed_all; pa.g.n = n; pa.g.L = L; pa.g.k = k;
    pa.latent_win = latent_win; pa.alpha = alpha; pa.use_tanh = use_tanh; pa.C = C; pa.N = N;
    pa.idx_out = idx_out; pa.win_logits = win_logits; pa.win_probs = win_probs;
    pa.forced = forced_idx; pa.uniform = sample_uniform; pa.inputs = inputs; pa.F = in_features;
    const int nxt = (k + 1) & 1;
    pa.a_hi_next = use_tc ? plan.hi[nxt] : nullptr; pa.a_lo_next = use_tc ? plan.lo[nxt] : nullptr;
    pa.a_ld = use_tc ? plan.ld : 0; pa.a_f16 = use_tc ? plan.f16 : 0;
    if (bahd) {
      float* qwp = qw_pointer + (int64_t)k * kH;                 // [n, K, kH]: kept for the dense-logits kernel
      if ((rc = launch_gemm_ffma(q_cur, (int64_t)K * kH, att_p + kOffWq, kH, att_p + kOffBq, nullptr, nullptr,
                                 GNNPN_ACT_NONE, qwp, (int64_t)K * kH, n, kH, kH, st))) return rc;
      pa.g.rows = e_p; pa.g.q = qwp; pa.g.q_ld = (int64_t)K * kH; pa.g.V = att_p + kOffV;
      pointer_general_kernel<true><<<(unsigned)n, kAttThreads, ptr_smem, st>>>(pa);
    } else {
      pa.g.rows = enc_out; pa.g.q = q_cur; pa.g.q_ld = (int64_t)K * kH; pa.g.V = nullptr;
      pointer_general_kernel<false><<<(unsigned)n, kAttThreads, ptr_smem, st>>>(pa);
    }
    if ((rc = after_launch())) return rc;
  }
  return GNNPN_OK;
}

int gnnpn_pn_ref_transform_f32(const float* enc_out, const float* att_block, int64_t rows, int hidden, float* E,
                               void* stream) {
  GNNPN_REQUIRE(enc_out && att_block && E, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && rows >= 0, GNNPN_ESHAPE);
  if (rows == 0) return GNNPN_OK;
  return launch_gemm_ffma(enc_out, kH, att_block + kOffWr, kH, att_block + kOffBr, nullptr, nullptr, GNNPN_ACT_NONE, E,
                          kH, rows, kH, kH, (cudaStream_t)stream);
}

int gnnpn_pn_query_transform_f32(const float* q, int64_t q_ld, const float* att_block, int64_t rows, int hidden,
                                 float* qw, int64_t qw_ld, void* stream) {
  GNNPN_REQUIRE(q && att_block && qw, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && rows >= 0 && q_ld >= kH && qw_ld >= kH, GNNPN_ESHAPE);
  if (rows == 0) return GNNPN_OK;
  return launch_gemm_ffma(q, q_ld, att_block + kOffWq, kH, att_block + kOffBq, nullptr, nullptr, GNNPN_ACT_NONE, qw,
                          qw_ld, rows, kH, kH, (cudaStream_t)stream);
}

int gnnpn_pn_full_logits_bahdanau_f32(const float* E, const float* qw, const float* att_block, const int32_t* idx,
                                      int use_tanh, float C, int64_t n, int L, int hidden, int K, float* logits_full,
                                      void* stream) {
  GNNPN_REQUIRE(E && qw && att_block && idx && logits_full, GNNPN_ENULL);
  GNNPN_REQUIRE(hidden == kH && K >= 1 && L >= 1, GNNPN_ESHAPE);
  const size_t smem = (size_t)K * kH * sizeof(float);
  GNNPN_REQUIRE(smem <= 200 * 1024, GNNPN_ESHAPE);
  GNNPN_REQUIRE(n < 65536ll * 32768ll, GNNPN_ERANGE);
  if (n == 0) return GNNPN_OK;
  cudaError_t e = cudaFuncSetAttribute(full_logits_bahdanau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)n, (unsigned)ceil_div(L, 32));
  full_logits_bahdanau_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(E, qw, att_block + kOffV, idx, use_tanh, C, n, L,
                                                                         K, logits_full);
  return after_launch();
}

}  // extern "C"
