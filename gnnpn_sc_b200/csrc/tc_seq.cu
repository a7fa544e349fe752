// Persistent tcgen05 LSTM scan: one launch = one whole recurrence (encoder: L steps; decoder: K steps with
// the pointer step fused in between).  Reference semantics: nn.LSTM cell, gate order i,f,g,o
// (modelPN.py:157-158,191,205) and the decode loop modelPN.py:204-239.
//
// A CTA owns 128 composition instances for the whole scan, so the step-to-step dependency is CTA-local:
//   * A operand  [h | x]  lives in SHARED MEMORY as fp16 hi/lo pairs (error-compensated 3xFP16, fp32 accumulate:
//     a.w ~= a_lo.w_hi + a_hi.w_hi + a_hi.w_lo) in the 128B-swizzled K-major UMMA layout; the epilogue of step t
//     writes h'(t) there directly -- h never goes through global memory between steps;
//   * B operand (the 4H x (H+F) folded weights, fp16 hi/lo, 1.1 MB, L2-resident) is streamed by TMA through a
//     ring of 16 KB slots, one (N tile, k block, hi|lo) box per slot;
//   * accumulators: two 128-column TMEM buffers (N tile = 128 gate columns = 32 hidden units), so the
//     epilogue of tile j overlaps the MMAs of tile j+1;
//   * h'(t) of tiles 0..6 cannot be written over A while the MMAs of step t still read it: it is staged in the
//     other 256 TMEM columns (tcgen05.st) and copied to shared memory once the last MMA of the step has retired.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = x producer (encoder inputs),
// 4..19 = epilogue (TMEM lane quarter = warp & 3, column group = (warp - 4) / 4).
//
// Encodings layout (template parameter NR, SeqParams::enc_layout):
//   NR == 0  row-major enc_out [n, L, H] (GNNPN_ENC_ROWMAJOR).  Encoder: h' tiles leave by TMA store.  Decoder: a
//            separate pointer phase after each step's cell epilogue (each epilogue warp owns 8 instances, window rows
//            streamed per instance, pointer.cuh::pointer_steps_batched) -- the step is MMA phase + pointer phase.
//   NR >= 1  blocked encodings (GNNPN_ENC_BLOCKED128): per block of 128 instances
//            [L][8 tiles][4 groups][128 instances][8 floats], unit = 32*tile + 8*group + i -- exactly the
//            (thread = instance, 8 units per tile) ownership of the epilogue.  The encoder writes it with coalesced
//            256-bit stores.  In the decoder (NR = row capacity of the window, N <= NR) the pointer dot products are
//            FUSED INTO THE CELL EPILOGUE: when thread (instance r, group g) has produced the 8 h' units of tile nt
//            it multiplies them into the matching 8 floats of each of the N window rows of ITS instance (a warp's
//            load is 1 KB contiguous; the otherwise idle warp 3 bulk-prefetches every 16 KB (row, tile) slice into L2
//            two tiles ahead), in pointer.cuh's canonical order.  The window rows stream from HBM UNDER the step's MMAs; after the last tile the four group
//            partials are combined through shared memory and one thread per instance finishes the step (C*tanh,
//            latent, softmax, first-max pick / draw, next input row) while the tensor cores already run the h parts
//            of the next step's first two tiles.  No separate pointer phase remains.
// Cross-CTA mbarrier arrives use .release.cta semantics on purpose (see mbar_arrive_cluster).
#include <stdio.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "tc_common.cuh"
#include "common.cuh"
#include "lstm_step.cuh"
#include "tc_lstm.cuh"
#include "tc_seq.cuh"
#include "pointer.cuh"
#include "tc_seq_dev.cuh"
#include "options.cuh"

namespace gnnpn {
namespace tc {
int make_map_2d(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f16);
}
namespace seq {

using namespace tc;

constexpr int BM = 128;                    // instances per CTA = UMMA M = TMEM lanes
constexpr int TILE_N = 128;                // gate columns per accumulator tile
constexpr int N_TILES = kG / TILE_N;       // 8
constexpr int KB_H = kH / 64;              // 4 k-blocks of 64 halfs (128-byte swizzle rows)
constexpr int BLK_BYTES = 128 * 128;       // one [128 rows x 64 halfs] block
constexpr int RING_BYTES = 4 * BLK_BYTES;  // weight ring: 4 slots of 16 KB (1 CTA) / 8 slots of 8 KB (CTA pair)
constexpr int MAX_RING = 8;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;             // [0,256): 2 accumulator buffers; [256,512): h' staging
constexpr int STAGE_COL0 = 256;

constexpr uint32_t OFF_A_HI = 0;
constexpr uint32_t OFF_A_LO = OFF_A_HI + KB_H * BLK_BYTES;
constexpr uint32_t OFF_AX_HI = OFF_A_LO + KB_H * BLK_BYTES;
constexpr uint32_t OFF_AX_LO = OFF_AX_HI + BM * XROW_BYTES;
constexpr uint32_t OFF_RING = OFF_AX_LO + BM * XROW_BYTES;
constexpr uint32_t OFF_HBUF = OFF_RING + RING_BYTES;      // [128 rows x 32 fp32] h' tile for the TMA store
constexpr uint32_t OFF_BIAS = OFF_HBUF + BLK_BYTES;              // two transformed bias sets
constexpr uint32_t OFF_BAR = OFF_BIAS + 2 * kG * 4;
constexpr uint32_t SMEM_USED = OFF_BAR + 256;
constexpr int SMEM_BYTES = SMEM_USED + 1024;                     // + alignment slack
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");


struct SeqParams {
  int64_t n;
  int steps;
  int L, F;
  const float* inputs; int64_t x_inst_ld;
  const float* bias0;      // [kG] gate-interleaved bias of step 0 (decoder: start-token bias)
  const float* bias;       // [kG] bias of steps >= 1
  float* c;                // [n, kH]
  int c_zero_init;
  const float* h0; int64_t h0_ld;          // initial hidden rows or nullptr (zeros)
  float* h_out; int64_t h_out_inst_ld;     // step t of instance m at h_out + m*ld + t*kH
  PointerStepArgs pa;      // decoder only
  int enc_layout;          // GNNPN_ENC_ROWMAJOR / GNNPN_ENC_BLOCKED128 of h_out (encoder) or pa.enc_out (decoder)
  unsigned long long* prof; // debug (GNNPN_SEQ_PROF): per-CTA wait-cycle counters, 16 per CTA, or nullptr
  float* c_scr;            // blocked cell-state scratch, 128*kH floats per CTA (coalesced 128-bit accesses)
};

// mbar_wait that adds the cycles spent waiting to *acc (debug counters; acc lives in a register)
template <bool CLUSTER = false>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, long long& acc) {
  const long long t0 = prof ? clock64() : 0;
  if (CLUSTER) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
  if (prof) acc += clock64() - t0;
}


// x block of the next decoder step: row rr, feature f -> halfs f and 8+f of the hi and lo 32-byte rows
__device__ __forceinline__ void store_ax(uint8_t* sgen, int rr, int f, float xv) {
  const __half hi = __float2half_rn(xv);
  const __half lo = __float2half_rn(xv - __half2float(hi));
  __half* ax_hi = reinterpret_cast<__half*>(sgen + OFF_AX_HI + rr * XROW_BYTES);
  __half* ax_lo = reinterpret_cast<__half*>(sgen + OFF_AX_LO + rr * XROW_BYTES);
  ax_hi[f] = hi; ax_hi[8 + f] = hi;
  ax_lo[f] = lo; ax_lo[8 + f] = lo;
}

// Pointer phase of decode step t for the 8 instances of one epilogue warp (rows rr0..rr0+7 of the CTA); see
// pointer_steps_batched in pointer.cuh.
__device__ __forceinline__ void pointer_phase(const SeqParams& p, int t, int rr0, int64_t m0, int lane, uint8_t* sgen
                                              ) {
  const PointerStepArgs& pa = p.pa;                 // stays in constant memory (p is a __grid_constant__ parameter)
  const int64_t b0 = m0 + rr0;
  const int64_t left = p.n - b0;
  const int count = left <= 0 ? 0 : (left < BM / EPI_WARPS ? (int)left : BM / EPI_WARPS);
  const bool feed_next = t + 1 < p.steps;
  const float* q_base = p.h_out + (int64_t)t * kH;
  // the raw rows of the picks (next decoder inputs) of a whole pass are fetched together and written
  // to the x block after the last pass, so their latency is paid once
  float pend_x[kWarpInstances];
#pragma unroll
  for (int r = 0; r < kWarpInstances; ++r) pend_x[r] = 0.f;
  auto run = [&](auto seg_tag, auto ch_tag) {
    constexpr int SEG = decltype(seg_tag)::value, CH = decltype(ch_tag)::value;
    constexpr int IPP = 32 / SEG;
    auto feed = [&](int si, int64_t b, int fed, int j) {
      if (!feed_next) return;
      const float xv = (si < count && j < p.F) ? __ldg(p.inputs + (b * p.L + fed) * (int64_t)p.F + j) : 0.f;
#pragma unroll
      for (int r = 0; r < kWarpInstances / IPP; ++r)
        if (si / IPP == r) pend_x[r] = xv;
    };
    pointer_steps_batched<SEG, CH>(pa, t, b0, count, q_base, p.h_out_inst_ld, lane, feed
                                   );
    if (feed_next) {
      const int seg_i = lane / SEG, seg_j = lane % SEG;
#pragma unroll
      for (int r = 0; r < kWarpInstances / IPP; ++r) {
        const int si = r * IPP + seg_i;
        if (si < count && seg_j < 8) store_ax(sgen, rr0 + si, seg_j, pend_x[r]);
      }
    }
  };
  using std::integral_constant;
  const bool five = pa.N % 5 == 0;
  if (pa.N <= 8)       { if (five) run(integral_constant<int, 8>{}, integral_constant<int, 5>{});
                         else      run(integral_constant<int, 8>{}, integral_constant<int, 4>{}); }
  else if (pa.N <= 16) { if (five) run(integral_constant<int, 16>{}, integral_constant<int, 5>{});
                         else      run(integral_constant<int, 16>{}, integral_constant<int, 4>{}); }
  else                 run(integral_constant<int, 32>{}, integral_constant<int, 4>{});
}


// ---- blocked encodings (GNNPN_ENC_BLOCKED128) ------------------------------------------------------------------
// float offset of (block of 128 instances, position l, tile nt, group g); + row * 8 inside (8 consecutive units)
constexpr int64_t ENC_BLK_ROW = 8 * 4 * BM * 8;                  // floats between consecutive positions l (= BM * kH)
constexpr int ENC_BLK_TILE = 4 * BM * 8;                         // floats of one (position, tile): 16 KB, contiguous
__host__ __device__ inline int64_t enc_blk_off(int64_t block, int L, int64_t l, int nt, int g) {
  return (block * L + l) * ENC_BLK_ROW + (int64_t)(nt * 4 + g) * (BM * 8);
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// L2 residency control.  The fused decoder streams 97 MB of window rows per step through L2 next to a 19 MB cell-state
// scratch that is re-read every step: the rows are loaded evict-first (read once), the scratch evict-last -- without
// the hints the scratch is evicted and costs 1.8 GB of extra DRAM traffic per launch (ncu, profiles/).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void ldg256_stream(const float* p, float4& a, float4& b, uint64_t pol) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p), "l"(pol));
}
__device__ __forceinline__ float4 ldg128_hint(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg128_hint(float* p, float a, float b, float c, float d, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Last part of a fused decode step (NR >= 1), executed by ONE thread per instance (the group-0 epilogue thread of row
// rr): d[j] = canonical dot <enc_out[b, kN+j], h'(k)>.  Same arithmetic, operation by operation, as
// pointer_finish_warp (C*tanhf, fmaf latent, expf(w - max), sequential sum in candidate order, first maximal
// probability, inverse-CDF draw) -- a thread-serial loop instead of warp shuffles.  Writes win_logits / win_probs /
// idx_out and the x block row of the next step.
template <int NR>
__device__ __forceinline__ void pointer_finish_thread(const SeqParams& p, int k, int rr, int64_t b, bool ok,
                                                      const float (&d)[NR], uint32_t sbase) {
  const PointerStepArgs& a = p.pa;
  const int N = a.N;
  const int64_t wpos = (ok ? b : 0) * a.L + (int64_t)k * N;
  float w[NR];
  float lat[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) lat[j] = (a.latent_win && j < N) ? __ldg(a.latent_win + wpos + j) : 0.f;
  const float uu = a.uniform ? __ldg(a.uniform + (int64_t)k * a.n + (ok ? b : 0)) : 0.f;
  const int forced = a.forced ? a.forced[(int64_t)k * a.n + (ok ? b : 0)] : -1;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    w[j] = -INFINITY;
    if (j < N) {
      const float l = a.use_tanh ? a.C * tanhf(d[j]) : d[j];
      if (ok) a.win_logits[wpos + j] = l;
      w[j] = a.latent_win ? fmaf(a.alpha, lat[j], l) : l;
      mx = fmaxf(mx, w[j]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (j < N) { w[j] = expf(w[j] - mx); s += w[j]; }
  }
  float best = -1.f, cum = 0.f;
  int best_j = 0, pick = -1, last_pos = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (j < N) {
      const float pj = w[j] / s;
      if (ok) a.win_probs[wpos + j] = pj;
      if (pj > best) { best = pj; best_j = j; }            // ascending j, strict >: first maximum (torch.max tie rule)
      cum += pj;
      if (pj > 0.f) last_pos = j;
      if (pick < 0 && uu < cum) pick = j;
    }
  }
  if (a.uniform) best_j = pick < 0 ? last_pos : pick;
  if (ok) a.idx_out[(int64_t)k * a.n + b] = k * N + best_j;
  if (k + 1 < p.steps) {
    const int fed = a.forced ? forced : k * N + best_j;
    float xv[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) xv[f] = (ok && f < p.F) ? __ldg(p.inputs + (b * p.L + fed) * (int64_t)p.F + f) : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = pack_h2(xv[2 * j], xv[2 * j + 1]);
      const float2 bk = unpack_h2(hi[j]);
      lo[j] = pack_h2(xv[2 * j] - bk.x, xv[2 * j + 1] - bk.y);
    }
    const uint32_t o = (uint32_t)rr * XROW_BYTES;
    st_shared_v4(sbase + OFF_AX_HI + o, hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(sbase + OFF_AX_HI + o + 16, hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(sbase + OFF_AX_LO + o, lo[0], lo[1], lo[2], lo[3]);
    st_shared_v4(sbase + OFF_AX_LO + o + 16, lo[0], lo[1], lo[2], lo[3]);
  }
}


// ------------------------------------------------------------------------------------------------
// CG = 1: one CTA per 128 instances, cta_group::1 MMAs (M=128, N=128).
// CG = 2: CTA pair (cluster of 2 = one TPC), cta_group::2 MMAs (M=256: 128 instances per CTA, N=128): each CTA
//         streams only HALF of every weight tile (64 of the 128 gate columns) and the tensor cores read the
//         other half from the peer -- half the L2->SM weight traffic, half the B-operand shared-memory reads,
//         twice the ring depth.  CTA rank 0 issues the MMAs for the pair; both CTAs run producer / epilogue.
template <bool DEC, int CG, int NR>
__global__ void __launch_bounds__(THREADS, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap map_wh_hi, const __grid_constant__ CUtensorMap map_wh_lo,
                const __grid_constant__ CUtensorMap map_wx_hi, const __grid_constant__ CUtensorMap map_wx_lo,
                const __grid_constant__ CUtensorMap map_h, const __grid_constant__ SeqParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));          // generic pointer to the aligned base
  constexpr int RING = 4 * CG;                          // slots
  constexpr int SLOT_BYTES = BLK_BYTES / CG;            // [128/CG gate columns x 64 halfs]
  constexpr int SLOT_ROWS = TILE_N / CG;
  const uint32_t bar0 = sbase + OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_RING + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_RING + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_RING + 2 + b); };
  const uint32_t a_ready_bar = bar0 + 8u * (2 * MAX_RING + 4);
  const uint32_t mma_done_bar = bar0 + 8u * (2 * MAX_RING + 5);
  const uint32_t hfull_bar = bar0 + 8u * (2 * MAX_RING + 6);
  const uint32_t hempty_bar = bar0 + 8u * (2 * MAX_RING + 7);
  const uint32_t tmem_slot = bar0 + 8u * (2 * MAX_RING + 8);
  const uint32_t h_ready_bar = bar0 + 8u * (2 * MAX_RING + 9);   // decoder: h'(t) is in the A tiles (the pointer phase follows)
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  // barriers the MMA issuer (rank 0) waits on collect arrivals from both CTAs of the pair
  const uint32_t a_ready_remote = CG == 2 ? mapa_rank(a_ready_bar, 0) : a_ready_bar;
  const uint32_t h_ready_remote = CG == 2 ? mapa_rank(h_ready_bar, 0) : h_ready_bar;
  auto arrive_leader = [&](uint32_t local_bar) {
    if (CG == 2) mbar_arrive_cluster(mapa_rank(local_bar, 0)); else mbar_arrive(local_bar);
  };
  float* sbias = reinterpret_cast<float*>(sgen + OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  constexpr bool BLK = NR > 0;                         // blocked encodings (see the file header)
  constexpr bool FUSED = DEC && BLK;                   // pointer dots fused into the cell epilogue
  const bool cta_ok = m0 < p.n;                        // a pair's second CTA may be all padding

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_wh_hi); tma_prefetch_desc(&map_wh_lo);
    tma_prefetch_desc(&map_wx_hi); tma_prefetch_desc(&map_wx_lo);
    if (!DEC && !BLK) tma_prefetch_desc(&map_h);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < RING; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), CG * EPI_WARPS); }
    mbar_init(a_ready_bar, CG * (EPI_WARPS + (DEC ? 0 : 1)));
    mbar_init(h_ready_bar, CG * EPI_WARPS);
    mbar_init(mma_done_bar, 1);
    mbar_init(hfull_bar, EPI_WARPS);
    mbar_init(hempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) { if (CG == 2) tmem_alloc_cg2(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }

  // ---- initial A operand: h(-1) split to fp16 hi/lo (zeros for the encoder), x block of step 0
  {
    for (int it = threadIdx.x; it < BM * 32; it += THREADS) {       // (row, 8-unit chunk)
      const int r = FUSED ? (it & (BM - 1)) : (it >> 5), ch = FUSED ? (it >> 7) : (it & 31);
      uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
      if (FUSED ? cta_ok : (p.h0 && m0 + r < p.n)) {
        float hv[8];
        if (FUSED) {
          // decoder start state = the encoder's last hidden state (modelPN.py:191,205): position L-1 of the block
          ldg256(p.pa.enc_out + enc_blk_off(blockIdx.x, p.L, p.L - 1, ch >> 2, ch & 3) + r * 8, hv);
        } else {
          ldg256(p.h0 + (m0 + r) * p.h0_ld + ch * 8, hv);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          hi[j] = pack_h2(hv[2 * j], hv[2 * j + 1]);
          const float2 bk = unpack_h2(hi[j]);
          lo[j] = pack_h2(hv[2 * j] - bk.x, hv[2 * j + 1] - bk.y);
        }
      }
      const uint32_t off = (uint32_t)(ch >> 3) * BLK_BYTES + sw128_off(r, ch & 7);
      st_shared_v4(sbase + OFF_A_HI + off, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(sbase + OFF_A_LO + off, lo[0], lo[1], lo[2], lo[3]);
    }
    // transformed biases: (i,f,o) * -log2e, g * -2log2e
    for (int i = threadIdx.x; i < 2 * kG; i += THREADS) {
      const int col = i & (kG - 1);
      const float b = __ldg((i < kG ? p.bias0 : p.bias) + col);
      sbias[i] = b * ((col & 3) == 2 ? -2.0f * kLog2e : -kLog2e);
    }
  }

  // x block writer (warp 3): row r -> 16 halfs = [x hi (8) | same again]; the weight block has zeros in
  // halfs 8..15, so the duplicate makes the 32B swizzle pattern irrelevant on the A side.
  auto write_x_rows = [&](const float (*xv)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lane + 32 * i;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = pack_h2(xv[i][2 * j], xv[i][2 * j + 1]);
        const float2 bk = unpack_h2(hi[j]);
        lo[j] = pack_h2(xv[i][2 * j] - bk.x, xv[i][2 * j + 1] - bk.y);
      }
      const uint32_t o = (uint32_t)r * XROW_BYTES;
      st_shared_v4(sbase + OFF_AX_HI + o, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(sbase + OFF_AX_HI + o + 16, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(sbase + OFF_AX_LO + o, lo[0], lo[1], lo[2], lo[3]);
      st_shared_v4(sbase + OFF_AX_LO + o + 16, lo[0], lo[1], lo[2], lo[3]);
    }
  };
  auto load_x_rows = [&](int t, float (*xv)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + lane + 32 * i;
#pragma unroll
      for (int f = 0; f < 8; ++f)
        xv[i][f] = (m < p.n && f < p.F) ? __ldg(p.inputs + m * p.x_inst_ld + (int64_t)t * p.F + f) : 0.f;
    }
  };
  if (warp == 3) {
    float xv[4][8];
    if (DEC) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int f = 0; f < 8; ++f) xv[i][f] = 0.f;
    } else {
      load_x_rows(0, xv);
    }
    write_x_rows(xv);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();     // barrier inits / TMEM / A operand visible pair-wide
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
  if constexpr (FUSED) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");   // role warpgroup hands registers to the epilogue
  if (warp == 0) {
    // ================= TMA producer: weights, the same 72 boxes every step (each CTA of a pair loads its
    // half of the gate columns of every box and signals the leader's barrier) =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const int row_off = (int)rank * SLOT_ROWS;
      auto load_h = [&](int it) {                      // the 8 h-part boxes of tile `it` (hi, lo per k-block)
        const int nt = it;
        for (int kb = 0; kb < KB_H; ++kb) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t dst = sbase + OFF_RING + s * SLOT_BYTES;
            if (CG == 2) {
              if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * SLOT_BYTES);
              tma_load_2d_cg2(dst, part ? &map_wh_lo : &map_wh_hi, mapa_rank(full_bar(s), 0), kb * 64,
                              nt * TILE_N + row_off);
            } else {
              mbar_arrive_expect_tx(full_bar(s), SLOT_BYTES);
              tma_load_2d(dst, part ? &map_wh_lo : &map_wh_hi, full_bar(s), kb * 64, nt * TILE_N);
            }
            if (++s == RING) { s = 0; ph ^= 1u; }
          }
        }
      };
      auto load_x = [&](int it) {                      // the x-part box (hi | lo) of tile `it`
        const int nt = it;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t dst = sbase + OFF_RING + s * SLOT_BYTES;
        if (CG == 2) {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * 2 * SLOT_ROWS * XROW_BYTES);
          const uint32_t fb = mapa_rank(full_bar(s), 0);
          tma_load_2d_cg2(dst, &map_wx_hi, fb, kH, nt * TILE_N + row_off);
          tma_load_2d_cg2(dst + SLOT_ROWS * XROW_BYTES, &map_wx_lo, fb, kH, nt * TILE_N + row_off);
        } else {
          mbar_arrive_expect_tx(full_bar(s), 2 * TILE_N * XROW_BYTES);
          tma_load_2d(dst, &map_wx_hi, full_bar(s), kH, nt * TILE_N);
          tma_load_2d(dst + TILE_N * XROW_BYTES, &map_wx_lo, full_bar(s), kH, nt * TILE_N);
        }
        if (++s == RING) { s = 0; ph ^= 1u; }
      };
      for (int t = 0; t < p.steps; ++t) {
        // decoder: the h parts of tiles 0 and 1 come first (issued while the pointer phase still runs), see the MMA warp
        const int first = DEC ? 2 : 0;
        if (DEC) { load_h(0); load_h(1); load_x(0); load_x(1); }
        for (int it = first; it < N_TILES; ++it) { load_h(it); load_x(it); }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp walks the pipeline (uniform control flow keeps the descriptors
    // in uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit =================
    if (rank == 0) {
      const uint32_t leader = elect_one();
      const uint32_t idesc = idesc_f16(CG * BM, TILE_N);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if (CG == 2) mma_f16_ss_cg2(d, a, b, idesc, acc); else mma_f16_ss(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t bar) { if (CG == 2) mma_commit_cg2(bar); else mma_commit(bar); };
      const uint64_t ax_hi = smem_desc_k_sw32(sbase + OFF_AX_HI), ax_lo = smem_desc_k_sw32(sbase + OFF_AX_LO);
      int s = 0; uint32_t ph = 0;
      uint32_t uses = 0;                                   // per-buffer use count = uses >> 1 (tiles alternate)
      const bool prof = p.prof != nullptr;
      long long w_aready = 0, w_tempty = 0, w_full = 0;
      const long long t_begin = clock64();
      auto mma_h = [&](uint32_t d) {                     // h part of one tile: 8 ring slots, 48 MMAs
        for (int kb = 0; kb < KB_H; ++kb) {
          const uint64_t a_hi = smem_desc_k_sw128(sbase + OFF_A_HI + kb * BLK_BYTES);
          const uint64_t a_lo = smem_desc_k_sw128(sbase + OFF_A_LO + kb * BLK_BYTES);
          mbar_wait_t(full_bar(s), ph, prof, w_full);
          tc_fence_after();
          const uint64_t b_hi = smem_desc_k_sw128(sbase + OFF_RING + s * SLOT_BYTES);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma(d, a_lo + (uint64_t)(ks * 2), b_hi + (uint64_t)(ks * 2), (uint32_t)((kb | ks) != 0));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma(d, a_hi + (uint64_t)(ks * 2), b_hi + (uint64_t)(ks * 2), 1u);
            commit(empty_bar(s));
          }
          __syncwarp();
          if (++s == RING) { s = 0; ph ^= 1u; }
          mbar_wait_t(full_bar(s), ph, prof, w_full);
          tc_fence_after();
          const uint64_t b_lo = smem_desc_k_sw128(sbase + OFF_RING + s * SLOT_BYTES);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma(d, a_hi + (uint64_t)(ks * 2), b_lo + (uint64_t)(ks * 2), 1u);
            commit(empty_bar(s));
          }
          __syncwarp();
          if (++s == RING) { s = 0; ph ^= 1u; }
        }
      };
      auto mma_x = [&](uint32_t d, int buf, bool last) {  // x part of one tile, then hand the accumulator to the epilogue
        mbar_wait_t(full_bar(s), ph, prof, w_full);
        tc_fence_after();
        const uint64_t bx_hi = smem_desc_k_sw32(sbase + OFF_RING + s * SLOT_BYTES);
        const uint64_t bx_lo = smem_desc_k_sw32(sbase + OFF_RING + s * SLOT_BYTES + SLOT_ROWS * XROW_BYTES);
        if (leader) {
          mma(d, ax_lo, bx_hi, 1u);
          mma(d, ax_hi, bx_hi, 1u);
          mma(d, ax_hi, bx_lo, 1u);
          commit(empty_bar(s));
          commit(tfull_bar(buf));
          if (last) commit(mma_done_bar);
        }
        __syncwarp();
        if (++s == RING) { s = 0; ph ^= 1u; }
      };
      for (int t = 0; t < p.steps; ++t) {
        int nt0 = 0;
        if (DEC) {
          // Only the K = 16 x part of a step depends on the pick.  TMEM has two accumulator buffers, so the h parts of
          // tiles 0 and 1 (96 of the step's 408 MMAs) are issued as soon as h'(t-1) is in the A tiles, i.e. while the
          // pointer phase of step t-1 still runs; their x parts follow once the picked rows are in the x block.
          if (t > 0) {
            mbar_wait_t<CG == 2>(h_ready_bar, (uint32_t)(t - 1) & 1u, prof, w_aready);
            tc_fence_after();
          }
          for (int b = 0; b < 2; ++b) {
            mbar_wait_t<CG == 2>(tempty_bar(b), (((uses + b) >> 1) & 1u) ^ 1u, prof, w_tempty);
            tc_fence_after();
            mma_h(tmem_base + (uint32_t)(b * TILE_N));
          }
          if (t > 0) {
            mbar_wait_t<CG == 2>(a_ready_bar, (uint32_t)(t - 1) & 1u, prof, w_aready);   // x(t) is in the x block
            tc_fence_after();
          }
          for (int b = 0; b < 2; ++b) mma_x(tmem_base + (uint32_t)(b * TILE_N), b, false);
          uses += 2;
          nt0 = 2;
        } else if (t > 0) {
          mbar_wait_t<CG == 2>(a_ready_bar, (uint32_t)(t - 1) & 1u, prof, w_aready);  // h'(t-1), x(t) are in smem
          tc_fence_after();
        }
        for (int nt = nt0; nt < N_TILES; ++nt, ++uses) {
          const int buf = nt & 1;
          mbar_wait_t<CG == 2>(tempty_bar(buf), ((uses >> 1) & 1u) ^ 1u, prof, w_tempty);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(buf * TILE_N);
          mma_h(d);
          mma_x(d, buf, nt == N_TILES - 1);
        }
      }
      if (prof && leader) {
        unsigned long long* o = p.prof + (size_t)blockIdx.x * 16;
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_aready; o[2] = w_tempty; o[3] = w_full;
      }
    }
  } else if (warp == 2) {
    // ================= h' store issuer (encoder): one TMA store per tile, [128 instances x 32 units] -> enc_out ====
    if (!DEC && !BLK && lane == 0) {
      uint32_t g = 0;
      for (int t = 0; t < p.steps; ++t) {
        for (int it = 0; it < N_TILES; ++it, ++g) {
          const int nt = it;
          mbar_wait(hfull_bar, g & 1u);
          tma_store_3d(&map_h, sbase + OFF_HBUF, nt * 32, t, (int)m0);
          bulk_commit();
          bulk_wait_read0();                             // the tile has been read out of shared memory
          mbar_arrive(hempty_bar);
        }
      }
      bulk_wait0();
    }
  } else if (warp == 3) {
    if (FUSED) {
      // ================= L2 prefetcher (fused decoder): the 16 KB slice (window row j, tile nt) of this CTA's block is
      // contiguous; it is requested two tiles before the epilogue multiplies it, paced by the MMA's tile barriers
      if (lane == 0 && cta_ok) {
        // tiles ahead.  N = 5 (QWS): 1 tile ahead 40.1k cycles / step, 2: 39.7k, 3: 51.8k, none: 46.0k (profiles/).  Wider
        // windows keep the prefetched-but-unused footprint (dist x N x 16 KB per CTA) at the same ~24 MB: at N = 10 two tiles
        // ahead made the kernel re-read 35 % of the rows from DRAM (ncu: 13.1 GB for 9.7 GB algorithmic)
        const int dist = p.pa.N > 5 ? 1 : 2;
        const float* const blk = p.pa.enc_out + enc_blk_off(blockIdx.x, p.L, 0, 0, 0);
        const int N = p.pa.N;
        for (int a = 0; a < dist; ++a)
          for (int j = 0; j < N; ++j) bulk_prefetch_l2(blk + (int64_t)j * ENC_BLK_ROW + a * ENC_BLK_TILE, ENC_BLK_TILE * 4);
        uint32_t uses = 0;
        for (int t = 0; t < p.steps; ++t) {
          for (int it = 0; it < N_TILES; ++it, ++uses) {
            mbar_wait(tfull_bar(it & 1), (uses >> 1) & 1u);      // tile `it` of step t is multiplied: the epilogue starts on it
            const int nt2 = (it + dist) & (N_TILES - 1);
            const int t2 = t + (it + dist >= N_TILES ? 1 : 0);
            if (t2 < p.steps)
              for (int j = 0; j < N; ++j)
                bulk_prefetch_l2(blk + ((int64_t)t2 * N + j) * ENC_BLK_ROW + nt2 * ENC_BLK_TILE, ENC_BLK_TILE * 4);
          }
        }
      }
    }
    // ================= x producer (encoder): raw input row of step t+1 -> fp16 hi/lo x block =================
    if (!DEC) {
      for (int t = 0; t + 1 < p.steps; ++t) {
        float xv[4][8];
        load_x_rows(t + 1, xv);
        mbar_wait(mma_done_bar, (uint32_t)t & 1u);        // the MMAs of step t no longer read the x block
        write_x_rows(xv);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(a_ready_remote); else mbar_arrive(a_ready_bar); }
      }
    }
  }
  } else {
    // ================= epilogue: thread = one instance (TMEM lane), 32 gate columns = 8 hidden units per tile ====
    // (fused decoder: a chunk of window-row slices, 40 registers, is in flight next to the cell arithmetic)
    if constexpr (FUSED) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = q * 32 + lane;                       // row inside the CTA
    const int64_t m = m0 + r;
    const bool ok = m < p.n;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* const c_row = p.c + (ok ? m : 0) * kH;
    float* const h_row = p.h_out + (ok ? m : 0) * p.h_out_inst_ld;
    // blocked scratch: float4 index (((cta*8 + nt)*4 + grp)*2 + half)*128 + row  -> a warp touches 512 contiguous bytes
    float* const c_blk = p.c_scr + ((int64_t)blockIdx.x * (BM * kH) + (int64_t)grp * (2 * BM * 4) + r * 4);
    // blocked encodings of this thread (8 consecutive units): + position * ENC_BLK_ROW + nt * ENC_BLK_TILE
    const int64_t blk_thread = enc_blk_off(blockIdx.x, p.L, 0, 0, grp) + r * 8;
    float* const e_out = BLK && !DEC ? p.h_out + blk_thread : nullptr;
    const float* const e_in = FUSED ? p.pa.enc_out + blk_thread : nullptr;
    const int N = DEC ? p.pa.N : 0;
    constexpr int NRA = NR > 0 ? NR : 1;
    constexpr int CHK = (NR % 5 == 0) ? 5 : 4;          // window rows in flight per thread (8 registers each)
    float acc[NRA];                                    // FUSED: running p[g] of the canonical dot, per window row
#pragma unroll
    for (int j = 0; j < NRA; ++j) acc[j] = 0.f;
    // FUSED: the first CHK window-row slices of the NEXT tile are requested as soon as the current tile's dots have
    // consumed the registers, i.e. a whole tile (barrier wait + cell arithmetic) before they are used
    float4 ea[CHK], eb[CHK];
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const bool keep_c = FUSED;
    const bool do_dot = FUSED && cta_ok;
    auto request_rows = [&](int t_, int nt_) {
#pragma unroll
      for (int u = 0; u < CHK; ++u)
        if (u < N) ldg256_stream(e_in + ((int64_t)t_ * N + u) * ENC_BLK_ROW + nt_ * ENC_BLK_TILE, ea[u], eb[u], pol_stream);
    };
    if (do_dot) request_rows(0, 0);
    uint32_t uses = 0;
    const bool prof = p.prof != nullptr;
    long long w_tfull = 0, w_hempty = 0, w_ptr = 0;
    const long long t_begin = clock64();
    for (int t = 0; t < p.steps; ++t) {
      const float4* bias4 = reinterpret_cast<const float4*>(sbias + (t == 0 ? 0 : kG));
      const bool last = t == p.steps - 1;
      const float* const e_win = FUSED ? e_in + (int64_t)t * N * ENC_BLK_ROW : nullptr;     // window rows of step t
      for (int it = 0; it < N_TILES; ++it, ++uses) {
        const int buf = it & 1;
        const int nt = it;                             // which 128 gate columns this tile holds
        const int u0 = nt * 32 + grp * 8;              // first hidden unit of this thread's chunk
        float* const c_t = c_blk + nt * (4 * 2 * BM * 4);
        float c_old[8];
        if (t > 0) {
          const float4 a = keep_c ? ldg128_hint(c_t, pol_keep) : ldg128(c_t);
          const float4 b = keep_c ? ldg128_hint(c_t + BM * 4, pol_keep) : ldg128(c_t + BM * 4);
          c_old[0] = a.x; c_old[1] = a.y; c_old[2] = a.z; c_old[3] = a.w;
          c_old[4] = b.x; c_old[5] = b.y; c_old[6] = b.z; c_old[7] = b.w;
        } else if (ok && !p.c_zero_init) {
          ldg256(c_row + u0, c_old);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) c_old[u] = 0.f;
        }
        mbar_wait_t(tfull_bar(buf), (uses >> 1) & 1u, prof, w_tfull);
        tc_fence_after();
        float v[32];
        tmem_ld_32x32_issue(t_lane + (uint32_t)(buf * TILE_N + grp * 32), v);
        tmem_ld_wait(v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(tempty_bar(buf));  // accumulators are in registers: MMA may reuse the buffer
        float cn[8], hn[8];
        lstm_cell8(v, bias4 + u0, c_old, cn, hn);
        if (!last && keep_c) {
          stg128_hint(c_t, cn[0], cn[1], cn[2], cn[3], pol_keep);
          stg128_hint(c_t + BM * 4, cn[4], cn[5], cn[6], cn[7], pol_keep);
        } else if (!last) {
          stg128(c_t, cn[0], cn[1], cn[2], cn[3]);
          stg128(c_t + BM * 4, cn[4], cn[5], cn[6], cn[7]);
        } else if (ok) {
          stg256(c_row + u0, cn);
        }
        if (DEC) {
          if (ok && p.h_out) stg256(h_row + (int64_t)t * kH + u0, hn);        // dec_h is optional in the fused decoder
        } else if (BLK) {
          // blocked encodings: the thread's 8 units of position t, coalesced across the warp (1 KB per store)
          if (cta_ok) stg256(e_out + (int64_t)t * ENC_BLK_ROW + nt * ENC_BLK_TILE, hn);
        } else {
          // fp32 h' -> 128B-swizzled [128 x 32] tile; the store issuer (warp 2) sends it to enc_out by TMA
          mbar_wait_t(hempty_bar, (uses & 1u) ^ 1u, prof, w_hempty);   // the previous tile's store has left shared memory
          const uint32_t hb = sbase + OFF_HBUF + (uint32_t)r * 128;
          st_shared_v4(hb + (uint32_t)(((2 * grp) ^ (r & 7)) << 4), __float_as_uint(hn[0]), __float_as_uint(hn[1]),
                       __float_as_uint(hn[2]), __float_as_uint(hn[3]));
          st_shared_v4(hb + (uint32_t)(((2 * grp + 1) ^ (r & 7)) << 4), __float_as_uint(hn[4]), __float_as_uint(hn[5]),
                       __float_as_uint(hn[6]), __float_as_uint(hn[7]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(hfull_bar);
        }
        uint32_t pk[8];                                  // {hi01, hi23, hi45, hi67, lo01, lo23, lo45, lo67}
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pk[j] = pack_h2(hn[2 * j], hn[2 * j + 1]);
          const float2 bk = unpack_h2(pk[j]);
          pk[4 + j] = pack_h2(hn[2 * j] - bk.x, hn[2 * j + 1] - bk.y);
        }
        if (it < N_TILES - 1) {
          tmem_st_32x32_x8(t_lane + (uint32_t)(STAGE_COL0 + nt * 32 + grp * 8), pk);
        } else {
          // tile 7's accumulators were committed after the last MMA of the step: nothing reads A any more.
          const uint32_t off7 = (uint32_t)(nt >> 1) * BLK_BYTES + sw128_off(r, (nt & 1) * 4 + grp);
          st_shared_v4(sbase + OFF_A_HI + off7, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(sbase + OFF_A_LO + off7, pk[4], pk[5], pk[6], pk[7]);
          tmem_st_wait();
#pragma unroll
          for (int jj = 1; jj < N_TILES; ++jj) {
            const int j = (nt + jj) & (N_TILES - 1);        // the seven tiles staged earlier in this step
            uint32_t sg[8];
            tmem_ld_32x32_x8(t_lane + (uint32_t)(STAGE_COL0 + j * 32 + grp * 8), sg);
            tmem_ld_wait8(sg);
            const uint32_t off = (uint32_t)(j >> 1) * BLK_BYTES + sw128_off(r, (j & 1) * 4 + grp);
            st_shared_v4(sbase + OFF_A_HI + off, sg[0], sg[1], sg[2], sg[3]);
            st_shared_v4(sbase + OFF_A_LO + off, sg[4], sg[5], sg[6], sg[7]);
          }
          if (DEC) {
            // h'(t) is in the A tiles: the MMA warp may start the h parts of the next step's first two tiles now
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(h_ready_remote); else mbar_arrive(h_ready_bar); }
          }
        }
        if (do_dot) {
          // canonical dot (pointer.cuh): s[nt][g] = fma chain over the 8 units, p[g] sequential over the tiles
#pragma unroll
          for (int j0 = 0; j0 < NR; j0 += CHK) {
            if (j0 > 0) {
#pragma unroll
              for (int u = 0; u < CHK; ++u) {
                if (j0 + u < NR && j0 + u < N)
                  ldg256_stream(e_win + (int64_t)(j0 + u) * ENC_BLK_ROW + nt * ENC_BLK_TILE, ea[u], eb[u], pol_stream);
              }
            }
#pragma unroll
            for (int u = 0; u < CHK; ++u) {
              if (j0 + u < NR && j0 + u < N) {
                // dot8's operand order: row element first, query element second
                float sj = __fmul_rn(ea[u].x, hn[0]);
                sj = fmaf(ea[u].y, hn[1], sj); sj = fmaf(ea[u].z, hn[2], sj); sj = fmaf(ea[u].w, hn[3], sj);
                sj = fmaf(eb[u].x, hn[4], sj); sj = fmaf(eb[u].y, hn[5], sj); sj = fmaf(eb[u].z, hn[6], sj);
                sj = fmaf(eb[u].w, hn[7], sj);
                acc[j0 + u] = it == 0 ? sj : __fadd_rn(acc[j0 + u], sj);
              }
            }
          }
          if (it + 1 < N_TILES) request_rows(t, it + 1);
          else if (!last) request_rows(t + 1, 0);
        }
      }
      if (DEC && !FUSED) {
        // ---- pointer step k = t: query = h'(t) (just written to dec_h by this CTA), window rows of enc_out
        const long long tp0 = prof ? clock64() : 0;
        __threadfence_block();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        pointer_phase(p, t, (warp - 4) * (BM / EPI_WARPS), m0, lane, sgen);
        if (prof) w_ptr += clock64() - tp0;
      }
      if (FUSED) {
        // ---- combine the four group partials of every instance and finish the step (one thread per instance)
        const long long tp0 = prof ? clock64() : 0;
        float* const part = reinterpret_cast<float*>(sgen + OFF_HBUF);      // [3 groups][NR][128 rows]
        if (grp > 0) {
#pragma unroll
          for (int j = 0; j < NR; ++j)
            if (j < N) part[((grp - 1) * NR + j) * BM + r] = acc[j];
        }
        named_bar_sync(1 + q, 128);                     // the four warps that own this quarter's rows
        if (grp == 0) {
          float d[NRA];
#pragma unroll
          for (int j = 0; j < NR; ++j)
            d[j] = j < N ? __fadd_rn(__fadd_rn(acc[j], part[(0 * NR + j) * BM + r]),
                                      __fadd_rn(part[(1 * NR + j) * BM + r], part[(2 * NR + j) * BM + r])) : 0.f;
          pointer_finish_thread<NRA>(p, t, r, m, ok, d, sbase);
        }
        if (prof) w_ptr += clock64() - tp0;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG == 2) mbar_arrive_cluster(a_ready_remote); else mbar_arrive(a_ready_bar); }
    }
    if (prof && warp == 4 && lane == 0) {
      unsigned long long* o = p.prof + (size_t)blockIdx.x * 16;
      o[4] = (unsigned long long)(clock64() - t_begin); o[5] = w_tfull; o[6] = w_hempty; o[7] = w_ptr;
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) { if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

template <bool DEC, int CG, int NR>
int launch_seq_cg(const float* packed, const SeqParams& p, cudaStream_t st) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return GNNPN_EUNSUPPORTED;
  CUtensorMap maps[4];
  const float* w_hi = packed + kOffTc16Hi;
  const float* w_lo = packed + kOffTc16Lo;
  const cuuint32_t box_rows = TILE_N / CG;
  for (int i = 0; i < 4; ++i) {
    // i<2: h part, 64 halfs (128 bytes) per gate column, 128B swizzle; i>=2: x part, 16 halfs, 32B swizzle
    cuuint64_t dims[2] = {(cuuint64_t)kKp16, (cuuint64_t)kG};
    cuuint64_t strides[1] = {(cuuint64_t)kKp16 * 2};
    cuuint32_t box[2] = {i < 2 ? 64u : 16u, box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)((i & 1) ? w_lo : w_hi), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, i < 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GNNPN_ESHAPE;
  }
  CUtensorMap map_h = maps[0];                         // only the row-major encoder stores through it
  if (!DEC && NR == 0) {
    // h_out as a 3-D tensor {unit, step, instance}; box = 32 units x 1 step x 128 instances, 128B swizzle
    const int64_t T = p.h_out_inst_ld / kH;
    cuuint64_t dims[3] = {(cuuint64_t)kH, (cuuint64_t)T, (cuuint64_t)p.n};
    cuuint64_t strides[2] = {(cuuint64_t)kH * 4, (cuuint64_t)p.h_out_inst_ld * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)BM};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&map_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.h_out, dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GNNPN_ESHAPE;
  }
  SeqParams pp = p;
  auto kern = lstm_seq_kernel<DEC, CG, NR>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const unsigned grid = (unsigned)(ceil_div(p.n, BM * CG) * CG);
  const int do_prof = options().prof.load(std::memory_order_relaxed);
  unsigned long long* prof = nullptr;
  if (do_prof) {                                       // debug only: synchronous, allocates
    if (cudaMalloc(&prof, (size_t)grid * 16 * 8) != cudaSuccess) return GNNPN_EUNSUPPORTED;
    cudaMemsetAsync(prof, 0, (size_t)grid * 16 * 8, st);
  }
  pp.prof = prof;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], map_h, pp);
  if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
  const int rc_launch = after_launch();
  if (do_prof) {
    unsigned long long* hbuf = (unsigned long long*)malloc((size_t)grid * 16 * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, prof, (size_t)grid * 16 * 8, cudaMemcpyDeviceToHost);
    double acc[8] = {0};
    unsigned leaders = 0;
    for (unsigned c = 0; c < grid; ++c) {
      if (hbuf[c * 16]) ++leaders;
      for (int i = 0; i < 8; ++i) acc[i] += (double)hbuf[c * 16 + i];
    }
    if (!leaders) leaders = 1;
    fprintf(stderr, "[seq prof %s cg=%d nr=%d steps=%d grid=%u] per-step cycles: mma total %.0f (wait a_ready %.0f, tmem_empty %.0f, "
            "B full %.0f) | epi total %.0f (wait tmem_full %.0f, h buf %.0f, pointer %.0f)\n", DEC ? "dec" : "enc", CG, NR,
            p.steps, grid, acc[0] / leaders / p.steps, acc[1] / leaders / p.steps, acc[2] / leaders / p.steps,
            acc[3] / leaders / p.steps, acc[4] / grid / p.steps, acc[5] / grid / p.steps, acc[6] / grid / p.steps,
            acc[7] / grid / p.steps);
    free(hbuf);
    cudaFree(prof);
  }
  return rc_launch;
}

}  // namespace seq

int tc_seq_encode(const SeqEncodeArgs& a, cudaStream_t st) {
  if (a.F < 1 || a.F > 8 || a.L < 1) return GNNPN_EUNSUPPORTED;
  seq::SeqParams p{};
  p.n = a.n; p.steps = a.L; p.L = a.L; p.F = a.F;
  p.inputs = a.inputs; p.x_inst_ld = (int64_t)a.L * a.F;
  p.bias0 = p.bias = a.packed + kOffBias;
  p.c = a.c_state; p.c_zero_init = 1;
  p.h0 = nullptr; p.h0_ld = 0;
  p.h_out = a.enc_out; p.h_out_inst_ld = (int64_t)a.L * kH;
  p.c_scr = a.c_scratch;
  p.enc_layout = a.enc_layout;
  if (a.enc_layout == GNNPN_ENC_BLOCKED128) return seq::launch_seq_cg<false, 2, 1>(a.packed, p, st);
  return seq::launch_seq_cg<false, 2, 0>(a.packed, p, st);
}

bool tc_seq_fused_decode_supported(int N) { return N >= 1 && N <= 10; }

int tc_seq_decode(const SeqDecodeArgs& a, cudaStream_t st) {
  if (a.F < 1 || a.F > 8 || a.K < 1 || a.N < 1 || a.N > kMaxWindow) return GNNPN_EUNSUPPORTED;
  seq::SeqParams p{};
  p.n = a.n; p.steps = a.K; p.L = a.L; p.F = a.F;
  p.inputs = a.inputs; p.x_inst_ld = (int64_t)a.L * a.F;
  p.bias0 = a.packed + kOffStart; p.bias = a.packed + kOffBias;
  p.c = a.c_state; p.c_zero_init = 0;
  p.h_out = a.dec_h; p.h_out_inst_ld = (int64_t)a.K * kH;
  p.pa.enc_out = a.enc_out; p.pa.enc_inst_ld = (int64_t)a.L * kH; p.pa.latent_win = a.latent_win;
  p.pa.alpha = a.alpha; p.pa.use_tanh = a.use_tanh; p.pa.C = a.C; p.pa.n = a.n; p.pa.L = a.L;
  p.pa.N = a.N; p.pa.idx_out = a.idx_out; p.pa.win_logits = a.win_logits; p.pa.win_probs = a.win_probs;
  p.pa.forced = a.forced_idx; p.pa.uniform = a.sample_uniform;
  p.c_scr = a.c_scratch;
  p.enc_layout = a.enc_layout;
  if (a.enc_layout == GNNPN_ENC_BLOCKED128) {
    // blocked encodings: the pointer dots ride in the cell epilogue (row capacity 5 / 8 / 10 of the window)
    if (!tc_seq_fused_decode_supported(a.N)) return GNNPN_EUNSUPPORTED;
    p.h0 = nullptr; p.h0_ld = 0;                         // read from the block's position L-1 inside the kernel
    if (a.N == 5) return seq::launch_seq_cg<true, 2, 5>(a.packed, p, st);
    if (a.N <= 8) return seq::launch_seq_cg<true, 2, 8>(a.packed, p, st);
    return seq::launch_seq_cg<true, 2, 10>(a.packed, p, st);
  }
  p.h0 = a.enc_out + (int64_t)(a.L - 1) * kH; p.h0_ld = (int64_t)a.L * kH;
  return seq::launch_seq_cg<true, 2, 0>(a.packed, p, st);
}

}  // namespace gnnpn
