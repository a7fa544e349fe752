// Host-side interface of the persistent tcgen05 LSTM sequence kernels (implemented in tc_seq.cu).
//
// One launch runs a whole LSTM scan (encoder: L steps; decoder: K steps with the pointer step fused
// between them).  A CTA owns 128 composition instances for the whole scan; the recurrence is CTA-local
// (h never leaves the SM between steps), so there is no grid-wide synchronisation.
#pragma once
#include <cuda.h>
#include "lstm_step.cuh"
#include "tc_lstm.cuh"

namespace gnnpn {

// cell-state scratch of the persistent kernels: 128 x kH floats per CTA (blocked layout), CTAs launched in pairs
inline size_t tc_seq_scratch_floats(int64_t n) { return (size_t)((n + 255) / 256) * 256 * kH; }

struct SeqEncodeArgs {
  const float* inputs;     // [n, L, F]
  int64_t n;
  int L;
  int F;                   // <= 8
  const float* packed;     // packed LSTM block (uses the bias and the fp16 hi/lo operand blocks)
  float* enc_out;          // [n, L, kH], or the blocked layout (gnnpn_pn_enc_out_floats) when enc_layout says so
  float* c_state;          // [n, kH] out
  float* c_scratch;        // tc_seq_scratch_floats(n) floats
  int enc_layout;          // GNNPN_ENC_ROWMAJOR / GNNPN_ENC_BLOCKED128 (CTA-pair scan only)
  // training (column-split scan only): per-step saves for the BPTT, or nullptr
  float* save_gates;       // [L, n, 4H] post-activation gates, columns 4j + {i,f,g,o}
  float* save_c;           // [L, n, H] cell state after every step
};
// returns GNNPN_EUNSUPPORTED when the shape is outside what the persistent kernel covers
int tc_seq_encode(const SeqEncodeArgs& a, cudaStream_t st);

// Small-batch (latency) variant, tc_colsplit.cu: a cluster of 8 CTAs shares 128 instances and splits the gate
// columns; `scratch` = tc_colsplit_scratch_bytes(n) bytes (1024-byte aligned) for the h' exchange.
bool tc_colsplit_wanted(int64_t n);          // fused decoder
bool tc_colsplit_wanted_encode(int64_t n);
int tc_colsplit_max_active_clusters();
size_t tc_colsplit_scratch_bytes(int64_t n);
int tc_colsplit_encode(const SeqEncodeArgs& a, void* scratch, cudaStream_t st);
struct SeqDecodeArgs;
int tc_colsplit_decode(const SeqDecodeArgs& a, void* scratch, cudaStream_t st);

struct SeqDecodeArgs {
  const float* inputs;     // [n, L, F]
  const float* enc_out;    // [n, L, kH]
  float* c_state;          // [n, kH] in/out
  const float* latent_win; // [n, L] or nullptr
  float alpha;
  const float* packed;     // decoder block
  int use_tanh;
  float C;
  int64_t n;
  int L, F, K, N;
  float* dec_h;            // [n, K, kH]
  int32_t* idx_out;        // [K, n]
  float* win_logits;       // [n, L]
  float* win_probs;        // [n, L]
  const int32_t* forced_idx;     // [K, n] or nullptr
  const float* sample_uniform;   // [K, n] or nullptr
  float* c_scratch;              // tc_seq_scratch_floats(n) floats
  int enc_layout;                // layout of enc_out
  float* save_gates;             // training (column-split scan only): [K, n, 4H] post-activation gates, or nullptr
  float* save_c;                 // [K, n, H] cell state after every step, or nullptr
};
bool tc_seq_fused_decode_supported(int N);     // blocked-layout decoder (pointer dots fused into the cell epilogue)
int tc_seq_decode(const SeqDecodeArgs& a, cudaStream_t st);

}  // namespace gnnpn
