// One LSTM time step for a whole batch of composition instances:
//   gates[m, 4j+g] = sum_k [h | x][m,k] * P[k][4j+g] + bias[4j+g]          (fp32 FFMA tile GEMM)
//   c' = sigm(f)*c + sigm(i)*tanh(g) ; h' = sigm(o)*tanh(c')                  (fused epilogue)
// P is the packed block written by gnnpn_pn_pack_lstm_f32 (gate-interleaved columns so one thread
// owns the four gates of a hidden unit and the cell update needs no exchange).
//
// Reference semantics: torch nn.LSTM cell, gate order i,f,g,o (modelPN.py:157-158,191,205).
#pragma once
#include "common.cuh"

namespace gnnpn {

constexpr int kH = 256;        // hidden_size of every PN ini section (environment.ini:24,39,54,70)
constexpr int kG = 4 * kH;     // gate columns
constexpr int kXPad = 32;      // raw input columns padded to two k-tiles of 16 (embedding_size 20 + 8 QoS/constraint columns = 28)

struct LstmStepArgs {
  const float* h_in;      // [M, kH] rows h_in_ld apart; ignored when first != 0
  int64_t h_in_ld;
  const float* x;         // raw PN rows [M, L, F]; row of instance m is x + m*x_inst_ld + row*F
  int64_t x_inst_ld;
  int x_row;              // fixed row (encoder step t) or <0: per-instance gather
  const int32_t* gather;  // [M] row per instance (decoder input = previously selected candidate)
  int F;                  // raw columns (<= kXPad)
  int use_x;              // 0: no input term (decoder start token is folded into bias)
  const float* P;         // packed [(kH + kXPad)][kG]
  const float* bias;      // [kG]
  float* c;               // [M, kH] in/out
  const float* c_in;      // optional: read the old cell state from here instead of `c` (c is then write-only)
  float* h_out;           // [M, kH] rows h_out_ld apart
  int64_t h_out_ld;
  int M;
  int first;              // 1: h_in = 0 and c = 0 (first encoder step)
  // training replay (pn_train.cu): optional per-step saves for the backward pass, or nullptr
  float* gates_out;       // [M, kG] post-activation gates, columns 4j+{i,f,g,o}
};

int launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream);

}  // namespace gnnpn
