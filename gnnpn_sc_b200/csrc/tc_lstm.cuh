// Host-side interface of the tcgen05 LSTM recurrence (implemented in tc_kernels.cu).
#pragma once
#include <cuda.h>
#include "lstm_step.cuh"

namespace gnnpn {

constexpr int kKp = kH + 32;     // [h | x | zero pad] columns of the tensor-core A operand (9 k-blocks of 32)

// float offsets inside one packed LSTM block (see gnnpn_pn_packed_lstm_floats)
constexpr size_t kOffBias = (size_t)(kH + kXPad) * kG;
constexpr size_t kOffStart = kOffBias + kG;
constexpr size_t kOffTcHi = kOffStart + kG;
constexpr size_t kOffTcLo = kOffTcHi + (size_t)kG * kKp;
constexpr size_t kPackedFloats = kOffTcLo + (size_t)kG * kKp;

struct TcLstmPlan {
  CUtensorMap a_hi[2], a_lo[2], b_hi, b_lo;
  float* hi[2];
  float* lo[2];
  int64_t n;
};

size_t tc_lstm_workspace_bytes(int64_t n);
// carve the workspace into the two ping-pong [n, kKp] hi/lo pairs and build the TMA descriptors
int tc_lstm_plan(TcLstmPlan* plan, void* workspace, size_t workspace_bytes, int64_t n, const float* packed);
// zero both ping-pong buffers and stage h = 0, x = inputs[:, row0, :] into buffer 0 (encoder start)
int tc_lstm_reset(const TcLstmPlan& plan, const float* inputs, int64_t x_inst_ld, int row0, int F, cudaStream_t st);
// buffer `dst` <- tf32 split of fp32 rows h[m*ld .. +kH), padding columns zeroed (decoder start)
int tc_lstm_load_h(const TcLstmPlan& plan, int dst, const float* h, int64_t ld, cudaStream_t st);
int tc_lstm_zero(const TcLstmPlan& plan, int which, cudaStream_t st);

struct TcLstmStep {
  int cur;                 // ping-pong buffer holding this step's [h | x]; the epilogue writes buffer cur^1
  int use_x;               // 0: contract over h only (decoder start token is folded into the bias)
  const float* bias;       // [kG]
  float* c;                // [n, kH]
  float* h_out;            // exact fp32 h', rows h_out_ld apart
  int64_t h_out_ld;
  const float* x_next;     // raw rows for the NEXT step (encoder) or nullptr (decoder: pointer step stages them)
  int64_t x_inst_ld;
  int x_row_next;
  int F;
  int first;
};
int tc_lstm_step(const TcLstmPlan& plan, const TcLstmStep& s, cudaStream_t st);

}  // namespace gnnpn
