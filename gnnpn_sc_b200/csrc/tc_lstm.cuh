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
// fp16-split operand blocks (default recurrence path): [4H][kKp16] halfs, weights pre-scaled by kW16Scale
constexpr int kKp16 = kH + 64;   // [h | x | zero pad] as 5 k-blocks of 64 halfs (128-byte swizzle rows)
constexpr float kW16Scale = 16.0f;   // |w| <~ 0.06 would push the fp16 "lo" parts into subnormals; 2^4 is exact to undo
constexpr size_t kOffTc16Hi = kOffTcLo + (size_t)kG * kKp;
constexpr size_t kOffTc16Lo = kOffTc16Hi + (size_t)kG * kKp16 / 2;
constexpr size_t kPackedFloats = kOffTc16Lo + (size_t)kG * kKp16 / 2;

struct TcLstmPlan {
  CUtensorMap a_hi[2], a_lo[2], b_hi, b_lo;
  void* hi[2];          // [n, ld] tf32-in-fp32 or fp16
  void* lo[2];
  int64_t n;
  int f16;              // 1: fp16 split (kKp16 columns), 0: tf32 split (kKp columns)
  int ld;               // elements per row
};
int tc_lstm_default_f16();     // GNNPN_TC_KIND=tf32 selects the tf32 split, anything else fp16

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
