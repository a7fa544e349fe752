// Probe of the fused decoder's pointer phase-1 loop in isolation: 148 CTAs x 16 warps, 8 chunks of 5 rows per warp
// per "step"; variants: V0 loads only, V1 + dot8 + 5 butterfly warp_sums, V2 + dot8 + transposing (halving) reduction.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float dot8(const float4 r0, const float4 r1, const float4 q0, const float4 q1) {
  float s = r0.x * q0.x;
  s = fmaf(r0.y, q0.y, s); s = fmaf(r0.z, q0.z, s); s = fmaf(r0.w, q0.w, s);
  s = fmaf(r1.x, q1.x, s); s = fmaf(r1.y, q1.y, s); s = fmaf(r1.z, q1.z, s); s = fmaf(r1.w, q1.w, s);
  return s;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int V>
__global__ void __launch_bounds__(512, 1) probe(const float* base, const float* qbase, long long inst_ld, int iters, float* out,
                                                long long* cyc) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const long long inst0 = ((long long)blockIdx.x * 16 + warp) * 8;
    const float* rows = base + inst0 * inst_ld + (long long)it * 5 * 256;
    const float* qr = qbase + inst0 * 256;
    float4 r0[5], r1[5], q0, q1;
    auto issue = [&]() {
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const float4* p = reinterpret_cast<const float4*>(rows + u * 256);
        r0[u] = ldg_stream(p + lane); r1[u] = ldg_stream(p + 32 + lane);
      }
      q0 = reinterpret_cast<const float4*>(qr)[lane]; q1 = reinterpret_cast<const float4*>(qr)[32 + lane];
    };
    issue();
    float dv0 = 0.f, dv1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float part[8];
#pragma unroll
      for (int u = 0; u < 5; ++u) part[u] = dot8(r0[u], r1[u], q0, q1);
      part[5] = part[6] = part[7] = 0.f;
      rows += inst_ld; qr += 256;
      if (c + 1 < 8) issue();
      if (V == 0) {
#pragma unroll
        for (int u = 0; u < 5; ++u) acc += part[u];
      } else if (V == 1) {
        const int base_l = (c & 3) * 8;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const float d = warp_sum(part[u]);
          if (lane == base_l + u) { if (c < 4) dv0 = d; else dv1 = d; }
        }
      } else {
        // halving reduction: same pairing order (16, 8, 4, 2, 1) as warp_sum -> same bits; value u ends in lanes with
        // bits (4,3,2) == u
        float a[4], b2[2], d;
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float keep = h16 ? part[u + 4] : part[u], send = h16 ? part[u] : part[u + 4];
          a[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float keep = h8 ? a[u + 2] : a[u], send = h8 ? a[u] : a[u + 2];
          b2[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const float keep = h4 ? b2[1] : b2[0], send = h4 ? b2[0] : b2[1];
          d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        // lane l now holds the sum of value ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1)
        if (c < 4) { if ((lane & 3) == (c & 3)) dv0 = d; } else { if ((lane & 3) == (c & 3)) dv1 = d; }
      }
    }
    acc += dv0 + dv1;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc + sm[0];
}
int main() {
  const int n = 148 * 128, L = 235, H = 256;
  const long long inst_ld = (long long)L * H;
  float* buf; cudaMalloc(&buf, (size_t)n * inst_ld * 4); cudaMemset(buf, 0, (size_t)n * inst_ld * 4);
  float* q; cudaMalloc(&q, (size_t)n * H * 4); cudaMemset(q, 0, (size_t)n * H * 4);
  float* out; cudaMalloc(&out, 4);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  for (int v = 0; v < 3; ++v)
    for (int grid : {148, 74, 16}) {
      auto k = v == 0 ? probe<0> : v == 1 ? probe<1> : probe<2>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
      const int iters = 40;
      k<<<grid, 512, 225 * 1024>>>(buf, q, inst_ld, iters, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      printf("variant %d grid %3d: %s  %.0f cycles per 8-instance phase (%.0f per chunk)\n", v, grid, cudaGetErrorString(e),
             s / grid / iters, s / grid / iters / 8);
    }
  return 0;
}
