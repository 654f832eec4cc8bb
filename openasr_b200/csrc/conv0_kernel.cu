// Row f2: the first layer of the conv-subsampling block that consumes the
// front-end's [B, T, D] features -- Conv2d(1 -> C, 3x3, stride (2, 1)) + ReLU of
// src/blocks/conv_layers.py:122-150 (Conv2dSubsampleV2, "subsample/conv0" + "subsample/relu0") --
// read straight from the feature tensor kernel B just wrote (L2 resident), without the
// unsqueeze / NCHW staging of the reference.
//
//   out[b, c, t1, d1] = max(0, bias[c] + sum_{i,j<3} w[c, 0, i, j] * x[b, 2 t1 + i, d1 + j])
//   T1 = (T - 3) / 2 + 1,  D1 = D - 2,  output layout [B, C, T1, D1] (what conv1 consumes).
//
// The output is C * D1 / (2 D) ~ 15.6x larger than the input, so the kernel is HBM-write bound:
// a thread keeps the 9 taps of 4 output positions in registers (packed in pairs for the fp32x2
// pipe), walks the C channels with the weights broadcast from shared memory (3 LDS.128 per
// channel), and every store instruction of a warp covers 128 contiguous bytes of one channel plane.
#include "fft_c2.cuh"
#include "spl_internal.cuh"

namespace spl {

constexpr int kConvThreads = 256;
constexpr int kConvPos = 4;                        // output positions per thread, 32 apart
constexpr int kConvTile = kConvThreads * kConvPos;  // positions per CTA
constexpr int kConvMaxC = 64;

__global__ void __launch_bounds__(kConvThreads) conv0_relu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, float* __restrict__ out,
                                                                  int T, int D, int T1, int D1, int C) {
  __shared__ float4 sw[kConvMaxC * 3];  // per channel: {w00 w01 w02 w10} {w11 w12 w20 w21} {w22 bias 0 0}
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C * 3; i += kConvThreads) {
    const int c = i / 3, q = i - 3 * c;
    const float* wc = w + (size_t)c * 9;
    float4 v;
    if (q == 0) v = make_float4(wc[0], wc[1], wc[2], wc[3]);
    else if (q == 1) v = make_float4(wc[4], wc[5], wc[6], wc[7]);
    else v = make_float4(wc[8], bias ? bias[c] : 0.f, 0.f, 0.f);
    sw[i] = v;
  }
  const int b = blockIdx.y;
  const int plane = T1 * D1;
  const int p0 = blockIdx.x * kConvTile + warp * (32 * kConvPos) + lane;
  float xv[kConvPos][9];
#pragma unroll
  for (int k = 0; k < kConvPos; ++k) {
    const int p = p0 + 32 * k;
    const int pc = p < plane ? p : plane - 1;  // clamp: loads stay in bounds, stores are predicated
    const int t1 = pc / D1, d1 = pc - t1 * D1;
    const float* xr = x + ((size_t)b * T + 2 * t1) * D + d1;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) xv[k][3 * i + j] = __ldg(xr + (size_t)i * D + j);
  }
  c2 x01[9], x23[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    x01[tap] = c2_make(xv[0][tap], xv[1][tap]);
    x23[tap] = c2_make(xv[2][tap], xv[3][tap]);
  }
  __syncthreads();
  float* ob = out + (size_t)b * C * plane;
#pragma unroll 2
  for (int c = 0; c < C; ++c) {
    const float4 wa = sw[3 * c], wb = sw[3 * c + 1], wc = sw[3 * c + 2];
    const float wt[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
    c2 a01 = c2_splat(wc.y), a23 = c2_splat(wc.y);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {  // same order as a row-major 3x3 dot product
      a01 = c2_fma(x01[tap], c2_splat(wt[tap]), a01);
      a23 = c2_fma(x23[tap], c2_splat(wt[tap]), a23);
    }
    float* oc = ob + (size_t)c * plane;
    const float r[4] = {c2_re(a01), c2_im(a01), c2_re(a23), c2_im(a23)};
#pragma unroll
    for (int k = 0; k < kConvPos; ++k)
      if (p0 + 32 * k < plane) oc[p0 + 32 * k] = r[k] > 0.f ? r[k] : (r[k] != r[k] ? r[k] : 0.f);  // torch.relu propagates NaN
  }
}

cudaError_t launch_conv0_relu(const float* x, const float* w, const float* bias, float* out, int B, int T, int D, int C,
                              cudaStream_t st) {
  const int T1 = (T - 3) / 2 + 1, D1 = D - 2;
  dim3 grid((T1 * D1 + kConvTile - 1) / kConvTile, B);
  conv0_relu_kernel<<<grid, kConvThreads, 0, st>>>(x, w, bias, out, T, D, T1, D1, C);
  return cudaGetLastError();
}

}  // namespace spl
