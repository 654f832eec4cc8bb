// In-register complex FFT butterflies (forward transform, e^{-2*pi*i*jk/N}), N = 2..32.
//
// Fully unrolled decimation-in-frequency radix-2 network with the trivial twiddles
// (1, -i, (+-1-i)/sqrt2) special-cased at compile time.  All array indices are
// compile-time constants after unrolling, so the data stays in registers and the
// twiddles become FFMA immediates.  Output X[k] lands at index bitrev<N>(k).
//
// Host+device so the network can be unit-tested on the CPU (tests/test_fft_host.py).
#pragma once

#ifdef __CUDACC__
#define SPL_HD __host__ __device__ __forceinline__
#else
#define SPL_HD inline
#endif

namespace spl {

// cos(2*pi*j/64), j = 0..16 ; everything else follows by symmetry.
SPL_HD constexpr float cos64(int j) {
  constexpr float t[17] = {1.0f,
                           0.99518472667219688624f,
                           0.98078528040323044913f,
                           0.95694033573220886494f,
                           0.92387953251128675613f,
                           0.88192126434835502971f,
                           0.83146961230254523708f,
                           0.77301045336273696081f,
                           0.70710678118654752440f,
                           0.63439328416364549822f,
                           0.55557023301960222474f,
                           0.47139673682599764856f,
                           0.38268343236508977173f,
                           0.29028467725446236764f,
                           0.19509032201612826785f,
                           0.09801714032956060199f,
                           0.0f};
  // j in [0, 64)
  return j <= 16 ? t[j] : (j <= 32 ? -t[32 - j] : (j <= 48 ? -t[j - 32] : t[64 - j]));
}
SPL_HD constexpr float sin64(int j) { return cos64((j + 48) & 63); }  // sin(x) = cos(x - pi/2)

template <int N>
SPL_HD constexpr int bitrev(int k) {
  int r = 0;
  for (int b = 1; b < N; b <<= 1) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

// One DIF butterfly: (a, b) -> (a + b, (a - b) * W_M^j),  W_M = e^{-2 pi i / M}.
template <int M, int J>
SPL_HD void dif_bfly(float& ar, float& ai, float& br, float& bi) {
  const float tr = ar - br, ti = ai - bi;
  ar = ar + br;
  ai = ai + bi;
  if constexpr (J == 0) {
    br = tr;
    bi = ti;
  } else if constexpr (4 * J == M) {  // W = -i
    br = ti;
    bi = -tr;
  } else if constexpr (8 * J == M) {  // W = (1 - i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    br = (tr + ti) * h;
    bi = (ti - tr) * h;
  } else if constexpr (8 * J == 3 * M) {  // W = (-1 - i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    br = (ti - tr) * h;
    bi = -(tr + ti) * h;
  } else {
    constexpr float c = cos64(J * (64 / M));
    constexpr float s = sin64(J * (64 / M));  // W = c - i s
    br = tr * c + ti * s;
    bi = ti * c - tr * s;
  }
}

// Butterfly whose second input is known to be zero: (a, 0) -> (a, a * W_M^j).
template <int M, int J>
SPL_HD void dif_bfly_bzero(float ar, float ai, float& br, float& bi) {
  if constexpr (J == 0) {
    br = ar;
    bi = ai;
  } else if constexpr (4 * J == M) {
    br = ai;
    bi = -ar;
  } else if constexpr (8 * J == M) {
    constexpr float h = 0.70710678118654752440f;
    br = (ar + ai) * h;
    bi = (ai - ar) * h;
  } else if constexpr (8 * J == 3 * M) {
    constexpr float h = 0.70710678118654752440f;
    br = (ai - ar) * h;
    bi = -(ar + ai) * h;
  } else {
    constexpr float c = cos64(J * (64 / M));
    constexpr float s = sin64(J * (64 / M));
    br = ar * c + ai * s;
    bi = ai * c - ar * s;
  }
}

// NZ: inputs with index >= NZ are known to be zero (zero-padded frames); only the first
// (widest) stage can exploit it, afterwards every slot is populated.
template <int N, int SPAN, int BASE, int J, int NZ>
struct DifStageJ {
  static SPL_HD void run(float (&re)[N], float (&im)[N]) {
    if constexpr (2 * SPAN == N && BASE + J + SPAN >= NZ) {
      static_assert(BASE + J < NZ || NZ == 0, "at least half of the inputs must be populated");
      dif_bfly_bzero<2 * SPAN, J>(re[BASE + J], im[BASE + J], re[BASE + J + SPAN], im[BASE + J + SPAN]);
    } else {
      dif_bfly<2 * SPAN, J>(re[BASE + J], im[BASE + J], re[BASE + J + SPAN], im[BASE + J + SPAN]);
    }
    if constexpr (J + 1 < SPAN) DifStageJ<N, SPAN, BASE, J + 1, NZ>::run(re, im);
  }
};

template <int N, int SPAN, int BASE, int NZ>
struct DifStageB {
  static SPL_HD void run(float (&re)[N], float (&im)[N]) {
    DifStageJ<N, SPAN, BASE, 0, NZ>::run(re, im);
    if constexpr (BASE + 2 * SPAN < N) DifStageB<N, SPAN, BASE + 2 * SPAN, NZ>::run(re, im);
  }
};

template <int N, int SPAN, int NZ>
struct DifAll {
  static SPL_HD void run(float (&re)[N], float (&im)[N]) {
    DifStageB<N, SPAN, 0, NZ>::run(re, im);
    if constexpr (SPAN > 1) DifAll<N, SPAN / 2, NZ>::run(re, im);
  }
};

// In-place forward DFT of N complex values held in registers; X[k] = out[bitrev<N>(k)].
// NZ (default N): inputs re/im[NZ..N) are known zeros and need not be initialised... they ARE
// written by the first stage, so the arrays are fully populated afterwards.
template <int N, int NZ = N>
SPL_HD void fft_dif(float (&re)[N], float (&im)[N]) {
  static_assert(N >= 2 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two <= 32");
  static_assert(NZ > N / 2 && NZ <= N, "pruning supports up to N/2 trailing zeros");
  DifAll<N, N / 2, NZ>::run(re, im);
}

}  // namespace spl
