// se_sortnet.h — Batcher's odd-even merge sort as a fully unrolled compare-exchange network over an array that lives
// in registers (N a power of two; 191 compare-exchanges for N = 32, 543 for N = 64 — the bitonic network needs 240 /
// 672).  The pair list is built at compile time and walked by ONE flat unrolled loop, so every index is a literal
// after unrolling (nested data-dependent loop bounds left the array on the local-memory stack).  `ce(lo, hi)` must
// leave min in `lo` and max in `hi`.  Host-compilable: tests/test_host_cpu.py sorts random and 0/1 inputs through the
// same template with g++.
#pragma once
#if defined(__CUDACC__)
#define SE_SORTNET_HD __host__ __device__ __forceinline__
#define SE_SORTNET_CX __host__ __device__ constexpr
#else
#define SE_SORTNET_HD inline
#define SE_SORTNET_CX constexpr
#endif

namespace se {

template <int N>
struct OddEvenNet {
  static_assert(N >= 1 && N <= 128 && (N & (N - 1)) == 0, "power of two, <= 128");
  static constexpr int kMax = N < 2 ? 1 : N * 10;  // >= the count for every N <= 128 (N = 64: 543, N = 128: 1471 > 1280 is not used)
  struct Pairs {
    unsigned char a[kMax];
    unsigned char b[kMax];
    int n;
  };
  static SE_SORTNET_CX Pairs make() {
    Pairs p{};
    p.n = 0;
    for (int q = 1; q < N; q <<= 1)
      for (int k = q; k >= 1; k >>= 1)
        for (int j = k % q; j + k < N; j += 2 * k)
          for (int i = 0; i < k; ++i)
            if (i + j + k < N && (i + j) / (2 * q) == (i + j + k) / (2 * q)) {
              p.a[p.n] = (unsigned char)(i + j);
              p.b[p.n] = (unsigned char)(i + j + k);
              ++p.n;
            }
    return p;
  }
};

template <int N, typename T, typename CE>
SE_SORTNET_HD void sortnet_oddeven(T (&v)[N], CE ce) {
  static_assert(N <= 64, "pair table sized for N <= 64");
  constexpr typename OddEvenNet<N>::Pairs P = OddEvenNet<N>::make();
#pragma unroll
  for (int c = 0; c < P.n; ++c) ce(v[P.a[c]], v[P.b[c]]);
}

}  // namespace se
