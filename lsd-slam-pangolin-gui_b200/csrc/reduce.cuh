// reduce.cuh -- the reduction of one partial record (shared by the SE3 and Sim3 trackers): per-thread sums of NW warps ->
// NF fp32 + ND fp64 numbers, in a fixed order (a pure function of the inputs: bit-reproducible).
#pragma once
#include <cuda_runtime.h>

namespace lsd {

// Step 1, inside each warp, registers only: a PACKED butterfly.  A plain butterfly spends five shuffles per value (190 per
// lane, most of them moving sums nobody needs); here, at the step with lane distance o, the two halves of every lane pair
// split the values still to be reduced -- the lane with bit o clear keeps the even ones and hands over the odd ones, its
// partner the other way round -- so the number of live values halves with every step: 17 + 9 + 5 + 3 + 2 shuffles for the 33
// floats, 8 for the 5 doubles.  Every value still goes through the full tree (L, L^16), (.,.^8), (.,.^4), (.,.^2), (.,.^1), and
// fp addition is commutative, so both lanes of a pair hold the same bits: the result is a pure function of the 32 inputs.
// Step 2: the warps' sums meet in shared memory (688 bytes per group, double-buffered by record parity so that ONE barrier
// per record suffices) and thread r adds them in warp order: v = (((0 + w0) + w1) + w2) + w3.
// r03g: replaces a 22 KB parking area, 38 STS + 44 LDS + 65 SHFL per lane and three barriers (1.4 us per record of a live
// evaluation; numbers for the SE3 tracker: 33 floats + 5 doubles, 4 warps).
template <typename T, int N, int O>
__device__ __forceinline__ void packed_step(const T (&a)[N], const int (&ia)[N], T (&b)[(N + 1) / 2], int (&ib)[(N + 1) / 2], const bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; i++) {
    const T keep = upper ? a[2 * i + 1] : a[2 * i];
    const T give = upper ? a[2 * i] : a[2 * i + 1];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, give, O);
    ib[i] = upper ? ia[2 * i + 1] : ia[2 * i];
  }
  if (N & 1) {  // the odd one out is kept by both halves
    b[N / 2] = a[N - 1] + __shfl_xor_sync(0xffffffffu, a[N - 1], O);
    ib[N / 2] = ia[N - 1];
  }
}

// Warp-wide sums of N per-lane values: on return out[j] is the complete sum of value number row[j] (every value ends up on
// at least one lane; lanes that hold the same value hold the same bits).
template <typename T, int N>
__device__ __forceinline__ void packed_warp_sum(const T (&v)[N], const int lane, T (&out)[(((((N + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2],
                                                int (&row)[(((((N + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2]) {
  constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
  int i0[N];
#pragma unroll
  for (int i = 0; i < N; i++) i0[i] = i;
  T b1[N1], b2[N2], b3[N3], b4[N4];
  int i1[N1], i2[N2], i3[N3], i4[N4];
  packed_step<T, N, 16>(v, i0, b1, i1, (lane & 16) != 0);
  packed_step<T, N1, 8>(b1, i1, b2, i2, (lane & 8) != 0);
  packed_step<T, N2, 4>(b2, i2, b3, i3, (lane & 4) != 0);
  packed_step<T, N3, 2>(b3, i3, b4, i4, (lane & 2) != 0);
  packed_step<T, N4, 1>(b4, i4, out, row, (lane & 1) != 0);
}

// per reducing group: the warps' sums of two consecutive records (double-buffered by record parity: one barrier per record)
template <int NF, int ND, int NW>
struct RecordRed {
  float f[2][NW][NF + 1];
  double d[2][NW][ND + 1];
};

// Record layout at dst: ND doubles, then NF floats.
// `parity`: which half of the scratch this record uses (the caller alternates it record by record, or separates records by a
// barrier of its own).  `bar`: the barrier of the NW * 32 threads that reduce the record.
template <int NF, int ND, int NW, class Bar>
__device__ __forceinline__ void reduce_record(const float (&acc)[NF], const double (&dacc)[ND], float *dst, RecordRed<NF, ND, NW> &sm, const int tid,
                                              const int parity, Bar bar) {
  const int lane = tid & 31, wid = tid >> 5;
  {
    constexpr int OF = (((((NF + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2;
    constexpr int OD = (((((ND + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2;
    float of[OF];
    int rf[OF];
    packed_warp_sum<float, NF>(acc, lane, of, rf);
#pragma unroll
    for (int j = 0; j < OF; j++) sm.f[parity][wid][rf[j]] = of[j];
    double od[OD];
    int rd[OD];
    packed_warp_sum<double, ND>(dacc, lane, od, rd);
#pragma unroll
    for (int j = 0; j < OD; j++) sm.d[parity][wid][rd[j]] = od[j];
  }
  bar();
  if (tid < NF) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < NW; k++) v += sm.f[parity][k][tid];
    dst[2 * ND + tid] = v;
  } else if (tid < NF + ND) {
    const int r = tid - NF;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < NW; k++) v += sm.d[parity][k][r];
    reinterpret_cast<double *>(dst)[r] = v;
  }
}

}  // namespace lsd
