// Standalone probe (run under gpurun): 2-D TMA tile load of a 36x12 box of 4-byte cells with out-of-bounds origin, descriptor taken
// (a) from a __grid_constant__ parameter, (b) from global memory + fence.proxy.tensormap, (c) from global memory without the fence.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define BW 36
#define BH 12
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ int g_spins;

template <int MODE>
__global__ void k_probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int x0, int y0, uint32_t *out, int *status) {
  __shared__ __align__(128) uint32_t tile[BH][BW];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void *tm = MODE == 0 ? (const void *)&pmap : (const void *)gmap;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)(BW * BH * 4)) : "memory");
    if (MODE == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(&tile[0][0])),
                 "l"(tm), "r"(x0), "r"(y0), "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned done = 0;
  int spin = 0;
  for (; !done && spin < (1 << 22); spin++) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  if (threadIdx.x == 0) { status[0] = (int)done; status[1] = spin; }
  __syncthreads();
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&tile[0][0])[i];
}

int main() {
  const int W = 640, H = 480;
  std::vector<uint32_t> h((size_t)W * H);
  for (int i = 0; i < W * H; i++) h[i] = 1000000u + i;
  uint32_t *d, *dout;
  int *dst;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&dout, BW * BH * 4);
  cudaMalloc(&dst, 8);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  printf("entry point: err=%d q=%d ptr=%p\n", (int)e, (int)q, fp);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap map;
  const cuuint64_t gdim[2] = {W, H}, gstr[1] = {W * 4};
  const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = ((Fn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CUtensorMap *gmap;
  cudaMalloc(&gmap, sizeof(map));
  cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 3; mode++)
    for (int org = 0; org < 2; org++) {
      const int x0 = org ? -2 : 30, y0 = org ? -2 : 6;
      cudaMemset(dout, 0xff, BW * BH * 4);
      cudaMemset(dst, 0, 8);
      if (mode == 0) k_probe<0><<<1, 256>>>(map, gmap, x0, y0, dout, dst);
      if (mode == 1) k_probe<1><<<1, 256>>>(map, gmap, x0, y0, dout, dst);
      if (mode == 2) k_probe<2><<<1, 256>>>(map, gmap, x0, y0, dout, dst);
      cudaError_t le = cudaDeviceSynchronize();
      int st[2] = {0, 0};
      std::vector<uint32_t> o(BW * BH);
      cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(o.data(), dout, BW * BH * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int yy = 0; yy < BH; yy++)
        for (int xx = 0; xx < BW; xx++) {
          const int gx = x0 + xx, gy = y0 + yy;
          const uint32_t want = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? 1000000u + gx + gy * W : 0u;
          bad += o[yy * BW + xx] != want;
        }
      printf("mode %d origin (%d,%d): sync=%s done=%d spins=%d mismatches=%d first=%u\n", mode, x0, y0, cudaGetErrorString(le), st[0], st[1], bad, o[0]);
      if (le != cudaSuccess) return 1;
    }
  return 0;
}
