// Hardware experiment (B200): does a tcgen05 shared-memory descriptor with SWIZZLE_128B accept a start address that is
// NOT aligned to the 1024 B swizzle atom (row-shifted windows of one halo tile), and a stride between 8-row groups that
// is not a multiple of 1024 B?  This decides whether a 3x3 convolution can feed all nine taps from ONE halo tile in
// shared memory (see DESIGN.md "halo reuse").
//
//   mode 0: A K-major  (rows = pixels, 64 bf16 of K per 128 B row), M row m -> tile row r0 + (m>>3)*G + (m&7)
//   mode 1: A and B MN-major (rows = K index = pixels, 64 channels per 128 B row), k -> row r0 + (k>>3)*G + (k&7)
//   variant 0: descriptor base_offset field = 0 ; variant 1: base_offset = (start_address >> 7) & 7
// Data is written to shared memory with the TMA 128B-swizzle pattern (16 B chunk index ^= (address >> 7) & 7).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I causaldiffae_b200/csrc -o gpurun_out/exp_desc tools/exp_desc.cu
// usage: exp_desc mode r0 G variant   -> prints PASS / FAIL(max err)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

namespace cdae { void set_error(const char*, ...) {} }
using namespace cdae::sm100;

constexpr int kRows = 320;          // rows of 128 B in each operand tile

__device__ __forceinline__ uint32_t swz(uint32_t byte_addr) { return byte_addr ^ (((byte_addr >> 7) & 7u) << 4); }

struct Params { int mode, r0, G, variant; };

__global__ void __launch_bounds__(128) exp_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                  float* __restrict__ d, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: A block0 [kRows][128B] | A block1 | B block0 | bars
  constexpr int kBlk = kRows * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * kBlk);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t base = smem_u32(smem);
  // fill smem with the swizzled image of the operands; a: [2][kRows][64], b: [kRows][64]
  for (int i = threadIdx.x; i < 3 * kRows * 8; i += blockDim.x) {
    const int blk = i / (kRows * 8), row = (i / 8) % kRows, ch = i % 8;       // 16 B chunk
    const __nv_bfloat16* src = (blk < 2 ? a + ((size_t)blk * kRows + row) * 64 : b + (size_t)row * 64) + ch * 8;
    const uint32_t addr = swz(base + blk * kBlk + row * 128 + ch * 16);
    *reinterpret_cast<uint4*>(smem + (addr - base)) = *reinterpret_cast<const uint4*>(src);
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sbo = p.G * 128;
    if (p.mode == 0) {
      const uint32_t a_start = base + p.r0 * 128;
      const uint64_t bo = p.variant ? (uint64_t)((a_start >> 7) & 7) << 49 : 0;
      const uint64_t adesc = (uint64_t)((a_start >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
                             (2ull << 61) | bo;
      const uint64_t bdesc = smem_desc_kmajor_sw128(base + 2 * kBlk);
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    } else {
      const uint32_t a_start = base + p.r0 * 128, b_start = base + 2 * kBlk + p.r0 * 128;
      const uint64_t boa = p.variant ? (uint64_t)((a_start >> 7) & 7) << 49 : 0;
      const uint64_t bob = p.variant ? (uint64_t)((b_start >> 7) & 7) << 49 : 0;
      const uint64_t adesc = smem_desc_mnmajor_sw128(a_start, kBlk, sbo) | boa;
      const uint64_t bdesc = smem_desc_mnmajor_sw128(b_start, kBlk, sbo) | bob;
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t step = (2 * sbo) >> 4;        // K = 16 = two 8-row groups
      for (int k = 0; k < 4; ++k) umma_f16(tmem, adesc + (uint64_t)step * k, bdesc + (uint64_t)step * k, idesc, k != 0);
    }
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  {
    uint32_t acc[32];
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c = 0; c < 64; c += 32) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, acc);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) d[row * 64 + c + j] = __uint_as_float(acc[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main(int argc, char** argv) {
  if (argc < 5) { printf("usage: exp_desc mode r0 G variant\n"); return 2; }
  Params p{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
  std::vector<__nv_bfloat16> ha(2 * kRows * 64), hb(kRows * 64);
  std::vector<float> fa(ha.size()), fb(hb.size());
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 16) % 5) - 2); };
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = rnd(); ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = rnd(); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db; float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, 128 * 64 * 4);
  const int smem = 3 * kRows * 128 + 1024 + 64;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  exp_kernel<<<1, 128, smem>>>(da, db, dd, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d r0 %2d G %2d var %d : CUDA ERROR %s\n", p.mode, p.r0, p.G, p.variant, cudaGetErrorString(e)); return 1; }
  std::vector<float> hd(128 * 64);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double ref = 0;
      if (p.mode == 0) {
        const int r = p.r0 + (m >> 3) * p.G + (m & 7);
        for (int k = 0; k < 64; ++k) ref += fa[(size_t)r * 64 + k] * fb[(size_t)n * 64 + k];
      } else {
        for (int k = 0; k < 64; ++k) {
          const int r = p.r0 + (k >> 3) * p.G + (k & 7);
          ref += fa[((size_t)(m >> 6) * kRows + r) * 64 + (m & 63)] * fb[(size_t)r * 64 + n];
        }
      }
      const double err = fabs(ref - hd[m * 64 + n]);
      if (err > maxerr) maxerr = err;
    }
  printf("mode %d r0 %2d G %2d var %d : %s (max err %.1f)\n", p.mode, p.r0, p.G, p.variant, maxerr == 0 ? "PASS" : "FAIL", maxerr);
  return 0;
}
