// Hardware probe (developer tool): does a K-major SWIZZLE_128B UMMA operand descriptor whose start address is only
// 128-byte aligned and whose stride between 8-row groups (SBO) is NOT a multiple of 1024 B read rows with the swizzle
// phase of their ABSOLUTE shared-memory address (what TMA used when it wrote them)?  If yes, one TMA halo tile can feed
// all k x k taps of a convolution through shifted descriptors.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>
#include "../vp_suite_b200/csrc/ptx.cuh"
using namespace vpk;

constexpr int kRows = 512;           // smem rows (pixels) of 128 B
constexpr int kN = 16;

__global__ void probe(const __nv_bfloat16* a_rows /*[kRows][64]*/, const __nv_bfloat16* b /*[kN][64]*/, float* out,
                      int start_row, int sbo_bytes, int base_offset_mode, int nk) {
  extern __shared__ uint8_t raw[];
  const uint32_t ra = ptx::smem_u32(raw);
  uint8_t* smem = raw + (((ra + 1023u) & ~1023u) - ra);
  uint8_t* sa = smem;                       // kRows * 128 B
  uint8_t* sb = smem + kRows * 128;         // kN * 128 B (1024-aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 2048);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // address-based 128B swizzle, as TMA writes it: 16-byte chunk c of row r lands at chunk (c ^ (r & 7))
  for (int i = threadIdx.x; i < kRows * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) * 16)) = *reinterpret_cast<const uint4*>(a_rows + r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < kN * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) * 16)) = *reinterpret_cast<const uint4*>(b + r * 64 + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    if (ptx::elect_one()) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_barrier_init(); }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(slot), 32);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    if (ptx::elect_one()) {
      const uint32_t a_addr = ptx::smem_u32(sa) + start_row * 128;
      uint64_t ad = 0;
      ad |= static_cast<uint64_t>((a_addr >> 4) & 0x3FFF);
      ad |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
      ad |= static_cast<uint64_t>(1) << 46;
      if (base_offset_mode) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      ad |= static_cast<uint64_t>(2) << 61;
      const uint64_t bd = ptx::smem_desc_sw128(ptx::smem_u32(sb));
      const uint32_t idesc = ptx::idesc_bf16_f32(128, kN);
      for (int k = 0; k < nk; ++k) ptx::mma_bf16_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, k > 0);
      ptx::mma_commit(ptx::smem_u32(bar));
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(bar), 0);
  ptx::tc_fence_after();
  if (warp < 4) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + threadIdx.x % 32) * kN + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<__nv_bfloat16> ha(kRows * 64), hb(kN * 64);
  std::vector<float> fa(kRows * 64), fb(kN * 64);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = v; }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * kN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + 4096 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cfgs[][2] = {{0, 1024}, {8, 1024}, {1, 1024}, {3, 1024}, {0, 1280}, {11, 1280}, {21, 1280}, {0, 1536}, {26, 1536}, {0, 2048}, {5, 2048}};
  for (auto& c : cfgs)
    for (int mode = 0; mode < 2; ++mode)
      for (int nk = 1; nk <= 4; nk += 3) {
        const int start = c[0], sbo = c[1];
        cudaMemset(dout, 0, 128 * kN * 4);
        probe<<<1, 128, smem>>>(da, db, dout, start, sbo, mode, nk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("start %d sbo %d mode %d: CUDA error %s\n", start, sbo, mode, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * kN);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = start + (m / 8) * (sbo / 128) + (m % 8);   // the smem row a shifted halo view should read
          for (int n = 0; n < kN; ++n) {
            double ref = 0;
            for (int k = 0; k < nk * 16; ++k) ref += fa[row * 64 + k] * fb[n * 64 + k];
            maxerr = fmax(maxerr, fabs(ref - out[m * kN + n]));
          }
        }
        printf("start_row %2d sbo %4d base_offset_mode %d nk %d : max err %.4f %s\n", start, sbo, mode, nk, maxerr,
               maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
  return 0;
}
