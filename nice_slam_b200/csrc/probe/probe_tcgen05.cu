// probe_tcgen05.cu -- stand-alone check of the tcgen05 building blocks the tensor-core MLP will use:
//   * canonical K-major (no swizzle) shared-memory operand layout + UMMA descriptors
//   * kind::tf32 MMA, M=128 N=32 K=8, accumulator in TMEM, tcgen05.commit -> mbarrier
//   * tcgen05.ld 32x32b epilogue
//   * 3xTF32 split (hi/lo) accuracy versus plain TF32
// Computes D[128][32] = A[128][K] * W[32][K]^T for K = 32 in both modes.  Built by `make probe` (not part of libnsb.so).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: [row/8][k/4][row%8][k%4]  (core matrix = 8 rows x 16 B, contiguous)
__device__ __forceinline__ int canon_idx(int r, int k, int K) { return (r >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  return d;                               // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ float to_tf32(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D1,
                                                      float* __restrict__ D3, float* __restrict__ D4, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* a_hi = reinterpret_cast<float*>(smem);
  float* a_lo = a_hi + 128 * K;
  float* w_hi = a_lo + 128 * K;
  float* w_lo = w_hi + 32 * K;
  float* m_hi = w_lo + 32 * K;      // MN-major copy of the same B: [k/8][n/4][k%8][n%4]
  float* m_lo = m_hi + 32 * K;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, k = i % K; const float v = A[i], h = to_tf32(v);
    a_hi[canon_idx(r, k, K)] = h; a_lo[canon_idx(r, k, K)] = to_tf32(v - h); }
  for (int i = tid; i < 32 * K; i += 128) { const int r = i / K, k = i % K; const float v = W[i], h = to_tf32(v);
    w_hi[canon_idx(r, k, K)] = h; w_lo[canon_idx(r, k, K)] = to_tf32(v - h);
    const int mi = (k >> 3) * (32 / 4) * 32 + (r >> 2) * 32 + (k & 7) * 4 + (r & 3);      // n = r
    m_hi[mi] = h; m_lo[mi] = v - h; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: D=F32, A=B=TF32, K-major both, N=32, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t lbo = 128, sbo = (uint32_t)(K >> 2) * 128;
  uint32_t parity = 0;
  for (int mode = 0; mode < 3; mode++) {           // 0: plain TF32 ; 1: 3xTF32 ; 2: 3xTF32 with an MN-major B operand
    if (tid == 0) {
      const uint32_t dcol = tmem + (mode == 1 ? 32u : 0u);
      const uint32_t idesc_mn = idesc | (1u << 16);          // b_major = MN
      uint32_t acc = 0;
      for (int ks = 0; ks < K / 8; ks++) {
        const uint64_t ah = make_desc(a_hi + ks * 64, lbo, sbo), al = make_desc(a_lo + ks * 64, lbo, sbo);
        const uint64_t wh = make_desc(w_hi + ks * 64, lbo, sbo), wl = make_desc(w_lo + ks * 64, lbo, sbo);
        auto mma = [&](uint64_t da, uint64_t db) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          acc = 1;
        };
        if (mode == 2) {
          // MN-major: SBO = stride between 4-wide n groups (128 B), LBO = stride between 8-deep k groups ((N/4)*128 B)
          const uint64_t mh = make_desc(m_hi + ks * (32 / 4) * 32, (32 / 4) * 128, 128), ml = make_desc(m_lo + ks * (32 / 4) * 32, (32 / 4) * 128, 128);
          auto mma2 = [&](uint64_t da, uint64_t db) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dcol), "l"(da), "l"(db), "r"(idesc_mn), "r"(acc) : "memory");
            acc = 1;
          };
          mma2(al, mh); mma2(ah, ml); mma2(ah, mh);
        } else {
          if (mode) { mma(al, wh); mma(ah, wl); }
          mma(ah, wh);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // everyone waits for the MMAs of this mode
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(&mbar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (mode == 1 ? 32u : 0u);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* D = mode == 0 ? D1 : (mode == 1 ? D3 : D4);
    for (int j = 0; j < 32; j++) D[tid * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  const int K = 32;
  float *hA = new float[128 * K], *hW = new float[32 * K], *h1 = new float[128 * 32], *h3 = new float[128 * 32], *h4 = new float[128 * 32];
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f * 2.0f - 1.0f; };
  for (int i = 0; i < 128 * K; i++) hA[i] = rnd() * 3.0f;
  for (int i = 0; i < 32 * K; i++) hW[i] = rnd();
  float *dA, *dW, *d1, *d3, *d4;
  cudaMalloc(&dA, 128 * K * 4); cudaMalloc(&dW, 32 * K * 4); cudaMalloc(&d1, 128 * 32 * 4); cudaMalloc(&d3, 128 * 32 * 4); cudaMalloc(&d4, 128 * 32 * 4);
  cudaMemcpy(dA, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, 32 * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(d1, 0, 128 * 32 * 4); cudaMemset(d3, 0, 128 * 32 * 4);
  const size_t smem = (size_t)(2 * 128 * K + 4 * 32 * K) * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dW, d1, d3, d4, K);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h1, d1, 128 * 32 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h3, d3, 128 * 32 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h4, d4, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e3 = 0, e4 = 0, ef = 0, mx = 0;
  for (int r = 0; r < 128; r++) for (int n = 0; n < 32; n++) {
    double ref = 0; float f32 = 0.f;
    for (int k = 0; k < K; k++) { ref += (double)hA[r * K + k] * (double)hW[n * K + k]; f32 = fmaf(hA[r * K + k], hW[n * K + k], f32); }
    e1 = fmax(e1, fabs(h1[r * 32 + n] - ref)); e3 = fmax(e3, fabs(h3[r * 32 + n] - ref)); e4 = fmax(e4, fabs(h4[r * 32 + n] - ref)); ef = fmax(ef, fabs((double)f32 - ref)); mx = fmax(mx, fabs(ref));
  }
  printf("max|ref| %.4f   max abs err: tf32 %.3e   3xtf32 %.3e   3xtf32 MN-major B %.3e   fp32-fma %.3e\n", mx, e1, e3, e4, ef);
  printf("sample D3[5][7]=%f D1[5][7]=%f\n", h3[5 * 32 + 7], h1[5 * 32 + 7]);
  return (e == cudaSuccess && e3 < 1e-4 * mx) ? 0 : 1;
}
