// probe_mnmajor.cu -- which shared-memory arrangement / descriptor assignment does tcgen05.mma (kind::tf32, no swizzle) accept for an
// MN-major B operand?  D[128 x 32] = A[128 x 32] * B^T with A K-major (known good) and B given MN-major in two arrangements x two
// (LBO, SBO) assignments.  Prints the max abs error of every combination against an fp64 host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mnmajor probe_mnmajor.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const float* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ int canon_idx(int r, int k, int K) { return ((r >> 3) * (K >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

constexpr int K = 32, N = 32, NCOMBO = 5;
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* a = reinterpret_cast<float*>(smem);        // K-major canonical [128 x K]
  float* wk = a + 128 * K;                           // K-major canonical [N x K]      (combo 0, sanity)
  float* wa = wk + N * K;                            // MN-major arrangement A: [k/8][n/4][k%8][n%4]
  float* wb = wa + N * K;                            // MN-major arrangement B: [n/4][k/8][k%8][n%4]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) a[canon_idx(i / K, i % K, K)] = A[i];
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K; const float v = W[i];
    wk[canon_idx(n, k, K)] = v;
    wa[(k >> 3) * (N / 4) * 32 + (n >> 2) * 32 + (k & 7) * 4 + (n & 3)] = v;
    wb[(n >> 2) * (K / 8) * 32 + (k >> 3) * 32 + (k & 7) * 4 + (n & 3)] = v;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc_k = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_mn = idesc_k | (1u << 16);
  uint32_t parity = 0;
  for (int combo = 0; combo < NCOMBO; combo++) {
    if (tid == 0) {
      for (int ks = 0; ks < K / 8; ks++) {
        const uint64_t da = make_desc(a + ks * 64, 128, (K >> 2) * 128);
        uint64_t db; uint32_t idesc = idesc_mn;
        const uint32_t kA = (N / 4) * 128, nA = 128;          // arrangement A strides (bytes): per 8 k, per 4 n
        const uint32_t kB = 128, nB = (K / 8) * 128;          // arrangement B
        switch (combo) {
          case 0: db = make_desc(wk + ks * 64, 128, (K >> 2) * 128); idesc = idesc_k; break;
          case 1: db = make_desc(wa + ks * (kA / 4), /*lbo*/ kA, /*sbo*/ nA); break;       // LBO = K stride, SBO = N stride
          case 2: db = make_desc(wa + ks * (kA / 4), /*lbo*/ nA, /*sbo*/ kA); break;       // swapped
          case 3: db = make_desc(wb + ks * (kB / 4), /*lbo*/ kB, /*sbo*/ nB); break;
          default: db = make_desc(wb + ks * (kB / 4), /*lbo*/ nB, /*sbo*/ kB); break;
        }
        const uint32_t acc = ks ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(&mbar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; j++) D[(combo * 128 + tid) * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  float *hA = new float[128 * K], *hW = new float[N * K], *hD = new float[NCOMBO * 128 * 32];
  uint32_t s = 777u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(int)(((s >> 8) & 0xFF) - 128) / 16.0f; };   // exactly representable in tf32
  for (int i = 0; i < 128 * K; i++) hA[i] = rnd();
  for (int i = 0; i < N * K; i++) hW[i] = rnd();
  float *dA, *dW, *dD;
  cudaMalloc(&dA, 128 * K * 4); cudaMalloc(&dW, N * K * 4); cudaMalloc(&dD, NCOMBO * 128 * 32 * 4);
  cudaMemcpy(dA, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, NCOMBO * 128 * 32 * 4);
  const size_t smem = (size_t)(128 * K + 3 * N * K) * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dW, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, NCOMBO * 128 * 32 * 4, cudaMemcpyDeviceToHost);
  const char* names[NCOMBO] = {"K-major (sanity)", "MN arrangement A [k/8][n/4][k%8][n%4], LBO=Kstride SBO=Nstride", "MN arrangement A, LBO=Nstride SBO=Kstride",
                               "MN arrangement B [n/4][k/8][k%8][n%4], LBO=Kstride SBO=Nstride", "MN arrangement B, LBO=Nstride SBO=Kstride"};
  for (int c = 0; c < NCOMBO; c++) {
    double err = 0, mx = 0;
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) {
      double ref = 0; for (int k = 0; k < K; k++) ref += (double)hA[r * K + k] * (double)hW[n * K + k];
      err = fmax(err, fabs(hD[(c * 128 + r) * 32 + n] - ref)); mx = fmax(mx, fabs(ref));
    }
    printf("combo %d  max abs err %.3e (max|ref| %.2f)  %s\n", c, err, mx, names[c]);
  }
  return 0;
}
