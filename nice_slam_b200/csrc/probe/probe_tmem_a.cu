// probe_tmem_a.cu -- tcgen05.mma (kind::tf32) with the A operand in TENSOR MEMORY (written by tcgen05.st, one row per thread) and B in
// shared memory (K-major canonical, as in nsb_tc.cuh).  D[128 x 32] = A[128 x 32] * B[32 x 32]^T, checked against an fp64 host reference.
// Also times (clock64) the round trip  tcgen05.st A -> fence -> barrier -> 4 MMAs -> commit -> mbarrier wait -> tcgen05.ld D  against the
// shared-memory A path (st.shared -> fence.proxy.async -> barrier -> ...), 64 repetitions each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tmem_a probe_tmem_a.cu
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 8; j++) v[j] = __uint_as_float(r[j]);
}

constexpr int K = 32, N = 32, REPS = 64;
// 512 threads: row = tid & 127, column group cg = tid >> 7 (8 columns each) -- the mapping of nsb_tc.cuh
__global__ void __launch_bounds__(512, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D, long long* __restrict__ cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* a_s = reinterpret_cast<float*>(smem);       // K-major canonical [128 x K] (shared-memory A path)
  float* w_s = a_s + 128 * K;                         // K-major canonical [N x K]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, cg = tid >> 7;
  for (int i = tid; i < N * K; i += 512) w_s[canon_idx(i / K, i % K, K)] = W[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
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
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  float av[8];
  for (int j = 0; j < 8; j++) av[j] = A[row * K + 8 * cg + j];
  uint32_t parity = 0;
  for (int mode = 0; mode < 2; mode++) {             // 0: A in shared memory, 1: A in tensor memory (columns [64, 96))
    long long t_begin = 0;
    for (int rep = 0; rep < REPS + 1; rep++) {
      if (rep == 1) { __syncthreads(); t_begin = clock64(); }
      if (mode == 0) {
        for (int k = 0; k < 2; k++)
          *reinterpret_cast<float4*>(a_s + canon_idx(row, 8 * cg + 4 * k, K)) = make_float4(av[4 * k], av[4 * k + 1], av[4 * k + 2], av[4 * k + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      } else {
        tmem_st8(tmem + 64u + lane_base + 8u * cg, av);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < K / 8; ks++) {
          const uint64_t db = make_desc(w_s + ks * 64, 128, (K >> 2) * 128);
          const uint32_t acc = ks ? 1u : 0u;
          if (mode == 0) {
            const uint64_t da = make_desc(a_s + ks * 64, 128, (K >> 2) * 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          } else {
            const uint32_t ta = tmem + 64u + 8u * ks;          // A: lane = row, 8 fp32 columns per K step
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
      }
      __syncwarp();
      mbar_wait(&mbar, parity); parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float dv[8];
      tmem_ld8(tmem + lane_base + 8u * cg, dv);
      if (rep == REPS) for (int j = 0; j < 8; j++) D[(mode * 128 + row) * 32 + 8 * cg + j] = dv[j];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) cyc[mode] = (clock64() - t_begin) / REPS;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
  float *hA = new float[128 * K], *hW = new float[N * K], *hD = new float[2 * 128 * 32];
  uint32_t s = 4242u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(int)(((s >> 8) & 0xFF) - 128) / 16.0f; };   // exactly representable in tf32
  for (int i = 0; i < 128 * K; i++) hA[i] = rnd();
  for (int i = 0; i < N * K; i++) hW[i] = rnd();
  float *dA, *dW, *dD; long long* dC; long long hC[2] = {0, 0};
  cudaMalloc(&dA, 128 * K * 4); cudaMalloc(&dW, N * K * 4); cudaMalloc(&dD, 2 * 128 * 32 * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 2 * 128 * 32 * 4);
  const size_t smem = (size_t)(128 * K + N * K) * 4;
  probe_kernel<<<1, 512, smem>>>(dA, dW, dD, dC);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, 2 * 128 * 32 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hC, dC, 16, cudaMemcpyDeviceToHost);
  const char* names[2] = {"A in shared memory (st.shared + fence.proxy.async)", "A in tensor memory (tcgen05.st)"};
  for (int c = 0; c < 2; c++) {
    double err = 0, mx = 0;
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) {
      double ref = 0; for (int k = 0; k < K; k++) ref += (double)hA[r * K + k] * (double)hW[n * K + k];
      err = fmax(err, fabs(hD[(c * 128 + r) * 32 + n] - ref)); mx = fmax(mx, fabs(ref));
    }
    printf("mode %d  max abs err %.3e (max|ref| %.2f)  %lld cycles per write->MMA->read round trip   %s\n", c, err, mx, hC[c], names[c]);
  }
  return 0;
}
