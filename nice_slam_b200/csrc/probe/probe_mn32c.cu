// probe_mn32c.cu (extends probe_mn32b: adds the byte-address reading of Swizzle<2,5,2> and a K-major read of the same tile) -- can tcgen05.mma (kind::tf32) contract over the ROWS of two row-major [128 points x 32 features] shared-memory tiles
// (MN-major A and B operands), i.e. compute dW[m][n] = sum_p X[p][m] * Y[p][n] without transposing anything?
// CUTLASS (sm100_common.inl:92) says MN-major tf32 operands exist only with the SWIZZLE_128B_BASE32B layout (descriptor layout type 1,
// Swizzle<2,5,2> on bits = within every 128-byte row the float index m is stored at m ^ ((m >> 2) & 3)).  This probe tries that layout with a
// few (LBO, SBO) assignments, the un-permuted row-major tile, and the ordinary SWIZZLE_128B layout (16-byte chunk c of row p at c ^ (p & 7)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mn32b probe_mn32b.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const float* smem, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}

constexpr int NCOMBO = 16;
struct Combo { int perm; uint32_t layout, lbo, sbo, kstep_bytes; int M; int mode; const char* name; };   // mode 0: dW (A, B MN-major, K = points); 1: chain (A K-major under test, B canonical K-major weights, K = features)
__constant__ Combo g_combo[NCOMBO];

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, int only) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // three arrangements of each tile, 16 KB each, 1024-byte aligned: 0 = plain row-major, 1 = 32B-base permutation, 2 = 128B chunk swizzle
  float* xa[4]; float* ya[4];
  for (int a = 0; a < 4; a++) { xa[a] = reinterpret_cast<float*>(smem) + a * 8192; ya[a] = xa[a] + 4096; }
  float* wk = reinterpret_cast<float*>(smem) + 4 * 8192;      // canonical K-major no-swizzle [32 n x 32 f] = first 32 rows of Y
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 32; i += 128) {
    const int p = i >> 5, m = i & 31;
    const float x = X[i], y = Y[i];
    xa[0][p * 32 + m] = x; ya[0][p * 32 + m] = y;
    const int m1 = m ^ ((m >> 2) & 3);
    xa[1][p * 32 + m1] = x; ya[1][p * 32 + m1] = y;
    const int m2 = (((m >> 2) ^ (p & 7)) << 2) | (m & 3);
    xa[2][p * 32 + m2] = x; ya[2][p * 32 + m2] = y;
    const int m3 = ((((m >> 3) ^ (p & 3)) & 3) << 3) | (m & 7);      // byte-address Swizzle<2,5,2>: 32-byte chunk ^= row & 3
    xa[3][p * 32 + m3] = x; ya[3][p * 32 + m3] = y;
    if (p < 32) wk[((p >> 3) * 8 + (m >> 2)) * 32 + (p & 7) * 4 + (m & 3)] = y;
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
  uint32_t parity = 0;
  for (int combo = 0; combo < NCOMBO; combo++) {
    if (only >= 0 && combo != only) continue;
    const Combo c = g_combo[combo];
    if (tid == 0) {
      // D = F32, A = B = TF32, both MN-major (bits 15, 16), N = 32, M = c.M
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (c.mode == 0 ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
      const int nks = c.mode == 0 ? 16 : 4;
      for (int ks = 0; ks < nks; ks++) {                           // mode 0: K = 128 points, 8 per MMA; mode 1: K = 32 features
        const uint64_t da = make_desc(xa[c.perm] + ks * (c.kstep_bytes / 4), c.lbo, c.sbo, c.layout);
        const uint64_t db = c.mode == 0 ? make_desc(ya[c.perm] + ks * (c.kstep_bytes / 4), c.lbo, c.sbo, c.layout) : make_desc(wk + ks * 64, 128, 1024, 0);
        const uint32_t acc = ks ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
      long long t0 = clock64();
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000ll) { if (tid == 0) printf("combo %d: timeout\n", combo); __trap(); }
      }
    }
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

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  // perm: 0 plain, 1 32B-base permutation, 2 128B chunk swizzle.  Row pitch 128 B; 4-row k-atom = 512 B; 8-row k-atom = 1024 B; one MMA (K = 8) = 1024 B.
  Combo h[NCOMBO] = {
      {1, 1, 0, 512, 1024, 128, 0, "dW: in-row perm m^((m>>2)&3), type 1, LBO 0, SBO 512"},
      {3, 1, 0, 512, 1024, 128, 0, "dW: chunk32 ^= p&3,  type 1, LBO 0,    SBO 512"},
      {3, 1, 16384, 512, 1024, 128, 0, "dW: chunk32 ^= p&3,  type 1, LBO 16K,  SBO 512"},
      {3, 1, 512, 16384, 1024, 128, 0, "dW: chunk32 ^= p&3,  type 1, LBO 512,  SBO 16K"},
      {3, 1, 512, 0, 1024, 128, 0, "dW: chunk32 ^= p&3,  type 1, LBO 512,  SBO 0"},
      {3, 1, 0, 512, 1024, 64, 0, "dW: chunk32 ^= p&3,  type 1, LBO 0,    SBO 512, M = 64"},
      {0, 1, 0, 512, 1024, 128, 0, "dW: plain rows,      type 1, LBO 0,    SBO 512"},
      {2, 2, 0, 1024, 1024, 128, 0, "dW: 128B swizzle,    type 2, LBO 0,    SBO 1024"},
      {2, 2, 1024, 0, 1024, 128, 0, "dW: 128B swizzle,    type 2, LBO 1024, SBO 0"},
      {0, 0, 0, 512, 1024, 128, 0, "dW: plain rows,      type 0"},
      {3, 1, 0, 1024, 32, 128, 1, "chain: K-major A, chunk32 ^= p&3, type 1, SBO 1024, k-step 32 B"},
      {3, 1, 16, 1024, 32, 128, 1, "chain: K-major A, chunk32 ^= p&3, type 1, LBO 16, SBO 1024"},
      {3, 1, 0, 512, 32, 128, 1, "chain: K-major A, chunk32 ^= p&3, type 1, SBO 512"},
      {2, 2, 0, 1024, 32, 128, 1, "chain: K-major A, 128B swizzle, type 2, SBO 1024 (known-good form)"},
      {1, 1, 0, 1024, 32, 128, 1, "chain: K-major A, in-row perm, type 1, SBO 1024"},
      {0, 1, 0, 1024, 32, 128, 1, "chain: K-major A, plain rows, type 1, SBO 1024"},
  };
  cudaMemcpyToSymbol(g_combo, h, sizeof(h));
  float *hX = new float[128 * 32], *hY = new float[128 * 32], *hD = new float[NCOMBO * 128 * 32];
  uint32_t s = 4242u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(int)(((s >> 8) & 0xFF) - 128) / 16.0f; };      // exactly representable in tf32
  for (int i = 0; i < 128 * 32; i++) { hX[i] = rnd(); hY[i] = rnd(); }
  float *dX, *dY, *dD;
  cudaMalloc(&dX, 128 * 32 * 4); cudaMalloc(&dY, 128 * 32 * 4); cudaMalloc(&dD, NCOMBO * 128 * 32 * 4);
  cudaMemcpy(dX, hX, 128 * 32 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dY, hY, 128 * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, NCOMBO * 128 * 32 * 4);
  const size_t smem = 200 * 1024;      // slack: descriptors with large LBO / SBO may read past the tiles
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dX, dY, dD, only);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, NCOMBO * 128 * 32 * 4, cudaMemcpyDeviceToHost);
  static double ref[32][32];
  double mx = 0;
  for (int m = 0; m < 32; m++) for (int n = 0; n < 32; n++) {
    double r = 0; for (int p = 0; p < 128; p++) r += (double)hX[p * 32 + m] * (double)hY[p * 32 + n];
    ref[m][n] = r; mx = fmax(mx, fabs(r));
  }
  static double ref2[128][32];
  double mx2 = 0;
  for (int p = 0; p < 128; p++) for (int n = 0; n < 32; n++) { double r = 0; for (int f = 0; f < 32; f++) r += (double)hX[p * 32 + f] * (double)hY[n * 32 + f]; ref2[p][n] = r; mx2 = fmax(mx2, fabs(r)); }
  for (int c = 0; c < NCOMBO; c++) {
    if (only >= 0 && c != only) continue;
    if (h[c].mode == 1) {
      double e2 = 0; for (int p = 0; p < 128; p++) for (int n = 0; n < 32; n++) e2 = fmax(e2, fabs(hD[(c * 128 + p) * 32 + n] - ref2[p][n]));
      printf("combo %d  max abs err %.3e (max|ref| %.1f)  %s\n", c, e2, mx2, h[c].name);
      continue;
    }
    double err = 0, errT = 0;
    for (int m = 0; m < 32; m++) for (int n = 0; n < 32; n++) {
      err = fmax(err, fabs(hD[(c * 128 + m) * 32 + n] - ref[m][n]));
      errT = fmax(errT, fabs(hD[(c * 128 + n) * 32 + m] - ref[m][n]));
    }
    // where do rows 32..63 come from? (replica of rows 0..31 when the M-atoms alias)
    double rep = 0;
    for (int m = 0; m < 32; m++) for (int n = 0; n < 32; n++) rep = fmax(rep, fabs(hD[(c * 128 + 32 + m) * 32 + n] - ref[m][n]));
    printf("combo %d  max abs err %.3e (transposed %.3e, rows 32-63 vs ref %.3e; max|ref| %.1f)  %s\n", c, err, errT, rep, mx, h[c].name);
  }
  return 0;
}
