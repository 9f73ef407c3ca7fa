// nsb_mlp.cuh -- warp-level evaluation (forward + hand-rolled backward) of NICE-SLAM's tiny decoders
// for one CHUNK of 16 sample points, with FP32 FMA register tiles (4 points x 4 features per lane).
//
// Reference semantics: MLP.forward (src/conv_onet/models/decoder.py:177-203) and MLP_no_xyz.forward
// (:262-274):  u_i = W_i x_i + b_i ; h_{i+1} = relu(u_i) + (Wc_i c + bc_i) ; x_3 = [first, h_3] ;
// out = Wo h_5 + bo, where first = sin(p @ B) (GaussianFourierFeatureTransform, :26-30) or, for the
// coarse decoder, the sampled grid feature c itself.  Backward = SURVEY.md section 8.1.
//
// Data layout: every activation is a ROW of 16 floats (one per point of the chunk) in the warp's private
// shared-memory area, swizzled (nsb_common.cuh: swz).  Weights are the CTA-shared packed image.
// Lane roles inside a warp: pg = lane>>3 owns points 4pg..4pg+3;
//   "T-own" tiles (forward GEMMs)  own features og+8j  (og = lane&7, j=0..3)
//   "N-own" tiles (backward GEMMs) own features 4ig+j  (ig = lane&7)
//
// Code-size discipline (round-1 ncu finding: the first version inlined one copy of every GEMM per layer and per
// decoder type -> 0.5 MB / 2.4 MB of SASS, and 'no instruction' was the top stall): the decoder is described by a
// RUNTIME descriptor (DecRT), layers run in real loops, and there is ONE software-pipelined body per GEMM kind.
#pragma once
#include "nsb_common.cuh"

namespace nsb {

// row bases inside a warp's activation area (all multiples of 8, required by the swizzle)
constexpr int R_E = 0;         // 32 rows: scratch block of the Fourier embedding (recomputed 32 features at a time)
constexpr int R_C = 32;        // 64 rows: sampled grid features (fine: [fine | middle])
constexpr int R_HA = 96;       // forward-only kernels: ping
constexpr int R_HB = 128;      //                        pong          -> 160 rows (10 KB) per warp
constexpr int R_S = 96;        // backward kernels: S1..S5 (5 x 32 rows): h_{i+1}, later g_{i+1}
constexpr int R_DU = 256;      // 32 rows
constexpr int R_DU3 = 288;     // 32 rows                               -> 320 rows (20 KB) per warp

struct LaneId { int lane, pg, og; int q[4]; };
__device__ __forceinline__ LaneId make_lane(int lane) {
  LaneId L; L.lane = lane; L.pg = lane >> 3; L.og = lane & 7;
#pragma unroll
  for (int x = 0; x < 4; x++) L.q[x] = ((L.pg ^ x) & 3) << 2;
  return L;
}

// runtime decoder descriptor (warp-uniform; filled from the compile-time layout Dec<LV>)
struct DecRT {
  int xyz, cd, no, firstp, pf, pc;
  int o_B, o_W0, o_W3E, o_WC, o_WO, o_b, o_bc, o_bo;
  int o_Wh[5];                 // hidden-part weights of layer i (i >= 1): W1, W2, W3H, W4
};
template <int LV>
__device__ __forceinline__ void fill_dec(DecRT& d) {
  using D = Dec<LV>;
  d.xyz = D::XYZ; d.cd = D::CD; d.no = D::NO; d.firstp = D::FIRSTP; d.pf = D::PF; d.pc = D::PC;
  d.o_B = D::o_B; d.o_W0 = D::o_W0; d.o_W3E = D::o_W3E; d.o_WC = D::o_WC; d.o_WO = D::o_WO;
  d.o_b = D::o_b; d.o_bc = D::o_bc; d.o_bo = D::o_bo;
  d.o_Wh[0] = 0; d.o_Wh[1] = D::o_W1; d.o_Wh[2] = D::o_W2; d.o_Wh[3] = D::o_W3H; d.o_Wh[4] = D::o_W4;
}
__device__ __forceinline__ DecRT make_dec(int lv) {
  DecRT d;
  switch (lv) { case 0: fill_dec<0>(d); break; case 1: fill_dec<1>(d); break; case 2: fill_dec<2>(d); break; default: fill_dec<3>(d); break; }
  return d;
}
__device__ __forceinline__ int dec_wh(const DecRT& d, int i) {      // register-friendly lookup (no local memory)
  return i == 1 ? d.o_Wh[1] : i == 2 ? d.o_Wh[2] : i == 3 ? d.o_Wh[3] : d.o_Wh[4];
}

#define NSB_FMA4(ACC, J, A, WV)                      \
  ACC[0][J] = fmaf((A).x, (WV), ACC[0][J]);          \
  ACC[1][J] = fmaf((A).y, (WV), ACC[1][J]);          \
  ACC[2][J] = fmaf((A).z, (WV), ACC[2][J]);          \
  ACC[3][J] = fmaf((A).w, (WV), ACC[3][J]);

struct Frag4 { float4 w[4]; float4 a[4]; };

// ---- gemm_t: acc[p][j] += sum_{k<K} A[k][4pg+p] * W[og+8j][k]     (W row-major [32][pitch]; K % 8 == 0)
__device__ __forceinline__ void gemm_t_load(Frag4& f, const float* __restrict__ A, const float* __restrict__ w0,
                                            const int pitch, const int k, const int qa, const int qb) {
#pragma unroll
  for (int j = 0; j < 4; j++) f.w[j] = *reinterpret_cast<const float4*>(w0 + j * 8 * pitch + k);
  f.a[0] = *reinterpret_cast<const float4*>(A + (k + 0) * kRowF + qa);
  f.a[1] = *reinterpret_cast<const float4*>(A + (k + 1) * kRowF + qa);
  f.a[2] = *reinterpret_cast<const float4*>(A + (k + 2) * kRowF + qb);
  f.a[3] = *reinterpret_cast<const float4*>(A + (k + 3) * kRowF + qb);
}
__device__ __forceinline__ void gemm_t_compute(float (&acc)[4][4], const Frag4& f) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    NSB_FMA4(acc, j, f.a[0], f.w[j].x)
    NSB_FMA4(acc, j, f.a[1], f.w[j].y)
    NSB_FMA4(acc, j, f.a[2], f.w[j].z)
    NSB_FMA4(acc, j, f.a[3], f.w[j].w)
  }
}
__device__ __forceinline__ void gemm_t(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ W,
                                       const int pitch, const int K, const LaneId& L) {
  const float* w0 = W + L.og * pitch;
  Frag4 fa, fb;
  gemm_t_load(fa, A, w0, pitch, 0, L.q[0], L.q[1]);
#pragma unroll 1
  for (int k = 0; k < K; k += 8) {
    gemm_t_load(fb, A, w0, pitch, k + 4, L.q[2], L.q[3]);          // rows k+4..k+7 (always inside: K % 8 == 0)
    gemm_t_compute(acc, fa);
    if (k + 8 < K) gemm_t_load(fa, A, w0, pitch, k + 8, L.q[0], L.q[1]);
    gemm_t_compute(acc, fb);
  }
}

// ---- gemm_n: acc[p][j] += sum_{k<K} A[k][4pg+p] * W[k][4ig+j]     (W points at column n0 of a [K][pitch] matrix)
struct FragN { float4 a[4]; float4 w[4]; };
__device__ __forceinline__ void gemm_n_load(FragN& f, const float* __restrict__ A, const float* __restrict__ w0,
                                            const int pitch, const int k, const int qa, const int qb) {
  f.a[0] = *reinterpret_cast<const float4*>(A + (k + 0) * kRowF + qa);
  f.a[1] = *reinterpret_cast<const float4*>(A + (k + 1) * kRowF + qa);
  f.a[2] = *reinterpret_cast<const float4*>(A + (k + 2) * kRowF + qb);
  f.a[3] = *reinterpret_cast<const float4*>(A + (k + 3) * kRowF + qb);
#pragma unroll
  for (int r = 0; r < 4; r++) f.w[r] = *reinterpret_cast<const float4*>(w0 + (k + r) * pitch);
}
__device__ __forceinline__ void gemm_n_compute(float (&acc)[4][4], const FragN& f) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    NSB_FMA4(acc, 0, f.a[r], f.w[r].x)
    NSB_FMA4(acc, 1, f.a[r], f.w[r].y)
    NSB_FMA4(acc, 2, f.a[r], f.w[r].z)
    NSB_FMA4(acc, 3, f.a[r], f.w[r].w)
  }
}
__device__ __forceinline__ void gemm_n(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ W,
                                       const int pitch, const int K, const LaneId& L) {
  const float* w0 = W + 4 * L.og;
  FragN fa, fb;
  gemm_n_load(fa, A, w0, pitch, 0, L.q[0], L.q[1]);
#pragma unroll 1
  for (int k = 0; k < K; k += 8) {
    gemm_n_load(fb, A, w0, pitch, k + 4, L.q[2], L.q[3]);
    gemm_n_compute(acc, fa);
    if (k + 8 < K) gemm_n_load(fa, A, w0, pitch, k + 8, L.q[0], L.q[1]);
    gemm_n_compute(acc, fb);
  }
}

__device__ __forceinline__ void zero_tile(float (&acc)[4][4]) {
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[p][j] = 0.0f;
}
// T-own tile -> rows og+8j ; N-own tile -> rows 4ig+j
__device__ __forceinline__ void store_tile_t(float* __restrict__ rows, const float (&acc)[4][4], const LaneId& L) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = L.og + 8 * j;
    *reinterpret_cast<float4*>(rows + n * kRowF + swz(n, L.pg)) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
  }
}
__device__ __forceinline__ void store_tile_n(float* __restrict__ rows, const float (&acc)[4][4], const LaneId& L) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = 4 * L.og + j;
    *reinterpret_cast<float4*>(rows + n * kRowF + swz(n, L.pg)) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
  }
}
__device__ __forceinline__ void load_tile_t(const float* __restrict__ rows, float (&acc)[4][4], const LaneId& L) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = L.og + 8 * j;
    const float4 v = *reinterpret_cast<const float4*>(rows + n * kRowF + swz(n, L.pg));
    acc[0][j] = v.x; acc[1][j] = v.y; acc[2][j] = v.z; acc[3][j] = v.w;
  }
}

// Range-reduced sine / cosine of the Fourier argument (|x| up to ~1e3): r = x - round(x/2pi)*2pi in two
// FMA steps, then the MUFU approximations on [-pi,pi] (abs err ~4e-7, far inside the 1e-4 parity budget).
__device__ __forceinline__ float reduce_2pi(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);
  return fmaf(-k, -1.7484555314695172e-7f, r);
}
__device__ __forceinline__ float fourier_arg(const float pf[3], const float* __restrict__ B, int f) {
  float x = pf[0] * B[f];                         // same association as the oracle: ((p0*b0) + p1*b1) + p2*b2
  x = fmaf(pf[1], B[kEmbPad + f], x);
  return fmaf(pf[2], B[2 * kEmbPad + f], x);
}

// Scratch rows [R_E, R_E+32) <- features [32*blk, 32*blk+32) of sin(p @ B) for the 16 points of the chunk (features
// >= 93 are zero).  pf = coordinates of this lane's point (lane & 15).  The embedding is recomputed block by block
// where it is consumed (layers 0 and 3, and their weight gradients) instead of living in 96 rows of shared memory:
// ~600 extra instructions per chunk buy 4 KB per warp, i.e. more resident warps per SM.
__device__ __forceinline__ void embed_block(float* __restrict__ act, const float* __restrict__ B, const float pf[3], int lane, int blk) {
  const int pt = lane & 15, half = lane >> 4;
#pragma unroll 4
  for (int it = 0; it < 16; it++) {
    const int fl = 2 * it + half, f = 32 * blk + fl;
    float v = 0.0f;
    if (f < kEmb) v = __sinf(reduce_2pi(fourier_arg(pf, B, f)));
    act[act_idx(R_E + fl, pt)] = v;
  }
}
// acc += first-input part of layer 0 / 3:  W[:, first] * first   (first = Fourier embedding, or the grid feature for coarse)
__device__ __forceinline__ void gemm_first(float (&acc)[4][4], const DecRT& d, float* __restrict__ act, const float* __restrict__ Wt,
                                           const int o_w, const float pf[3], const LaneId& L) {
  if (d.xyz) {
#pragma unroll 1
    for (int blk = 0; blk < 3; blk++) {
      embed_block(act, Wt + d.o_B, pf, L.lane, blk);
      __syncwarp();
      gemm_t(acc, act + R_E * kRowF, Wt + o_w + 32 * blk, d.pf, 32, L);
      __syncwarp();
    }
  } else gemm_t(acc, act + R_C * kRowF, Wt + o_w, d.pf, 32, L);
}

// relu masks of the five layers, 16 bits each (bit 4p+j <-> T-own element [p][j]), packed in three registers
struct Masks { uint32_t m01, m23, m4; };
__device__ __forceinline__ void set_mask(Masks& M, int i, uint32_t m) {
  if (i == 0) M.m01 = m; else if (i == 1) M.m01 |= m << 16; else if (i == 2) M.m23 = m; else if (i == 3) M.m23 |= m << 16; else M.m4 = m;
}
__device__ __forceinline__ uint32_t get_mask(const Masks& M, int i) {
  const uint32_t w = i < 2 ? M.m01 : (i < 4 ? M.m23 : M.m4);
  return (w >> ((i & 1) * 16)) & 0xffffu;
}

// ---------------------------------------------------------------------------------------------
// Forward.  keep=false: ping-pong HA/HB (forward-only kernel).  keep=true: h_{i+1} -> S_{i+1} and the relu
// masks are returned.  On return out[o] (o < NO) holds the decoder output of point (lane & 15) in every lane.
// ---------------------------------------------------------------------------------------------
template <bool KEEP>
__device__ __forceinline__ void mlp_forward(const DecRT& d, const float* __restrict__ Wt, float* __restrict__ act, const LaneId& L,
                                            const float pf[3], Masks& masks, float (&out)[4]) {
  const float* crow = act + R_C * kRowF;
  masks.m01 = masks.m23 = masks.m4 = 0u;
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float b = Wt[d.o_b + i * 32 + L.og + 8 * j];
      acc[0][j] = b; acc[1][j] = b; acc[2][j] = b; acc[3][j] = b;
    }
    const int rin = KEEP ? (R_S + (i - 1) * 32) : ((i & 1) ? R_HA : R_HB);     // rows of h_i (i >= 1)
    const int rout = KEEP ? (R_S + i * 32) : ((i & 1) ? R_HB : R_HA);         // rows of h_{i+1}
    if (i == 0 || i == 3) gemm_first(acc, d, act, Wt, i == 0 ? d.o_W0 : d.o_W3E, pf, L);
    if (i >= 1) gemm_t(acc, act + rin * kRowF, Wt + dec_wh(d, i), Dec<1>::PH, 32, L);
    uint32_t m = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (acc[p][j] > 0.0f) m |= 1u << (4 * p + j); else acc[p][j] = 0.0f;
      }
    set_mask(masks, i, m);
    if (d.xyz) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float b = Wt[d.o_bc + i * 32 + L.og + 8 * j];
        acc[0][j] += b; acc[1][j] += b; acc[2][j] += b; acc[3][j] += b;
      }
      gemm_t(acc, crow, Wt + d.o_WC + i * 32 * d.pc, d.pc, d.cd, L);
    }
    store_tile_t(act + rout * kRowF, acc, L);
    __syncwarp();
  }
  // output layer: out[pt][o] = bo[o] + sum_k h5[k][pt] Wo[o][k]   (lane: pt = lane&15, k-half = lane>>4)
  const float* h5 = act + (KEEP ? (R_S + 4 * 32) : R_HA) * kRowF;
  const int pt = L.lane & 15, half = L.lane >> 4;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int kk = 0; kk < 16; kk++) {
    const int k = half * 16 + kk;
    const float hv = h5[k * kRowF + swz(k, pt >> 2) + (pt & 3)];
#pragma unroll
    for (int o = 0; o < 4; o++) part[o] = fmaf(hv, Wt[d.o_WO + o * Dec<1>::PH + k], part[o]);      // rows >= NO are zero
  }
#pragma unroll
  for (int o = 0; o < 4; o++) {
    float v = part[o] + __shfl_xor_sync(0xffffffffu, part[o], 16);
    v += Wt[d.o_bo + o];                                                                          // zero for o >= NO
    out[o] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient helpers (only used for decoders whose parameters are being optimised: the colour decoder in
// stage 'color', src/Mapper.py:339-341).  Gradients of one chunk are reduced over its 16 points in registers and
// added to a global image with the PACKED layout by 16-byte vector reductions; the unpack kernel later folds that
// image into the canonical flat order.
// ---------------------------------------------------------------------------------------------
// dW[o][k] += sum_pt A[o][pt] * X[k][pt]   for o<32, k<K  (dst row-major [32][pitch]; K % 16 == 0)
__device__ __forceinline__ void wgrad_nt(float* __restrict__ dst, const int pitch, const float* __restrict__ A,
                                         const float* __restrict__ X, const int K, const LaneId& L) {
#pragma unroll 1
  for (int kb = 0; kb < K; kb += 16) {
    float acc[4][4];   // [j: o = og+8j][c: k = kb+4pg+c]
    zero_tile(acc);
#pragma unroll 2
    for (int q = 0; q < 4; q++) {
      float4 a[4], x[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { const int r = L.og + 8 * j; a[j] = *reinterpret_cast<const float4*>(A + r * kRowF + swz(r, q)); }
#pragma unroll
      for (int c = 0; c < 4; c++) { const int r = kb + 4 * L.pg + c; x[c] = *reinterpret_cast<const float4*>(X + r * kRowF + swz(r, q)); }
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          acc[j][c] = fmaf(a[j].x, x[c].x, acc[j][c]); acc[j][c] = fmaf(a[j].y, x[c].y, acc[j][c]);
          acc[j][c] = fmaf(a[j].z, x[c].z, acc[j][c]); acc[j][c] = fmaf(a[j].w, x[c].w, acc[j][c]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      red_add_v4(dst + (L.og + 8 * j) * pitch + kb + 4 * L.pg, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  }
}
// db[o] += sum_pt A[o][pt]  (o = lane)
__device__ __forceinline__ void bgrad(float* __restrict__ dst, const float* __restrict__ A, int lane) {
  float s = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; q++) { const float4 v = *reinterpret_cast<const float4*>(A + lane * kRowF + 4 * q); s += (v.x + v.y) + (v.z + v.w); }
  atomicAdd(dst + lane, s);
}

// ---------------------------------------------------------------------------------------------
// Backward for one chunk.  Preconditions: mlp_forward<true> just ran (S1..S5 = h1..h5, masks set).
// g_out[o]: dL/d out[o] of point (lane & 15) (identical in both half-warps).
// pfq[p][a]: un-normalised f32 coordinates of this lane's four points 4pg+p (for the embedding chain).
// Results:  C rows <- dL/dc (CD rows);  dpe[p][a]: dL/dp through the Fourier embedding for points 4pg+p
// (complete in every lane after the in-function reduction).  dWp: packed-layout gradient image in global
// memory, or nullptr when this decoder's parameters get no gradient.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mlp_backward(const DecRT& d, const float* __restrict__ Wt, float* __restrict__ act, const LaneId& L,
                                             const float pf[3], const Masks& masks, const float (&g_out)[4],
                                             const float (&pfq)[4][3], float (&dpe)[4][3], float* __restrict__ dWp) {
  float* S = act + R_S * kRowF;
  float* DU = act + R_DU * kRowF;
  float* DU3 = act + R_DU3 * kRowF;
  float* C = act + R_C * kRowF;
  const int pt = L.lane & 15, half = L.lane >> 4;
  const bool wgrad = dWp != nullptr;
  constexpr int PH = Dec<1>::PH;

  // ---- output layer: (wgrad) dWo, dbo ; g5 = Wo^T g_out  -> S5 (overwrites h5)
  if (wgrad) {
    if (half == 0) {      // stash g_out as rows DU[o][pt] so the helpers can read it
#pragma unroll
      for (int o = 0; o < 4; o++) DU[act_idx(o, pt)] = g_out[o];
    }
    __syncwarp();
    {  // dWo[o][k] (k = lane) ; dbo[o]
      const int k = L.lane;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 h = *reinterpret_cast<const float4*>(S + (4 * 32 + k) * kRowF + swz(k, q));
#pragma unroll
        for (int o = 0; o < 4; o++) {
          const float4 g = *reinterpret_cast<const float4*>(DU + o * kRowF + swz(o, q));
          s[o] += g.x * h.x + g.y * h.y + g.z * h.z + g.w * h.w;
          sb[o] += (g.x + g.y) + (g.z + g.w);
        }
      }
      for (int o = 0; o < d.no; o++) {
        atomicAdd(dWp + d.o_WO + o * PH + k, s[o]);
        if (k == 0) atomicAdd(dWp + d.o_bo + o, sb[o]);
      }
    }
    __syncwarp();
  }
  {
    float g5[16];
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      const int k = half * 16 + kk;
      float v = 0.0f;
#pragma unroll
      for (int o = 0; o < 4; o++) v = fmaf(Wt[d.o_WO + o * PH + k], g_out[o], v);     // g_out[o >= NO] == 0, rows >= NO are zero
      g5[kk] = v;
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) { const int k = half * 16 + kk; S[act_idx(4 * 32 + k, pt)] = g5[kk]; }
    __syncwarp();
  }

  // ---- hidden layers 4..0
#pragma unroll 1
  for (int i = 4; i >= 0; i--) {
    float* G = S + i * 32 * kRowF;                   // g_{i+1}
    float* du = (i == 3) ? DU3 : DU;
    if (wgrad && d.xyz) {                            // fc_c.i : dWc = G C^T, dbc = sum G
      wgrad_nt(dWp + d.o_WC + i * 32 * d.pc, d.pc, G, C, d.cd, L);
      bgrad(dWp + d.o_bc + i * 32, G, L.lane);
    }
    {  // du_i = relu'(u_i) * g_{i+1}   (T-own lanes hold the masks)
      float t[4][4];
      load_tile_t(G, t, L);
      const uint32_t m = get_mask(masks, i);
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 4; j++) if (!((m >> (4 * p + j)) & 1u)) t[p][j] = 0.0f;
      store_tile_t(du, t, L);
    }
    __syncwarp();
    if (wgrad) {                                     // pts_linears.i : dW = du x_i^T, db = sum du
      if (i == 0 || i == 3) {
        float* dst = dWp + (i == 0 ? d.o_W0 : d.o_W3E);
        if (d.xyz) {
#pragma unroll 1
          for (int blk = 0; blk < 3; blk++) {           // recompute the embedding block as the X operand
            embed_block(act, Wt + d.o_B, pf, L.lane, blk);
            __syncwarp();
            wgrad_nt(dst + 32 * blk, d.pf, du, act + R_E * kRowF, 32, L);
            __syncwarp();
          }
        } else wgrad_nt(dst, d.pf, du, C, 32, L);
      }
      if (i >= 1) wgrad_nt(dWp + dec_wh(d, i), PH, du, S + (i - 1) * 32 * kRowF, 32, L);
      bgrad(dWp + d.o_b + i * 32, du, L.lane);
      __syncwarp();
    }
    if (i >= 1) {                                    // g_i = W_i[:, hidden part]^T du_i  -> S_i (overwrites h_i)
      float acc[4][4];
      zero_tile(acc);
      gemm_n(acc, du, Wt + dec_wh(d, i), PH, 32, L);
      store_tile_n(S + (i - 1) * 32 * kRowF, acc, L);
      __syncwarp();
    }
  }

  // ---- dL/dc through fc_c: dc = sum_i Wc_i^T g_{i+1}  (one K=160 GEMM over S1..S5)  -> C rows
  if (d.xyz) {
#pragma unroll 1
    for (int n0 = 0; n0 < d.cd; n0 += 32) {
      float acc[4][4];
      zero_tile(acc);
      gemm_n(acc, S, Wt + d.o_WC + n0, d.pc, 160, L);
      __syncwarp();
      store_tile_n(C + n0 * kRowF, acc, L);
    }
    __syncwarp();
  }

  // ---- gradient w.r.t. the first-layer / skip input: dfirst = W3E^T du_3 + W0^T du_0
#pragma unroll
  for (int p = 0; p < 4; p++) { dpe[p][0] = 0.f; dpe[p][1] = 0.f; dpe[p][2] = 0.f; }
#pragma unroll 1
  for (int n0 = 0; n0 < d.firstp; n0 += 32) {
    float acc[4][4];
    zero_tile(acc);
    gemm_n(acc, DU3, Wt + d.o_W3E + n0, d.pf, 32, L);
    gemm_n(acc, DU, Wt + d.o_W0 + n0, d.pf, 32, L);
    if (d.xyz) {
      const float* B = Wt + d.o_B;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int f = n0 + 4 * L.og + j;
        float dB0 = 0.f, dB1 = 0.f, dB2 = 0.f;
        if (f < kEmb) {
          const float b0 = B[f], b1 = B[kEmbPad + f], b2 = B[2 * kEmbPad + f];
#pragma unroll
          for (int p = 0; p < 4; p++) {
            float x = pfq[p][0] * b0; x = fmaf(pfq[p][1], b1, x); x = fmaf(pfq[p][2], b2, x);
            const float dx = __cosf(reduce_2pi(x)) * acc[p][j];
            dpe[p][0] = fmaf(b0, dx, dpe[p][0]); dpe[p][1] = fmaf(b1, dx, dpe[p][1]); dpe[p][2] = fmaf(b2, dx, dpe[p][2]);
            dB0 = fmaf(pfq[p][0], dx, dB0); dB1 = fmaf(pfq[p][1], dx, dB1); dB2 = fmaf(pfq[p][2], dx, dB2);
          }
        }
        if (wgrad) {     // reduce dB over the four point-groups (lanes differing in pg), one atomic per (a,f)
          dB0 += __shfl_xor_sync(0xffffffffu, dB0, 8); dB0 += __shfl_xor_sync(0xffffffffu, dB0, 16);
          dB1 += __shfl_xor_sync(0xffffffffu, dB1, 8); dB1 += __shfl_xor_sync(0xffffffffu, dB1, 16);
          dB2 += __shfl_xor_sync(0xffffffffu, dB2, 8); dB2 += __shfl_xor_sync(0xffffffffu, dB2, 16);
          if (L.pg == 0 && f < kEmb) {
            atomicAdd(dWp + d.o_B + f, dB0); atomicAdd(dWp + d.o_B + kEmbPad + f, dB1); atomicAdd(dWp + d.o_B + 2 * kEmbPad + f, dB2);
          }
        }
      }
    } else {
      // coarse decoder: the first input IS the grid feature -> this is dL/dc
      __syncwarp();
      store_tile_n(C + n0 * kRowF, acc, L);
    }
  }
  if (d.xyz) {        // reduce the embedding chain over the 8 feature-lanes of each point group
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        float v = dpe[p][a];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        dpe[p][a] = v;
      }
  }
  __syncwarp();
}

}  // namespace nsb
