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
#pragma once
#include "nsb_common.cuh"

namespace nsb {

// row bases inside a warp's activation area (all multiples of 8, required by the swizzle)
constexpr int R_E = 0;         // 96 rows: Fourier embedding (rows 93..95 are zero)
constexpr int R_C = 96;        // 64 rows: sampled grid features (fine: [fine | middle])
constexpr int R_HA = 160;      // forward-only kernels: ping
constexpr int R_HB = 192;      //                        pong
constexpr int R_S = 160;       // backward kernels: S1..S5 (5 x 32 rows): h_{i+1}, later g_{i+1}
constexpr int R_DU = 320;      // 32 rows
constexpr int R_DU3 = 352;     // 32 rows

struct LaneId { int lane, pg, og; int q[4]; };
__device__ __forceinline__ LaneId make_lane(int lane) {
  LaneId L; L.lane = lane; L.pg = lane >> 3; L.og = lane & 7;
#pragma unroll
  for (int x = 0; x < 4; x++) L.q[x] = ((L.pg ^ x) & 3) << 2;
  return L;
}

#define NSB_FMA4(ACC, J, A, WV)                      \
  ACC[0][J] = fmaf((A).x, (WV), ACC[0][J]);          \
  ACC[1][J] = fmaf((A).y, (WV), ACC[1][J]);          \
  ACC[2][J] = fmaf((A).z, (WV), ACC[2][J]);          \
  ACC[3][J] = fmaf((A).w, (WV), ACC[3][J]);

// acc[p][j] += sum_{k<K} A[k][4pg+p] * W[og+8j][k]        (W row-major [32][pitch])
template <int K>
__device__ __forceinline__ void gemm_t(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ W,
                                       const int pitch, const LaneId& L) {
  static_assert(K % 8 == 0, "K must be a multiple of 8");
  const float* w0 = W + L.og * pitch;
#pragma unroll 1
  for (int k = 0; k < K; k += 8) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; j++) w[j] = *reinterpret_cast<const float4*>(w0 + j * 8 * pitch + k + 4 * h);
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int rr = 4 * h + kk;                                  // row within the 8-row group
        const float4 a = *reinterpret_cast<const float4*>(A + (k + rr) * kRowF + L.q[(rr >> 1) & 3]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float wv = kk == 0 ? w[j].x : kk == 1 ? w[j].y : kk == 2 ? w[j].z : w[j].w;
          NSB_FMA4(acc, j, a, wv)
        }
      }
    }
  }
}

// acc[p][j] += sum_{k<K} A[k][4pg+p] * W[k][4ig+j]          (W points at column n0 of a [K][pitch] matrix)
template <int K>
__device__ __forceinline__ void gemm_n(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ W,
                                       const int pitch, const LaneId& L) {
  static_assert(K % 8 == 0, "K must be a multiple of 8");
  const float* w0 = W + 4 * L.og;
#pragma unroll 1
  for (int k = 0; k < K; k += 8) {
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
      const float4 a = *reinterpret_cast<const float4*>(A + (k + rr) * kRowF + L.q[(rr >> 1) & 3]);
      const float4 w = *reinterpret_cast<const float4*>(w0 + (k + rr) * pitch);
      NSB_FMA4(acc, 0, a, w.x)
      NSB_FMA4(acc, 1, a, w.y)
      NSB_FMA4(acc, 2, a, w.z)
      NSB_FMA4(acc, 3, a, w.w)
    }
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

// E rows <- sin(p @ B) for the 16 points of the chunk.  pf = this lane's point (lane&15) coordinates.
__device__ __forceinline__ void embed_chunk(float* __restrict__ act, const float* __restrict__ B, const float pf[3], int lane) {
  const int pt = lane & 15, half = lane >> 4;
#pragma unroll 4
  for (int it = 0; it < kEmbPad / 2; it++) {
    const int f = 2 * it + half;
    float v = 0.0f;
    if (f < kEmb) v = __sinf(reduce_2pi(fourier_arg(pf, B, f)));
    act[act_idx(R_E + f, pt)] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Forward.  KEEP=false: ping-pong HA/HB (forward-only kernel).  KEEP=true: h_{i+1} -> S_{i+1} and the relu
// masks are returned bit-packed (bit 4p+j of masks[i] <-> T-own element [p][j]).
// On return out[o] (o < NO) holds the decoder output of point (lane & 15) in every lane.
// ---------------------------------------------------------------------------------------------
template <int LV, bool KEEP>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ Wt, float* __restrict__ act, const LaneId& L,
                                            uint32_t (&masks)[5], float (&out)[4]) {
  using D = Dec<LV>;
  const float* first = act + (D::XYZ ? R_E : R_C) * kRowF;
  const float* crow = act + R_C * kRowF;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float b = Wt[D::o_b + i * 32 + L.og + 8 * j];
      acc[0][j] = b; acc[1][j] = b; acc[2][j] = b; acc[3][j] = b;
    }
    const int rin = KEEP ? (R_S + (i - 1) * 32) : ((i & 1) ? R_HA : R_HB);     // rows of h_i (i >= 1)
    const int rout = KEEP ? (R_S + i * 32) : ((i & 1) ? R_HB : R_HA);         // rows of h_{i+1}
    if (i == 0) gemm_t<D::FIRSTP>(acc, first, Wt + D::o_W0, D::PF, L);
    else if (i == 3) {
      gemm_t<D::FIRSTP>(acc, first, Wt + D::o_W3E, D::PF, L);
      gemm_t<32>(acc, act + rin * kRowF, Wt + D::o_W3H, D::PH, L);
    } else {
      const int ow = i == 1 ? D::o_W1 : i == 2 ? D::o_W2 : D::o_W4;
      gemm_t<32>(acc, act + rin * kRowF, Wt + ow, D::PH, L);
    }
    uint32_t m = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (acc[p][j] > 0.0f) m |= 1u << (4 * p + j); else acc[p][j] = 0.0f;
      }
    masks[i] = m;
    if (D::XYZ) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float b = Wt[D::o_bc + i * 32 + L.og + 8 * j];
        acc[0][j] += b; acc[1][j] += b; acc[2][j] += b; acc[3][j] += b;
      }
      gemm_t<D::CD>(acc, crow, Wt + D::o_WC + i * 32 * D::PC, D::PC, L);
    }
    store_tile_t(act + rout * kRowF, acc, L);
    __syncwarp();
  }
  // output layer: out[pt][o] = bo[o] + sum_k h5[k][pt] Wo[o][k]   (lane: pt = lane&15, k-half = lane>>4)
  const float* h5 = act + (KEEP ? (R_S + 4 * 32) : R_HA) * kRowF;
  const int pt = L.lane & 15, half = L.lane >> 4;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < 16; kk++) {
    const int k = half * 16 + kk;
    const float hv = h5[k * kRowF + swz(k, pt >> 2) + (pt & 3)];
#pragma unroll
    for (int o = 0; o < D::NO; o++) part[o] = fmaf(hv, Wt[D::o_WO + o * D::PH + k], part[o]);
  }
#pragma unroll
  for (int o = 0; o < 4; o++) {
    float v = 0.0f;
    if (o < D::NO) { v = part[o] + __shfl_xor_sync(0xffffffffu, part[o], 16); v += Wt[D::o_bo + o]; }
    out[o] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient helpers (only instantiated for decoders whose parameters are being optimised:
// the colour decoder in stage 'color', src/Mapper.py:339-341).  Gradients of one chunk are reduced
// over its 16 points in registers and added to a global image with the PACKED layout by 16-byte
// vector reductions; nsb_unpack_grads later folds that image into the canonical flat order.
// ---------------------------------------------------------------------------------------------
// dW[o][k] += sum_pt A[o][pt] * X[k][pt]   for o<32, k<K  (dst row-major [32][pitch])
template <int K>
__device__ __forceinline__ void wgrad_nt(float* __restrict__ dst, const int pitch, const float* __restrict__ A,
                                         const float* __restrict__ X, const LaneId& L) {
  static_assert(K % 16 == 0, "K must be a multiple of 16");
#pragma unroll 1
  for (int kb = 0; kb < K; kb += 16) {
    float acc[4][4];   // [j: o = og+8j][c: k = kb+4pg+c]
    zero_tile(acc);
#pragma unroll
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
// Backward for one chunk.  Preconditions: mlp_forward<LV,true> just ran (S1..S5 = h1..h5, masks set).
// g_out[o]: dL/d out[o] of point (lane & 15) (identical in both half-warps).
// pfq[p][a]: un-normalised f32 coordinates of this lane's four points 4pg+p (for the embedding chain).
// Results:  C rows <- dL/dc (CD rows);  dpe[p][a] (valid in lanes with og == 0): dL/dp through the
// Fourier embedding for points 4pg+p.  dWp: packed-layout gradient image in global memory (WGRAD only).
// ---------------------------------------------------------------------------------------------
template <int LV, bool WGRAD>
__device__ __forceinline__ void mlp_backward(const float* __restrict__ Wt, float* __restrict__ act, const LaneId& L,
                                             const uint32_t (&masks)[5], const float (&g_out)[4],
                                             const float (&pfq)[4][3], float (&dpe)[4][3], float* __restrict__ dWp) {
  using D = Dec<LV>;
  float* S = act + R_S * kRowF;
  float* DU = act + R_DU * kRowF;
  float* DU3 = act + R_DU3 * kRowF;
  float* C = act + R_C * kRowF;
  const float* first = act + (D::XYZ ? R_E : R_C) * kRowF;
  const int pt = L.lane & 15, half = L.lane >> 4;

  // ---- output layer: (WGRAD) dWo, dbo ; g5 = Wo^T g_out  -> S5 (overwrites h5)
  if (WGRAD) {
    // stash g_out as rows DU[o][pt] so the generic helpers can read it
    if (half == 0) {
#pragma unroll
      for (int o = 0; o < 4; o++) DU[act_idx(o, pt)] = o < D::NO ? g_out[o] : 0.0f;
    }
    __syncwarp();
    {  // dWo[o][k] (k = lane) ; dbo[o]
      const int k = L.lane;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 h = *reinterpret_cast<const float4*>(S + (4 * 32 + k) * kRowF + swz(k, q));
#pragma unroll
        for (int o = 0; o < D::NO; o++) {
          const float4 g = *reinterpret_cast<const float4*>(DU + o * kRowF + swz(o, q));
          s[o] += g.x * h.x + g.y * h.y + g.z * h.z + g.w * h.w;
          sb[o] += (g.x + g.y) + (g.z + g.w);
        }
      }
#pragma unroll
      for (int o = 0; o < D::NO; o++) {
        atomicAdd(dWp + D::o_WO + o * D::PH + k, s[o]);
        if (k == 0) atomicAdd(dWp + D::o_bo + o, sb[o]);
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
      for (int o = 0; o < D::NO; o++) v = fmaf(Wt[D::o_WO + o * D::PH + k], g_out[o], v);
      g5[kk] = v;
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) { const int k = half * 16 + kk; S[act_idx(4 * 32 + k, pt)] = g5[kk]; }
    __syncwarp();
  }

  // ---- hidden layers 4..0
#pragma unroll
  for (int i = 4; i >= 0; i--) {
    float* G = S + i * 32 * kRowF;                   // g_{i+1}
    float* du = (i == 3) ? DU3 : DU;
    if (WGRAD && D::XYZ) {                           // fc_c.i : dWc = G C^T, dbc = sum G
      wgrad_nt<D::CD>(dWp + D::o_WC + i * 32 * D::PC, D::PC, G, C, L);
      bgrad(dWp + D::o_bc + i * 32, G, L.lane);
    }
    {  // du_i = relu'(u_i) * g_{i+1}   (T-own lanes hold the masks)
      float t[4][4];
      load_tile_t(G, t, L);
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 4; j++) if (!((masks[i] >> (4 * p + j)) & 1u)) t[p][j] = 0.0f;
      store_tile_t(du, t, L);
    }
    __syncwarp();
    if (WGRAD) {                                     // pts_linears.i : dW = du x_i^T, db = sum du
      if (i == 0) wgrad_nt<D::FIRSTP>(dWp + D::o_W0, D::PF, du, first, L);
      else if (i == 3) {
        wgrad_nt<D::FIRSTP>(dWp + D::o_W3E, D::PF, du, first, L);
        wgrad_nt<32>(dWp + D::o_W3H, D::PH, du, S + 2 * 32 * kRowF, L);
      } else {
        const int ow = i == 1 ? D::o_W1 : i == 2 ? D::o_W2 : D::o_W4;
        wgrad_nt<32>(dWp + ow, D::PH, du, S + (i - 1) * 32 * kRowF, L);
      }
      bgrad(dWp + D::o_b + i * 32, du, L.lane);
      __syncwarp();
    }
    if (i >= 1) {                                    // g_i = W_i[:, hidden part]^T du_i  -> S_i (overwrites h_i)
      float acc[4][4];
      zero_tile(acc);
      const int ow = i == 1 ? D::o_W1 : i == 2 ? D::o_W2 : i == 3 ? D::o_W3H : D::o_W4;
      gemm_n<32>(acc, du, Wt + ow, D::PH, L);
      store_tile_n(S + (i - 1) * 32 * kRowF, acc, L);
      __syncwarp();
    }
  }

  // ---- dL/dc through fc_c: dc = sum_i Wc_i^T g_{i+1}  (one K=160 GEMM over S1..S5)  -> C rows
  if (D::XYZ) {
#pragma unroll
    for (int n0 = 0; n0 < D::CD; n0 += 32) {
      float acc[4][4];
      zero_tile(acc);
      gemm_n<160>(acc, S, Wt + D::o_WC + n0, D::PC, L);
      __syncwarp();
      store_tile_n(C + n0 * kRowF, acc, L);
    }
    __syncwarp();
  }

  // ---- gradient w.r.t. the first-layer / skip input: dfirst = W3E^T du_3 + W0^T du_0
#pragma unroll
  for (int p = 0; p < 4; p++) { dpe[p][0] = 0.f; dpe[p][1] = 0.f; dpe[p][2] = 0.f; }
#pragma unroll 1
  for (int n0 = 0; n0 < D::FIRSTP; n0 += 32) {
    float acc[4][4];
    zero_tile(acc);
    gemm_n<32>(acc, DU3, Wt + D::o_W3E + n0, D::PF, L);
    gemm_n<32>(acc, DU, Wt + D::o_W0 + n0, D::PF, L);
    if (D::XYZ) {
      const float* B = Wt + D::o_B;
      float dB[4][3];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int f = n0 + 4 * L.og + j;
        dB[j][0] = dB[j][1] = dB[j][2] = 0.0f;
        if (f < kEmb) {
          const float b0 = B[f], b1 = B[kEmbPad + f], b2 = B[2 * kEmbPad + f];
#pragma unroll
          for (int p = 0; p < 4; p++) {
            float x = pfq[p][0] * b0; x = fmaf(pfq[p][1], b1, x); x = fmaf(pfq[p][2], b2, x);
            const float dx = __cosf(reduce_2pi(x)) * acc[p][j];
            dpe[p][0] = fmaf(b0, dx, dpe[p][0]); dpe[p][1] = fmaf(b1, dx, dpe[p][1]); dpe[p][2] = fmaf(b2, dx, dpe[p][2]);
            if (WGRAD) { dB[j][0] = fmaf(pfq[p][0], dx, dB[j][0]); dB[j][1] = fmaf(pfq[p][1], dx, dB[j][1]); dB[j][2] = fmaf(pfq[p][2], dx, dB[j][2]); }
          }
        }
      }
      if (WGRAD) {     // reduce dB over the four point-groups (lanes differing in pg), one atomic per (a,f)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int a = 0; a < 3; a++) {
            float v = dB[j][a];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            const int f = n0 + 4 * L.og + j;
            if (L.pg == 0 && f < kEmb) atomicAdd(dWp + D::o_B + a * kEmbPad + f, v);
          }
      }
    } else {
      // coarse decoder: the first input IS the grid feature -> this is dL/dc
      __syncwarp();
      store_tile_n(C + n0 * kRowF, acc, L);
    }
  }
  if (D::XYZ) {        // reduce the embedding chain over the 8 feature-lanes of each point group
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
