/*
 * ORACLE (test infrastructure, NOT product code): scalar C restatement of NICE-SLAM's per-iteration
 * render-and-backprop path, with the backward pass written out by hand instead of
 * relying on autograd.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED against the real reference -- tests/test_oracle_c.py compares every output and
 * gradient of this file with fixtures produced by the unmodified reference (tests/make_golden.py, which
 * imports /root/reference) and, in the build container, with the reference itself.
 *
 * Each function cites the reference lines (under /root/reference) it follows:
 *   ray_far_bb, sample_z        src/utils/Renderer.py:82-170     (dtype flow: near f32, far/z f64)
 *   point / in-bound mask       src/utils/Renderer.py:172-174, 43-46
 *   normalise                   src/common.py:269-284            (f64, then .float())
 *   trilinear gather / grads    src/conv_onet/models/decoder.py:168-175 -> torch F.grid_sample
 *                               (ATen/native/GridSampler.h:27-33 unnormalise, :58-60 clip, :66-82 clip grad)
 *   mlp forward/backward        src/conv_onet/models/decoder.py:177-203 (MLP), :262-274 (MLP_no_xyz)
 *   stage dispatch              src/conv_onet/models/decoder.py:312-342, src/utils/Renderer.py:57
 *   composite                   src/common.py:204-245 (occupancy branch)
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HID 32
#define EMB 93
#define MAXS 512

typedef struct { const float* data; int D, H, W; long long sc, sd, sh, sw; } nsbo_grid;

typedef struct {
  int stage, n_rays, n_samples, n_surface;
  double bound[6], coarse_bound[6];
  const float *rays_o, *rays_d, *gt_depth;
  const float* t_uniform;   /* f32[n_samples]  torch.linspace(0,1,n_samples)          */
  const double* t_surface;  /* f64[n_surface]  torch.linspace(0,1,n_surface).double() */
  nsbo_grid grid[4];
  const float* flat[4];     /* canonical flat decoder parameters (include/nice_slam_b200.h) */
} nsbo_inputs;

/* ------------------------------------------------------------------ decoder shapes / flat layout */
static int dec_xyz(int l) { return l != 0; }
static int dec_cdim(int l) { return l == 2 ? 64 : 32; }
static int dec_nout(int l) { return l == 3 ? 4 : 1; }
static int dec_in(int l, int i) {
  if (l == 0) return i == 3 ? 64 : 32;
  return i == 0 ? EMB : (i == 3 ? EMB + HID : HID);
}
/* kind: 0=B 1=W 2=b 3=Wc 4=bc 5=Wo 6=bo */
long long nsbo_flat_offset(int l, int kind, int layer) {
  long long off = 0;
  if (dec_xyz(l)) { if (kind == 0) return off; off += 3 * EMB; }
  for (int i = 0; i < 5; i++) {
    if (kind == 1 && layer == i) return off; off += (long long)HID * dec_in(l, i);
    if (kind == 2 && layer == i) return off; off += HID;
  }
  if (dec_xyz(l)) for (int i = 0; i < 5; i++) {
    if (kind == 3 && layer == i) return off; off += (long long)HID * dec_cdim(l);
    if (kind == 4 && layer == i) return off; off += HID;
  }
  if (kind == 5) return off; off += (long long)dec_nout(l) * HID;
  if (kind == 6) return off; off += dec_nout(l);
  return off; /* kind 7: total */
}
long long nsbo_flat_floats(int l) { return nsbo_flat_offset(l, 7, 0); }

/* ------------------------------------------------------------------ sampling (Renderer.py:82-170) */
static double nanmax(double a, double b) { return (a > b || a != a) ? a : b; }   /* torch.max propagates NaN */
static double nanmin(double a, double b) { return (a < b || a != a) ? a : b; }

static double ray_far_bb(const double* bound, const float* o, const float* d) {
  double far = 0; /* min over axes of max over (lo,hi) of (bound - o)/d, all in f64 (bound is f64) */
  for (int a = 0; a < 3; a++) {
    double t0 = (bound[2 * a] - (double)o[a]) / (double)d[a];
    double t1 = (bound[2 * a + 1] - (double)o[a]) / (double)d[a];
    double m = nanmax(t0, t1);
    far = a == 0 ? m : nanmin(far, m);
  }
  return far;
}

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* z[S] sorted ascending; returns S */
static int sample_z(const nsbo_inputs* in, int r, int has_gt, float gtmax, float gtmax12, double* z) {
  const float* o = in->rays_o + 3 * r; const float* d = in->rays_d + 3 * r;
  int ns = in->n_samples, nf = has_gt ? in->n_surface : 0;
  double far_bb = ray_far_bb(in->bound, o, d) + 0.01;
  float near; double far;
  if (has_gt) {
    float gt = in->gt_depth[r];
    near = gt * 0.01f;                                  /* f32 tensor * python scalar -> f32 */
    far = nanmax(far_bb, 0.0); far = nanmin(far, (double)gtmax12);   /* clamp(far_bb, 0, max(gt*1.2)) */
  } else { near = 0.01f; far = far_bb; }
  for (int i = 0; i < ns; i++) {
    float t = in->t_uniform[i];
    float a = near * (1.0f - t);                        /* f32 */
    z[i] = (double)a + far * (double)t;                 /* f64 (far is f64) */
  }
  if (nf > 0) {
    float gt = in->gt_depth[r];
    for (int j = 0; j < nf; j++) {
      double ts = in->t_surface[j];
      if (gt > 0) z[ns + j] = (double)(0.95f * gt) * (1.0 - ts) + (double)(1.05f * gt) * ts;
      else        z[ns + j] = 0.001 * (1.0 - ts) + (double)gtmax * ts;
    }
    qsort(z, ns + nf, sizeof(double), cmp_double);      /* torch.sort of the concatenation */
  }
  return ns + nf;
}

/* ------------------------------------------------------------------ trilinear (F.grid_sample) */
typedef struct { int i0[3]; float f[3]; float clipg[3]; } tri_t;   /* f = unnormalised clipped coord */

static void tri_setup(const nsbo_grid* g, const float xn[3], tri_t* t) {
  int size[3] = { g->W, g->H, g->D };
  for (int a = 0; a < 3; a++) {
    float u = ((xn[a] + 1.0f) / 2.0f) * (float)(size[a] - 1);       /* align_corners=True */
    float mx = (float)(size[a] - 1);
    if (u <= 0.0f) { t->clipg[a] = 0.0f; u = 0.0f; }
    else if (u >= mx) { t->clipg[a] = 0.0f; u = mx; }
    else t->clipg[a] = 1.0f;
    t->f[a] = u; t->i0[a] = (int)floorf(u);
  }
}
/* corner order tnw,tne,tsw,tse,bnw,bne,bsw,bse : k bit0 -> +x, bit1 -> +y, bit2 -> +z */
static void tri_weights(const tri_t* t, float w[8]) {
  float x0 = (float)t->i0[0], y0 = (float)t->i0[1], z0 = (float)t->i0[2];
  float wx[2] = { (x0 + 1.0f) - t->f[0], t->f[0] - x0 };
  float wy[2] = { (y0 + 1.0f) - t->f[1], t->f[1] - y0 };
  float wz[2] = { (z0 + 1.0f) - t->f[2], t->f[2] - z0 };
  for (int k = 0; k < 8; k++) w[k] = wx[k & 1] * wy[(k >> 1) & 1] * wz[(k >> 2) & 1];
}
static int tri_inside(const nsbo_grid* g, const tri_t* t, int k, long long* off) {
  int x = t->i0[0] + (k & 1), y = t->i0[1] + ((k >> 1) & 1), z = t->i0[2] + ((k >> 2) & 1);
  if (x < 0 || x >= g->W || y < 0 || y >= g->H || z < 0 || z >= g->D) return 0;
  *off = z * g->sd + y * g->sh + x * g->sw; return 1;
}
static void tri_gather(const nsbo_grid* g, const tri_t* t, float* c /*[32]*/) {
  float w[8]; tri_weights(t, w);
  for (int ch = 0; ch < 32; ch++) c[ch] = 0.0f;
  for (int k = 0; k < 8; k++) { long long off; if (!tri_inside(g, t, k, &off)) continue;
    for (int ch = 0; ch < 32; ch++) c[ch] += g->data[off + ch * g->sc] * w[k]; }
}
/* backward: scatter dc into dgrid (if non-NULL) and return d/d(normalised coord) in gx[3] */
static void tri_backward(const nsbo_grid* g, const tri_t* t, const float* dc, float* dgrid, float gx[3]) {
  float w[8]; tri_weights(t, w);
  float x0 = (float)t->i0[0], y0 = (float)t->i0[1], z0 = (float)t->i0[2];
  float wx[2] = { (x0 + 1.0f) - t->f[0], t->f[0] - x0 };
  float wy[2] = { (y0 + 1.0f) - t->f[1], t->f[1] - y0 };
  float wz[2] = { (z0 + 1.0f) - t->f[2], t->f[2] - z0 };
  float gi[3] = { 0, 0, 0 };
  for (int k = 0; k < 8; k++) { long long off; if (!tri_inside(g, t, k, &off)) continue;
    float dot = 0.0f;
    for (int ch = 0; ch < 32; ch++) {
      if (dgrid) {
#pragma omp atomic
        dgrid[off + ch * g->sc] += w[k] * dc[ch];
      }
      dot += g->data[off + ch * g->sc] * dc[ch];
    }
    int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
    gi[0] += (bx ? 1.0f : -1.0f) * wy[by] * wz[bz] * dot;
    gi[1] += (by ? 1.0f : -1.0f) * wx[bx] * wz[bz] * dot;
    gi[2] += (bz ? 1.0f : -1.0f) * wx[bx] * wy[by] * dot;
  }
  int size[3] = { g->W, g->H, g->D };
  for (int a = 0; a < 3; a++) gx[a] = t->clipg[a] * ((float)(size[a] - 1) / 2.0f) * gi[a];
}

/* ------------------------------------------------------------------ decoders */
typedef struct {            /* per-point activations kept for the backward pass */
  float c[64]; float e[EMB]; float xarg[EMB];
  float h[6][HID];          /* h[i] = input of layer i (i>=1), h[5] = input of output layer */
  unsigned char m[5][HID];  /* relu masks */
  float out[4];
} act_t;

static void mlp_forward(int l, const float* F, const float pf[3], act_t* A) {
  const int xyz = dec_xyz(l), cd = dec_cdim(l), no = dec_nout(l);
  if (xyz) { const float* B = F + nsbo_flat_offset(l, 0, 0);
    for (int j = 0; j < EMB; j++) {                       /* x = p @ B ; e = sin(x)  (decoder.py:26-30) */
      float x = pf[0] * B[j]; x = fmaf(pf[1], B[EMB + j], x); x = fmaf(pf[2], B[2 * EMB + j], x);
      A->xarg[j] = x; A->e[j] = sinf(x); } }
  const float* first = xyz ? A->e : A->c; const int nfirst = xyz ? EMB : 32;
  float x[EMB + HID];
  for (int i = 0; i < 5; i++) {
    int nin = dec_in(l, i);
    if (i == 0) memcpy(x, first, sizeof(float) * nfirst);
    else if (i == 3) { memcpy(x, first, sizeof(float) * nfirst); memcpy(x + nfirst, A->h[3], sizeof(float) * HID); }
    else memcpy(x, A->h[i], sizeof(float) * HID);
    const float* W = F + nsbo_flat_offset(l, 1, i); const float* b = F + nsbo_flat_offset(l, 2, i);
    for (int o = 0; o < HID; o++) {
      float u = b[o];
      for (int k = 0; k < nin; k++) u += W[o * nin + k] * x[k];
      A->m[i][o] = u > 0.0f; float r = u > 0.0f ? u : 0.0f;
      if (xyz) { const float* Wc = F + nsbo_flat_offset(l, 3, i); const float* bc = F + nsbo_flat_offset(l, 4, i);
        float s = bc[o]; for (int k = 0; k < cd; k++) s += Wc[o * cd + k] * A->c[k]; r += s; }
      A->h[i + 1][o] = r;
    }
  }
  const float* Wo = F + nsbo_flat_offset(l, 5, 0); const float* bo = F + nsbo_flat_offset(l, 6, 0);
  for (int o = 0; o < no; o++) { float u = bo[o]; for (int k = 0; k < HID; k++) u += Wo[o * HID + k] * A->h[5][k]; A->out[o] = u; }
}

/* g_out[n_out] -> dc[c_dim], dpf[3] (+= ), dF (flat weight grads, += under atomic) */
static void mlp_backward(int l, const float* F, const float pf[3], const act_t* A, const float* g_out,
                         float* dc, float dpf[3], float* dF) {
  const int xyz = dec_xyz(l), cd = dec_cdim(l), no = dec_nout(l);
  const int nfirst = xyz ? EMB : 32;
  const float* first = xyz ? A->e : A->c;
  float gh[HID], dfirst[EMB];
  for (int k = 0; k < nfirst; k++) dfirst[k] = 0.0f;
  for (int k = 0; k < cd; k++) dc[k] = 0.0f;
  const float* Wo = F + nsbo_flat_offset(l, 5, 0);
  for (int k = 0; k < HID; k++) { float s = 0; for (int o = 0; o < no; o++) s += Wo[o * HID + k] * g_out[o]; gh[k] = s; }
  if (dF) { float* dWo = dF + nsbo_flat_offset(l, 5, 0); float* dbo = dF + nsbo_flat_offset(l, 6, 0);
    for (int o = 0; o < no; o++) {
#pragma omp atomic
      dbo[o] += g_out[o];
      for (int k = 0; k < HID; k++) {
#pragma omp atomic
        dWo[o * HID + k] += g_out[o] * A->h[5][k]; } } }
  for (int i = 4; i >= 0; i--) {
    int nin = dec_in(l, i);
    float x[EMB + HID];
    if (i == 0) memcpy(x, first, sizeof(float) * nfirst);
    else if (i == 3) { memcpy(x, first, sizeof(float) * nfirst); memcpy(x + nfirst, A->h[3], sizeof(float) * HID); }
    else memcpy(x, A->h[i], sizeof(float) * HID);
    if (xyz) { const float* Wc = F + nsbo_flat_offset(l, 3, i);
      for (int k = 0; k < cd; k++) { float s = 0; for (int o = 0; o < HID; o++) s += Wc[o * cd + k] * gh[o]; dc[k] += s; }
      if (dF) { float* dWc = dF + nsbo_flat_offset(l, 3, i); float* dbc = dF + nsbo_flat_offset(l, 4, i);
        for (int o = 0; o < HID; o++) {
#pragma omp atomic
          dbc[o] += gh[o];
          for (int k = 0; k < cd; k++) {
#pragma omp atomic
            dWc[o * cd + k] += gh[o] * A->c[k]; } } } }
    float du[HID]; for (int o = 0; o < HID; o++) du[o] = A->m[i][o] ? gh[o] : 0.0f;
    const float* W = F + nsbo_flat_offset(l, 1, i);
    if (dF) { float* dW = dF + nsbo_flat_offset(l, 1, i); float* db = dF + nsbo_flat_offset(l, 2, i);
      for (int o = 0; o < HID; o++) {
#pragma omp atomic
        db[o] += du[o];
        for (int k = 0; k < nin; k++) {
#pragma omp atomic
          dW[o * nin + k] += du[o] * x[k]; } } }
    float dx[EMB + HID];
    for (int k = 0; k < nin; k++) { float s = 0; for (int o = 0; o < HID; o++) s += W[o * nin + k] * du[o]; dx[k] = s; }
    if (i == 0) { for (int k = 0; k < nfirst; k++) dfirst[k] += dx[k]; }
    else if (i == 3) { for (int k = 0; k < nfirst; k++) dfirst[k] += dx[k]; for (int k = 0; k < HID; k++) gh[k] = dx[nfirst + k]; }
    else for (int k = 0; k < HID; k++) gh[k] = dx[k];
  }
  if (xyz) { const float* B = F + nsbo_flat_offset(l, 0, 0); float* dB = dF ? dF + nsbo_flat_offset(l, 0, 0) : 0;
    for (int j = 0; j < EMB; j++) { float dxj = cosf(A->xarg[j]) * dfirst[j];
      for (int a = 0; a < 3; a++) { dpf[a] += B[a * EMB + j] * dxj;
        if (dB) {
#pragma omp atomic
          dB[a * EMB + j] += pf[a] * dxj; } } } }
  else for (int k = 0; k < 32; k++) dc[k] += dfirst[k];
}

/* ------------------------------------------------------------------ per-point evaluation */
typedef struct { double p[3]; float pf[3]; int inb; float xn[3], xnc[3]; } point_t;

static void make_point(const nsbo_inputs* in, int r, double z, point_t* P) {
  const float* o = in->rays_o + 3 * r; const float* d = in->rays_d + 3 * r;
  P->inb = 1;
  for (int a = 0; a < 3; a++) {
    double p = (double)o[a] + (double)d[a] * z;            /* Renderer.py:172-174, f64 */
    P->p[a] = p; P->pf[a] = (float)p;
    double lo = in->bound[2 * a], hi = in->bound[2 * a + 1];
    if (!(p < hi && p > lo)) P->inb = 0;                   /* strict, Renderer.py:43-46 */
    P->xn[a] = (float)(((p - lo) / (hi - lo)) * 2 - 1.0);  /* common.py:269-284 then .float() */
    double clo = in->coarse_bound[2 * a], chi = in->coarse_bound[2 * a + 1];
    P->xnc[a] = (float)(((p - clo) / (chi - clo)) * 2 - 1.0);
  }
}

/* stage -> decoders evaluated, in the reference's order (decoder.py:317-342) */
static int stage_decoders(int stage, int dl[3]) {
  switch (stage) { case 0: dl[0] = 0; return 1; case 1: dl[0] = 1; return 1;
    case 2: dl[0] = 2; dl[1] = 1; return 2; default: dl[0] = 2; dl[1] = 3; dl[2] = 1; return 3; }
}

static void eval_decoder(const nsbo_inputs* in, int l, const point_t* P, act_t* A, tri_t* T /*[2]*/) {
  const float* xn = l == 0 ? P->xnc : P->xn;
  tri_setup(&in->grid[l], xn, &T[0]); tri_gather(&in->grid[l], &T[0], A->c);
  if (l == 2) { tri_setup(&in->grid[1], xn, &T[1]); tri_gather(&in->grid[1], &T[1], A->c + 32); }  /* no_grad concat */
  mlp_forward(l, in->flat[l], P->pf, A);
}

static void batch_max(const nsbo_inputs* in, float* gmax, float* gmax12) {
  float m = 0, m12 = 0;
  for (int i = 0; i < in->n_rays; i++) { float g = in->gt_depth[i]; float g12 = g * 1.2f;
    if (i == 0 || g > m) m = g; if (i == 0 || g12 > m12) m12 = g12; }
  *gmax = m; *gmax12 = m12;
}

static float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

int nsbo_forward(const nsbo_inputs* in, double* depth, double* var, float* rgb,
                 double* z_vals, float* raw_out, float* weights_out, int* corner_idx) {
  int has_gt = in->gt_depth != 0 && in->stage != 0;
  int S = in->n_samples + (has_gt ? in->n_surface : 0);
  if (S > MAXS) return -1;
  float gmax = 0, gmax12 = 0; if (has_gt) batch_max(in, &gmax, &gmax12);
  int dl[3]; int nd = stage_decoders(in->stage, dl);
#pragma omp parallel for schedule(dynamic, 4)
  for (int r = 0; r < in->n_rays; r++) {
    double z[MAXS]; float raw[MAXS][4];
    sample_z(in, r, has_gt, gmax, gmax12, z);
    for (int s = 0; s < S; s++) {
      point_t P; make_point(in, r, z[s], &P);
      float occ = 0, col[3] = { 0, 0, 0 };
      for (int q = 0; q < nd; q++) { act_t A; tri_t T[2]; eval_decoder(in, dl[q], &P, &A, T);
        if (dl[q] == 3) { col[0] = A.out[0]; col[1] = A.out[1]; col[2] = A.out[2]; } else occ += A.out[0];
        if (q == 0 && corner_idx) for (int a = 0; a < 3; a++) corner_idx[((long long)r * S + s) * 3 + a] = T[0].i0[a]; }
      if (!P.inb) occ = 100.0f;                                          /* Renderer.py:57 */
      raw[s][0] = col[0]; raw[s][1] = col[1]; raw[s][2] = col[2]; raw[s][3] = occ;
    }
    /* raw2outputs_nerf_color, occupancy branch (common.py:233-244) */
    float T = 1.0f; float c3[3] = { 0, 0, 0 }; double dsum = 0; float w[MAXS];
    for (int s = 0; s < S; s++) { float al = sigmoidf(10.0f * raw[s][3]); w[s] = al * T; T = T * (1.0f - al + 1e-10f);
      for (int a = 0; a < 3; a++) c3[a] += w[s] * raw[s][a]; dsum += (double)w[s] * z[s]; }
    double v = 0; for (int s = 0; s < S; s++) { double t = z[s] - dsum; v += (double)w[s] * t * t; }
    depth[r] = dsum; var[r] = v; for (int a = 0; a < 3; a++) rgb[3 * r + a] = c3[a];
    for (int s = 0; s < S; s++) {
      if (z_vals) z_vals[(long long)r * S + s] = z[s];
      if (weights_out) weights_out[(long long)r * S + s] = w[s];
      if (raw_out) for (int a = 0; a < 4; a++) raw_out[((long long)r * S + s) * 4 + a] = raw[s][a]; }
  }
  return 0;
}

/* Backward of nsbo_forward (what loss.backward() does, Tracker.py:125 / Mapper.py:503).  Seeds g_depth[N]
 * (f64), g_var[N] (f64 or NULL), g_rgb[N,3] (f32 or NULL).  Outputs are ACCUMULATED into d_grid / d_flat
 * (caller zeroes) and written to d_rays_o / d_rays_d. */
int nsbo_backward(const nsbo_inputs* in, const double* g_depth, const double* g_var, const float* g_rgb,
                  float* d_rays_o, float* d_rays_d, float* const d_grid[4], float* const d_flat[4]) {
  int has_gt = in->gt_depth != 0 && in->stage != 0;
  int S = in->n_samples + (has_gt ? in->n_surface : 0);
  if (S > MAXS) return -1;
  float gmax = 0, gmax12 = 0; if (has_gt) batch_max(in, &gmax, &gmax12);
  int dl[3]; int nd = stage_decoders(in->stage, dl);
#pragma omp parallel for schedule(dynamic, 4)
  for (int r = 0; r < in->n_rays; r++) {
    double z[MAXS]; float raw[MAXS][4]; int inb[MAXS];
    sample_z(in, r, has_gt, gmax, gmax12, z);
    /* recompute forward raw (activations are recomputed again below, per point) */
    for (int s = 0; s < S; s++) { point_t P; make_point(in, r, z[s], &P); float occ = 0, col[3] = { 0, 0, 0 };
      for (int q = 0; q < nd; q++) { act_t A; tri_t T[2]; eval_decoder(in, dl[q], &P, &A, T);
        if (dl[q] == 3) { col[0] = A.out[0]; col[1] = A.out[1]; col[2] = A.out[2]; } else occ += A.out[0]; }
      inb[s] = P.inb; if (!P.inb) occ = 100.0f;
      raw[s][0] = col[0]; raw[s][1] = col[1]; raw[s][2] = col[2]; raw[s][3] = occ; }
    /* composite forward quantities */
    float al[MAXS], Tr[MAXS], w[MAXS]; float T = 1.0f; double D = 0;
    for (int s = 0; s < S; s++) { al[s] = sigmoidf(10.0f * raw[s][3]); Tr[s] = T; w[s] = al[s] * T; T = T * (1.0f - al[s] + 1e-10f); D += (double)w[s] * z[s]; }
    double gD = g_depth ? g_depth[r] : 0.0, gV = g_var ? g_var[r] : 0.0;
    float gC[3] = { 0, 0, 0 }; if (g_rgb) for (int a = 0; a < 3; a++) gC[a] = g_rgb[3 * r + a];
    double swt = 0; for (int s = 0; s < S; s++) swt += (double)w[s] * (z[s] - D);
    double gDe = gD + gV * (-2.0 * swt);                    /* var depends on depth through tmp = z - depth */
    float gw[MAXS];
    for (int s = 0; s < S; s++) { double t = z[s] - D;
      gw[s] = (float)(gDe * z[s] + gV * t * t) + gC[0] * raw[s][0] + gC[1] * raw[s][1] + gC[2] * raw[s][2]; }
    /* cumprod backward, division form: dL/dalpha_i = T_i g_w_i - (sum_{k>i} g_w_k w_k)/q_i */
    float gocc[MAXS]; float R = 0.0f;
    for (int s = S - 1; s >= 0; s--) { float q = 1.0f - al[s] + 1e-10f; float ga = Tr[s] * gw[s] - R / q; R += gw[s] * w[s];
      gocc[s] = inb[s] ? 10.0f * al[s] * (1.0f - al[s]) * ga : 0.0f; }      /* OOB logits were overwritten */
    double dro[3] = { 0, 0, 0 }, drd[3] = { 0, 0, 0 };
    for (int s = 0; s < S; s++) {
      point_t P; make_point(in, r, z[s], &P);
      double dp[3] = { 0, 0, 0 };
      for (int q = 0; q < nd; q++) { int l = dl[q]; act_t A; tri_t Tt[2]; eval_decoder(in, l, &P, &A, Tt);
        float g_out[4] = { 0, 0, 0, 0 };
        if (l == 3) { for (int a = 0; a < 3; a++) g_out[a] = w[s] * gC[a]; } else g_out[0] = gocc[s];
        float dc[64], dpf[3] = { 0, 0, 0 }, gx[3];
        mlp_backward(l, in->flat[l], P.pf, &A, g_out, dc, dpf, d_flat ? d_flat[l] : 0);
        tri_backward(&in->grid[l], &Tt[0], dc, d_grid ? d_grid[l] : 0, gx);
        const double* bb = l == 0 ? in->coarse_bound : in->bound;
        for (int a = 0; a < 3; a++) dp[a] += (double)dpf[a] + ((double)gx[a] * 2.0) / (bb[2 * a + 1] - bb[2 * a]);
      }
      for (int a = 0; a < 3; a++) { dro[a] += dp[a]; drd[a] += dp[a] * z[s]; }
    }
    for (int a = 0; a < 3; a++) { if (d_rays_o) d_rays_o[3 * r + a] = (float)dro[a]; if (d_rays_d) d_rays_d[3 * r + a] = (float)drd[a]; }
  }
  return 0;
}
