// K6 — PDE residual of sampled fields and its gradient (the guidance term of the sampler).
//
// Replaces, with one launch each, the torch expression graphs of
//   SweFvLoss.calculate_loss / f_t_swp1d     models/pde_loss.py:129-165, :199-215   (≈45 elementwise launches)
//   SweFvLoss.forward(return_d=True)          models/pde_loss.py:231-242             (autograd over the same graph)
//   DarcyLoss.calculate_loss / forward        models/pde_loss.py:30-56, :81-84
// including the casts and the inverse normalisation of their callers (PlMcedm.get_pde_loss models/mcedm.py:468-499,
// PlCondDdim.get_pde_loss / get_dx_pde models/ddim.py:1388-1450, Normalizer(inverse=True) models/normalizer.py:26-27).
//
// The residual kernels evaluate every float32 operation in the order torch evaluates it, with explicit
// round-to-nearest intrinsics (no FMA contraction), so the loss matrix is bit-identical to the reference's; only the
// final sum is taken in a different (fixed, float64) order.  The gradient is the analytic adjoint of the FORCE step
// (autograd in the reference), float32, staged through shared memory one image row per CTA.
//
// HBM-bound: 8-16 B read and 8 B written per cell; the FORCE stencil only spans the two neighbouring cells.
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

namespace mcedm {

// one channel of a [B,T,X] field: element (b,t,x) at p[b*sb + t*st + x*sx], float32 or float64, normalised
struct Plane {
  const void* p;
  long long sb, st, sx;
  int f64;
  float div, sub;  // un-normalised value = v*div + sub
  int apply;       // 0: v is already un-normalised
};

__device__ __forceinline__ float load_plane(const Plane& pl, long long b, int t, int x) {
  const long long i = b * pl.sb + (long long)t * pl.st + (long long)x * pl.sx;
  const float v = pl.f64 ? (float)reinterpret_cast<const double*>(pl.p)[i] : reinterpret_cast<const float*>(pl.p)[i];
  return pl.apply ? __fadd_rn(__fmul_rn(v, pl.div), pl.sub) : v;
}

struct SweConst {
  float c;    // float32(0.5*dt)
  float dx;   // float32 grid spacing
  float eps;  // 1e-8f
  float hg;   // float32(0.5*g)
};

// hu**2 / (h + eps) + 0.5*g*h**2
__device__ __forceinline__ float swe_flux(float q, float h, const SweConst& k) {
  return __fadd_rn(__fdiv_rn(__fmul_rn(q, q), __fadd_rn(h, k.eps)), __fmul_rn(k.hg, __fmul_rn(h, h)));
}
// 0.5*(a0 + a1) - 0.5*dt*(b1 - b0)/dx
__device__ __forceinline__ float swe_half(float a0, float a1, float b0, float b1, const SweConst& k) {
  return __fsub_rn(__fmul_rn(0.5f, __fadd_rn(a0, a1)), __fdiv_rn(__fmul_rn(k.c, __fsub_rn(b1, b0)), k.dx));
}

// (H, U) of cell x after one FORCE step of the row whose cells x-1, x, x+1 (replicated at the ends) are given
__device__ __forceinline__ void swe_step_cell(const float h[3], const float u[3], const SweConst& k, float& H, float& U) {
  float q[3], F[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    q[i] = __fmul_rn(u[i], h[i]);
    F[i] = swe_flux(q[i], h[i], k);
  }
  const float hm0 = swe_half(h[0], h[1], q[0], q[1], k), hm1 = swe_half(h[1], h[2], q[1], q[2], k);
  const float qm0 = swe_half(q[0], q[1], F[0], F[1], k), qm1 = swe_half(q[1], q[2], F[1], F[2], k);
  H = swe_half(hm0, hm1, qm0, qm1, k);
  const float G0 = swe_flux(qm0, hm0, k), G1 = swe_flux(qm1, hm1, k);
  const float Q = swe_half(qm0, qm1, G0, G1, k);
  U = __fdiv_rn(Q, __fadd_rn(H, k.eps));
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < nw; ++i) s += sh[i];
  return s;
}

// typed load of one plane element (compile-time dtype / normalisation: no per-load branches)
template <bool F64, bool APPLY>
__device__ __forceinline__ float load_row(const void* row, long long off, float div, float sub) {
  const float v = F64 ? (float)reinterpret_cast<const double*>(row)[off] : reinterpret_cast<const float*>(row)[off];
  return APPLY ? __fadd_rn(__fmul_rn(v, div), sub) : v;
}

// loss[b,t,x,:] = (with_ic[b,t,x,:] - gt[b,t,x,:])^2 / scale ; with_ic[t] = step(pred[t-1]) (t>0), pred[0] (t=0)
// One CTA walks `rows_per_cta` consecutive (b,t) rows; thread x owns cell x.  The cell a thread loads for the
// comparison at row t is the centre cell of the step that feeds row t+1, its two neighbours come from the adjacent
// lanes (warp-edge lanes load them), so every plane element is fetched from memory once (+1/16 at warp edges).
template <bool HF64, bool UF64, bool APPLY>
__global__ void __launch_bounds__(1024) swe_fv_loss_kernel(Plane ph, Plane pu, const float* __restrict__ gt, int T, int X,
                                                           long long n_rows, int rows_per_cta, SweConst k, float sc_h,
                                                           float sc_u, float* __restrict__ loss,
                                                           double* __restrict__ cta_sums) {
  __shared__ double sh[32];
  const int x = threadIdx.x, lane = threadIdx.x & 31;
  const bool act = x < X;
  const int xc = act ? x : X - 1;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, n_rows);
  const size_t hsz = HF64 ? 8 : 4, usz = UF64 ? 8 : 4;
  float ch = 0.f, cu = 0.f;   // cell x of the previous row (un-normalised); valid when have_prev
  bool have_prev = false;
  double mine = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const long long b = r / T;
    const int t = (int)(r - b * T);
    const char* hrow = reinterpret_cast<const char*>(ph.p) + (size_t)(b * ph.sb + (long long)t * ph.st) * hsz;
    const char* urow = reinterpret_cast<const char*>(pu.p) + (size_t)(b * pu.sb + (long long)t * pu.st) * usz;
    // this row's own cell: the comparison target when gt == NULL, and the centre cell of the next row's step
    const float oh = load_row<HF64, APPLY>(hrow, (long long)xc * ph.sx, ph.div, ph.sub);
    const float ou = load_row<UF64, APPLY>(urow, (long long)xc * pu.sx, pu.div, pu.sub);
    float vh, vu;
    if (t == 0) {
      vh = oh;
      vu = ou;
    } else {
      if (!have_prev) {   // first row of this CTA: fetch the centre cell of row t-1
        ch = load_row<HF64, APPLY>(hrow - (size_t)ph.st * hsz, (long long)xc * ph.sx, ph.div, ph.sub);
        cu = load_row<UF64, APPLY>(urow - (size_t)pu.st * usz, (long long)xc * pu.sx, pu.div, pu.sub);
      }
      float h[3], u[3];
      h[1] = ch;
      u[1] = cu;
      h[0] = __shfl_up_sync(0xffffffffu, ch, 1);
      u[0] = __shfl_up_sync(0xffffffffu, cu, 1);
      h[2] = __shfl_down_sync(0xffffffffu, ch, 1);
      u[2] = __shfl_down_sync(0xffffffffu, cu, 1);
      if (lane == 0) {
        const int xm = max(xc - 1, 0);
        h[0] = load_row<HF64, APPLY>(hrow - (size_t)ph.st * hsz, (long long)xm * ph.sx, ph.div, ph.sub);
        u[0] = load_row<UF64, APPLY>(urow - (size_t)pu.st * usz, (long long)xm * pu.sx, pu.div, pu.sub);
      }
      if (lane == 31 || x >= X - 1) {
        const int xp = min(xc + 1, X - 1);
        h[2] = load_row<HF64, APPLY>(hrow - (size_t)ph.st * hsz, (long long)xp * ph.sx, ph.div, ph.sub);
        u[2] = load_row<UF64, APPLY>(urow - (size_t)pu.st * usz, (long long)xp * pu.sx, pu.div, pu.sub);
      }
      swe_step_cell(h, u, k, vh, vu);
    }
    if (vh != vh) vh = 0.f;  // pred_next_with_ic[isnan] = 0   (pde_loss.py:211; the initial-condition row included)
    if (vu != vu) vu = 0.f;
    ch = oh;
    cu = ou;
    have_prev = t + 1 < T;   // the next row of this CTA belongs to the same sample
    if (act) {
      float gh = oh, gu = ou;
      const long long o = (r * (long long)X + x) * 2;
      if (gt) {
        const float2 g = *reinterpret_cast<const float2*>(gt + o);
        gh = g.x;
        gu = g.y;
      }
      const float dh = __fsub_rn(vh, gh), du = __fsub_rn(vu, gu);
      const float lh = __fdiv_rn(__fmul_rn(dh, dh), sc_h), lu = __fdiv_rn(__fmul_rn(du, du), sc_u);
      if (loss) *reinterpret_cast<float2*>(loss + o) = make_float2(lh, lu);
      mine += (double)lh + (double)lu;
    }
  }
  const double s = block_sum(mine, sh);
  if (threadIdx.x == 0) cta_sums[blockIdx.x] = s;
}

// out[0] = sum of n partial sums, fixed order (one CTA)
__global__ void __launch_bounds__(1024) pde_sum_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  const double s = block_sum(v, sh);
  if (threadIdx.x == 0) out[0] = s;
}

// d mean(loss) / d pred for one (b,t) row per CTA.  Shared arrays are indexed by padded cell (X+4), midpoint (X+3)
// and node (X+2) exactly as pde_loss.py:139-158; the adjoint runs the same stages backwards.
// mode 0: out[B,T,X,2] ; 1: out[B,T,X] = (g_h + g_u)/2 (channel mean, ddim.py:1445-1446) ; 2: out[B,T,X] = g_h + g_u
__global__ void __launch_bounds__(1024) swe_fv_grad_kernel(Plane ph, Plane pu, const float* __restrict__ gt, int T, int X,
                                                           SweConst k, float sc_h, float sc_u, float inv_n, float g,
                                                           int mode, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int n4 = X + 4;
  float* s_h = smem;             // padded cells
  float* s_u = s_h + n4;
  float* s_q = s_u + n4;
  float* s_hm = s_q + n4;        // midpoints (X+3), reused for g_h
  float* s_qm = s_hm + n4;       // reused for g_u
  float* s_a = s_qm + n4;        // g_hn (X+2) then g_hm (X+3)
  float* s_b = s_a + n4;         // g_qn then g_qm
  float* s_c = s_b + n4;         // scratch for the second adjoint stage
  float* s_d = s_c + n4;
  const long long b = blockIdx.x / T;
  const int t = blockIdx.x - (int)b * T;
  const int i = threadIdx.x;
  const float cdx = k.c / k.dx;
  const bool live = t < T - 1;   // the step of the last row is dropped (pred_next[:, :-1])

  float h_i = 0.f, u_i = 0.f, q_i = 0.f, F_i = 0.f;
  if (i < n4) {
    const int xi = min(max(i - 2, 0), X - 1);
    h_i = load_plane(ph, b, t, xi);
    u_i = load_plane(pu, b, t, xi);
    q_i = __fmul_rn(u_i, h_i);
    F_i = swe_flux(q_i, h_i, k);  // forward values bit-identical to the residual kernel's
    s_h[i] = h_i;
    s_u[i] = F_i;  // s_u carries F during the forward stages
    s_q[i] = q_i;
  }
  __syncthreads();
  float hm = 0.f, qm = 0.f, G = 0.f;
  if (i < X + 3) {
    hm = swe_half(s_h[i], s_h[i + 1], s_q[i], s_q[i + 1], k);
    qm = swe_half(s_q[i], s_q[i + 1], s_u[i], s_u[i + 1], k);
    G = swe_flux(qm, hm, k);
    s_hm[i] = hm;
    s_qm[i] = qm;
    s_c[i] = G;
  }
  __syncthreads();
  if (i < X + 2) {
    float g_hn = 0.f, g_qn = 0.f;
    if (live && i >= 1 && i <= X) {
      const float H = swe_half(s_hm[i], s_hm[i + 1], s_qm[i], s_qm[i + 1], k);
      const float Q = swe_half(s_qm[i], s_qm[i + 1], s_c[i], s_c[i + 1], k);
      const float He = __fadd_rn(H, k.eps);
      const float U = __fdiv_rn(Q, He);
      const int x = i - 1;
      float gh, gu;
      if (gt) {
        const float2 v = *reinterpret_cast<const float2*>(gt + ((b * T + t + 1) * (long long)X + x) * 2);
        gh = v.x;
        gu = v.y;
      } else {
        gh = load_plane(ph, b, t + 1, x);
        gu = load_plane(pu, b, t + 1, x);
      }
      const float gH = (H != H) ? 0.f : 2.f * __fsub_rn(H, gh) / sc_h * inv_n;
      const float gU = (U != U) ? 0.f : 2.f * __fsub_rn(U, gu) / sc_u * inv_n;
      g_qn = gU / He;
      g_hn = gH - gU * Q / (He * He);
    }
    s_a[i] = g_hn;
    s_b[i] = g_qn;
  }
  __syncthreads();
  // adjoint of node <- midpoints
  float g_hm = 0.f, g_qm = 0.f;
  if (i < X + 3) {
    const float hn0 = i >= 1 ? s_a[i - 1] : 0.f, hn1 = i <= X + 1 ? s_a[i] : 0.f;
    const float qn0 = i >= 1 ? s_b[i - 1] : 0.f, qn1 = i <= X + 1 ? s_b[i] : 0.f;
    g_hm = 0.5f * (hn0 + hn1);
    g_qm = cdx * (hn1 - hn0) + 0.5f * (qn0 + qn1);
    const float g_G = cdx * (qn1 - qn0);
    const float he = hm + k.eps;
    g_qm += g_G * 2.f * qm / he;
    g_hm += g_G * (-(qm * qm) / (he * he) + g * hm);
  }
  __syncthreads();
  if (i < X + 3) {
    s_c[i] = g_hm;
    s_d[i] = g_qm;
  }
  __syncthreads();
  // adjoint of midpoints <- padded cells
  if (i < n4) {
    const float a0 = i >= 1 ? s_c[i - 1] : 0.f, a1 = i <= X + 2 ? s_c[i] : 0.f;
    const float b0 = i >= 1 ? s_d[i - 1] : 0.f, b1 = i <= X + 2 ? s_d[i] : 0.f;
    float g_h = 0.5f * (a0 + a1);
    float g_q = cdx * (a1 - a0) + 0.5f * (b0 + b1);
    const float g_F = cdx * (b1 - b0);
    const float he = h_i + k.eps;
    g_q += g_F * 2.f * q_i / he;
    g_h += g_F * (-(q_i * q_i) / (he * he) + g * h_i);
    s_hm[i] = g_h + g_q * u_i;
    s_qm[i] = g_q * h_i;
  }
  __syncthreads();
  if (i < X) {
    float g_h = s_hm[i + 2], g_u = s_qm[i + 2];
    if (i == 0) {
      g_h += s_hm[0] + s_hm[1];
      g_u += s_qm[0] + s_qm[1];
    }
    if (i == X - 1) {
      g_h += s_hm[X + 2] + s_hm[X + 3];
      g_u += s_qm[X + 2] + s_qm[X + 3];
    }
    if (t == 0 && gt) {  // row 0 is compared with gt directly; with gt == pred the term is exactly zero
      const float2 v = *reinterpret_cast<const float2*>(gt + ((b * T) * (long long)X + i) * 2);
      g_h += 2.f * (s_h[i + 2] - v.x) / sc_h * inv_n;
      g_u += 2.f * (load_plane(pu, b, 0, i) - v.y) / sc_u * inv_n;
    }
    if (g_h != g_h) g_h = 0.f;  // dloss[isnan(dloss)] = 0   (pde_loss.py:241)
    if (g_u != g_u) g_u = 0.f;
    const long long o = (b * T + t) * (long long)X + i;
    if (mode == 0)
      *reinterpret_cast<float2*>(out + o * 2) = make_float2(g_h, g_u);
    else if (mode == 1)
      out[o] = __fdiv_rn(__fadd_rn(g_h, g_u), 2.f);
    else
      out[o] = __fadd_rn(g_h, g_u);
  }
}

// Darcy residual: loss[b,i-2,j-2] = (-(d/dx(a du/dx) + d/dy(a du/dy)) - 1)^2 / ((S-4)^2), central differences
// (pde_loss.py:30-56, :81-84); pa = permeability a, pu = solution u, both [B,S,S] planes (t = first index).
__global__ void __launch_bounds__(1024) darcy_loss_kernel(Plane pa, Plane pu, int S, float two_dx, float tn,
                                                          float* __restrict__ loss, double* __restrict__ row_sums) {
  __shared__ double sh[32];
  const int n = S - 4;
  const long long b = blockIdx.x / n;
  const int i = blockIdx.x - (int)b * n + 2;
  const int j = threadIdx.x + 2;
  double mine = 0.0;
  if (threadIdx.x < n) {
    auto U = [&](int p, int q) { return load_plane(pu, b, p, q); };
    auto A = [&](int p, int q) { return load_plane(pa, b, p, q); };
    const float aux1 = __fmul_rn(A(i + 1, j), __fdiv_rn(__fsub_rn(U(i + 2, j), U(i, j)), two_dx));
    const float aux0 = __fmul_rn(A(i - 1, j), __fdiv_rn(__fsub_rn(U(i, j), U(i - 2, j)), two_dx));
    const float auy1 = __fmul_rn(A(i, j + 1), __fdiv_rn(__fsub_rn(U(i, j + 2), U(i, j)), two_dx));
    const float auy0 = __fmul_rn(A(i, j - 1), __fdiv_rn(__fsub_rn(U(i, j), U(i, j - 2)), two_dx));
    const float auxx = __fdiv_rn(__fsub_rn(aux1, aux0), two_dx);
    const float auyy = __fdiv_rn(__fsub_rn(auy1, auy0), two_dx);
    const float r = __fsub_rn(-__fadd_rn(auxx, auyy), 1.0f);
    const float l = __fdiv_rn(__fmul_rn(r, r), tn);
    if (loss) loss[(b * n + (i - 2)) * (long long)n + (j - 2)] = l;
    mine = (double)l;
  }
  const double s = block_sum(mine, sh);
  if (threadIdx.x == 0) row_sums[blockIdx.x] = s;
}

static Plane make_plane(const void* p, int f64, long long sb, long long st, long long sx, int apply, float div, float sub) {
  Plane pl;
  pl.p = p;
  pl.f64 = f64;
  pl.sb = sb;
  pl.st = st;
  pl.sx = sx;
  pl.apply = apply;
  pl.div = div;
  pl.sub = sub;
  return pl;
}

static inline int round32(int n) { return (n + 31) / 32 * 32; }

}  // namespace mcedm

extern "C" int mcedm_swe_fv_loss(const void* h, int h_f64, const long long* h_strides, const void* u, int u_f64,
                                 const long long* u_strides, int apply_norm, float h_div, float h_sub, float u_div,
                                 float u_sub, const float* gt, int B, int T, int X, float half_dt, float dx, float g,
                                 float* loss, double* row_sums, double* total, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && T >= 2 && X >= 2 && X <= 1024, "swe_fv_loss: unsupported field %d x %d x %d", B, T, X);
  MCEDM_REQUIRE(row_sums != nullptr, "swe_fv_loss: row_sums workspace [B*T] is required");
  MCEDM_REQUIRE((long long)B * T < (1ll << 31), "swe_fv_loss: too many rows");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const Plane ph = make_plane(h, h_f64, h_strides[0], h_strides[1], h_strides[2], apply_norm, h_div, h_sub);
  const Plane pu = make_plane(u, u_f64, u_strides[0], u_strides[1], u_strides[2], apply_norm, u_div, u_sub);
  const SweConst k{half_dt, dx, 1e-8f, 0.5f * g};
  const long long n_rows = (long long)B * T;
  const int R = T % 8 == 0 ? 8 : 1;                       // rows per CTA (never straddling two samples' step chains)
  const unsigned grid = (unsigned)((n_rows + R - 1) / R);
  const int threads = round32(X);
  const float sh2 = h_div * h_div, su2 = u_div * u_div;
#define MCEDM_SWE_LOSS(HF, UF, AP) \
  swe_fv_loss_kernel<HF, UF, AP><<<grid, threads, 0, st>>>(ph, pu, gt, T, X, n_rows, R, k, sh2, su2, loss, row_sums)
  const int sel = (h_f64 ? 4 : 0) | (u_f64 ? 2 : 0) | (apply_norm ? 1 : 0);
  switch (sel) {
    case 0: MCEDM_SWE_LOSS(false, false, false); break;
    case 1: MCEDM_SWE_LOSS(false, false, true); break;
    case 2: MCEDM_SWE_LOSS(false, true, false); break;
    case 3: MCEDM_SWE_LOSS(false, true, true); break;
    case 4: MCEDM_SWE_LOSS(true, false, false); break;
    case 5: MCEDM_SWE_LOSS(true, false, true); break;
    case 6: MCEDM_SWE_LOSS(true, true, false); break;
    default: MCEDM_SWE_LOSS(true, true, true); break;
  }
#undef MCEDM_SWE_LOSS
  MCEDM_CUDA(cudaGetLastError());
  if (total) {
    pde_sum_kernel<<<1, 1024, 0, st>>>(row_sums, (int)grid, total);
    MCEDM_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int mcedm_swe_fv_grad(const void* h, int h_f64, const long long* h_strides, const void* u, int u_f64,
                                 const long long* u_strides, int apply_norm, float h_div, float h_sub, float u_div,
                                 float u_sub, const float* gt, int B, int T, int X, float half_dt, float dx, float g,
                                 int mode, float* out, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && T >= 2 && X >= 2 && X + 4 <= 1024, "swe_fv_grad: unsupported field %d x %d x %d", B, T, X);
  MCEDM_REQUIRE(mode >= 0 && mode <= 2, "swe_fv_grad: mode %d", mode);
  MCEDM_REQUIRE((long long)B * T < (1ll << 31), "swe_fv_grad: too many rows");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const Plane ph = make_plane(h, h_f64, h_strides[0], h_strides[1], h_strides[2], apply_norm, h_div, h_sub);
  const Plane pu = make_plane(u, u_f64, u_strides[0], u_strides[1], u_strides[2], apply_norm, u_div, u_sub);
  const SweConst k{half_dt, dx, 1e-8f, 0.5f * g};
  const float inv_n = (float)(1.0 / ((double)B * T * X * 2));
  const size_t smem = (size_t)(X + 4) * 9 * sizeof(float);
  swe_fv_grad_kernel<<<B * T, round32(X + 4), smem, st>>>(ph, pu, gt, T, X, k, h_div * h_div, u_div * u_div, inv_n, g,
                                                          mode, out);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_darcy_loss(const void* a, int a_f64, const long long* a_strides, const void* u, int u_f64,
                                const long long* u_strides, int apply_norm, float a_div, float a_sub, float u_div,
                                float u_sub, int B, int S, float D, float* loss, double* row_sums, double* total,
                                void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && S >= 5 && S - 4 <= 1024, "darcy_loss: unsupported field %d x %d x %d", B, S, S);
  MCEDM_REQUIRE(row_sums != nullptr, "darcy_loss: row_sums workspace [B*(S-4)] is required");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const Plane pa = make_plane(a, a_f64, a_strides[0], a_strides[1], a_strides[2], apply_norm, a_div, a_sub);
  const Plane pu = make_plane(u, u_f64, u_strides[0], u_strides[1], u_strides[2], apply_norm, u_div, u_sub);
  const int n = S - 4;
  const float two_dx = (float)(2.0 * ((double)D / S));
  darcy_loss_kernel<<<B * n, round32(n), 0, st>>>(pa, pu, S, two_dx, (float)(n * n), loss, row_sums);
  MCEDM_CUDA(cudaGetLastError());
  if (total) {
    pde_sum_kernel<<<1, 1024, 0, st>>>(row_sums, B * n, total);
    MCEDM_CUDA(cudaGetLastError());
  }
  return 0;
}
