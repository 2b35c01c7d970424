// Host-side emulation of the device math headers (TEST INFRASTRUCTURE ONLY).
// Compiles nefii_b200/csrc/*_math.cuh with g++ (-ffp-contract=off) so the exact code the CUDA
// kernels run per ray can be checked against the oracle on a machine without a GPU, in float and
// in double.  Never loaded by the product package.
#include <cstring>
#include "sg_math.cuh"

using namespace nefii::sgm;

template <typename T>
static void sg_render_fwd_t(int n_rays, int n_sg, int n_mat, const T* lgt, const T* spec, const T* rough,
                            const T* albedo, const T* normal, const T* view, const T* blend,
                            T* out_rgb, T* out_spec, T* out_diff) {
  LightSG<T>* L = new LightSG<T>[n_sg];
  for (int m = 0; m < n_sg; ++m) load_light(lgt + 7 * m, L[m]);
  for (int r = 0; r < n_rays; ++r) {
    const T* n = normal + 3 * r;
    const T* v = view + 3 * r;
    T a_pi[3];
    for (int c = 0; c < 3; ++c) a_pi[c] = albedo[3 * r + c] * (T(1) / K<T>::pi());
    T s_acc[3] = {0, 0, 0}, d_acc[3] = {0, 0, 0};
    for (int k = 0; k < n_mat; ++k) {
      BrdfLobe<T> B;
      make_brdf_lobe(n, v, rough[k], spec + 3 * k, B);
      T sk[3] = {0, 0, 0};
      for (int m = 0; m < n_sg; ++m) {
        T t[3];
        specular_term(n, L[m], B, t);
        for (int c = 0; c < 3; ++c) sk[c] += t[c];
      }
      T wk = blend ? blend[r * n_mat + k] : T(1);
      for (int c = 0; c < 3; ++c) s_acc[c] += sk[c] * wk;
    }
    for (int m = 0; m < n_sg; ++m) {
      T t[3];
      diffuse_term(n, L[m], a_pi, t);
      for (int c = 0; c < 3; ++c) d_acc[c] += t[c];
    }
    for (int c = 0; c < 3; ++c) {
      T s = clamp_min(s_acc[c], T(0));
      T d = clamp_min(d_acc[c] * T(n_mat), T(0));
      out_spec[3 * r + c] = s;
      out_diff[3 * r + c] = d;
      out_rgb[3 * r + c] = s + d;
    }
  }
  delete[] L;
}

extern "C" {
void emu_sg_render_fwd_f32(int n_rays, int n_sg, int n_mat, const float* lgt, const float* spec, const float* rough,
                           const float* albedo, const float* normal, const float* view, const float* blend,
                           float* out_rgb, float* out_spec, float* out_diff) {
  sg_render_fwd_t<float>(n_rays, n_sg, n_mat, lgt, spec, rough, albedo, normal, view, blend, out_rgb, out_spec, out_diff);
}
void emu_sg_render_fwd_f64(int n_rays, int n_sg, int n_mat, const double* lgt, const double* spec, const double* rough,
                           const double* albedo, const double* normal, const double* view, const double* blend,
                           double* out_rgb, double* out_spec, double* out_diff) {
  sg_render_fwd_t<double>(n_rays, n_sg, n_mat, lgt, spec, rough, albedo, normal, view, blend, out_rgb, out_spec, out_diff);
}
}

// ---- backward of render_with_sg: the hand-derived adjoint of csrc/sg_adjoint_math.cuh, in the loop structure of the CUDA
// kernel (csrc/sg_render.cu: sg_render_bwd_kernel) ----
#include "sg_adjoint_math.cuh"

extern "C" void emu_sg_render_bwd_f64(int n_rays, int n_sg, int n_mat, const double* lgt, const double* spec, const double* rough,
                                      const double* albedo, const double* normal, const double* view,
                                      const double* g_spec, const double* g_diff, double* acc /*[M,7]*/, double* g_rough /*[K]*/,
                                      double* g_specrefl /*[K,3]*/, double* g_albedo /*[N,3]*/, double* g_normal /*[N,3]*/,
                                      const double* blend /*[N,K] or null*/, double* g_blend /*[N,K] or null*/) {
  namespace sga = nefii::sga;
  const double inv_pi = 1.0 / K<double>::pi();
  for (int r = 0; r < n_rays; ++r) {
    // the reference clamps the summed specular / diffuse radiance at 0: recompute the sums for the masks
    double out_rgb[3], out_s[3], out_d[3];
    sg_render_fwd_t<double>(1, n_sg, n_mat, lgt, spec, rough, albedo + 3 * r, normal + 3 * r, view + 3 * r,
                            blend ? blend + (size_t)r * n_mat : nullptr, out_rgb, out_s, out_d);
    double gs[3], gd[3];
    for (int c = 0; c < 3; ++c) {
      gs[c] = out_s[c] > 0 ? g_spec[3 * r + c] : 0.0;
      gd[c] = out_d[c] > 0 ? g_diff[3 * r + c] * n_mat : 0.0;     // the reference sums the diffuse term over the K axis
    }
    const double* n = normal + 3 * r;
    const double* v = view + 3 * r;
    double n_bar[3] = {0, 0, 0}, ga[3] = {0, 0, 0};
    for (int k = 0; k < n_mat; ++k) {
      BrdfLobe<double> B;
      make_brdf_lobe(n, v, rough[k], spec + 3 * k, B);
      double b_bar[3] = {0, 0, 0}, beta_bar = 0, nu_bar[3] = {0, 0, 0}, sk[3] = {0, 0, 0};
      const double wk = blend ? blend[(size_t)r * n_mat + k] : 1.0;
      const double gsw[3] = {gs[0] * wk, gs[1] * wk, gs[2] * wk};
      for (int m = 0; m < n_sg; ++m) {
        LightSG<double> L;
        load_light(lgt + 7 * m, L);
        double w = 0;
        for (int c = 0; c < 3; ++c) w += gsw[c] * L.amp[c] * B.amp[c];
        double a_bar[3] = {0, 0, 0}, l_bar = 0;
        const double phi = sga::specular_phi_vjp(n, L.axis, L.sharp, B.axis, B.sharp, w, a_bar, l_bar, b_bar, beta_bar, n_bar);
        for (int i = 0; i < 3; ++i) acc[7 * m + i] += a_bar[i];
        acc[7 * m + 3] += l_bar;
        for (int c = 0; c < 3; ++c) {
          acc[7 * m + 4 + c] += gsw[c] * B.amp[c] * phi;
          nu_bar[c] += gsw[c] * L.amp[c] * phi;
          sk[c] += L.amp[c] * B.amp[c] * phi;
        }
      }
      if (g_blend) g_blend[(size_t)r * n_mat + k] = gs[0] * sk[0] + gs[1] * sk[1] + gs[2] * sk[2];
      sga::brdf_lobe_vjp(n, v, rough[k], spec + 3 * k, b_bar, beta_bar, nu_bar, n_bar, g_rough[k], g_specrefl + 3 * k);
    }
    for (int m = 0; m < n_sg; ++m) {
      LightSG<double> L;
      load_light(lgt + 7 * m, L);
      double w = 0;
      for (int c = 0; c < 3; ++c) w += gd[c] * L.amp[c] * albedo[3 * r + c] * inv_pi;
      double a_bar[3] = {0, 0, 0}, l_bar = 0;
      const double psi = sga::psi_vjp(n, L.axis, L.sharp, w, n_bar, a_bar, l_bar);
      for (int i = 0; i < 3; ++i) acc[7 * m + i] += a_bar[i];
      acc[7 * m + 3] += l_bar;
      for (int c = 0; c < 3; ++c) {
        acc[7 * m + 4 + c] += gd[c] * albedo[3 * r + c] * inv_pi * psi;
        ga[c] += gd[c] * L.amp[c] * psi * inv_pi;
      }
    }
    for (int c = 0; c < 3; ++c) {
      g_albedo[3 * r + c] = ga[c];
      if (g_normal) g_normal[3 * r + c] = n_bar[c];
    }
  }
}
