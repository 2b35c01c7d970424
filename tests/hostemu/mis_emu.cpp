// Host-side emulation of csrc/mis_math.cuh (TEST INFRASTRUCTURE ONLY, see sg_emu.cpp).
#include <vector>
#include "mis_math.cuh"

using namespace nefii::mism;

template <typename T>
static void mis_sample_t(int n, int n_sg, const T* lgt, const T* rough, const T* normal, const T* view, const T* u,
                         T* wi, T* pdf, T* weight, T* mat) {
  std::vector<MixLobe<T>> L(n_sg);
  for (int m = 0; m < n_sg; ++m) load_mix_lobe(lgt + 7 * m, L[m]);
  for (int i = 0; i < n; ++i) {
    T w[3][3], p[3], m[3][3], wt[3];
    sample_point(L.data(), n_sg, normal + 3 * i, view + 3 * i, rough[i], u + 7 * i, w, p, m, wt);
    for (int s = 0; s < 3; ++s) {
      for (int c = 0; c < 3; ++c) wi[((size_t)s * n + i) * 3 + c] = w[s][c];
      pdf[(size_t)s * n + i] = p[s];
      weight[(size_t)s * n + i] = wt[s];
      for (int j = 0; j < 3; ++j) mat[((size_t)s * 3 + j) * n + i] = m[s][j];
    }
  }
}

template <typename T>
static void mis_shade_t(int n, int n_sg, const T* lgt, const T* spec, int spec_stride, const T* rough, const T* albedo,
                        const T* normal, const T* view, const T* wi, const T* pdf, const T* weight, const T* vis,
                        const T* indirect, T* out_rgb, T* out_spec, T* out_diff,
                        const T* g_rgb, T* g_rough, T* g_albedo, T* g_specrefl, T* g_indirect, T* g_lgt_acc, T* g_normal) {
  std::vector<MixLobe<T>> L(n_sg);
  for (int m = 0; m < n_sg; ++m) load_mix_lobe(lgt + 7 * m, L[m]);
  for (int i = 0; i < n; ++i) {
    const T* nn = normal + 3 * i; const T* vv = view + 3 * i;
    T rgb[3] = {0, 0, 0}, st[3] = {0, 0, 0}, dt[3] = {0, 0, 0};
    T gr = 0, ga[3] = {0, 0, 0}, gsr[3] = {0, 0, 0}, gn[3] = {0, 0, 0};
    for (int s = 0; s < 3; ++s) {
      const size_t si = (size_t)s * n + i;
      T light[3];
      env_light(L.data(), n_sg, wi + si * 3, light);
      ShadeGeom<T> g;
      shade_geom(nn, vv, wi + si * 3, g);
      T sp[3], df[3];
      shade_sample(g, rough[i], spec + (size_t)i * spec_stride, albedo + 3 * i, light, vis[si], indirect + si * 3,
                   weight[si], pdf[si], sp, df);
      for (int c = 0; c < 3; ++c) { st[c] += sp[c]; dt[c] += df[c]; rgb[c] += sp[c] + df[c]; }
      if (g_rgb) {
        T gs[3] = {g_rgb[3 * i], g_rgb[3 * i + 1], g_rgb[3 * i + 2]};
        T glight[3], gind[3], gdots[3];
        shade_sample_bwd(g, rough[i], spec + (size_t)i * spec_stride, albedo + 3 * i, light, vis[si], indirect + si * 3,
                         weight[si], pdf[si], gs, gs, gr, ga, gsr, glight, gind, gdots);
        shade_geom_bwd_normal(g, vv, wi + si * 3, gdots, gn);
        for (int c = 0; c < 3; ++c) g_indirect[si * 3 + c] = gind[c];
        const T* w = wi + si * 3;
        for (int k = 0; k < n_sg; ++k) {
          const MixLobe<T>& Lk = L[k];
          const T dm1 = dot3(w, Lk.axis) - T(1);
          const T e = m_exp(Lk.sharp * dm1);
          const T tsum = (glight[0] * Lk.amp[0] + glight[1] * Lk.amp[1] + glight[2] * Lk.amp[2]) * e;
          for (int c = 0; c < 3; ++c) g_lgt_acc[k * 7 + c] += tsum * Lk.sharp * w[c];
          g_lgt_acc[k * 7 + 3] += tsum * dm1;
          for (int c = 0; c < 3; ++c) g_lgt_acc[k * 7 + 4 + c] += glight[c] * e;
        }
      }
    }
    for (int c = 0; c < 3; ++c) { out_rgb[3 * i + c] = rgb[c]; out_spec[3 * i + c] = st[c]; out_diff[3 * i + c] = dt[c]; }
    if (g_rgb) {
      g_rough[i] = gr;
      for (int c = 0; c < 3; ++c) { g_albedo[3 * i + c] = ga[c]; g_specrefl[3 * i + c] = gsr[c]; if (g_normal) g_normal[3 * i + c] = gn[c]; }
    }
  }
}

extern "C" {
void emu_mis_sample_f64(int n, int n_sg, const double* lgt, const double* rough, const double* normal, const double* view,
                        const double* u, double* wi, double* pdf, double* weight, double* mat) {
  mis_sample_t<double>(n, n_sg, lgt, rough, normal, view, u, wi, pdf, weight, mat);
}
void emu_mis_sample_f32(int n, int n_sg, const float* lgt, const float* rough, const float* normal, const float* view,
                        const float* u, float* wi, float* pdf, float* weight, float* mat) {
  mis_sample_t<float>(n, n_sg, lgt, rough, normal, view, u, wi, pdf, weight, mat);
}
void emu_mis_shade_f64(int n, int n_sg, const double* lgt, const double* spec, int spec_stride, const double* rough,
                       const double* albedo, const double* normal, const double* view, const double* wi, const double* pdf,
                       const double* weight, const double* vis, const double* indirect, double* out_rgb, double* out_spec,
                       double* out_diff, const double* g_rgb, double* g_rough, double* g_albedo, double* g_specrefl,
                       double* g_indirect, double* g_lgt_acc, double* g_normal) {
  mis_shade_t<double>(n, n_sg, lgt, spec, spec_stride, rough, albedo, normal, view, wi, pdf, weight, vis, indirect, out_rgb,
                      out_spec, out_diff, g_rgb, g_rough, g_albedo, g_specrefl, g_indirect, g_lgt_acc, g_normal);
}
}
