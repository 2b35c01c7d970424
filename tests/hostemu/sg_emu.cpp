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
    for (int c = 0; c < 3; ++c) a_pi[c] = albedo[3 * r + c] * (T(1) / K<T>::pi);
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
