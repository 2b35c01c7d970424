// Importance sampling, MIS weights and GGX shading of the near-field integrator, shared by the CUDA
// kernels (float, csrc/mis.cu) and the host emulation used in tests (float / double).
// Reference: code/model/path_tracing_render.py:12-33,61-271,390-413,1406-1476.  Operation order follows
// the reference term by term (see sg_math.cuh for the rules); the translation unit is built with
// -fmad=false.
#pragma once
#include "sg_math.cuh"

namespace nefii {
namespace mism {

using sgm::K;
using sgm::clamp_min;
using sgm::dot3;
using sgm::m_exp;
using sgm::m_sqrt;
using sgm::m_pow;
using sgm::m_abs;

NEFII_HD float m_acos(float x) { return acosf(x); }
NEFII_HD double m_acos(double x) { return acos(x); }
NEFII_HD float m_atan(float x) { return atanf(x); }
NEFII_HD double m_atan(double x) { return atan(x); }
NEFII_HD float m_sin(float x) { return sinf(x); }
NEFII_HD double m_sin(double x) { return sin(x); }
NEFII_HD float m_cos(float x) { return cosf(x); }
NEFII_HD double m_cos(double x) { return cos(x); }
NEFII_HD float m_log(float x) { return logf(x); }
NEFII_HD double m_log(double x) { return log(x); }

// per-light quantities that do not depend on the ray
template <typename T> struct MixLobe {
  T axis[3];
  T sharp;
  T amp[3];
  T energy;   // sum of amplitudes
  T c;        // sharp / (2 pi (1 - exp(-2 sharp)))
};

template <typename T> NEFII_HD void load_mix_lobe(const T* raw7, MixLobe<T>& L) {
  sgm::unit3(raw7, L.axis);
  L.sharp = m_abs(raw7[3]);
  L.amp[0] = m_abs(raw7[4]); L.amp[1] = m_abs(raw7[5]); L.amp[2] = m_abs(raw7[6]);
  L.energy = sgm::sum3(L.amp[0], L.amp[1], L.amp[2]);
  L.c = L.sharp / (K<T>::two_pi() * (T(1) - m_exp(T(-2) * L.sharp)));
}

NEFII_HD float m_fma(float a, float b, float c) { return fmaf(a, b, c); }
NEFII_HD double m_fma(double a, double b, double c) { return fma(a, b, c); }

// torch.cross is ONE kernel (a1*b2 - a2*b1 per component), which nvcc contracts to fma(a1, b2, -(a2*b1));
// everything else in the reference is one torch op per arithmetic operation (no contraction).
template <typename T> NEFII_HD void cross3(const T* a, const T* b, T* c) {
  c[0] = m_fma(a[1], b[2], -(a[2] * b[1]));
  c[1] = m_fma(a[2], b[0], -(a[0] * b[2]));
  c[2] = m_fma(a[0], b[1], -(a[1] * b[0]));
}

// rotate_to_normal: local (z = n) -> world
template <typename T> NEFII_HD void to_frame(const T* local, const T* n, T* out) {
  T up[3] = {T(1), T(0), T(0)};
  if (n[0] > T(0.9)) { up[0] = T(0); up[1] = T(1); }
  T t[3], tu[3], s[3];
  cross3(up, n, t);
  sgm::unit3(t, tu);
  cross3(tu, n, s);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = (local[0] * tu[i] + local[1] * s[i]) + local[2] * n[i];
}

template <typename T> NEFII_HD void spherical(T theta, T phi, T* out) {
  const T st = m_sin(theta);
  out[0] = st * m_cos(phi);
  out[1] = st * m_sin(phi);
  out[2] = m_cos(theta);
}

template <typename T> NEFII_HD T pdf_cos(const T* wi, const T* n) {
  return clamp_min(dot3(wi, n), K<T>::eps()) * (T(1) / K<T>::pi());
}

template <typename T> NEFII_HD T pdf_ggx(const T* wi, const T* n, const T* v, T rough) {
  T h[3] = {wi[0] + v[0], wi[1] + v[1], wi[2] + v[2]};
  const T hn = sgm::norm3(h);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    h[i] = h[i] / hn;
    if (h[i] != h[i]) h[i] = n[i];   // wi == -view: NaN half vector -> normal (path_tracing_render.py:110-111)
  }
  const T c = clamp_min(dot3(h, n), K<T>::eps());
  const T r4 = m_pow(rough, T(4));                         // roughness ** 4 (torch: powf for exponents other than 2, 3)
  const T c2 = c * c;
  const T root = c2 + (T(1) - c2) / r4;
  const T pdf_h = c / (((K<T>::pi() * r4) * root) * root);
  const T hv = clamp_min(dot3(h, v), K<T>::eps());
  return pdf_h / (T(4) * hv);
}

// normalisation of the mixture weights for a given normal: sum_k energy_k * max(n . axis_k, 1e-6)
template <typename T> NEFII_HD T mix_weight(const MixLobe<T>& L, const T* n) {
  return L.energy * clamp_min(dot3(n, L.axis), K<T>::eps());
}

template <typename T> NEFII_HD T pdf_mix(const MixLobe<T>* lobes, int n_sg, const T* wi, const T* n, T wsum) {
  T acc = T(0);
  for (int k = 0; k < n_sg; ++k) {
    const MixLobe<T>& L = lobes[k];
    const T alpha = mix_weight(L, n) / wsum;
    acc += (alpha * L.c) * m_exp(L.sharp * (dot3(wi, L.axis) - T(1)));
  }
  return acc;
}

// The three importance-sampled directions of one surface point and their MIS data.
// u: the 7 uniforms in the reference's draw order.  Outputs: wi[3][3], pdf[3] (clamped), mat[3][3]
// (mat[i][j] = pdf of strategy j at direction i), weight[3] = power heuristic.
template <typename T>
NEFII_HD void sample_point(const MixLobe<T>* lobes, int n_sg, const T* n, const T* v, T rough, const T* u,
                           T wi[3][3], T pdf[3], T mat[3][3], T weight[3]) {
  T local[3];
  // cosine-weighted (path_tracing_render.py:128-156)
  {
    const T theta = m_acos(m_sqrt(T(1) - u[0]));
    const T phi = K<T>::two_pi() * u[1];
    spherical(theta, phi, local);
    to_frame(local, n, wi[0]);
    pdf[0] = m_cos(theta) * (T(1) / K<T>::pi());
  }
  // GGX half vector (:61-103)
  {
    const T theta = m_atan((rough * rough) * m_sqrt(u[2] / (T(1) - u[2])));
    const T phi = K<T>::two_pi() * u[3];
    spherical(theta, phi, local);
    T h[3];
    to_frame(local, n, h);
    const T two_vh = T(2) * dot3(v, h);
#pragma unroll
    for (int i = 0; i < 3; ++i) wi[1][i] = two_vh * h[i] - v[i];
    pdf[1] = pdf_ggx(wi[1], n, v, rough);
  }
  // mixture of light SGs (:168-242)
  T wsum = T(0);
  for (int k = 0; k < n_sg; ++k) wsum += mix_weight(lobes[k], n);
  {
    int pick = -1;
    T right = T(0);
    for (int k = 0; k < n_sg; ++k) {
      const T alpha = mix_weight(lobes[k], n) / wsum;
      right += alpha;
      const T hi = (k == n_sg - 1) ? T(1) : right;
      const T lo = (k == 0) ? T(0) : right - alpha;
      if (pick < 0 && u[4] >= lo && u[4] < hi) pick = k;
    }
    if (pick < 0) pick = 0;   // torch.max over an all-false row returns index 0
    const MixLobe<T>& L = lobes[pick];
    const T inner = clamp_min(T(1) - (L.sharp * u[5]) / (K<T>::two_pi() * L.c), K<T>::eps());
    const T theta = m_acos(((T(1) / L.sharp) * m_log(inner)) + T(1));
    const T phi = K<T>::two_pi() * u[6];
    spherical(theta, phi, local);
    to_frame(local, L.axis, wi[2]);
    pdf[2] = pdf_mix(lobes, n_sg, wi[2], n, wsum);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) pdf[i] = clamp_min(pdf[i], K<T>::eps());
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mat[i][0] = (i == 0) ? pdf[0] : pdf_cos(wi[i], n);
    mat[i][1] = (i == 1) ? pdf[1] : pdf_ggx(wi[i], n, v, rough);
    mat[i][2] = (i == 2) ? pdf[2] : pdf_mix(lobes, n_sg, wi[i], n, wsum);
    const T total = ((T(0) + mat[i][0] * mat[i][0]) + mat[i][1] * mat[i][1]) + mat[i][2] * mat[i][2];
    weight[i] = (mat[i][i] * mat[i][i]) / clamp_min(total, K<T>::eps());
  }
}

// environment radiance along w: sum_k amp_k exp(sharp_k (w . axis_k - 1))
template <typename T> NEFII_HD void env_light(const MixLobe<T>* lobes, int n_sg, const T* w, T* out3) {
  T acc[3] = {T(0), T(0), T(0)};
  for (int k = 0; k < n_sg; ++k) {
    const MixLobe<T>& L = lobes[k];
    const T e = m_exp(L.sharp * (dot3(w, L.axis) - T(1)));
    acc[0] += L.amp[0] * e; acc[1] += L.amp[1] * e; acc[2] += L.amp[2] * e;
  }
  out3[0] = acc[0]; out3[1] = acc[1]; out3[2] = acc[2];
}

// geometry-only factors of one shading sample (no gradient flows through them)
template <typename T> struct ShadeGeom {
  T nh, E, d1, d2, cosn;
  T hu[3];             // unit half vector (for the gradient w.r.t. the normal)
  T nh_raw, d1_raw, d2_raw;   // the three dot products with the normal before clamp(min=0)
};

template <typename T> NEFII_HD void shade_geom(const T* n, const T* v, const T* w, ShadeGeom<T>& g) {
  T h[3] = {w[0] + v[0], w[1] + v[1], w[2] + v[2]};
  sgm::unit3(h, g.hu);
  g.nh_raw = dot3(n, g.hu);
  g.nh = clamp_min(g.nh_raw, T(0));
  const T vh = clamp_min(dot3(v, g.hu), T(0));
  g.E = m_pow(T(2), -(T(5.55473) * vh + T(6.8316)) * vh);
  g.d1_raw = dot3(v, n);
  g.d2_raw = dot3(w, n);
  g.d1 = clamp_min(g.d1_raw, T(0));
  g.d2 = clamp_min(g.d2_raw, T(0));
  g.cosn = g.d2;
}

// d/d normal from the gradients w.r.t. the three clamped dot products (n.h, v.n, w.n); torch.clamp(min=0) passes the
// gradient where its input is >= 0.  Only needed when the geometry trains (the normal is d sdf/dx with a graph).
template <typename T>
NEFII_HD void shade_geom_bwd_normal(const ShadeGeom<T>& g, const T* v, const T* w, const T* g_dots, T* g_n) {
  const T a = (g.nh_raw >= T(0)) ? g_dots[0] : T(0);
  const T b = (g.d1_raw >= T(0)) ? g_dots[1] : T(0);
  const T c = (g.d2_raw >= T(0)) ? g_dots[2] : T(0);
#pragma unroll
  for (int i = 0; i < 3; ++i) g_n[i] += (a * g.hu[i] + b * v[i]) + c * w[i];
}

// One sample of the estimator (path_tracing_render.py:1406-1476): spec[3], diff[3] after the clamp.
template <typename T>
NEFII_HD void shade_sample(const ShadeGeom<T>& g, T rough, const T* spec_refl, const T* albedo, const T* light, T vis,
                           const T* indirect, T weight, T pdf, T* spec3, T* diff3) {
  const T r2 = rough * rough;
  const T r4 = r2 * r2;
  const T nh2 = g.nh * g.nh;
  const T root = nh2 + (T(1) - nh2) / r4;
  const T D = T(1) / (((K<T>::pi() * r4) * root) * root);
  const T k = ((rough + T(1)) * (rough + T(1))) * T(0.125);
  const T G1 = g.d1 / ((g.d1 * (T(1) - k) + k) + K<T>::eps());
  const T G2 = g.d2 / ((g.d2 * (T(1) - k) + k) + K<T>::eps());
  const T G = G1 * G2;
  const T den = (T(4) * g.d1) * g.d2 + K<T>::eps();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const T F = spec_refl[c] + (T(1) - spec_refl[c]) * g.E;
    const T fs = ((F * D) * G) / den;
    const T la = light[c] * vis + (T(1) - vis) * indirect[c];
    const T s = (((weight * la) * fs) * g.cosn) / pdf;
    const T d = (((weight * la) * (albedo[c] * (T(1) / K<T>::pi()))) * g.cosn) / pdf;
    spec3[c] = clamp_min(s, T(0));
    diff3[c] = clamp_min(d, T(0));
  }
}

// Backward of one sample.  gs/gd: upstream gradients of the (clamped) specular / diffuse estimates.
// Accumulates d/d roughness, d/d albedo, d/d specular reflectance; returns d/d light and d/d indirect.
template <typename T>
NEFII_HD void shade_sample_bwd(const ShadeGeom<T>& g, T rough, const T* spec_refl, const T* albedo, const T* light, T vis,
                               const T* indirect, T weight, T pdf, const T* gs, const T* gd, T& g_rough, T* g_albedo,
                               T* g_spec_refl, T* g_light, T* g_indirect, T* g_dots = nullptr) {
  const T r2 = rough * rough;
  const T r4 = r2 * r2;
  const T nh2 = g.nh * g.nh;
  const T root = nh2 + (T(1) - nh2) / r4;
  const T D = T(1) / (((K<T>::pi() * r4) * root) * root);
  const T k = ((rough + T(1)) * (rough + T(1))) * T(0.125);
  const T den1 = (g.d1 * (T(1) - k) + k) + K<T>::eps();
  const T den2 = (g.d2 * (T(1) - k) + k) + K<T>::eps();
  const T G1 = g.d1 / den1, G2 = g.d2 / den2;
  const T G = G1 * G2;
  const T den = (T(4) * g.d1) * g.d2 + K<T>::eps();
  const T q = (weight * g.cosn) / pdf;
  T g_DG = T(0), g_den = T(0), g_cosn = T(0);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const T F = spec_refl[c] + (T(1) - spec_refl[c]) * g.E;
    const T fs = ((F * D) * G) / den;
    const T la = light[c] * vis + (T(1) - vis) * indirect[c];
    const T a_pi = albedo[c] * (T(1) / K<T>::pi());
    const T s = (((weight * la) * fs) * g.cosn) / pdf;
    const T d = (((weight * la) * a_pi) * g.cosn) / pdf;
    const T ms = (s >= T(0)) ? gs[c] : T(0);       // clamp(min=0) passes the gradient where the input is >= 0
    const T md = (d >= T(0)) ? gd[c] : T(0);
    const T g_la = ms * q * fs + md * q * a_pi;
    g_light[c] = g_la * vis;
    g_indirect[c] = g_la * (T(1) - vis);
    g_albedo[c] += md * q * la * (T(1) / K<T>::pi());
    const T g_fs = ms * q * la;
    const T g_F = g_fs * (D * G) / den;
    g_spec_refl[c] += g_F * (T(1) - g.E);
    g_DG += g_fs * F / den;
    g_den += g_fs * (-fs / den);
    g_cosn += (ms * fs + md * a_pi) * ((weight * la) / pdf);
  }
  const T g_D = g_DG * G, g_G = g_DG * D;
  const T dD_dr4 = -D / r4 + T(2) * D * (T(1) - nh2) / (root * r4 * r4);
  const T dG_dk = (-g.d1 * (T(1) - g.d1) / (den1 * den1)) * G2 + G1 * (-g.d2 * (T(1) - g.d2) / (den2 * den2));
  g_rough += g_D * dD_dr4 * (T(4) * rough * r2) + g_G * dG_dk * ((rough + T(1)) * T(0.25));
  if (g_dots != nullptr) {
    // gradients w.r.t. the clamped dot products n.h (through D), v.n (through G1 and the denominator) and w.n (G2, the
    // denominator and the cosine factor)
    const T droot_dnh = T(2) * g.nh * (T(1) - T(1) / r4);
    g_dots[0] = g_D * (T(-2) * D / root) * droot_dnh;
    g_dots[1] = g_G * G2 * ((k + K<T>::eps()) / (den1 * den1)) + g_den * (T(4) * g.d2);
    g_dots[2] = g_G * G1 * ((k + K<T>::eps()) / (den2 * den2)) + g_den * (T(4) * g.d1) + g_cosn;
  }
}

}  // namespace mism
}  // namespace nefii
