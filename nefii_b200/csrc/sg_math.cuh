// Spherical-Gaussian shading math shared by the CUDA kernels (float) and by the host-side
// emulation build used in tests (float / double, see tests/hostemu/).
//
// Every function replicates the per-term floating-point operation order of the reference
// (code/model/sg_render.py:112-158,164-295): the SG integrals cancel two large terms, so the
// translation units that include this header are compiled with -fmad=false (no FMA
// contraction), IEEE division / square root and libdevice exp/pow -- the same primitives
// torch's CUDA elementwise kernels use.  `x / python_scalar` in the reference is a multiply by
// the float reciprocal on CUDA (ATen div_true_kernel_cuda) and `python_scalar / x` is
// reciprocal(x) * scalar (Tensor.__rtruediv__); both are reproduced literally below.
#pragma once
#include <cmath>
#include <type_traits>

#if defined(__CUDACC__)
#define NEFII_HD __host__ __device__ __forceinline__
#else
#define NEFII_HD inline
#endif

namespace nefii {
namespace sgm {

template <typename T> struct K {   // functions, not static members: T may be a dual number inside device code
  static NEFII_HD T eps() { return T(1e-6); }
  static NEFII_HD T cos_mu() { return T(32.7080); }
  static NEFII_HD T cos_lambda() { return T(0.0315); }
  static NEFII_HD T cos_alpha() { return T(31.7003); }
  static NEFII_HD T two_pi() { return T(6.283185307179586); }
  static NEFII_HD T pi() { return T(3.141592653589793); }
};

NEFII_HD float m_exp(float x) { return expf(x); }
NEFII_HD double m_exp(double x) { return exp(x); }
NEFII_HD float m_sqrt(float x) { return sqrtf(x); }
NEFII_HD double m_sqrt(double x) { return sqrt(x); }
NEFII_HD float m_pow(float a, float b) { return powf(a, b); }
NEFII_HD double m_pow(double a, double b) { return pow(a, b); }
NEFII_HD float m_abs(float x) { return fabsf(x); }
NEFII_HD double m_abs(double x) { return fabs(x); }

// torch.clamp(x, min=lo) / (max=hi) / torch.min(a, b): NaN propagates
template <typename T> NEFII_HD T clamp_min(T x, T lo) { return x < lo ? lo : x; }
template <typename T> NEFII_HD T clamp_max(T x, T hi) { return x > hi ? hi : x; }
template <typename T> NEFII_HD T nan_min(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// torch.sum / torch.norm over a contiguous innermost axis of 3 reduce as a tree: (x0 + x2) + x1
// (measured on B200 with torch 2.11, tools/diag_gpu.py dotorder: 100 % bit-exact against this association)
template <typename T> NEFII_HD T sum3(T x0, T x1, T x2) { return (x0 + x2) + x1; }
template <typename T> NEFII_HD T dot3(const T* a, const T* b) { return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <typename T> NEFII_HD T norm3(const T* a) { return m_sqrt(sum3(a[0] * a[0], a[1] * a[1], a[2] * a[2])); }

// v / (|v| + 1e-6)
template <typename T> NEFII_HD void unit3(const T* v, T* out, T eps = K<T>::eps()) {
  T d = norm3(v) + eps;
  out[0] = v[0] / d; out[1] = v[1] / d; out[2] = v[2] / d;
}

// ---- per-light-SG quantities that do not depend on the ray (hoisted; identical values) -------
template <typename T> struct LightSG {
  T axis[3];
  T sharp;
  T amp[3];
  // hemisphere_int(sharp, .) pieces that depend on `sharp` only (sg_render.py:113-119,135-136)
  T h_t, h_ea, h_lower, h_upper;
};

template <typename T> struct HemiCoef { T t, ea, lower, upper; };

template <typename T> NEFII_HD HemiCoef<T> hemi_coef(T sharp_in) {
  HemiCoef<T> c;
  T sharp = sharp_in + K<T>::eps();
  T inv = T(1) / sharp;
  c.t = m_sqrt(sharp) * (T(1.6988) + T(10.8438) * inv) / ((T(1) + T(6.2201) * inv) + (T(10.2415) * inv) * inv);
  c.ea = m_exp(-c.t);
  T rs = (T(1) / sharp) * K<T>::two_pi();      // 2*pi / sharp  == reciprocal(sharp) * 2pi
  c.lower = rs * (m_exp(-sharp) - m_exp(T(-2) * sharp));
  c.upper = rs * (T(1) - m_exp(-sharp));
  return c;
}

template <typename T> NEFII_HD T hemi_eval(const HemiCoef<T>& c, T cos_beta) {
  T front = (cos_beta >= T(0)) ? T(1) : T(0);
  T eb = m_exp(-c.t * clamp_min(cos_beta, T(0)));
  T s_front = (T(1) - c.ea * eb) / (((T(1) - c.ea) + eb) - c.ea * eb);
  T b = m_exp(c.t * clamp_max(cos_beta, T(0)));
  T s_back = (b - c.ea) / ((T(1) - c.ea) * (b + T(1)));
  T s = front * s_front + (T(1) - front) * s_back;
  return c.lower * (T(1) - s) + c.upper * s;
}

template <typename T> NEFII_HD void load_light(const T* raw7, LightSG<T>& L) {
  unit3(raw7, L.axis);
  L.sharp = m_abs(raw7[3]);
  L.amp[0] = m_abs(raw7[4]); L.amp[1] = m_abs(raw7[5]); L.amp[2] = m_abs(raw7[6]);
  HemiCoef<T> c = hemi_coef(L.sharp);
  L.h_t = c.t; L.h_ea = c.ea; L.h_lower = c.lower; L.h_upper = c.upper;
}

// lambda_trick geometry (sg_render.py:141-158) given the already-formed ratio sharp1/sharp2.
template <typename T>
NEFII_HD void sg_product(T ratio, const T* axis1, const T* axis2, T sharp2, T* axis3, T& sharp3, T& eshift) {
  T cosang = dot3(axis1, axis2);
  T scale = m_sqrt((ratio * ratio + T(1)) + (T(2) * ratio) * cosang);
  scale = nan_min(scale, ratio + T(1));
  sharp3 = sharp2 * scale;
  T w1 = ratio / scale;
  T w2 = T(1) / scale;
  T shift = sharp2 * ((scale - ratio) - T(1));
  axis3[0] = w1 * axis1[0] + w2 * axis2[0];
  axis3[1] = w1 * axis1[1] + w2 * axis2[1];
  axis3[2] = w1 * axis1[2] + w2 * axis2[2];
  eshift = m_exp(shift);
}

// (SG(axis,sharp,amp[3]) x clamped cosine around `n`) integrated over the hemisphere,
// sg_render.py:243-252.  `hc2` = hemi_coef(sharp) (may be hoisted by the caller).
template <typename T>
NEFII_HD void cosine_lobe_integral(const T* n, const T* axis, T sharp, const T* amp, const HemiCoef<T>& hc2, T* out3) {
  T ratio = (T(1) / sharp) * K<T>::cos_lambda();   // python float / tensor
  T axis_p[3], sharp_p, e;
  sg_product(ratio, n, axis, sharp, axis_p, sharp_p, e);
  T c1 = dot3(axis_p, n);
  T c2 = dot3(axis, n);
  T h1 = hemi_eval(hemi_coef(sharp_p), c1);
  T h2 = hemi_eval(hc2, c2);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    T amp_p = (K<T>::cos_mu() * amp[c]) * e;
    out3[c] = amp_p * h1 - (amp[c] * K<T>::cos_alpha()) * h2;
  }
}

// Per-ray, per-base-material warped BRDF lobe (sg_render.py:199-237): independent of the light SG.
template <typename T> struct BrdfLobe {
  T axis[3];
  T sharp;
  T amp[3];
};

template <typename T>
NEFII_HD void make_brdf_lobe(const T* n, const T* v, T rough, const T* spec3, BrdfLobe<T>& B) {
  T inv_r4 = T(1) / (((rough * rough) * rough) * rough);
  T b_sharp = T(2) * inv_r4;
  T b_amp = inv_r4 * (T(1) / K<T>::pi());          // tensor / np.pi on CUDA
  T nv = clamp_min(dot3(n, v), T(0));
  T w[3];
  T two_nv = T(2) * nv;
  w[0] = two_nv * n[0] - v[0]; w[1] = two_nv * n[1] - v[1]; w[2] = two_nv * n[2] - v[2];
  unit3(w, B.axis);
  B.sharp = b_sharp / (T(4) * nv + K<T>::eps());
  T h[3] = {B.axis[0] + v[0], B.axis[1] + v[1], B.axis[2] + v[2]};
  T hu[3];
  unit3(h, hu);
  T vh = clamp_min(dot3(v, hu), T(0));
  T fexp = m_pow(T(2), -(T(5.55473) * vh + T(6.8316)) * vh);
  T d1 = clamp_min(dot3(B.axis, n), T(0));
  T d2 = clamp_min(dot3(v, n), T(0));
  T k = ((rough + T(1)) * (rough + T(1))) * T(0.125);
  T g1 = d1 / ((d1 * (T(1) - k) + k) + K<T>::eps());
  T g2 = d2 / ((d2 * (T(1) - k) + k) + K<T>::eps());
  T g = g1 * g2;
  T den = (T(4) * d1) * d2 + K<T>::eps();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    T f = spec3[c] + (T(1) - spec3[c]) * fexp;
    T moi = (f * g) / den;
    B.amp[c] = b_amp * moi;
  }
}

// specular contribution of one (light SG, BRDF lobe) pair, before the sum over M / clamp
template <typename T>
NEFII_HD void specular_term(const T* n, const LightSG<T>& L, const BrdfLobe<T>& B, T* out3) {
  T ratio = L.sharp / B.sharp;
  T axis[3], sharp, e;
  sg_product(ratio, L.axis, B.axis, B.sharp, axis, sharp, e);
  T amp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) amp[c] = (L.amp[c] * B.amp[c]) * e;
  cosine_lobe_integral(n, axis, sharp, amp, hemi_coef(sharp), out3);
}

// diffuse contribution of one light SG; albedo_over_pi = albedo * (1/pi)
template <typename T>
NEFII_HD void diffuse_term(const T* n, const LightSG<T>& L, const T* albedo_over_pi, T* out3) {
  T amp[3] = {L.amp[0] * albedo_over_pi[0], L.amp[1] * albedo_over_pi[1], L.amp[2] * albedo_over_pi[2]};
  HemiCoef<T> hc{L.h_t, L.h_ea, L.h_lower, L.h_upper};
  cosine_lobe_integral(n, L.axis, L.sharp, amp, hc, out3);
}

}  // namespace sgm
}  // namespace nefii
