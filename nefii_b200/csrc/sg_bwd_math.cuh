// Backward of render_with_sg (reference code/model/sg_render.py:164-295) per (ray, light SG, base material) term.
// The closed-form SG integrals are differentiated by running the forward code of sg_math.cuh on dual numbers:
//   specular term  = amp_light[c] * amp_brdf[c] * phi(axis_light, sharp_light, sharp_brdf)      (phi is channel independent)
//   diffuse term   = amp_light[c] * albedo[c] / pi * phi_d(axis_light, sharp_light)
// so 5 (resp. 4) tangents give every partial derivative.  Light-SG gradients are accumulated in the "unit"
// parametrisation {d axis(3), d sharpness, d amplitude(3)} and converted by nefii_sg_param_grad (abs(), normalisation).
#pragma once
#include "dual.cuh"

namespace nefii {
namespace sgb {

using sgm::K;

// acc7 += contributions of this term; g_rough / g_specrefl accumulate d/d roughness[k], d/d specular_reflectance[k]
template <typename S>
NEFII_HD void specular_term_bwd(const S* n, const S* v, const S* raw7, S rough, const S* spec3, const S* g3, S* acc7, S& g_rough,
                                S* g_specrefl) {
  typedef Dual<S, 5> D;
  // light lobe in the unit parametrisation (value semantics identical to sgm::load_light)
  S axis[3];
  sgm::unit3(raw7, axis);
  const S sharp = sgm::m_abs(raw7[3]);
  const S amp[3] = {sgm::m_abs(raw7[4]), sgm::m_abs(raw7[5]), sgm::m_abs(raw7[6])};
  // BRDF lobe (value) and its dependence on roughness / specular reflectance
  sgm::BrdfLobe<S> B;
  sgm::make_brdf_lobe(n, v, rough, spec3, B);
  // phi with tangents: axis (0..2), light sharpness (3), BRDF sharpness (4); amplitudes set to one
  sgm::LightSG<D> Ld;
  for (int i = 0; i < 3; ++i) Ld.axis[i] = D::variable(axis[i], i);
  Ld.sharp = D::variable(sharp, 3);
  Ld.amp[0] = Ld.amp[1] = Ld.amp[2] = D(S(1));
  sgm::BrdfLobe<D> Bd;
  for (int i = 0; i < 3; ++i) Bd.axis[i] = D(B.axis[i]);
  Bd.sharp = D::variable(B.sharp, 4);
  Bd.amp[0] = Bd.amp[1] = Bd.amp[2] = D(S(1));
  D nd[3] = {D(n[0]), D(n[1]), D(n[2])};
  D out3[3];
  sgm::specular_term(nd, Ld, Bd, out3);
  const D phi = out3[0];
  // upstream gradient folded with the amplitudes
  S w_sum = S(0);   // sum_c g_c amp_c ampB_c
  for (int c = 0; c < 3; ++c) {
    const S gab = g3[c] * amp[c] * B.amp[c];
    w_sum += gab;
    acc7[4 + c] += g3[c] * B.amp[c] * phi.v;            // d / d amp_light[c]
  }
  for (int i = 0; i < 3; ++i) acc7[i] += w_sum * phi.d[i];
  acc7[3] += w_sum * phi.d[3];
  // roughness: through B.sharp (= 2 r^-4 / (4 n.v + eps)) and B.amp (= r^-4/pi * F G / den)
  const S dsharp_dr = S(-4) * B.sharp / rough;
  S gr = w_sum * phi.d[4] * dsharp_dr;
  // B.amp[c] = b_amp * (F_c * G) / den with b_amp ~ r^-4, G = G1 G2 depending on k = (r+1)^2/8
  const S nv = sgm::clamp_min(sgm::dot3(n, v), S(0));
  (void)nv;
  S hvec[3] = {B.axis[0] + v[0], B.axis[1] + v[1], B.axis[2] + v[2]};
  S hu[3];
  sgm::unit3(hvec, hu);
  const S vh = sgm::clamp_min(sgm::dot3(v, hu), S(0));
  const S E = sgm::m_pow(S(2), -(S(5.55473) * vh + S(6.8316)) * vh);
  const S d1 = sgm::clamp_min(sgm::dot3(B.axis, n), S(0));
  const S d2 = sgm::clamp_min(sgm::dot3(v, n), S(0));
  const S k = ((rough + S(1)) * (rough + S(1))) * S(0.125);
  const S den1 = (d1 * (S(1) - k) + k) + K<S>::eps(), den2 = (d2 * (S(1) - k) + k) + K<S>::eps();
  const S G1 = d1 / den1, G2 = d2 / den2;
  const S G = G1 * G2;
  const S dG_dk = (-d1 * (S(1) - d1) / (den1 * den1)) * G2 + G1 * (-d2 * (S(1) - d2) / (den2 * den2));
  const S dk_dr = (rough + S(1)) * S(0.25);
  for (int c = 0; c < 3; ++c) {
    const S g_ampB = g3[c] * amp[c] * phi.v;             // d / d B.amp[c]
    const S F = spec3[c] + (S(1) - spec3[c]) * E;
    // d B.amp / d r = -4 B.amp / r  +  B.amp / G * dG/dk * dk/dr      (G may be 0 -> use the explicit product form)
    const S b_amp_over_den = (G != S(0)) ? B.amp[c] / (F * G) : S(0);   // = b_amp / den
    gr += g_ampB * (S(-4) * B.amp[c] / rough + b_amp_over_den * F * dG_dk * dk_dr);
    g_specrefl[c] += g_ampB * b_amp_over_den * G * (S(1) - E);
  }
  g_rough += gr;
}

template <typename S>
NEFII_HD void diffuse_term_bwd(const S* n, const S* raw7, const S* albedo3, S n_mat, const S* g3, S* acc7, S* g_albedo) {
  typedef Dual<S, 4> D;
  S axis[3];
  sgm::unit3(raw7, axis);
  const S sharp = sgm::m_abs(raw7[3]);
  const S amp[3] = {sgm::m_abs(raw7[4]), sgm::m_abs(raw7[5]), sgm::m_abs(raw7[6])};
  D ax[3] = {D::variable(axis[0], 0), D::variable(axis[1], 1), D::variable(axis[2], 2)};
  const D sh = D::variable(sharp, 3);
  D one3[3] = {D(S(1)), D(S(1)), D(S(1))};
  D nd[3] = {D(n[0]), D(n[1]), D(n[2])};
  D out3[3];
  sgm::cosine_lobe_integral(nd, ax, sh, one3, sgm::hemi_coef(sh), out3);
  const D phi = out3[0];
  const S inv_pi = S(1) / K<S>::pi();
  S w_sum = S(0);
  for (int c = 0; c < 3; ++c) {
    const S a_pi = albedo3[c] * inv_pi;
    const S gc = g3[c] * n_mat;                           // the reference sums the diffuse term over the K axis
    w_sum += gc * amp[c] * a_pi;
    acc7[4 + c] += gc * a_pi * phi.v;
    g_albedo[c] += gc * amp[c] * phi.v * inv_pi;
  }
  for (int i = 0; i < 3; ++i) acc7[i] += w_sum * phi.d[i];
  acc7[3] += w_sum * phi.d[3];
}

}  // namespace sgb
}  // namespace nefii
