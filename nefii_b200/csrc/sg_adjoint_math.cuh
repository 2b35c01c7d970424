// Reverse-mode (hand-derived adjoint) of the render_with_sg term math in sg_math.cuh -- what autograd does through the
// reference's code/model/sg_render.py:112-158 (hemisphere_int, lambda_trick) and :199-295 (warped GGX lobe, the cosine-lobe
// integrals).  One scalar function carries every (ray, light SG, base material) term:
//
//   specular term[c] = amp_light[c] * amp_brdf[c] * Phi,     Phi = E(a, lambda, b, beta) * Psi(n, p, sigma)
//   diffuse  term[c] = amp_light[c] * albedo[c] / pi * Psi(n, a, lambda)
//
// (a, lambda: light axis / sharpness; b, beta: warped BRDF lobe; (p, sigma, E) = their SG product; Psi = the cosine-lobe
// integral, channel independent), so ONE reverse sweep per term yields the gradients w.r.t. the light SG, the BRDF lobe and
// the normal.  The BRDF lobe's own dependence on (normal, roughness, specular reflectance) is swept once per (ray, material)
// after the lobe adjoints have been summed over the light SGs (brdf_lobe_vjp).
//
// Shared by the CUDA kernel (float) and the host emulation (double, tests/hostemu/sg_emu.cpp: checked against autograd through
// the oracle).  Values are recomputed with the forward's formulas; torch's subgradient conventions are followed: clamp(min=0)
// passes the gradient where x >= 0, torch.min(scale, ratio + 1) to the smaller argument.
#pragma once
#include "sg_math.cuh"

namespace nefii {
namespace sga {

using sgm::K;

// H(sharp, c) = hemi_eval(hemi_coef(sharp), c) with dH/dsharp and dH/dc  (sg_render.py:112-137)
template <typename S> NEFII_HD void hemi_vg(S sharp_in, S c, S& H, S& dH_dsharp, S& dH_dc) {
  const S s = sharp_in + K<S>::eps();
  const S inv = S(1) / s;
  const S rt = sgm::m_sqrt(s);
  const S P = S(1.6988) + S(10.8438) * inv;
  const S Q = (S(1) + S(6.2201) * inv) + (S(10.2415) * inv) * inv;
  const S t = rt * P / Q;
  const S dP = -S(10.8438) * inv * inv;
  const S dQ = -(S(6.2201) + S(2) * S(10.2415) * inv) * inv * inv;
  const S dt = t * (S(0.5) * inv) + rt * (dP * Q - P * dQ) / (Q * Q);
  const S ea = sgm::m_exp(-t);
  const S rs = inv * K<S>::two_pi();
  const S e1 = sgm::m_exp(-s), e2 = sgm::m_exp(S(-2) * s);
  const S lower = rs * (e1 - e2), upper = rs * (S(1) - e1);
  const S dlower = -lower * inv + rs * (S(2) * e2 - e1);
  const S dupper = -upper * inv + rs * e1;
  S Sv, dS_dt, dS_dc;
  if (c >= S(0)) {
    const S eb = sgm::m_exp(-t * c);
    const S u = ea * eb;
    const S D = ((S(1) - ea) + eb) - u;
    Sv = (S(1) - u) / D;
    const S u_t = -(S(1) + c) * u;
    const S D_t = (ea - c * eb) - u_t;
    dS_dt = (-u_t * D - (S(1) - u) * D_t) / (D * D);
    const S u_c = -t * u;
    const S D_c = -t * eb - u_c;
    dS_dc = (-u_c * D - (S(1) - u) * D_c) / (D * D);
  } else {
    const S b = sgm::m_exp(t * c);
    const S Nn = b - ea;
    const S Dn = (S(1) - ea) * (b + S(1));
    Sv = Nn / Dn;
    const S N_t = c * b + ea;
    const S D_t = ea * (b + S(1)) + (S(1) - ea) * c * b;
    dS_dt = (N_t * Dn - Nn * D_t) / (Dn * Dn);
    const S N_c = t * b;
    const S D_c = (S(1) - ea) * t * b;
    dS_dc = (N_c * Dn - Nn * D_c) / (Dn * Dn);
  }
  H = lower * (S(1) - Sv) + upper * Sv;
  dH_dsharp = dlower * (S(1) - Sv) + dupper * Sv + (upper - lower) * dS_dt * dt;
  dH_dc = (upper - lower) * dS_dc;
}

// lambda_trick (sg_render.py:141-158) with what its reverse sweep needs
template <typename S> struct SgProd {
  S axis[3], sharp, E;     // outputs
  S s, cosang;             // scale (after the min), a1 . a2
  bool clamped;            // the min() took ratio + 1
};

template <typename S> NEFII_HD void sg_prod_fwd(S ratio, const S* a1, const S* a2, S sharp2, SgProd<S>& o) {
  o.cosang = sgm::dot3(a1, a2);
  const S s0 = sgm::m_sqrt((ratio * ratio + S(1)) + (S(2) * ratio) * o.cosang);
  o.clamped = !(s0 <= ratio + S(1));
  o.s = o.clamped ? ratio + S(1) : s0;
  o.sharp = sharp2 * o.s;
  const S w1 = ratio / o.s, w2 = S(1) / o.s;
  for (int i = 0; i < 3; ++i) o.axis[i] = w1 * a1[i] + w2 * a2[i];
  o.E = sgm::m_exp(sharp2 * ((o.s - ratio) - S(1)));
}

// adjoints of (axis, sharp, E) -> accumulated into (ratio_bar, a1_bar, a2_bar, sharp2_bar)
template <typename S>
NEFII_HD void sg_prod_bwd(S ratio, const S* a1, const S* a2, S sharp2, const SgProd<S>& o, const S* axis_bar, S sharp_bar, S E_bar,
                          S& ratio_bar, S* a1_bar, S* a2_bar, S& sharp2_bar) {
  const S shift_bar = E_bar * o.E;
  const S w1_bar = sgm::dot3(axis_bar, a1), w2_bar = sgm::dot3(axis_bar, a2);
  const S inv_s = S(1) / o.s;
  const S s_bar = (shift_bar + sharp_bar) * sharp2 - (w1_bar * ratio + w2_bar) * inv_s * inv_s;
  sharp2_bar += shift_bar * ((o.s - ratio) - S(1)) + sharp_bar * o.s;
  ratio_bar += w1_bar * inv_s - shift_bar * sharp2;
  const S w1 = ratio * inv_s, w2 = inv_s;
  S c_bar = S(0);
  if (o.clamped) {
    ratio_bar += s_bar;
  } else {
    ratio_bar += s_bar * (ratio + o.cosang) * inv_s;
    c_bar = s_bar * ratio * inv_s;
  }
  for (int i = 0; i < 3; ++i) {
    a1_bar[i] += w1 * axis_bar[i] + c_bar * a2[i];
    a2_bar[i] += w2 * axis_bar[i] + c_bar * a1[i];
  }
}

// Psi(n, p, sigma) = cos_mu e2 H(tau, q . n) - cos_alpha H(sigma, p . n)   ((q, tau, e2) = lambda_trick(n, cos lobe; p, sigma)),
// the channel-independent part of cosine_lobe_integral (sg_render.py:243-252).  Returns Psi; with psi_bar != 0 accumulates
// psi_bar * dPsi/d(n, p, sigma).
template <typename S> NEFII_HD S psi_vjp(const S* n, const S* p, S sigma, S psi_bar, S* n_bar, S* p_bar, S& sigma_bar) {
  const S r2 = (S(1) / sigma) * K<S>::cos_lambda();
  SgProd<S> q;
  sg_prod_fwd(r2, n, p, sigma, q);
  const S c1 = sgm::dot3(q.axis, n), c2 = sgm::dot3(p, n);
  S h1, h1_s, h1_c, h2, h2_s, h2_c;
  hemi_vg(q.sharp, c1, h1, h1_s, h1_c);
  hemi_vg(sigma, c2, h2, h2_s, h2_c);
  const S psi = (K<S>::cos_mu() * q.E) * h1 - K<S>::cos_alpha() * h2;
  const S e2_bar = psi_bar * K<S>::cos_mu() * h1;
  const S h1_bar = psi_bar * K<S>::cos_mu() * q.E;
  const S h2_bar = -psi_bar * K<S>::cos_alpha();
  const S tau_bar = h1_bar * h1_s, c1_bar = h1_bar * h1_c, c2_bar = h2_bar * h2_c;
  sigma_bar += h2_bar * h2_s;
  S q_bar[3];
  for (int i = 0; i < 3; ++i) {
    q_bar[i] = c1_bar * n[i];
    n_bar[i] += c1_bar * q.axis[i] + c2_bar * p[i];
    p_bar[i] += c2_bar * n[i];
  }
  S r2_bar = S(0);
  sg_prod_bwd(r2, n, p, sigma, q, q_bar, tau_bar, e2_bar, r2_bar, n_bar, p_bar, sigma_bar);
  sigma_bar += -r2_bar * r2 / sigma;
  return psi;
}

// One specular term.  a, lambda: light lobe (unit parametrisation); b, beta: BRDF lobe; w = sum_c g_c amp_light_c amp_brdf_c.
// Returns Phi; accumulates w * dPhi/d(a, lambda, b, beta, n).
template <typename S>
NEFII_HD S specular_phi_vjp(const S* n, const S* a, S lambda, const S* b, S beta, S w, S* a_bar, S& lambda_bar, S* b_bar, S& beta_bar,
                            S* n_bar) {
  const S r = lambda / beta;
  SgProd<S> pr;
  sg_prod_fwd(r, a, b, beta, pr);
  S p_bar[3] = {S(0), S(0), S(0)};
  S sigma_bar = S(0);
  const S psi = psi_vjp(n, pr.axis, pr.sharp, w * pr.E, n_bar, p_bar, sigma_bar);
  S r_bar = S(0);
  sg_prod_bwd(r, a, b, beta, pr, p_bar, sigma_bar, w * psi, r_bar, a_bar, b_bar, beta_bar);
  lambda_bar += r_bar / beta;
  beta_bar += -r_bar * r / beta;
  return pr.E * psi;
}

// Reverse sweep of make_brdf_lobe (sg_render.py:199-237): adjoints of the lobe (axis, sharp, amp[3]) -> n_bar (accumulated),
// d/d roughness, d/d specular reflectance[3].  The view direction carries no gradient on this path.
template <typename S>
NEFII_HD void brdf_lobe_vjp(const S* n, const S* v, S rough, const S* spec3, const S* axis_bar, S sharp_bar, const S* amp_bar, S* n_bar,
                            S& rough_bar, S* spec_bar) {
  const S eps = K<S>::eps();
  const S inv_r4 = S(1) / (((rough * rough) * rough) * rough);
  const S b_sharp = S(2) * inv_r4;
  const S b_amp = inv_r4 * (S(1) / K<S>::pi());
  const S nv0 = sgm::dot3(n, v);
  const S nv = sgm::clamp_min(nv0, S(0));
  S wv[3], b[3];
  for (int i = 0; i < 3; ++i) wv[i] = (S(2) * nv) * n[i] - v[i];
  const S nw = sgm::norm3(wv), dw = nw + eps;
  for (int i = 0; i < 3; ++i) b[i] = wv[i] / dw;
  const S dsh = S(4) * nv + eps;
  const S beta = b_sharp / dsh;
  S h[3], hu[3];
  for (int i = 0; i < 3; ++i) h[i] = b[i] + v[i];
  const S nh = sgm::norm3(h), dh = nh + eps;
  for (int i = 0; i < 3; ++i) hu[i] = h[i] / dh;
  const S vh0 = sgm::dot3(v, hu);
  const S vh = sgm::clamp_min(vh0, S(0));
  const S fexp = sgm::m_pow(S(2), -(S(5.55473) * vh + S(6.8316)) * vh);
  const S d10 = sgm::dot3(b, n), d20 = sgm::dot3(v, n);
  const S d1 = sgm::clamp_min(d10, S(0)), d2 = sgm::clamp_min(d20, S(0));
  const S k = ((rough + S(1)) * (rough + S(1))) * S(0.125);
  const S den1 = (d1 * (S(1) - k) + k) + eps, den2 = (d2 * (S(1) - k) + k) + eps;
  const S g1 = d1 / den1, g2 = d2 / den2;
  const S g = g1 * g2;
  const S den = (S(4) * d1) * d2 + eps;
  // ---- reverse ----
  S b_amp_bar = S(0), g_bar = S(0), den_bar = S(0), fexp_bar = S(0);
  for (int c = 0; c < 3; ++c) {
    const S f = spec3[c] + (S(1) - spec3[c]) * fexp;
    const S moi = (f * g) / den;
    b_amp_bar += amp_bar[c] * moi;
    const S moi_bar = amp_bar[c] * b_amp;
    const S f_bar = moi_bar * g / den;
    g_bar += moi_bar * f / den;
    den_bar += -moi_bar * moi / den;
    spec_bar[c] += f_bar * (S(1) - fexp);
    fexp_bar += f_bar * (S(1) - spec3[c]);
  }
  S b_bar[3] = {axis_bar[0], axis_bar[1], axis_bar[2]};
  // fexp = 2^x, x = -(5.55473 vh + 6.8316) vh
  const S vh_bar = fexp_bar * fexp * S(0.6931471805599453) * (-(S(2) * S(5.55473) * vh + S(6.8316)));
  const S vh0_bar = (vh0 >= S(0)) ? vh_bar : S(0);
  {  // hu = h / (|h| + eps), h = b + v
    S hu_bar[3] = {vh0_bar * v[0], vh0_bar * v[1], vh0_bar * v[2]};
    const S hd = sgm::dot3(hu_bar, h);
    const S coef = nh > S(0) ? hd / (nh * dh * dh) : S(0);
    for (int i = 0; i < 3; ++i) b_bar[i] += hu_bar[i] / dh - h[i] * coef;
  }
  const S kk = k + eps;
  S d1_bar = g_bar * g2 * kk / (den1 * den1) + den_bar * S(4) * d2;
  S d2_bar = g_bar * g1 * kk / (den2 * den2) + den_bar * S(4) * d1;
  const S k_bar = g_bar * (g2 * (-d1 * (S(1) - d1)) / (den1 * den1) + g1 * (-d2 * (S(1) - d2)) / (den2 * den2));
  const S d10_bar = (d10 >= S(0)) ? d1_bar : S(0);
  const S d20_bar = (d20 >= S(0)) ? d2_bar : S(0);
  for (int i = 0; i < 3; ++i) {
    b_bar[i] += d10_bar * n[i];
    n_bar[i] += d10_bar * b[i] + d20_bar * v[i];
  }
  // beta = b_sharp / (4 nv + eps)
  const S b_sharp_bar = sharp_bar / dsh;
  S nv_bar = -sharp_bar * beta * S(4) / dsh;
  {  // b = wv / (|wv| + eps), wv = 2 nv n - v
    const S bd = sgm::dot3(b_bar, wv);
    const S coef = nw > S(0) ? bd / (nw * dw * dw) : S(0);
    S wv_bar[3];
    for (int i = 0; i < 3; ++i) wv_bar[i] = b_bar[i] / dw - wv[i] * coef;
    nv_bar += S(2) * sgm::dot3(wv_bar, n);
    for (int i = 0; i < 3; ++i) n_bar[i] += (S(2) * nv) * wv_bar[i];
  }
  const S nv0_bar = (nv0 >= S(0)) ? nv_bar : S(0);
  for (int i = 0; i < 3; ++i) n_bar[i] += nv0_bar * v[i];
  const S inv_r4_bar = b_sharp_bar * S(2) + b_amp_bar * (S(1) / K<S>::pi());
  rough_bar += inv_r4_bar * (S(-4) * inv_r4 / rough) + k_bar * (rough + S(1)) * S(0.25);
}

}  // namespace sga
}  // namespace nefii
