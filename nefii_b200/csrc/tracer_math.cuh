// Per-ray arithmetic of the IDR sphere tracer, shared by the CUDA kernels (csrc/tracer.cu) and the host emulation
// used in the CPU tests (tests/hostemu/tracer_emu.cpp).
//
// Reference: code/model/ray_tracing.py:104-193 (sphere_tracing), code/utils/rend_util.py:200-221
// (get_sphere_intersection), :90-142 (get_camera_params / lift).  The reference marches all rays in lock step; its
// loop exits are batch-wide but every update of a finished ray is a no-op, so a ray can run through ITS OWN sequence
// of iterations and line-search steps: `march_advance` is that per-ray state machine.  It is called each time all the
// SDF values a ray has asked for are available and returns the next request(s).
//
// Every value that feeds a comparison is a rounded multiply followed by a rounded add, as torch evaluates it: the
// including translation units are built with -fmad=false / -ffp-contract=off.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define NEFII_TRM_HD __host__ __device__ __forceinline__
#else
#define NEFII_TRM_HD inline
#endif

namespace nefii {
namespace trm {

enum : unsigned char { F_HIT = 1, F_UNF_S = 2, F_UNF_E = 4, F_PEND_S = 8, F_PEND_E = 16, F_NET = 32, F_SAMP = 64, F_MIN = 128 };
enum { REQ_S = 1, REQ_E = 2 };

struct March {
  float acc_s, acc_e;     // accumulated distances of the two marching ends
  float cur_s, cur_e;     // step taken in the current iteration (0 for a finished end)
  float nxt_s, nxt_e;     // SDF at the points reached by that step
  unsigned char flags;    // F_HIT | F_UNF_*
  unsigned char iter;     // iterations started (reference: `iters`)
  unsigned char ls;       // line-search steps taken in this iteration (reference: `not_proj_iters`)
};

NEFII_TRM_HD float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

NEFII_TRM_HD void along(const float* o, float t, const float* d, float* p) {
  p[0] = o[0] + t * d[0];
  p[1] = o[1] + t * d[1];
  p[2] = o[2] + t * d[2];
}

// rend_util.get_sphere_intersection for one ray: t0 <= t1 clamped at 0.01; returns mask_intersect
NEFII_TRM_HD bool sphere_intersection(const float* o, const float* d, float radius, float& t0, float& t1) {
  const float b = dot3(d, o);
  const float onorm = sqrtf((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]);
  const float under = b * b - (onorm * onorm - radius * radius);
  const bool hits = under > 0.f;
  t0 = 0.f; t1 = 0.f;
  if (hits) {
    const float root = sqrtf(under);
    t0 = root * -1.0f - b;
    t1 = root * 1.0f - b;
    t0 = fmaxf(t0, 0.01f);
    t1 = fmaxf(t1, 0.01f);
  }
  return hits;
}

// ray_tracing.py:112-134: both ends start on the bounding sphere and ask for their first SDF value
NEFII_TRM_HD int march_begin(March& m, bool hits, float t0, float t1) {
  m.acc_s = t0; m.acc_e = t1;
  m.cur_s = m.cur_e = m.nxt_s = m.nxt_e = 0.f;
  m.flags = hits ? (unsigned char)(F_HIT | F_UNF_S | F_UNF_E) : (unsigned char)0;
  m.iter = 0; m.ls = 0;
  return hits ? (REQ_S | REQ_E) : 0;
}

// One transition of the per-ray state machine (ray_tracing.py:136-191).  m.nxt_* hold the SDF values of the last
// request.  back0 = 1 - line_search_step.  Returns the ends whose new position (m.acc_*) needs an SDF value, 0 when the
// ray has left the march (converged, crossed, or out of iterations).
NEFII_TRM_HD int march_advance(March& m, float thr, float back0, int ls_iters, int max_iters) {
  bool unf_s = (m.flags & F_UNF_S) != 0, unf_e = (m.flags & F_UNF_E) != 0;
  if (m.iter > 0) {
    // :170-188 step back where the march crossed the surface
    const bool bad_s = m.nxt_s < 0.f, bad_e = m.nxt_e < 0.f;
    if ((bad_s || bad_e) && (int)m.ls < ls_iters) {
      const float back = ldexpf(back0, -(int)m.ls);
      int req = 0;
      if (bad_s) { m.acc_s = m.acc_s - back * m.cur_s; req |= REQ_S; }
      if (bad_e) { m.acc_e = m.acc_e + back * m.cur_e; req |= REQ_E; }
      m.ls = (unsigned char)(m.ls + 1);
      return req;
    }
    // :190-191 end of the iteration
    const bool crossed = m.acc_s < m.acc_e;
    unf_s = unf_s && crossed;
    unf_e = unf_e && crossed;
  }
  // :138-150 head of the next iteration
  m.cur_s = unf_s ? m.nxt_s : 0.f;
  if (m.cur_s <= thr) m.cur_s = 0.f;
  m.cur_e = unf_e ? m.nxt_e : 0.f;
  if (m.cur_e <= thr) m.cur_e = 0.f;
  unf_s = unf_s && (m.cur_s > thr);
  unf_e = unf_e && (m.cur_e > thr);
  m.flags = (unsigned char)((m.flags & ~(F_UNF_S | F_UNF_E)) | (unf_s ? F_UNF_S : 0) | (unf_e ? F_UNF_E : 0));
  if (!(unf_s || unf_e) || (int)m.iter == max_iters) return 0;
  // :152-168 make the step
  m.iter = (unsigned char)(m.iter + 1);
  m.ls = 0;
  m.acc_s = m.acc_s + m.cur_s;
  m.acc_e = m.acc_e - m.cur_e;
  m.nxt_s = 0.f; m.nxt_e = 0.f;
  return (unf_s ? REQ_S : 0) | (unf_e ? REQ_E : 0);
}

// Analytic test scene (oracle/tracer.py analytic_sdf): union of spheres and boxes, fixed operation order.
// prims: [n, 8] rows = kind (0 sphere, 1 box), centre xyz, radius | half extents, pad.
NEFII_TRM_HD float analytic_sdf(const float* prims, int n_prims, const float* p) {
  float best = 0.f;
  for (int k = 0; k < n_prims; ++k) {
    const float* P = prims + k * 8;
    const float q0 = p[0] - P[1], q1 = p[1] - P[2], q2 = p[2] - P[3];
    float val;
    if (P[0] == 0.f) {
      val = sqrtf((q0 * q0 + q1 * q1) + q2 * q2) - P[4];
    } else {
      const float a0 = fabsf(q0) - P[4], a1 = fabsf(q1) - P[5], a2 = fabsf(q2) - P[6];
      const float m0 = fmaxf(a0, 0.f), m1 = fmaxf(a1, 0.f), m2 = fmaxf(a2, 0.f);
      const float outside = sqrtf((m0 * m0 + m1 * m1) + m2 * m2);
      const float inside = fminf(fmaxf(a0, fmaxf(a1, a2)), 0.f);
      val = outside + inside;
    }
    best = (k == 0) ? val : fminf(best, val);
  }
  return best;
}

// get_camera_params + lift (rend_util.py:90-142) for one pixel: unit world direction of the ray through uv.
// pose: row-major 4x4 camera-to-world; K: row-major 4x4 intrinsics.  Operation order of the reference:
//   x_lift = (x - cx + cy*sk/fy - sk*y/fy) / fx * z ; y_lift = (y - cy) / fy * z ; z = 1
//   world = pose @ [x_lift, y_lift, z, 1] (torch.bmm: sums in index order) ; dir = normalize(world - cam_loc)
// F.normalize divides by max(||v||, 1e-12).
NEFII_TRM_HD void camera_ray(const float* pose, const float* K, float u, float v, float* dir) {
  const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
  const float z = 1.0f;
  const float x_lift = (u - cx + cy * sk / fy - sk * v / fy) / fx * z;
  const float y_lift = (v - cy) / fy * z;
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* P = pose + 4 * r;
    w[r] = ((P[0] * x_lift + P[1] * y_lift) + P[2] * z) + P[3] * 1.0f;
  }
  const float c0 = w[0] - pose[3], c1 = w[1] - pose[7], c2 = w[2] - pose[11];
  const float nrm = sqrtf((c0 * c0 + c1 * c1) + c2 * c2);
  const float den = fmaxf(nrm, 1e-12f);
  dir[0] = c0 / den; dir[1] = c1 / den; dir[2] = c2 / den;
}

}  // namespace trm
}  // namespace nefii
