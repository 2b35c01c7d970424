// Host-side emulation of csrc/tracer_math.cuh (TEST INFRASTRUCTURE ONLY): the per-ray march state machine driven in rounds
// exactly as csrc/tracer.cu drives it (every ray advances as soon as the values it asked for are there), with the analytic
// test SDF.  tests/test_hostemu_tracer.py compares it bit for bit with oracle/tracer.py's lock-step sphere_tracing.
#include <vector>
#include "tracer_math.cuh"

using namespace nefii::trm;

extern "C" {

// o, d [n,3]; outputs acc_s, acc_e, min_dis, max_dis [n], flags [n] (F_HIT | F_UNF_S | F_UNF_E), stats[0] = SDF evaluations,
// stats[1] = rounds with work
void emu_sphere_trace(int n, const float* o, const float* d, const float* prims, int n_prims, float radius, float thr,
                      float line_search_step, int ls_iters, int max_iters, float* acc_s, float* acc_e, float* min_dis,
                      float* max_dis, unsigned char* flags, long long* stats) {
  std::vector<March> M(n);
  std::vector<int> req(n);
  const float back0 = (float)(1.0 - (double)line_search_step);
  long long evals = 0, rounds = 0;
  bool any = false;
  for (int r = 0; r < n; ++r) {
    float t0, t1;
    const bool hits = sphere_intersection(o + 3 * r, d + 3 * r, radius, t0, t1);
    req[r] = march_begin(M[r], hits, t0, t1);
    min_dis[r] = M[r].acc_s; max_dis[r] = M[r].acc_e;
    any = any || req[r];
  }
  while (any) {
    ++rounds;
    any = false;
    for (int r = 0; r < n; ++r) {
      if (!req[r]) continue;
      float p[3];
      if (req[r] & REQ_S) { along(o + 3 * r, M[r].acc_s, d + 3 * r, p); M[r].nxt_s = analytic_sdf(prims, n_prims, p); ++evals; }
      if (req[r] & REQ_E) { along(o + 3 * r, M[r].acc_e, d + 3 * r, p); M[r].nxt_e = analytic_sdf(prims, n_prims, p); ++evals; }
      req[r] = march_advance(M[r], thr, back0, ls_iters, max_iters);
      any = any || req[r];
    }
  }
  for (int r = 0; r < n; ++r) { acc_s[r] = M[r].acc_s; acc_e[r] = M[r].acc_e; flags[r] = M[r].flags; }
  stats[0] = evals; stats[1] = rounds;
}

void emu_analytic_sdf(int n, const float* x, const float* prims, int n_prims, float* out) {
  for (int i = 0; i < n; ++i) out[i] = analytic_sdf(prims, n_prims, x + 3 * i);
}

}
