// CUDA-graph front end of the tracer: `ray_trace` captures the launch sequence of ray_trace_enqueue once per argument tuple
// (shapes, pointers, tracer configuration, accuracy settings) and replays it afterwards.  The loops of a trace -- march
// rounds, sampler chunks, bisection steps, min-SDF chunks -- become conditional WHILE nodes whose condition the last CTA of
// the loop's own kernels sets from the device-side control block (cudaGraphSetConditional), so a trace costs ONE host call,
// runs exactly as many rounds as its slowest ray needs and never waits for the host.
//
// Replaces the reference's Python while-loops with `mask.sum() > 0` host round trips (code/model/ray_tracing.py:136-191,
// 213-216, 264-277, 328-330).
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <vector>
#include "mlp_gemm.cuh"
#include "tracer.cuh"

namespace nefii {

namespace {

int env_graph_mode() {
  const char* e = getenv("NEFII_TRACE_GRAPH");
  return (e && e[0] == '0') ? 0 : 1;
}
int g_graph_mode = env_graph_mode();

struct Key {
  TraceConfig cfg;
  const void* net; const void* prims; int n_prims;
  int n_batch, n_pix, flags, device;
  const void *cam_loc, *ray_dirs, *object_mask, *linspace, *uniforms, *workspace, *points, *hit, *dists;
  size_t ws_bytes;
  long long gemm_epoch;
  int march_flush, bulk_flush, quad_rows, bisect_depth, grid_cap, pad_;
};

struct Entry {
  Key key;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;       // kernels enqueued during capture (one trip per loop)
};

std::mutex g_mu;
std::list<Entry> g_cache;        // most recently used first
long long g_captures = 0;        // graphs captured so far (a capture + instantiate costs tens of milliseconds)
constexpr size_t kMaxGraphs = 24;

struct DevStreams { cudaStream_t cap = nullptr, body = nullptr; };
DevStreams g_streams[64];

bool same_key(const Key& a, const Key& b) { return memcmp(&a, &b, sizeof(Key)) == 0; }

void destroy(Entry& e) {
  if (e.exec) cudaGraphExecDestroy(e.exec);
  if (e.graph) cudaGraphDestroy(e.graph);
  e.exec = nullptr; e.graph = nullptr;
}

class GraphLoops : public TraceLoops {
 public:
  explicit GraphLoops(cudaStream_t body) : body_(body) {}
  int error() const { return err_; }

  unsigned long long next_handle(cudaStream_t stream) override {
    cudaStreamCaptureStatus st;
    cudaGraph_t g = nullptr;
    if (cudaStreamGetCaptureInfo(stream, &st, nullptr, &g, nullptr, nullptr) != cudaSuccess || st != cudaStreamCaptureStatusActive || !g) {
      err_ = set_error(NEFII_ERR_CUDA, "trace graph: stream is not capturing");
      return 0ull;
    }
    cudaGraphConditionalHandle h;
    const cudaError_t e = cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) {
      err_ = set_error(NEFII_ERR_CUDA, "cudaGraphConditionalHandleCreate failed: %s", cudaGetErrorString(e));
      return 0ull;
    }
    pending_ = (unsigned long long)h;
    return pending_;
  }

  int begin(cudaStream_t stream, cudaStream_t* body_stream, unsigned long long* handle) override {
    if (err_) return err_;
    NEFII_CHECK_ARG(pending_ != 0ull, "trace graph: loop without a condition handle");
    cudaStreamCaptureStatus st;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    NEFII_CUDA(cudaStreamGetCaptureInfo(stream, &st, nullptr, &g, &deps, &n_deps));
    NEFII_CHECK_ARG(st == cudaStreamCaptureStatusActive && g, "trace graph: stream is not capturing");
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = (cudaGraphConditionalHandle)pending_;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    NEFII_CUDA(cudaGraphAddNode(&node, g, deps, n_deps, &p));
    NEFII_CUDA(cudaStreamUpdateCaptureDependencies(stream, &node, 1, cudaStreamSetCaptureDependencies));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    NEFII_CUDA(cudaStreamBeginCaptureToGraph(body_, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    *body_stream = body_;
    *handle = pending_;
    pending_ = 0ull;
    return NEFII_OK;
  }

  int end(cudaStream_t, cudaStream_t body_stream) override {
    cudaGraph_t g = nullptr;
    NEFII_CUDA(cudaStreamEndCapture(body_stream, &g));
    return NEFII_OK;
  }

 private:
  cudaStream_t body_;
  unsigned long long pending_ = 0ull;
  int err_ = 0;
};

}  // namespace

int trace_set_graph_mode(int mode) {
  NEFII_CHECK_ARG(mode == 0 || mode == 1, "trace_set_graph_mode: 0 (fixed schedule) or 1 (CUDA graph)");
  g_graph_mode = mode;
  return NEFII_OK;
}
int trace_graph_mode() { return g_graph_mode; }

int trace_graph_clear() {
  std::lock_guard<std::mutex> lock(g_mu);
  for (auto& e : g_cache) destroy(e);
  g_cache.clear();
  return NEFII_OK;
}

long long trace_graph_captures() {
  std::lock_guard<std::mutex> lock(g_mu);
  return g_captures;
}

int ray_trace(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
              const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
              const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
              unsigned char* hit, float* dists, long long* stats) {
  const long long n_rays_ll = (long long)n_batch * n_pix;
  if (stats) for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (n_rays_ll <= 0) return NEFII_OK;
  int rc;
  if ((rc = gemm_prepare_device())) return rc;
  if (!g_graph_mode || gemm_profile_active()) {
    if ((rc = ray_trace_enqueue(stream, cfg, src, n_batch, n_pix, cam_loc, ray_dirs, object_mask, flags, linspace, uniforms,
                                workspace, ws_bytes, points, hit, dists, nullptr)))
      return rc;
    return trace_read_stats(stream, src, (int)n_rays_ll, cfg.n_steps, workspace, stats);
  }

  int device = 0;
  NEFII_CUDA(cudaGetDevice(&device));
  NEFII_CHECK_ARG(device >= 0 && device < 64, "ray_trace: device index out of range");
  Key key;
  memset(&key, 0, sizeof(key));
  key.cfg = cfg; key.net = src.net; key.prims = src.prims; key.n_prims = src.n_prims;
  key.n_batch = n_batch; key.n_pix = n_pix; key.flags = flags; key.device = device;
  key.cam_loc = cam_loc; key.ray_dirs = ray_dirs; key.object_mask = object_mask; key.linspace = linspace; key.uniforms = uniforms;
  key.workspace = workspace; key.points = points; key.hit = hit; key.dists = dists; key.ws_bytes = ws_bytes;
  key.gemm_epoch = gemm_config_epoch();
  const TraceTiers tiers = trace_tiers();
  key.march_flush = tiers.march_flush; key.bulk_flush = tiers.bulk_flush; key.quad_rows = trace_quad_rows(); key.bisect_depth = trace_bisect_depth();
  key.grid_cap = gemm_grid_cap();

  std::lock_guard<std::mutex> lock(g_mu);
  Entry* found = nullptr;
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it) {
    if (same_key(it->key, key)) {
      g_cache.splice(g_cache.begin(), g_cache, it);
      found = &g_cache.front();
      break;
    }
  }
  if (!found) {
    DevStreams& ds = g_streams[device];
    if (!ds.cap) {
      NEFII_CUDA(cudaStreamCreateWithFlags(&ds.cap, cudaStreamNonBlocking));
      NEFII_CUDA(cudaStreamCreateWithFlags(&ds.body, cudaStreamNonBlocking));
    }
    Entry e;
    e.key = key;
    const long long l0 = launches();
    NEFII_CUDA(cudaStreamBeginCapture(ds.cap, cudaStreamCaptureModeThreadLocal));
    GraphLoops loops(ds.body);
    rc = ray_trace_enqueue(ds.cap, cfg, src, n_batch, n_pix, cam_loc, ray_dirs, object_mask, flags, linspace, uniforms, workspace,
                           ws_bytes, points, hit, dists, &loops);
    if (!rc) rc = loops.error();
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ds.cap, &g);
    if (rc) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      return rc;
    }
    if (ce != cudaSuccess) {
      cudaGetLastError();
      return set_error(NEFII_ERR_CUDA, "trace graph: cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
    }
    e.graph = g;
    e.launches = launches() - l0;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, g, 0);
    if (ie != cudaSuccess) {
      cudaGraphDestroy(g);
      cudaGetLastError();
      return set_error(NEFII_ERR_CUDA, "trace graph: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    }
    add_launches(-e.launches);      // the capture itself launched nothing; every replay is counted below
    ++g_captures;
    g_cache.push_front(e);
    while (g_cache.size() > kMaxGraphs) {
      destroy(g_cache.back());
      g_cache.pop_back();
    }
    found = &g_cache.front();
  }
  NEFII_CUDA(cudaGraphLaunch(found->exec, stream));
  add_launches(found->launches);
  return trace_read_stats(stream, src, (int)n_rays_ll, cfg.n_steps, workspace, stats);
}

}  // namespace nefii
