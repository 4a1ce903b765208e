// C-ABI plumbing shared by every entry point: thread-local error string, launch counter, device probe.
#include "nn.cuh"
#include <atomic>
#include <list>
#include <map>
#include <unordered_map>
#include <mutex>
#include <string>
#include <stdlib.h>
#include <string.h>
#include <vector>

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void qeb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void qeb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

QEB_API const char* qeb_last_error(void) { return g_err; }

QEB_API int qeb_abi_version(void) { return 1; }

QEB_API long long qeb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
long long qeb_launch_count_now() { return g_launches.load(std::memory_order_relaxed); }

QEB_API void qeb_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

// 0 when the current device is an sm_100 part this library was compiled for
QEB_API int qeb_check_device(void) {
  int dev = 0;
  QEB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  QEB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    qeb_set_error("qeb kernels are built for sm_100a only; device %d is sm_%d%d (%s)", dev, p.major, p.minor, p.name);
    return QEB_ERR_UNSUPPORTED;
  }
  return QEB_OK;
}

// ---- per-kernel profiling (bench.py's roofline leg): CUDA events around every launch of the library, recorded on the
// launching stream, with the launch site's algorithmic FLOPs / bytes. Off by default (two event records per launch).
namespace {
struct ProfRec {
  std::string tag;
  cudaEvent_t e0, e1;
  double flops, bytes;
};
std::atomic<int> g_prof{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_recs;
}  // namespace

int qeb_prof_on() { return g_prof.load(std::memory_order_relaxed); }

int qeb_prof_begin(const char* tag, cudaStream_t st, double flops, double bytes) {
  ProfRec r;
  r.tag = tag; r.flops = flops; r.bytes = bytes;
  if (g_prof.load(std::memory_order_relaxed) == 2) {  // detail mode: one bucket per (tag, work size)
    char suffix[64];
    snprintf(suffix, sizeof(suffix), "[%.3gGF,%.3gMB]", flops * 1e-9, bytes * 1e-6);
    r.tag += suffix;
  }
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_recs.push_back(r);
  return (int)g_recs.size() - 1;
}

void qeb_prof_end(int idx, cudaStream_t st) {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[idx].e1, st);
}

QEB_API void qeb_prof_enable(int on) { g_prof.store(on, std::memory_order_relaxed); }

// Synchronises, aggregates the records per tag and writes a JSON object
// {"tag": {"launches": n, "ms": total, "flops": total, "bytes": total}, ...} into buf; clears the records.
QEB_API int qeb_prof_report(char* buf, int cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    Agg& a = agg[r.tag];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += line;
    first = false;
  }
  out += "}";
  if ((int)out.size() + 1 > cap) {
    qeb_set_error("prof_report: buffer too small (%d needed)", (int)out.size() + 1);
    return QEB_ERR_INVALID;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return QEB_OK;
}

// ---- side stream (nn.cuh)
namespace {
struct SideRes {
  int device = -1;
  cudaStream_t side = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr, mark_ev = nullptr, mark2_ev = nullptr;
};
thread_local SideRes g_side;
}  // namespace

int SideStream::init(cudaStream_t main_stream) {
  main = main_stream;
  static const bool want = !(getenv("QEB_SIDE_STREAM") && atoi(getenv("QEB_SIDE_STREAM")) == 0);
  enabled = false;
  // per-kernel profiling (qeb_prof_enable) times every launch with events on its own stream: keep the launches serialised
  // then, like ncu does, so that a kernel's duration is its own and not its share of an SM it had to split
  if (!want || qeb_prof_on()) return QEB_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(main_stream, &cap) != cudaSuccess) { cudaGetLastError(); return QEB_OK; }
  int dev = 0;
  QEB_CUDA(cudaGetDevice(&dev));
  if (g_side.device != dev) {
    if (g_side.side) {
      cudaStreamDestroy(g_side.side); cudaEventDestroy(g_side.fork_ev); cudaEventDestroy(g_side.join_ev); cudaEventDestroy(g_side.mark_ev);
      cudaEventDestroy(g_side.mark2_ev);
    }
    g_side = SideRes();
    QEB_CUDA(cudaStreamCreateWithFlags(&g_side.side, cudaStreamNonBlocking));
    QEB_CUDA(cudaEventCreateWithFlags(&g_side.fork_ev, cudaEventDisableTiming));
    QEB_CUDA(cudaEventCreateWithFlags(&g_side.join_ev, cudaEventDisableTiming));
    QEB_CUDA(cudaEventCreateWithFlags(&g_side.mark_ev, cudaEventDisableTiming));
    QEB_CUDA(cudaEventCreateWithFlags(&g_side.mark2_ev, cudaEventDisableTiming));
    g_side.device = dev;
  }
  side = g_side.side; fork_ev = g_side.fork_ev; join_ev = g_side.join_ev; mark_ev = g_side.mark_ev; mark2_ev = g_side.mark2_ev;
  enabled = true;
  return QEB_OK;
}

int SideStream::fork() {
  if (!enabled) return QEB_OK;
  QEB_CUDA(cudaEventRecord(fork_ev, main));
  QEB_CUDA(cudaStreamWaitEvent(side, fork_ev, 0));
  dirty = true;
  return QEB_OK;
}

int SideStream::join() {
  if (!enabled || !dirty) return QEB_OK;
  QEB_CUDA(cudaEventRecord(join_ev, side));
  QEB_CUDA(cudaStreamWaitEvent(main, join_ev, 0));
  dirty = false;
  return QEB_OK;
}

int SideStream::mark() {
  if (!enabled) return QEB_OK;
  QEB_CUDA(cudaEventRecord(mark_ev, side));
  marked = true;
  return QEB_OK;
}

int SideStream::wait_mark() {
  if (!enabled || !marked) return QEB_OK;
  QEB_CUDA(cudaStreamWaitEvent(main, mark_ev, 0));
  marked = false;
  return QEB_OK;
}

int SideStream::mark2() {
  if (!enabled) return QEB_OK;
  QEB_CUDA(cudaEventRecord(mark2_ev, side));
  marked2 = true;
  return QEB_OK;
}

int SideStream::wait_mark2() {
  if (!enabled || !marked2) return QEB_OK;
  QEB_CUDA(cudaStreamWaitEvent(main, mark2_ev, 0));
  marked2 = false;
  return QEB_OK;
}

// ---- per-call CUDA graphs (nn.cuh) -------------------------------------------------------------------------------------------
namespace {
struct CallGraph {
  std::string key;
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
};
struct CallGraphCache {
  std::list<CallGraph> lru;                                            // most recently used first
  std::unordered_map<std::string, std::list<CallGraph>::iterator> index;
  cudaStream_t capture_stream = nullptr;
  int device = -1;
  ~CallGraphCache() {}   // graphs live until the process ends (destroying them at thread exit would race with the CUDA teardown)
};
thread_local CallGraphCache g_calls;
constexpr size_t kMaxCallGraphs = 24;
}  // namespace

int qeb_run_cached(const CallKey& key, cudaStream_t st, const std::function<int(cudaStream_t)>& body) {
  static const int mode = getenv("QEB_CALL_GRAPHS") ? atoi(getenv("QEB_CALL_GRAPHS")) : 1;
  if (!mode || qeb_prof_on() || qeb_dbg_skip_active() || qeb_debug_timeline() != nullptr) return body(st);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); return body(st); }
  if (cap != cudaStreamCaptureStatusNone) return body(st);             // already inside somebody's capture (GraphedStep)
  int dev = 0;
  QEB_CUDA(cudaGetDevice(&dev));
  CallGraphCache& c = g_calls;
  if (c.device != dev) {   // first use on this thread (or another device: start over)
    c.lru.clear(); c.index.clear();
    if (cudaStreamCreateWithFlags(&c.capture_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return body(st); }
    c.device = dev;
  }
  std::string k = key.bytes;
  k.append(reinterpret_cast<const char*>(&dev), sizeof(dev));
  auto hit = c.index.find(k);
  if (hit != c.index.end()) {
    c.lru.splice(c.lru.begin(), c.lru, hit->second);
    qeb_count_launch((int)hit->second->launches);
    QEB_CUDA(cudaGraphLaunch(hit->second->exec, st));
    return QEB_OK;
  }
  const long long n0 = qeb_launch_count_now();
  if (cudaStreamBeginCapture(c.capture_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return body(st); }
  const int rc = body(c.capture_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c.capture_stream, &graph);
  if (rc != QEB_OK) {   // argument error inside the body: nothing ran, report it
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || !graph) {   // not capturable on this driver / configuration: run it plainly from now on
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return body(st);
  }
  CallGraph g;
  g.key = k;
  g.launches = qeb_launch_count_now() - n0;
  const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { cudaGetLastError(); return body(st); }
  if (c.lru.size() >= kMaxCallGraphs) {
    cudaGraphExecDestroy(c.lru.back().exec);
    c.index.erase(c.lru.back().key);
    c.lru.pop_back();
  }
  c.lru.push_front(g);
  c.index[k] = c.lru.begin();
  QEB_CUDA(cudaGraphLaunch(c.lru.front().exec, st));
  return QEB_OK;
}

namespace {
std::vector<std::string> parse_skip() {
  std::vector<std::string> v;
  const char* e = getenv("QEB_DBG_SKIP");
  if (!e) return v;
  std::string s(e), cur;
  for (char c : s) {
    if (c == ',') { if (!cur.empty()) v.push_back(cur); cur.clear(); }
    else cur.push_back(c);
  }
  if (!cur.empty()) v.push_back(cur);
  return v;
}
const std::vector<std::string>& skip_list() {
  static const std::vector<std::string> v = parse_skip();
  return v;
}
thread_local int g_skip_flag = 0;
}  // namespace
bool qeb_dbg_skip_active() { return !skip_list().empty(); }
bool qeb_dbg_skip_tag(const char* tag) {
  for (const auto& p : skip_list())
    if (strncmp(tag, p.c_str(), p.size()) == 0) return true;
  return false;
}
int& qeb_skip_flag() { return g_skip_flag; }

bool qeb_pdl_enabled() {
  static const bool on = !(getenv("QEB_PDL") && atoi(getenv("QEB_PDL")) == 0);
  return on;
}

// ---- delayed scales of the fp16 gradient shadows (nn.cuh GradShadow / GradScales) ---------------------------------------------
namespace {
struct ScaleSet {
  unsigned* dev = nullptr;   // [n] running maxima (fp32 bit patterns) | [n] scales | [n] reciprocal scales
  int n = 0;
  long long calls = 0;       // backward calls issued with these slots
};
std::mutex g_scale_mu;
// key: (network kind, current device, first parameter pointer) - the same pointer value may exist on two devices of a process
struct ScaleKey {
  int kind, device;
  const void* ptr;
  bool operator<(const ScaleKey& o) const {
    if (kind != o.kind) return kind < o.kind;
    if (device != o.device) return device < o.device;
    return ptr < o.ptr;
  }
};
std::map<ScaleKey, ScaleSet> g_scale_sets;
ScaleKey scale_key(int kind, const void* ptr) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return ScaleKey{kind, dev, ptr};
}

// S = 2^(4 - ceil(log2(max))): the largest magnitude of the previous call lands in [8, 16] (2^12 below fp16's largest value). A maximum of 0, inf or NaN (nothing
// recorded, or a diverged step) keeps the previous scale.
__global__ void grad_scale_update_kernel(unsigned* amax, float* scale, float* inv, int n) {
  qeb_pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = __uint_as_float(amax[i]);
  float s = scale[i];
  if (a > 0.f && a < INFINITY) {
    int e;
    frexpf(a, &e);                       // a = m * 2^e, m in [0.5, 1)
    e = min(max(4 - e, -60), 100);
    s = ldexpf(1.f, e);
  }
  if (!(s > 0.f)) s = 1.f;
  scale[i] = s;
  inv[i] = 1.f / s;
  amax[i] = 0u;
}
}  // namespace

int grad_scales_state(int net_kind, const void* key) {
  std::lock_guard<std::mutex> lk(g_scale_mu);
  auto it = g_scale_sets.find(scale_key(net_kind, key));
  if (it == g_scale_sets.end()) return -1;
  return it->second.calls > 0 ? 1 : 0;
}

void grad_scales_commit(int net_kind, const void* key) {
  std::lock_guard<std::mutex> lk(g_scale_mu);
  auto it = g_scale_sets.find(scale_key(net_kind, key));
  if (it != g_scale_sets.end()) it->second.calls += 1;
}

// allocates the slots of (net_kind, key) on first use - unless `st` is being captured: allocation is not a stream operation
void grad_scales_prepare(int net_kind, const void* key, int n, cudaStream_t st) {
  static const bool allow = !(getenv("QEB_FP16_BWD") && atoi(getenv("QEB_FP16_BWD")) == 0);
  if (!allow) return;
  std::lock_guard<std::mutex> lk(g_scale_mu);
  auto it = g_scale_sets.find(scale_key(net_kind, key));
  if (it != g_scale_sets.end()) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); return; }
  if (cap != cudaStreamCaptureStatusNone) return;
  if (g_scale_sets.size() >= 256) return;   // a process that keeps creating networks: stay on the tf32 path
  unsigned* dev = nullptr;
  if (cudaMalloc(&dev, (size_t)3 * n * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return; }
  if (cudaMemset(dev, 0, (size_t)3 * n * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); cudaFree(dev); return; }
  ScaleSet set;
  set.dev = dev; set.n = n; set.calls = 0;
  g_scale_sets[scale_key(net_kind, key)] = set;
}

int grad_scales_begin(int net_kind, const void* key, int n, cudaStream_t st, GradScales* out) {
  *out = GradScales();
  ScaleSet set;
  {
    std::lock_guard<std::mutex> lk(g_scale_mu);
    auto it = g_scale_sets.find(scale_key(net_kind, key));
    if (it == g_scale_sets.end() || it->second.n != n) return QEB_OK;   // no slots: the tf32 path
    set = it->second;
  }
  out->amax = set.dev;
  out->scale = reinterpret_cast<float*>(set.dev + n);
  out->inv = reinterpret_cast<float*>(set.dev + 2 * n);
  out->n = n;
  out->valid = set.calls > 0;
  ProfScope prof("grad_scale", st);
  QEB_CUDA(qeb_launch(grad_scale_update_kernel, qeb_cdiv(n, 64), 64, 0, st, out->amax, out->scale, out->inv, n));
  qeb_count_launch();
  return QEB_OK;
}
