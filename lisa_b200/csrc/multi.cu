// lisa_b200/csrc/multi.cu — sample-space partition over the GPUs of one box in ONE process (include/lisa_rt.h: lisa_multi_*).
//
// every GPU holds the full scene and its own BVH, renders a disjoint block of subframes over the
// full image into its own float4 accumulators, and ONE ncclReduce (sum, fp32, root 0; ncclCommInitAll, ncclGroupStart/End
// around the per-device calls) combines them over NVLink.  Built on the single-GPU C ABI only (lisa_create with
// options.device, lisa_render_subframes, lisa_accum_device_ptr): a lisa_multi is G contexts + G communicators + G streams.
// NCCL is resolved at run time (dlopen of libnccl.so.2) so that liblisa_rt.so loads on a box without it; the reduce then
// goes through lisa_accum_add_peer (one kernel per peer reading over the NVLink peer mapping).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lisa_rt.h"
#include "internal.h"

namespace {

// the few NCCL declarations used (ABI of nccl.h 2.x; the library is loaded at run time)
typedef struct ncclComm* ncclComm_t;
typedef int              ncclResult_t;  // ncclSuccess = 0
enum { kNcclFloat = 7, kNcclSum = 0 };  // ncclFloat32, ncclSum

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  if (const char* e = getenv("LISA_NCCL")) if (!strcmp(e, "0")) return n;  // ablation: force the peer-kernel reduce
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) return n;
  auto sym = [&](const char* s) { return dlsym(n.lib, s); };
  n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
  n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
  n.Reduce = (decltype(n.Reduce))sym("ncclReduce");
  n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
  n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
  n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
  n.ok = n.CommInitAll && n.CommDestroy && n.Reduce && n.GroupStart && n.GroupEnd && n.GetErrorString;
  return n;
}

thread_local char g_merr[512] = "";

}  // namespace

struct lisa_multi {
  std::vector<lisa_ctx*>    ctx;
  std::vector<int>          dev;
  std::vector<ncclComm_t>   comm;
  std::vector<cudaStream_t> stream;
  bool                      use_nccl = false;
  uint32_t                  width = 0, height = 0;
  double                    render_ms = 0, reduce_ms = 0;
  std::string               err;
};

// errors of the multi layer are reported through lisa_last_error() like every other call: keep the text of the failing
// single-GPU call (it was set on a worker thread) and re-raise it on the caller's thread via a failing lisa_* call
static int mfail(lisa_multi* m, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_merr, sizeof(g_merr), fmt, ap);
  va_end(ap);
  if (m) m->err = g_merr;
  lisa_internal_set_last_error(g_merr);
  return code;
}

extern "C" void lisa_multi_destroy(lisa_multi* m) {
  if (!m) return;
  for (size_t g = 0; g < m->comm.size(); g++)
    if (m->comm[g]) { cudaSetDevice(m->dev[g]); nccl().CommDestroy(m->comm[g]); }
  for (size_t g = 0; g < m->stream.size(); g++)
    if (m->stream[g]) { cudaSetDevice(m->dev[g]); cudaStreamDestroy(m->stream[g]); }
  for (lisa_ctx* c : m->ctx) lisa_destroy(c);
  delete m;
}

extern "C" int lisa_multi_create(const lisa_scene_desc* sd, const lisa_options* opt, int num_gpus, lisa_multi** out) {
  if (!sd || !out) return mfail(nullptr, LISA_ERR_ARG, "lisa_multi_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return mfail(nullptr, LISA_ERR_CUDA, "no CUDA device: lisa_rt has no CPU fallback");
  if (num_gpus <= 0) num_gpus = ndev;
  if (num_gpus > ndev) return mfail(nullptr, LISA_ERR_ARG, "%d GPUs asked for, %d visible", num_gpus, ndev);
  lisa_multi* m = new lisa_multi();
  m->ctx.assign(num_gpus, nullptr);
  m->dev.resize(num_gpus);
  m->comm.assign(num_gpus, nullptr);
  m->stream.assign(num_gpus, nullptr);
  m->width = sd->width; m->height = sd->height;
  std::vector<std::string> err(num_gpus);
  std::vector<int>         rc(num_gpus, LISA_OK);
  std::vector<std::thread> th;
  for (int g = 0; g < num_gpus; g++) {
    m->dev[g] = g;
    th.emplace_back([&, g] {  // upload + BVH build of every replica at the same time
      lisa_options o{};
      o.struct_size = sizeof(o);
      if (opt) memcpy(&o, opt, opt->struct_size && opt->struct_size < sizeof(o) ? opt->struct_size : sizeof(o));
      o.struct_size = sizeof(o);
      o.device = g;
      rc[g] = lisa_create(sd, &o, &m->ctx[g]);
      if (rc[g] != LISA_OK) err[g] = lisa_last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < num_gpus; g++)
    if (rc[g] != LISA_OK) {
      const int code = rc[g];
      const std::string msg = "GPU " + std::to_string(g) + ": " + err[g];
      lisa_multi_destroy(m);
      return mfail(nullptr, code, "%s", msg.c_str());
    }
  for (int g = 0; g < num_gpus; g++) {
    if (cudaSetDevice(g) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[g], cudaStreamNonBlocking) != cudaSuccess) {
      lisa_multi_destroy(m);
      return mfail(nullptr, LISA_ERR_CUDA, "cannot create a stream on GPU %d", g);
    }
  }
  if (num_gpus > 1 && nccl().ok) {
    const ncclResult_t r = nccl().CommInitAll(m->comm.data(), num_gpus, m->dev.data());
    if (r == 0) {
      m->use_nccl = true;
      // the first collective of a communicator sets up its channels (tens of ms): pay that here, not in the first render
      std::vector<float*> tiny(num_gpus, nullptr);
      ncclResult_t w = 0;
      for (int g = 0; g < num_gpus; g++) { cudaSetDevice(g); if (cudaMalloc((void**)&tiny[g], 64) != cudaSuccess) w = 1; else cudaMemsetAsync(tiny[g], 0, 64, m->stream[g]); }
      if (w == 0) {
        w = nccl().GroupStart();
        for (int g = 0; g < num_gpus && w == 0; g++) { cudaSetDevice(g); w = nccl().Reduce(tiny[g], tiny[g], 16, kNcclFloat, kNcclSum, 0, m->comm[g], m->stream[g]); }
        const ncclResult_t w2 = nccl().GroupEnd();
        if (w == 0) w = w2;
      }
      for (int g = 0; g < num_gpus; g++) { cudaSetDevice(g); cudaStreamSynchronize(m->stream[g]); if (tiny[g]) cudaFree(tiny[g]); }
      if (w != 0) fprintf(stderr, "lisa_multi: NCCL warm-up reduce failed (%s)\n", nccl().GetErrorString(w));
    } else {  // keep going with the peer-kernel reduce, and say so
      fprintf(stderr, "lisa_multi: ncclCommInitAll failed (%s); reducing through peer-memory kernels\n", nccl().GetErrorString(r));
      m->comm.assign(num_gpus, nullptr);
    }
  }
  *out = m;
  return LISA_OK;
}

extern "C" int         lisa_multi_num_gpus(const lisa_multi* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" lisa_ctx*   lisa_multi_root(lisa_multi* m) { return m && !m->ctx.empty() ? m->ctx[0] : nullptr; }
extern "C" lisa_ctx*   lisa_multi_ctx(lisa_multi* m, int g) { return m && g >= 0 && g < (int)m->ctx.size() ? m->ctx[g] : nullptr; }
extern "C" const char* lisa_multi_backend(const lisa_multi* m) { return !m ? "" : m->ctx.size() < 2 ? "single" : m->use_nccl ? "nccl" : "peer"; }

extern "C" int lisa_multi_reset_accum(lisa_multi* m) {
  if (!m) return mfail(nullptr, LISA_ERR_ARG, "null lisa_multi");
  for (lisa_ctx* c : m->ctx) {
    const int rc = lisa_reset_accum(c);
    if (rc != LISA_OK) return rc;
  }
  return LISA_OK;
}

// sums every GPU's accumulators onto GPU 0 and clears the others
static int reduce_to_root(lisa_multi* m, const std::vector<uint32_t>& subframes, const std::vector<uint64_t>& samples) {
  const int G = (int)m->ctx.size();
  if (G < 2) return LISA_OK;
  const size_t nfloat = (size_t)m->width * m->height * 4;
  if (m->use_nccl) {
    Nccl& n = nccl();
    ncclResult_t r = n.GroupStart();
    for (int g = 0; g < G && r == 0; g++) {
      cudaSetDevice(m->dev[g]);
      void* buf = lisa_accum_device_ptr(m->ctx[g]);
      r = n.Reduce(buf, buf, nfloat, kNcclFloat, kNcclSum, 0, m->comm[g], m->stream[g]);  // in place on the root
    }
    const ncclResult_t r2 = n.GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) return mfail(m, LISA_ERR_CUDA, "ncclReduce: %s", n.GetErrorString(r));
    for (int g = 0; g < G; g++) {
      cudaSetDevice(m->dev[g]);
      const cudaError_t e = cudaStreamSynchronize(m->stream[g]);
      if (e != cudaSuccess) return mfail(m, LISA_ERR_CUDA, "reduce on GPU %d: %s", g, cudaGetErrorString(e));
    }
    for (int g = 1; g < G; g++) lisa_accum_note_merged(m->ctx[0], subframes[g], samples[g]);
  } else {
    for (int g = 1; g < G; g++) {
      const int rc = lisa_accum_add_peer(m->ctx[0], m->ctx[g]);  // also moves the counters
      if (rc != LISA_OK) return rc;
    }
  }
  for (int g = 1; g < G; g++) {
    const int rc = lisa_reset_accum(m->ctx[g]);
    if (rc != LISA_OK) return rc;
  }
  return LISA_OK;
}

// GPU g renders subframes [f[g], f[g] + n[g]) with spp[g] samples each
static int render_blocks(lisa_multi* m, const std::vector<uint32_t>& f, const std::vector<uint32_t>& n, const std::vector<uint32_t>& spp) {
  const int G = (int)m->ctx.size();
  std::vector<int>         rc(G, LISA_OK);
  std::vector<std::string> err(G);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++) {
    if (!n[g]) continue;
    th.emplace_back([&, g] {
      rc[g] = lisa_render_subframes(m->ctx[g], f[g], n[g], spp[g]);
      if (rc[g] != LISA_OK) err[g] = lisa_last_error();
    });
  }
  for (auto& t : th) t.join();
  const auto t1 = std::chrono::steady_clock::now();
  for (int g = 0; g < G; g++)
    if (rc[g] != LISA_OK) return mfail(m, rc[g], "GPU %d: %s", g, err[g].c_str());
  std::vector<uint32_t> subframes(G);
  std::vector<uint64_t> samples(G);
  for (int g = 0; g < G; g++) { subframes[g] = n[g]; samples[g] = (uint64_t)m->width * m->height * n[g] * spp[g]; }
  const int r = reduce_to_root(m, subframes, samples);
  const auto t2 = std::chrono::steady_clock::now();
  m->render_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  m->reduce_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
  return r;
}

extern "C" int lisa_multi_render_subframes(lisa_multi* m, uint32_t first, uint32_t count, uint32_t spp) {
  if (!m) return mfail(nullptr, LISA_ERR_ARG, "null lisa_multi");
  if (!count || !spp) return mfail(m, LISA_ERR_ARG, "count and spp must be positive");
  const uint32_t G = (uint32_t)m->ctx.size(), base = count / G, rem = count % G;
  std::vector<uint32_t> f(G), n(G), s(G, spp);
  for (uint32_t g = 0; g < G; g++) {  // contiguous blocks: ONE render call per GPU, all of its chains resident together
    n[g] = base + (g < rem ? 1u : 0u);
    f[g] = first + g * base + (g < rem ? g : rem);
  }
  return render_blocks(m, f, n, s);
}

extern "C" int lisa_multi_render_samples(lisa_multi* m, uint32_t first, uint32_t num_samples) {
  if (!m) return mfail(nullptr, LISA_ERR_ARG, "null lisa_multi");
  if (!num_samples) return mfail(m, LISA_ERR_ARG, "num_samples must be positive");
  const uint32_t G = (uint32_t)m->ctx.size(), base = num_samples / G, rem = num_samples % G;
  std::vector<uint32_t> f(G), n(G), s(G);
  for (uint32_t g = 0; g < G; g++) {  // GPU g: subframe first + g with floor/ceil(N / G) samples — N in total, exactly
    s[g] = base + (g < rem ? 1u : 0u);
    n[g] = s[g] ? 1u : 0u;
    f[g] = first + g;
    if (!s[g]) s[g] = 1;
  }
  return render_blocks(m, f, n, s);
}

extern "C" int lisa_multi_last_times(const lisa_multi* m, double* render_ms, double* reduce_ms) {
  if (!m) return LISA_ERR_ARG;
  if (render_ms) *render_ms = m->render_ms;
  if (reduce_ms) *reduce_ms = m->reduce_ms;
  return LISA_OK;
}
