/* gsi_capi.cu -- the C ABI of libgsl_scattered_cuda (include/gsl_scattered_cuda.h).
 *
 * Host side of the drop-in boundary: workspace life cycle, host build -> flatten -> upload,
 * replication of the slab to the selected GPUs (peer copy), chunked evaluation with the
 * deferred-list protocol of gsi_kernels.cu, and the pinned/pageable host <-> device
 * pipeline behind eval_e.  There is no CPU evaluation path in this library: every eval entry
 * point fails with GSL_SCATTERED_EFAILED when no CUDA device is usable.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gsl_scattered_cuda.h"
#include "gsi_builder.h"
#include "gsi_flatten.h"
#include "gsi_kernels.h"
#include "gsi_layout.h"

namespace
{

gsl_scattered_cuda_error_handler_t *g_handler = nullptr;

int
report(const char *reason, const char *file, int line, int code)
{
  if (g_handler) g_handler(reason, file, line, code);
  return code;
}

#define GSI_ERROR(reason, code) return report((reason), __FILE__, __LINE__, (code))
#define GSI_ERROR_NULL(reason, code)                                                           \
  do                                                                                           \
    {                                                                                          \
      report((reason), __FILE__, __LINE__, (code));                                            \
      return nullptr;                                                                          \
    }                                                                                          \
  while (0)

#define CU(call)                                                                               \
  do                                                                                           \
    {                                                                                          \
      cudaError_t e_ = (call);                                                                 \
      if (e_ != cudaSuccess)                                                                   \
        {                                                                                      \
          std::string m_ = std::string("CUDA: ") + cudaGetErrorString(e_) + " in " #call;      \
          return report(m_.c_str(), __FILE__, __LINE__,                                        \
                        e_ == cudaErrorMemoryAllocation ? GSL_SCATTERED_ENOMEM                 \
                                                        : GSL_SCATTERED_EFAILED);              \
        }                                                                                      \
    }                                                                                          \
  while (0)

/* No C++ exception may cross the C ABI: every extern "C" body that can allocate (std::vector,
 * std::string, std::thread) runs between these two macros and turns bad_alloc into
 * GSL_SCATTERED_ENOMEM and anything else into GSL_SCATTERED_EFAILED, reported through the hook. */
#define GSI_GUARD_BEGIN                                                                        \
  try                                                                                          \
    {
#define GSI_GUARD_END                                                                          \
    }                                                                                          \
  catch (const std::bad_alloc &)                                                               \
    {                                                                                          \
      GSI_ERROR("out of host memory", GSL_SCATTERED_ENOMEM);                                   \
    }                                                                                          \
  catch (...)                                                                                  \
    {                                                                                          \
      GSI_ERROR("internal error (C++ exception)", GSL_SCATTERED_EFAILED);                      \
    }

/* The library switches the current CUDA device to the replica it works on; the caller's current
 * device is restored when the entry point returns. */
struct device_guard
{
  int prev = -1;
  device_guard()
  {
    if (cudaGetDevice(&prev) != cudaSuccess)
      {
        cudaGetLastError();
        prev = -1;
      }
  }
  ~device_guard()
  {
    if (prev >= 0 && cudaSetDevice(prev) != cudaSuccess) cudaGetLastError();
  }
  device_guard(const device_guard &) = delete;
  device_guard &operator=(const device_guard &) = delete;
};

double
now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

struct stage_buf
{
  double *x = nullptr, *y = nullptr, *z = nullptr;
  int *leaf = nullptr;
  cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
};

/* pinned host staging for callers that hold pageable (malloc'd) arrays, as the reference's
 * gsl_matrix / gsl_vector do */
struct host_stage
{
  double *x = nullptr, *y = nullptr, *z = nullptr;
  int *leaf = nullptr;
  cudaEvent_t ev_out = nullptr;
};

/* Everything ONE evaluation call in flight needs besides the (read-only) slab: streams, deferred
 * list, counters, staging buffers, reorder scratch.  Contexts are pooled per replica, so eval*
 * is re-entrant across host threads: concurrent calls on one workspace get
 * different contexts and never share mutable device state. */
struct eval_ctx
{
  cudaStream_t s_k = nullptr, s_in = nullptr, s_out = nullptr;
  unsigned *list = nullptr;
  size_t list_cap = 0;
  size_t list_want = 0; /* entries the last launches would have needed (read_counters): the list grows to it */
  unsigned *h_deferred = nullptr; /* pinned: deferred count of the last launch, copied back behind it on its stream,
                                     so that asynchronous callers (who never read the counters) size the list too */
  unsigned *ovf_mask = nullptr; /* overflow bitmask, one bit per query of a launch */
  size_t ovf_words = 0;
  unsigned *counters = nullptr;
  unsigned long long *scratch2 = nullptr; /* 2 x u64 for checksum / compare */
  unsigned long long *trace = nullptr;    /* param "walk_trace": 4 x u64 per CTA of the last k_walkp launch */
  size_t trace_ctas = 0;
  stage_buf st[2];
  size_t st_cap = 0, st_fields = 0;
  host_stage hs[2];
  size_t hs_cap = 0, hs_fields = 0;
  gsi_reorder_scratch ro; /* reorder pipeline scratch (lazily sized) */
  size_t ro_fields = 0, ro_runs_entries = 0;
  /* profiling (param "profile"): timed event pairs around every kernel launch */
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_locate, prof_exact;
  double locate_seconds = 0, exact_seconds = 0;
  unsigned long long n_launches = 0, n_reordered = 0;
  /* recorded on the stream that used the context last: the next user's stream waits for it, so a
     context can be handed on while an asynchronous call on a caller's stream is still running */
  cudaEvent_t ev_done = nullptr;
  bool busy = false;
};

struct replica
{
  int dev = 0;
  int sm_count = 148;
  char *slab = nullptr;
  bool owns_slab = true;
  /* slab backed by the virtual-memory-management API (explicit 2 MB physical granularity) */
  bool slab_vmm = false;
  unsigned long long slab_handle = 0;
  size_t slab_mapped = 0;
  gsi_view view;
  std::mutex mu; /* guards the context pool, nothing else */
  std::vector<eval_ctx *> ctxs;
};

/* the builder's batched locate on a CUDA device (gsi_builder.h, k_build_locate) */
class cuda_build_locator : public gsi_build_locator
{
public:
  explicit cuda_build_locator(int dev) : dev_(dev) {}
  ~cuda_build_locator() override
  {
    device_guard restore;
    if (cudaSetDevice(dev_) != cudaSuccess) return;
    for (void *p : {(void *) d_rec_, (void *) d_lnk_, (void *) d_type_, (void *) d_px_, (void *) d_py_, (void *) d_node_,
                    (void *) d_c0_, (void *) d_c1_, (void *) d_last_})
      if (p) cudaFree(p);
    cudaGetLastError();
  }
  int begin(const gsi_xform &xf, const double *px, const double *py, size_t n, size_t max_nodes) override
  {
    device_guard restore;
    if (cudaSetDevice(dev_) != cudaSuccess) return fail();
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev_) != cudaSuccess) return fail();
    sm_count_ = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    const size_t per_round = n / 2 + 1;
    const size_t need = max_nodes * (sizeof(gsi_rec) + 13) + n * 16 + per_round * 24;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || need + ((size_t) 1 << 30) > free_b) return fail();
    xf_ = xf;
    max_nodes_ = max_nodes;
    cap_ = per_round;
    if (cudaMalloc(&d_rec_, max_nodes * sizeof(gsi_rec)) != cudaSuccess || cudaMalloc(&d_lnk_, max_nodes * 12) != cudaSuccess
        || cudaMalloc(&d_type_, max_nodes) != cudaSuccess || cudaMalloc(&d_px_, n * 8) != cudaSuccess
        || cudaMalloc(&d_py_, n * 8) != cudaSuccess || cudaMalloc(&d_node_, per_round * 4) != cudaSuccess
        || cudaMalloc(&d_c0_, per_round * 8) != cudaSuccess || cudaMalloc(&d_c1_, per_round * 8) != cudaSuccess
        || cudaMalloc(&d_last_, per_round * 4) != cudaSuccess)
      return fail();
    if (cudaMemcpy(d_px_, px, n * 8, cudaMemcpyHostToDevice) != cudaSuccess
        || cudaMemcpy(d_py_, py, n * 8, cudaMemcpyHostToDevice) != cudaSuccess)
      return fail();
    return 0;
  }
  int locate(const gsi_rec *rec, const int32_t *lnk, const uint8_t *type, size_t n_nodes, size_t n_prev, size_t first,
             size_t count, int32_t *node, double *c0, double *c1, int32_t *last) override
  {
    if (n_nodes > max_nodes_ || count > cap_ || n_prev > n_nodes) return 1;
    device_guard restore;
    if (cudaSetDevice(dev_) != cudaSuccess) return fail();
    /* LU records never change: only the new ones travel; links and types of old nodes do change */
    if (cudaMemcpyAsync(d_rec_ + n_prev, rec + n_prev, (n_nodes - n_prev) * sizeof(gsi_rec), cudaMemcpyHostToDevice, 0)
            != cudaSuccess
        || cudaMemcpyAsync(d_lnk_, lnk, n_nodes * 12, cudaMemcpyHostToDevice, 0) != cudaSuccess
        || cudaMemcpyAsync(d_type_, type, n_nodes, cudaMemcpyHostToDevice, 0) != cudaSuccess)
      return fail();
    if (gsi_launch_build_locate(d_rec_, d_lnk_, d_type_, xf_, d_px_, d_py_, first, count, d_node_, d_c0_, d_c1_, d_last_,
                                sm_count_, 0)
        != cudaSuccess)
      return fail();
    if (cudaMemcpyAsync(node, d_node_, count * 4, cudaMemcpyDeviceToHost, 0) != cudaSuccess
        || cudaMemcpyAsync(c0, d_c0_, count * 8, cudaMemcpyDeviceToHost, 0) != cudaSuccess
        || cudaMemcpyAsync(c1, d_c1_, count * 8, cudaMemcpyDeviceToHost, 0) != cudaSuccess
        || cudaMemcpyAsync(last, d_last_, count * 4, cudaMemcpyDeviceToHost, 0) != cudaSuccess
        || cudaStreamSynchronize(0) != cudaSuccess)
      return fail();
    return 0;
  }

private:
  int fail()
  {
    cudaGetLastError();
    return 1;
  }
  int dev_, sm_count_ = 148;
  gsi_xform xf_;
  size_t max_nodes_ = 0, cap_ = 0;
  gsi_rec *d_rec_ = nullptr;
  int32_t *d_lnk_ = nullptr, *d_node_ = nullptr, *d_last_ = nullptr;
  uint8_t *d_type_ = nullptr;
  double *d_px_ = nullptr, *d_py_ = nullptr, *d_c0_ = nullptr, *d_c1_ = nullptr;
};

} // namespace

static int
default_copy_threads()
{
  const unsigned hw = std::thread::hardware_concurrency();
  const unsigned t = hw ? (3 * hw) / 4 : 4;
  return (int) (t < 4 ? 4 : (t > 16 ? 16 : t));
}

struct gsl_scattered_cuda
{
  size_t n_cap = 0;
  size_t n_fields = 1;
  gsi_host_tree tree;
  std::vector<double> xa, ya;
  bool built = false, uploaded = false;
  gsi_flatten_params fprm;
  int mode = GSI_MODE_FAST, stats = 0, max_steps = 256, ctas_per_sm = 8, profile = 0;
  int kernel = GSI_KERNEL_AUTO, walk_ctas_per_sm = 5;
  int l2_fetch = 0; /* cudaLimitMaxL2FetchGranularity, device-wide: 0 = leave the host application's setting
                       alone (measured without effect on B200: an L2 miss fills the whole 128-byte line) */
  /* reorder pipeline (gsi_kernels.cu): 0 = off (default: measured no faster than the direct walk on B200,
     DESIGN.md section 4.3), 1 = always for query arrays, 2 = auto: large batches of few fields on a
     triangulation whose fast-path structure exceeds the L2 */
  int reorder = 0;
  size_t reorder_window = (size_t) 1 << 27, reorder_min = (size_t) 1 << 24;
  size_t reorder_min_leaves = 600000; /* ~58 MB of leaf records: about half the L2 */
  int reorder_chunk = GSI_BIN_CHUNK_MAX, tile_ranges = 0, tile_threads = GSI_TILE_THREADS;
  size_t reorder_max_fields = 2;
  int exact_entry = 1; /* exact replay starts at the cell's certified entry node (0: at the root) */
  int exact_coop = 1;  /* deferred queries are replayed by four lanes each (1, default; 2 = groups refilled as they
                          finish: measured slower; 0 = one thread each) */
  int walk_variant = 2; /* 1: k_walkp (two walks per lane in rotation behind L2 prefetches), 0: k_walk */
  int walk_dynamic = 1; /* k_walkp draws its chunks from a device counter (0: round-robin) */
  int walk_l1 = -1;       /* k_walkc: cp.async.ca (1) / .cg (0) for the leaf lines, -1 = by structure size */
  int walk_carveout = 72; /* k_walkc / k_walkf: preferred shared-memory carve-out in percent (-1: the driver's choice).
                             164 KB of shared memory leave 92 KB of L1: with 196 KB or more (60 KB of L1 or less)
                             the walk kernels lose a quarter of their rate, profiles/r2u_carveout_and_pitch.txt */
  int walk_pf = 1;      /* k_walkp: 0 no prefetch, 1..4 = flavours of gsi_prefetch_leaf_l2 */
  int walk_trace = 0;   /* diagnostic: per-CTA {smid, start, end, queries} of the last k_walkp launch */
  int build_device = 1; /* the builder locates its points in batches on the GPU when there is one (0: host only) */
  int vmm = 1;         /* slab through cuMemCreate/cuMemMap with the recommended granularity (0: cudaMalloc) */
  size_t walk_chunk = 0;
  size_t chunk = (size_t) 0xfffffff0u;
  double list_fraction = 1.0 / 8;
  size_t host_chunk = (size_t) 1 << 22;
  /* host threads moving pageable caller arrays to / from the pinned staging: three quarters of the host's
     hardware threads, between 4 and 16 (measured on a 16-core box: 0.69 / 1.03 / 1.29 / 1.32e9 queries/s
     end to end with 4 / 8 / 12 / 16 threads) */
  int copy_threads = default_copy_threads();
  size_t pageable_min = (size_t) 1 << 18; /* smaller pageable batches go straight through cudaMemcpyAsync */
  std::vector<int> devs;
  std::vector<replica *> reps;
  gsi_slab_header hdr;
  double flatten_seconds = 0, upload_seconds = 0;
};

namespace
{

/* ---- slab allocation --------------------------------------------------------------------------
 * The fast path gathers at random from ~224 MB of the slab at 10^6 nodes, right at the reach of
 * the GPU's TLB with 2 MB pages.  cudaMalloc does not promise a page size, and on B200 the same
 * kernel was measured at 38 ms or 47 ms (or worse) per 10^9 queries from process to process
 * (profiles/r1e_*).  The slab is therefore allocated through the driver's virtual-memory-
 * management API with the recommended granularity, so that every physical chunk is a whole large
 * page.  The entry points are resolved at run time (cudaGetDriverEntryPoint): the library does
 * not link to libcuda and still loads on a machine without a driver.  Any failure falls back to
 * cudaMalloc. */
struct vmm_api
{
  CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  bool ok = false;
};

const vmm_api &
vmm()
{
  static vmm_api a;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char *name, void **fn) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        {
          cudaGetLastError();
          *fn = nullptr;
        }
      return *fn != nullptr;
    };
    a.ok = get("cuMemGetAllocationGranularity", (void **) &a.getGranularity) && get("cuMemCreate", (void **) &a.create)
           && get("cuMemAddressReserve", (void **) &a.reserve) && get("cuMemMap", (void **) &a.map)
           && get("cuMemSetAccess", (void **) &a.setAccess) && get("cuMemUnmap", (void **) &a.unmap)
           && get("cuMemRelease", (void **) &a.release) && get("cuMemAddressFree", (void **) &a.addressFree);
  });
  return a;
}

void
slab_free(replica *r)
{
  if (!r->slab) return;
  if (r->slab_vmm)
    {
      const vmm_api &v = vmm();
      v.unmap((CUdeviceptr) r->slab, r->slab_mapped);
      v.release((CUmemGenericAllocationHandle) r->slab_handle);
      v.addressFree((CUdeviceptr) r->slab, r->slab_mapped);
    }
  else if (r->owns_slab)
    cudaFree(r->slab);
  r->slab = nullptr;
  r->slab_vmm = false;
  r->slab_handle = 0;
  r->slab_mapped = 0;
}

/* the current device must be r->dev (make_replica leaves it so) */
cudaError_t
slab_alloc(replica *r, size_t bytes, int use_vmm)
{
  const vmm_api &v = vmm();
  if (use_vmm && v.ok)
    {
      CUmemAllocationProp prop;
      std::memset(&prop, 0, sizeof prop);
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = r->dev;
      size_t gran = 0;
      CUmemGenericAllocationHandle h = 0;
      CUdeviceptr p = 0;
      if (v.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran)
        {
          const size_t size = (bytes + gran - 1) / gran * gran;
          if (v.create(&h, size, &prop, 0) == CUDA_SUCCESS)
            {
              if (v.reserve(&p, size, gran, 0, 0) == CUDA_SUCCESS)
                {
                  CUmemAccessDesc acc;
                  std::memset(&acc, 0, sizeof acc);
                  acc.location = prop.location;
                  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
                  if (v.map(p, size, 0, h, 0) == CUDA_SUCCESS)
                    {
                      if (v.setAccess(p, size, &acc, 1) == CUDA_SUCCESS)
                        {
                          r->slab = (char *) p;
                          r->slab_vmm = true;
                          r->slab_handle = (unsigned long long) h;
                          r->slab_mapped = size;
                          return cudaSuccess;
                        }
                      v.unmap(p, size);
                    }
                  v.addressFree(p, size);
                }
              v.release(h);
            }
        }
    }
  r->slab_vmm = false;
  return cudaMalloc(&r->slab, bytes);
}

void
free_reorder(eval_ctx *c)
{
  gsi_reorder_scratch &o = c->ro;
  if (o.bx) cudaFree(o.bx);
  if (o.by) cudaFree(o.by);
  if (o.bi) cudaFree(o.bi);
  if (o.runs) cudaFree(o.runs);
  if (o.work) cudaFree(o.work);
  if (o.zb) cudaFree(o.zb);
  if (o.lb) cudaFree(o.lb);
  o = gsi_reorder_scratch();
  c->ro_fields = c->ro_runs_entries = 0;
}

void
free_host_stage(eval_ctx *c)
{
  for (int b = 0; b < 2; b++)
    {
      host_stage &h = c->hs[b];
      if (h.x) cudaFreeHost(h.x);
      if (h.y) cudaFreeHost(h.y);
      if (h.z) cudaFreeHost(h.z);
      if (h.leaf) cudaFreeHost(h.leaf);
      h.x = h.y = h.z = nullptr;
      h.leaf = nullptr;
    }
  c->hs_cap = c->hs_fields = 0;
}

/* the current device must be the context's */
void
free_ctx(eval_ctx *c)
{
  if (!c) return;
  for (auto &p : c->prof_locate) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto &p : c->prof_exact) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  if (c->list) cudaFree(c->list);
  if (c->ovf_mask) cudaFree(c->ovf_mask);
  if (c->counters) cudaFree(c->counters);
  if (c->scratch2) cudaFree(c->scratch2);
  if (c->h_deferred) cudaFreeHost(c->h_deferred);
  if (c->trace) cudaFree(c->trace);
  free_reorder(c);
  free_host_stage(c);
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = c->st[b];
      if (s.x) cudaFree(s.x);
      if (s.y) cudaFree(s.y);
      if (s.z) cudaFree(s.z);
      if (s.leaf) cudaFree(s.leaf);
      if (s.ev_in) cudaEventDestroy(s.ev_in);
      if (s.ev_k) cudaEventDestroy(s.ev_k);
      if (s.ev_out) cudaEventDestroy(s.ev_out);
      if (c->hs[b].ev_out) cudaEventDestroy(c->hs[b].ev_out);
    }
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->s_k) cudaStreamDestroy(c->s_k);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
}

void
free_replica(replica *r)
{
  if (!r) return;
  cudaSetDevice(r->dev);
  cudaDeviceSynchronize();
  for (eval_ctx *c : r->ctxs) free_ctx(c);
  r->ctxs.clear();
  slab_free(r);
  cudaGetLastError();
  delete r;
}

void
drop_replicas(gsl_scattered_cuda *w)
{
  for (replica *r : w->reps) free_replica(r);
  w->reps.clear();
  w->uploaded = false;
}

int
make_ctx(eval_ctx **out)
{
  eval_ctx *c = new (std::nothrow) eval_ctx();
  if (!c) GSI_ERROR("out of memory", GSL_SCATTERED_ENOMEM);
  *out = c;
  CU(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  CU(cudaMalloc(&c->counters, GSI_CNT_WORDS * sizeof(unsigned)));
  CU(cudaMemset(c->counters, 0, GSI_CNT_WORDS * sizeof(unsigned)));
  CU(cudaMalloc(&c->scratch2, 2 * sizeof(unsigned long long)));
  CU(cudaMallocHost(&c->h_deferred, sizeof(unsigned)));
  *c->h_deferred = 0;
  CU(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  for (int b = 0; b < 2; b++)
    {
      CU(cudaEventCreateWithFlags(&c->st[b].ev_in, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->st[b].ev_k, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->st[b].ev_out, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->hs[b].ev_out, cudaEventDisableTiming));
    }
  return 0;
}

/* Take a context of replica r for one evaluation call (current device = r->dev).  `user` is the
 * caller's stream or NULL; *st receives the stream the call launches on, already ordered after
 * whatever used the context before. */
int
acquire_ctx(replica *r, cudaStream_t user, eval_ctx **out, cudaStream_t *st)
{
  eval_ctx *c = nullptr;
  {
    std::lock_guard<std::mutex> lock(r->mu);
    for (eval_ctx *k : r->ctxs)
      if (!k->busy)
        {
          c = k;
          c->busy = true;
          break;
        }
  }
  if (!c)
    {
      const int s = make_ctx(&c);
      if (s)
        {
          free_ctx(c);
          return s;
        }
      c->busy = true;
      std::lock_guard<std::mutex> lock(r->mu);
      r->ctxs.push_back(c);
    }
  *out = c;
  *st = user ? user : c->s_k;
  CU(cudaStreamWaitEvent(*st, c->ev_done, 0));
  return 0;
}

void
release_ctx(replica *r, eval_ctx *c, cudaStream_t st)
{
  if (!c) return;
  if (cudaEventRecord(c->ev_done, st) != cudaSuccess) cudaGetLastError();
  std::lock_guard<std::mutex> lock(r->mu);
  c->busy = false;
}

/* RAII: a context for the duration of one entry point */
struct ctx_lease
{
  replica *r = nullptr;
  eval_ctx *c = nullptr;
  cudaStream_t st = nullptr;
  int status = 0;
  ctx_lease(replica *rep, cudaStream_t user) : r(rep) { status = acquire_ctx(r, user, &c, &st); }
  ~ctx_lease()
  {
    if (status == 0) release_ctx(r, c, st);
  }
  ctx_lease(const ctx_lease &) = delete;
  ctx_lease &operator=(const ctx_lease &) = delete;
};

int
make_replica(int dev, replica **out)
{
  replica *r = new (std::nothrow) replica();
  if (!r) GSI_ERROR("out of memory", GSL_SCATTERED_ENOMEM);
  r->dev = dev;
  *out = r;
  CU(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  r->sm_count = prop.multiProcessorCount;
  return 0;
}

/* Optional device-wide hint (cudaLimitMaxL2FetchGranularity), param "l2_fetch"; off by default:
 * the library does not touch device-wide state of its host application unless asked. */
void
apply_l2_fetch(const gsl_scattered_cuda *w)
{
  if (w->l2_fetch > 0)
    {
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) w->l2_fetch) != cudaSuccess)
        cudaGetLastError();
    }
}

replica *
find_replica(const gsl_scattered_cuda *w, int dev)
{
  for (replica *r : w->reps)
    if (r->dev == dev) return r;
  return nullptr;
}

int
ensure_list(eval_ctx *c, size_t cap, size_t per_launch)
{
  if (c->list_cap < cap)
    {
      if (c->list) cudaFree(c->list);
      c->list = nullptr;
      c->list_cap = 0;
      /* twice: the second half receives what the vertex-memo pass leaves of the list (k_memo_list) */
      CU(cudaMalloc(&c->list, 2 * cap * sizeof(unsigned)));
      c->list_cap = cap;
    }
  const size_t words = (per_launch + 31) / 32;
  if (c->ovf_words < words)
    {
      if (c->ovf_mask) cudaFree(c->ovf_mask);
      c->ovf_mask = nullptr;
      c->ovf_words = 0;
      CU(cudaMalloc(&c->ovf_mask, words * sizeof(unsigned)));
      CU(cudaMemset(c->ovf_mask, 0, words * sizeof(unsigned))); /* kernels leave it zero */
      c->ovf_words = words;
    }
  return 0;
}

gsi_launch_cfg
cfg_of(const gsl_scattered_cuda *w, const replica *r, const eval_ctx *c)
{
  gsi_launch_cfg g;
  g.mode = w->mode;
  g.stats = w->stats;
  g.max_steps = w->max_steps;
  g.sm_count = r->sm_count;
  g.locate_ctas_per_sm = w->ctas_per_sm;
  g.kernel = w->kernel;
  g.walk_ctas_per_sm = w->walk_ctas_per_sm;
  g.exact_coop = w->exact_coop;
  g.walk_chunk = w->walk_chunk;
  g.walk_variant = w->walk_variant;
  g.walk_dynamic = w->walk_dynamic;
  g.walk_carveout = w->walk_carveout;
  g.walk_l1 = w->walk_l1;
  g.walk_pf = w->walk_pf;
  g.walk_trace = w->walk_trace ? c->trace : nullptr;
  g.walk_trace_ctas = w->walk_trace ? c->trace_ctas : 0;
  g.list = c->list;
  g.list2 = c->list ? c->list + c->list_cap : nullptr;
  g.list_capacity = c->list_cap;
  g.ovf_mask = c->ovf_mask;
  g.counters = c->counters;
  return g;
}

/* Does a batch of nq incoherent queries go through the reorder pipeline?  Auto (2): the batch is
 * large enough to amortise three streaming passes and the leaf records do not fit the L2, so the
 * direct walk would pay an HBM line fill per visited leaf. */
bool
want_reorder(const gsl_scattered_cuda *w, size_t nq)
{
  if (w->mode != GSI_MODE_FAST || w->reorder == 0 || nq == 0) return false;
  if (w->hdr.tile_bits < 1 || w->hdr.tile_bits > 12) return false;
  if (w->reorder == 1) return true;
  return nq >= w->reorder_min && w->n_fields <= w->reorder_max_fields && w->hdr.n_leaves >= w->reorder_min_leaves;
}

int
ensure_reorder(const gsl_scattered_cuda *w, eval_ctx *c, size_t window, bool want_z, bool want_leaf)
{
  gsi_reorder_scratch &o = c->ro;
  const int bits = (int) w->hdr.tile_bits;
  const int S = std::max(32, std::min(w->reorder_chunk, GSI_BIN_CHUNK_MAX)) & ~31;
  const size_t nt = (size_t) 1 << bits;
  const size_t n_chunks = (window + (size_t) S - 1) / (size_t) S;
  const size_t runs_entries = n_chunks * nt;
  const size_t F = w->n_fields;
  if (o.cap < window || o.chunk != S)
    {
      free_reorder(c);
      CU(cudaMalloc(&o.bx, window * sizeof(double)));
      CU(cudaMalloc(&o.by, window * sizeof(double)));
      CU(cudaMalloc(&o.bi, window * sizeof(unsigned short)));
      CU(cudaMalloc(&o.work, 64));
      o.cap = window;
      o.chunk = S;
    }
  if (c->ro_runs_entries < runs_entries)
    {
      if (o.runs) cudaFree(o.runs);
      o.runs = nullptr;
      c->ro_runs_entries = 0;
      CU(cudaMalloc(&o.runs, runs_entries * sizeof(unsigned)));
      c->ro_runs_entries = runs_entries;
    }
  if (want_z && (!o.zb || c->ro_fields < F))
    {
      if (o.zb) cudaFree(o.zb);
      o.zb = nullptr;
      CU(cudaMalloc(&o.zb, o.cap * F * sizeof(double)));
      c->ro_fields = F;
    }
  if (want_leaf && !o.lb) CU(cudaMalloc(&o.lb, o.cap * sizeof(int)));
  o.tile_max_leaves = (int) w->hdr.tile_max_leaves;
  o.ranges = w->tile_ranges;
  o.tile_threads = w->tile_threads;
  return 0;
}

/* a block of whole raster columns: pixel (i, j), i in [i_first, i_first + n_cols), j in [0, ny) */
struct raster_spec
{
  double x0, dx, y0, dy;
  size_t ny;
  size_t i_first, n_cols;
};

int
prepare_launch(const gsl_scattered_cuda *w, const replica *r, eval_ctx *c, cudaEvent_t ev[4], gsi_launch_cfg *cfg)
{
  *cfg = cfg_of(w, r, c);
  if (w->profile)
    {
      for (int i = 0; i < 4; i++) CU(cudaEventCreate(&ev[i]));
      if (w->mode == GSI_MODE_FAST)
        {
          cfg->ev_locate0 = ev[0]; cfg->ev_locate1 = ev[1];
          c->prof_locate.emplace_back(ev[0], ev[1]);
        }
      else
        {
          cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        }
      cfg->ev_exact0 = ev[2]; cfg->ev_exact1 = ev[3];
      c->prof_exact.emplace_back(ev[2], ev[3]);
    }
  c->n_launches += (w->mode == GSI_MODE_FAST) ? 4 : 1;
  return 0;
}

/* chunked launch of the locate+blend kernels on device data: either query arrays (ras == NULL)
 * or a block of raster columns.  d_z / d_leaf point at the first query / pixel of the block. */
int
eval_on_device(const gsl_scattered_cuda *w, const replica *r, eval_ctx *c, const double *d_x, const double *d_y,
               const raster_spec *ras, size_t nq, double *d_z, size_t ld, int *d_leaf, cudaStream_t st)
{
  if (nq == 0) return 0;
  size_t chunk = std::max<size_t>(1, std::min(w->chunk, (size_t) 0xfffffff0u));
  const bool reorder = !ras && want_reorder(w, nq);
  if (reorder)
    {
      chunk = std::min(chunk, std::max<size_t>(w->reorder_window, 32));
      const int s = ensure_reorder(w, c, std::min(chunk, nq), d_z != nullptr, d_leaf != nullptr);
      if (s) return s;
    }
  size_t cols_per_chunk = 0;
  if (ras)
    {
      if (ras->ny == 0 || ras->ny > (size_t) 0xfffffff0u) GSI_ERROR("raster column too long", GSL_SCATTERED_EINVAL);
      cols_per_chunk = std::max<size_t>(1, chunk / ras->ny);
      chunk = cols_per_chunk * ras->ny;
    }
  if (w->mode == GSI_MODE_FAST)
    {
      /* deferred list: a fraction of the launch size (overflow goes to a bitmask), or what the last
         launch on this context turned out to need (possibly still running: then the value is the one
         before it, which is just as good a hint) */
      const size_t per_launch = std::min(chunk, nq);
      {
        const size_t seen = *(volatile unsigned *) c->h_deferred;
        if (seen > c->list_cap) c->list_want = std::max(c->list_want, seen + seen / 4);
      }
      const size_t want = std::max<size_t>(std::max<size_t>(std::min<size_t>(per_launch, 65536),
                                                            (size_t) (w->list_fraction * (double) per_launch)),
                                           std::min(c->list_want, per_launch));
      const int s = ensure_list(c, want, per_launch);
      if (s) return s;
      const size_t ctas = (size_t) r->sm_count * 16;
      if (w->walk_trace && c->trace_ctas < ctas)
        {
          if (c->trace) cudaFree(c->trace);
          c->trace = nullptr;
          c->trace_ctas = 0;
          CU(cudaMalloc(&c->trace, ctas * 4 * sizeof(unsigned long long)));
          CU(cudaMemset(c->trace, 0, ctas * 4 * sizeof(unsigned long long)));
          c->trace_ctas = ctas;
        }
    }
  for (size_t off = 0; off < nq; off += chunk)
    {
      const size_t n = std::min(chunk, nq - off);
      gsi_launch_cfg cfg;
      cudaEvent_t ev[4];
      const int s = prepare_launch(w, r, c, ev, &cfg);
      if (s) return s;
      if (reorder)
        {
          cfg.reorder = &c->ro;
          c->n_reordered += n;
          c->n_launches += 2; /* k_bin + k_tile_walk + k_unpermute instead of one locate kernel */
        }
      cudaError_t e;
      if (ras)
        e = gsi_launch_eval_raster(r->view, ras->x0, ras->dx, ras->y0, ras->dy, ras->ny,
                                   ras->i_first + off / ras->ny, n / ras->ny,
                                   d_z ? d_z + off : nullptr, d_leaf ? d_leaf + off : nullptr, ld,
                                   cfg, st);
      else
        e = gsi_launch_eval_arrays(r->view, d_x + off, d_y + off, n, d_z ? d_z + off : nullptr,
                                   d_leaf ? d_leaf + off : nullptr, ld, cfg, st);
      CU(e);
    }
  if (w->mode == GSI_MODE_FAST)
    CU(cudaMemcpyAsync(c->h_deferred, c->counters + GSI_CNT_DEFERRED, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  return 0;
}

int
read_counters(eval_ctx *c, unsigned out[GSI_CNT_WORDS])
{
  CU(cudaMemcpy(out, c->counters, GSI_CNT_WORDS * sizeof(unsigned), cudaMemcpyDeviceToHost));
  /* the last launch deferred more queries than the list holds (the rest went through the overflow
     bitmask, which is slower): size the list for such launches from now on */
  const size_t deferred = out[GSI_CNT_DEFERRED];
  if (deferred > c->list_cap) c->list_want = std::max(c->list_want, deferred + deferred / 4);
  return 0;
}

int
ensure_stage(eval_ctx *c, size_t cap, size_t fields)
{
  if (c->st_cap >= cap && c->st_fields >= fields) return 0;
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = c->st[b];
      if (s.x) cudaFree(s.x);
      if (s.y) cudaFree(s.y);
      if (s.z) cudaFree(s.z);
      if (s.leaf) cudaFree(s.leaf);
      s.x = s.y = s.z = nullptr;
      s.leaf = nullptr;
    }
  c->st_cap = c->st_fields = 0;
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = c->st[b];
      CU(cudaMalloc(&s.x, cap * sizeof(double)));
      CU(cudaMalloc(&s.y, cap * sizeof(double)));
      CU(cudaMalloc(&s.z, cap * fields * sizeof(double)));
      CU(cudaMalloc(&s.leaf, cap * sizeof(int)));
    }
  c->st_cap = cap;
  c->st_fields = fields;
  return 0;
}

int
ensure_host_stage(eval_ctx *c, size_t cap, size_t fields)
{
  if (c->hs_cap >= cap && c->hs_fields >= fields) return 0;
  free_host_stage(c);
  for (int b = 0; b < 2; b++)
    {
      host_stage &h = c->hs[b];
      CU(cudaHostAlloc((void **) &h.x, cap * sizeof(double), cudaHostAllocDefault));
      CU(cudaHostAlloc((void **) &h.y, cap * sizeof(double), cudaHostAllocDefault));
      CU(cudaHostAlloc((void **) &h.z, cap * fields * sizeof(double), cudaHostAllocDefault));
      CU(cudaHostAlloc((void **) &h.leaf, cap * sizeof(int), cudaHostAllocDefault));
    }
  c->hs_cap = cap;
  c->hs_fields = fields;
  return 0;
}

/* is p ordinary pageable host memory (not pinned / registered / managed)? */
bool
is_pageable(const void *p)
{
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
      cudaGetLastError();
      return true;
    }
  return a.type == cudaMemoryTypeUnregistered;
}

/* several host threads share one large copy (a single core moves ~10 GB/s, the PCIe link 50) */
void
parallel_copy(int threads, const std::vector<std::pair<void *, const void *>> &jobs, const std::vector<size_t> &bytes)
{
  size_t total = 0;
  for (size_t b : bytes) total += b;
  if (threads <= 1 || total < ((size_t) 1 << 20))
    {
      for (size_t j = 0; j < jobs.size(); j++) std::memcpy(jobs[j].first, jobs[j].second, bytes[j]);
      return;
    }
  auto part = [&](int t) {
    for (size_t j = 0; j < jobs.size(); j++)
      {
        const size_t lo = bytes[j] * (size_t) t / (size_t) threads, hi = bytes[j] * (size_t) (t + 1) / (size_t) threads;
        std::memcpy((char *) jobs[j].first + lo, (const char *) jobs[j].second + lo, hi - lo);
      }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; t++) th.emplace_back(part, t);
  part(0);
  for (auto &t : th) t.join();
}

/* host-buffer pipeline on one replica: H2D (s_in) | kernels (s_k) | D2H (s_out), two chunk
 * buffers in flight.  Pinned caller buffers (gsl_scattered_cuda_host_alloc, cudaHostRegister) are
 * used in place.  Pageable ones (plain malloc, as the reference's gsl_matrix / gsl_vector hold)
 * would make every cudaMemcpyAsync synchronous and staged by the driver through one small bounce
 * buffer; they are moved through the context's own pinned staging by a few host threads instead,
 * copy-in of chunk k+1 and copy-out of chunk k-1 overlapping the GPU work on chunk k. */
int
eval_host_slice(const gsl_scattered_cuda *w, replica *r, const double *xq, const double *yq,
                const raster_spec *ras, size_t nq, double *zq, size_t ld, int *leaf,
                unsigned *outside_before_after)
{
  CU(cudaSetDevice(r->dev));
  ctx_lease lease(r, nullptr);
  if (lease.status) return lease.status;
  eval_ctx *c = lease.c;
  const size_t F = w->n_fields;
  size_t C = std::max<size_t>(1, std::min(w->host_chunk, nq));
  if (ras) C = std::max<size_t>(1, w->host_chunk / ras->ny) * ras->ny; /* whole columns */
  if (C > nq) C = nq;
  int s = ensure_stage(c, C, F);
  if (s) return s;
  const bool staged = nq >= w->pageable_min
                      && ((!ras && (is_pageable(xq) || is_pageable(yq))) || is_pageable(zq) || is_pageable(leaf));
  if (staged && (s = ensure_host_stage(c, C, F))) return s;
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  CU(cudaStreamSynchronize(c->s_k));
  if ((s = read_counters(c, c0))) return s;
  auto copy_out = [&](size_t k) { /* staged mode: pinned staging -> caller arrays of chunk k */
    const size_t off = k * C, n = std::min(C, nq - off);
    host_stage &h = c->hs[k & 1];
    std::vector<std::pair<void *, const void *>> jobs;
    std::vector<size_t> bytes;
    if (leaf) { jobs.emplace_back(leaf + off, h.leaf); bytes.push_back(n * sizeof(int)); }
    if (zq)
      for (size_t f = 0; f < F; f++) { jobs.emplace_back(zq + f * ld + off, h.z + f * C); bytes.push_back(n * sizeof(double)); }
    parallel_copy(w->copy_threads, jobs, bytes);
  };
  size_t k = 0;
  for (size_t off = 0; off < nq; off += C, k++)
    {
      const size_t n = std::min(C, nq - off);
      stage_buf &b = c->st[k & 1];
      host_stage &h = c->hs[k & 1];
      if (!ras)
        {
          const double *sx = xq + off, *sy = yq + off;
          if (staged)
            {
              /* h was drained by copy_out(k - 2) in the previous iteration */
              parallel_copy(w->copy_threads, {{h.x, sx}, {h.y, sy}}, {n * sizeof(double), n * sizeof(double)});
              sx = h.x;
              sy = h.y;
            }
          if (k >= 2) CU(cudaStreamWaitEvent(c->s_in, b.ev_k, 0));
          CU(cudaMemcpyAsync(b.x, sx, n * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
          CU(cudaMemcpyAsync(b.y, sy, n * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
          CU(cudaEventRecord(b.ev_in, c->s_in));
          CU(cudaStreamWaitEvent(c->s_k, b.ev_in, 0));
        }
      if (k >= 2) CU(cudaStreamWaitEvent(c->s_k, b.ev_out, 0));
      raster_spec rs;
      if (ras)
        {
          rs = *ras;
          rs.i_first += off / ras->ny;
          rs.n_cols = n / ras->ny;
        }
      s = eval_on_device(w, r, c, b.x, b.y, ras ? &rs : nullptr, n, zq ? b.z : nullptr, C,
                         leaf ? b.leaf : nullptr, c->s_k);
      if (s) return s;
      CU(cudaEventRecord(b.ev_k, c->s_k));
      CU(cudaStreamWaitEvent(c->s_out, b.ev_k, 0));
      if (staged)
        {
          if (leaf) CU(cudaMemcpyAsync(h.leaf, b.leaf, n * sizeof(int), cudaMemcpyDeviceToHost, c->s_out));
          if (zq) CU(cudaMemcpyAsync(h.z, b.z, ((F - 1) * C + n) * sizeof(double), cudaMemcpyDeviceToHost, c->s_out));
          CU(cudaEventRecord(b.ev_out, c->s_out));
          CU(cudaEventRecord(h.ev_out, c->s_out));
          if (k >= 1)
            {
              CU(cudaEventSynchronize(c->hs[(k - 1) & 1].ev_out));
              copy_out(k - 1);
            }
        }
      else
        {
          if (leaf)
            CU(cudaMemcpyAsync(leaf + off, b.leaf, n * sizeof(int), cudaMemcpyDeviceToHost, c->s_out));
          if (zq)
            for (size_t f = 0; f < F; f++)
              CU(cudaMemcpyAsync(zq + f * ld + off, b.z + f * C, n * sizeof(double),
                                 cudaMemcpyDeviceToHost, c->s_out));
          CU(cudaEventRecord(b.ev_out, c->s_out));
        }
    }
  CU(cudaStreamSynchronize(c->s_out));
  CU(cudaStreamSynchronize(c->s_k));
  if (staged && k >= 1) copy_out(k - 1);
  if ((s = read_counters(c, c1))) return s;
  *outside_before_after = c1[GSI_CNT_OUTSIDE] - c0[GSI_CNT_OUTSIDE];
  return 0;
}

int
eval_host(const gsl_scattered_cuda *w, const double *xq, const double *yq, const raster_spec *ras,
          size_t nq, double *zq, int *leaf)
{
  if (!w || !w->uploaded || w->reps.empty())
    GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  if (nq == 0) return GSL_SCATTERED_SUCCESS;
  device_guard restore;
  const size_t R = w->reps.size();
  std::vector<int> status(R, 0);
  std::vector<unsigned> outside(R, 0);
  auto run = [&](size_t i) {
    try
      {
        size_t q0 = nq * i / R, q1 = nq * (i + 1) / R;
        raster_spec rs;
        if (ras)
          {
            /* split a raster by whole columns */
            const size_t c0 = ras->n_cols * i / R, c1 = ras->n_cols * (i + 1) / R;
            rs = *ras;
            rs.i_first += c0;
            rs.n_cols = c1 - c0;
            q0 = c0 * ras->ny;
            q1 = c1 * ras->ny;
          }
        status[i] = eval_host_slice(w, w->reps[i], ras ? nullptr : xq + q0, ras ? nullptr : yq + q0,
                                    ras ? &rs : nullptr, q1 - q0, zq ? zq + q0 : nullptr, nq,
                                    leaf ? leaf + q0 : nullptr, &outside[i]);
      }
    catch (const std::bad_alloc &)
      {
        status[i] = report("out of host memory", __FILE__, __LINE__, GSL_SCATTERED_ENOMEM);
      }
    catch (...)
      {
        status[i] = report("internal error (C++ exception)", __FILE__, __LINE__, GSL_SCATTERED_EFAILED);
      }
  };
  if (R == 1)
    run(0);
  else
    {
      std::vector<std::thread> th;
      for (size_t i = 1; i < R; i++) th.emplace_back(run, i);
      run(0);
      for (auto &t : th) t.join();
    }
  unsigned long long out_total = 0;
  for (size_t i = 0; i < R; i++)
    {
      if (status[i]) return status[i];
      out_total += outside[i];
    }
  if (out_total) GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
  return GSL_SCATTERED_SUCCESS;
}

/* upload a host slab image to every selected device and fill the seed grid on the device */
int
install_image(gsl_scattered_cuda *w, const std::vector<char> &image)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    GSI_ERROR("no CUDA device available: libgsl_scattered_cuda has no CPU evaluation path",
              GSL_SCATTERED_EFAILED);
  device_guard restore;
  drop_replicas(w);
  std::memcpy(&w->hdr, image.data(), sizeof w->hdr);
  std::vector<int> devs = w->devs;
  if (devs.empty()) devs.push_back(restore.prev >= 0 ? restore.prev : 0);
  for (int d : devs)
    if (d < 0 || d >= ndev) GSI_ERROR("device id out of range", GSL_SCATTERED_EINVAL);
  for (int d : devs)
    {
      replica *r = nullptr;
      int s = make_replica(d, &r);
      if (r) w->reps.push_back(r);
      if (s) return s;
      apply_l2_fetch(w);
      CU(slab_alloc(r, w->hdr.total_bytes, w->vmm && devs.size() == 1)); /* peer copies: plain cudaMalloc */
      r->view = gsi_view_of(r->slab, &w->hdr);
      if (!w->exact_entry) r->view.entry = nullptr;
    }
  replica *r0 = w->reps[0];
  CU(cudaSetDevice(r0->dev));
  cudaStream_t st = nullptr;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaError_t e = cudaMemcpyAsync(r0->slab, image.data(), image.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = gsi_launch_leaf_rcp(r0->view, r0->slab + w->hdr.off_leaf, r0->sm_count, st);
  if (e == cudaSuccess) e = gsi_launch_entry_cells(r0->view, (int32_t *) (r0->slab + w->hdr.off_entry), r0->sm_count, st);
  if (e == cudaSuccess && w->hdr.memo_bits)
    e = gsi_launch_memo_fill(r0->view, (gsi_memo *) (r0->slab + w->hdr.off_memo), r0->sm_count, st);
  if (e == cudaSuccess) e = gsi_launch_seed_cells(r0->view, (int32_t *) (r0->slab + w->hdr.off_seeds), r0->sm_count, st);
  if (e == cudaSuccess)
    e = gsi_launch_pack_seeds(r0->view, (const int32_t *) (r0->slab + w->hdr.off_seeds), (uint16_t *) (r0->slab + w->hdr.off_seed16),
                              (int32_t *) (r0->slab + w->hdr.off_seedbase), r0->sm_count, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  for (size_t i = 1; e == cudaSuccess && i < w->reps.size(); i++)
    {
      replica *r = w->reps[i];
      e = cudaMemcpyPeerAsync(r->slab, r->dev, r0->slab, r0->dev, w->hdr.total_bytes, st);
    }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  CU(e);
  w->uploaded = true;
  return 0;
}

/* FNV-1a over the image with the checksum slot (header.reserved[0]) taken as zero, 8 bytes at
 * a time (images are multiples of 256 bytes) */
uint64_t
slab_checksum(const char *image, size_t bytes)
{
  const size_t slot = offsetof(gsi_slab_header, reserved);
  uint64_t hsh = 0xcbf29ce484222325ull;
  for (size_t i = 0; i + 8 <= bytes; i += 8)
    {
      uint64_t v = 0;
      if (i != slot) std::memcpy(&v, image + i, 8);
      hsh = (hsh ^ v) * 0x100000001b3ull;
    }
  return hsh;
}

/* a + b*c without wrapping: false on overflow */
bool
extent_ok(uint64_t off, uint64_t count, uint64_t elem, uint64_t total)
{
  if (off > total) return false;
  if (elem && count > (total - off) / elem) return false;
  return true;
}

/* take over a complete slab image (device or host memory; the seed grid is part of it) */
int
adopt_slab(gsl_scattered_cuda *w, const gsi_slab_header &h, const void *src, size_t bytes)
{
  if (h.magic != GSI_SLAB_MAGIC || h.version != GSI_SLAB_VERSION || h.total_bytes != bytes)
    GSI_ERROR("slab header mismatch", GSL_SCATTERED_EINVAL);
  if (h.n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  const uint64_t cells = (uint64_t) h.grid_g * h.grid_g;
  if (h.n_dag == 0 || h.grid_g == 0 || h.grid_g > 32768 || h.n_dag > 0x7fffffffull || h.n_leaves > 0x7fffffffull
      || h.n_leaves > h.n_dag || (h.leaf_stride != 96 && h.leaf_stride != 128) || h.tile_bits < 1 || h.tile_bits > 12
      || !extent_ok(h.off_dag, h.n_dag, sizeof(gsi_rec), bytes)
      || !extent_ok(h.off_leaf, h.n_leaves, h.leaf_stride, bytes)
      || !extent_ok(h.off_z, std::max<uint64_t>(1, h.n_points), (uint64_t) h.n_fields * sizeof(double), bytes)
      || !extent_ok(h.off_seeds, cells, sizeof(int32_t), bytes) || !extent_ok(h.off_entry, cells, sizeof(int32_t), bytes)
      || !extent_ok(h.off_tiles, ((uint64_t) 1 << h.tile_bits) + 1, sizeof(int32_t), bytes)
      || !extent_ok(h.off_seed16, cells, sizeof(uint16_t), bytes)
      || h.memo_bits > 31
      || (h.memo_bits && (h.off_memo % 32 != 0 || !extent_ok(h.off_memo, (uint64_t) 1 << h.memo_bits, sizeof(gsi_memo), bytes)))
      || h.seed_shift < 3 || h.seed_shift > 15 || gsi_seed_blocks(h.grid_g, h.seed_shift) > GSI_SEED_GB_MAX
      || !extent_ok(h.off_seedbase, (uint64_t) gsi_seed_blocks(h.grid_g, h.seed_shift) * gsi_seed_blocks(h.grid_g, h.seed_shift),
                    sizeof(int32_t), bytes)
      || (h.off_dag | h.off_leaf | h.off_z | h.off_seeds | h.off_entry | h.off_tiles | h.off_seed16 | h.off_seedbase) % 32 != 0
      || !(h.inv_du > 0) || !(h.inv_dv > 0) || !(h.delta > 0))
    GSI_ERROR("slab sections out of range", GSL_SCATTERED_EINVAL);
  device_guard restore;
  drop_replicas(w);
  w->built = false;
  w->hdr = h;
  std::vector<int> devs = w->devs;
  if (devs.empty()) devs.push_back(restore.prev >= 0 ? restore.prev : 0);
  for (int d : devs)
    {
      replica *r = nullptr;
      int s = make_replica(d, &r);
      if (r) w->reps.push_back(r);
      if (s) return s;
      apply_l2_fetch(w);
      CU(slab_alloc(r, bytes, w->vmm && devs.size() == 1));
      r->view = gsi_view_of(r->slab, &w->hdr);
      if (!w->exact_entry) r->view.entry = nullptr;
      CU(cudaMemcpy(r->slab, src, bytes, cudaMemcpyDefault));
      /* links, neighbour slots, seeds, entry nodes and tile boundaries must stay inside their
         sections: one pass on the device before any evaluation kernel trusts them */
      unsigned long long *bad = nullptr;
      CU(cudaMalloc(&bad, sizeof *bad));
      cudaError_t e = gsi_launch_validate_slab(r->view, bad, r->sm_count, nullptr);
      if (e == cudaSuccess) e = gsi_launch_leaf_rcp(r->view, r->slab + w->hdr.off_leaf, r->sm_count, nullptr);
      unsigned long long n_bad = 0;
      if (e == cudaSuccess) e = cudaMemcpy(&n_bad, bad, sizeof n_bad, cudaMemcpyDeviceToHost);
      cudaFree(bad);
      CU(e);
      if (n_bad)
        {
          drop_replicas(w);
          GSI_ERROR("slab image holds out-of-range links", GSL_SCATTERED_EINVAL);
        }
    }
  w->uploaded = true;
  return GSL_SCATTERED_SUCCESS;
}

} // namespace

/* ============================================================================================ */

extern "C" {

const char *
gsl_scattered_cuda_version(void)
{
  return "gsl_scattered_cuda 0.2 (sm_100a)";
}

gsl_scattered_cuda_error_handler_t *
gsl_scattered_cuda_set_error_handler(gsl_scattered_cuda_error_handler_t *handler)
{
  gsl_scattered_cuda_error_handler_t *old = g_handler;
  g_handler = handler;
  return old;
}

int
gsl_scattered_cuda_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
  return n;
}

gsl_scattered_cuda *
gsl_scattered_cuda_alloc(size_t n_nodes, size_t n_fields)
{
  if (n_fields == 0) GSI_ERROR_NULL("n_fields must be positive", GSL_SCATTERED_EINVAL);
  if (n_nodes > (size_t) 200000000) GSI_ERROR_NULL("n_nodes too large", GSL_SCATTERED_EINVAL);
  gsl_scattered_cuda *w = nullptr;
  try
    {
      w = new gsl_scattered_cuda();
    }
  catch (...)
    {
      GSI_ERROR_NULL("failed to allocate workspace", GSL_SCATTERED_ENOMEM);
    }
  w->n_cap = n_nodes;
  w->n_fields = n_fields;
  std::memset(&w->hdr, 0, sizeof w->hdr);
  return w;
}

void
gsl_scattered_cuda_free(gsl_scattered_cuda *w)
{
  if (!w) return;
  {
    device_guard restore;
    drop_replicas(w);
  }
  delete w;
}

int
gsl_scattered_cuda_set_devices(gsl_scattered_cuda *w, const int dev[], size_t ndev)
{
  if (!w || (ndev && !dev)) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (w->uploaded) GSI_ERROR("set_devices must precede init/upload", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  std::vector<int> d(dev, dev + ndev);
  for (size_t i = 0; i < d.size(); i++)
    for (size_t j = i + 1; j < d.size(); j++)
      if (d[i] == d[j]) GSI_ERROR("duplicate device id", GSL_SCATTERED_EINVAL);
  w->devs = d;
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_set_param(gsl_scattered_cuda *w, const char *name, double value)
{
  if (!w || !name) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  const std::string k(name);
  if (k == "cells_per_leaf" && value > 0) w->fprm.cells_per_leaf = value;
  else if (k == "grid_g" && value >= 0 && value <= 32768) w->fprm.grid_g = (uint32_t) value;
  else if (k == "max_steps" && value >= 1) w->max_steps = (int) value;
  else if (k == "delta" && value > 0) w->fprm.delta = value;
  else if (k == "leaf_stride" && (value == 96 || value == 128)) w->fprm.leaf_stride = (uint32_t) value;
  else if (k == "vertex_memo" && (value == 0 || value == 1)) w->fprm.vertex_memo = (int) value;
  else if (k == "mode" && (value == 0 || value == 1)) w->mode = (int) value;
  else if (k == "stats") w->stats = value != 0;
  else if (k == "chunk" && value >= 1) w->chunk = (size_t) value;
  else if (k == "host_chunk" && value >= 1) w->host_chunk = (size_t) value;
  else if (k == "ctas_per_sm" && value >= 1) w->ctas_per_sm = (int) value;
  else if (k == "profile") w->profile = value != 0;
  else if (k == "kernel" && (value == 0 || value == 1 || value == 2)) w->kernel = (int) value;
  else if (k == "l2_fetch" && (value == 0 || value == 32 || value == 64 || value == 128)) w->l2_fetch = (int) value;
  else if (k == "walk_ctas_per_sm" && value >= 1) w->walk_ctas_per_sm = (int) value;
  else if (k == "walk_chunk" && value >= 0) w->walk_chunk = ((size_t) value + 31) & ~(size_t) 31;
  else if (k == "list_fraction" && value > 0 && value <= 1) w->list_fraction = value;
  else if (k == "reorder" && (value == 0 || value == 1 || value == 2)) w->reorder = (int) value;
  else if (k == "reorder_window" && value >= 32) w->reorder_window = (size_t) value;
  else if (k == "reorder_min" && value >= 0) w->reorder_min = (size_t) value;
  else if (k == "reorder_min_leaves" && value >= 0) w->reorder_min_leaves = (size_t) value;
  else if (k == "reorder_chunk" && value >= 32 && value <= GSI_BIN_CHUNK_MAX) w->reorder_chunk = (int) value;
  else if (k == "reorder_max_fields" && value >= 0) w->reorder_max_fields = (size_t) value;
  else if (k == "tile_ranges" && value >= 0 && value <= 64) w->tile_ranges = (int) value;
  else if (k == "tile_threads" && value >= 32 && value <= GSI_TILE_THREADS) w->tile_threads = (int) value;
  else if (k == "tile_leaves" && value >= 1) w->fprm.tile_leaves = (uint32_t) value;
  else if (k == "vmm" && (value == 0 || value == 1)) w->vmm = (int) value;
  else if (k == "build_device" && (value == 0 || value == 1)) w->build_device = (int) value;
  else if (k == "exact_coop" && (value == 0 || value == 1 || value == 2)) w->exact_coop = (int) value;
  else if (k == "walk_variant" && value >= 0 && value <= 3) w->walk_variant = (int) value;
  else if (k == "walk_dynamic" && (value == 0 || value == 1)) w->walk_dynamic = (int) value;
  else if (k == "walk_carveout" && value >= -1 && value <= 100) w->walk_carveout = (int) value;
  else if (k == "walk_l1" && (value == -1 || value == 0 || value == 1)) w->walk_l1 = (int) value;
  else if (k == "walk_pf" && value >= 0 && value <= 4) w->walk_pf = (int) value;
  else if (k == "walk_trace" && (value == 0 || value == 1)) w->walk_trace = (int) value;
  else if (k == "copy_threads" && value >= 1 && value <= 64) w->copy_threads = (int) value;
  else if (k == "pageable_min" && value >= 0) w->pageable_min = (size_t) value;
  else if (k == "exact_entry" && (value == 0 || value == 1))
    {
      w->exact_entry = (int) value;
      for (replica *r : w->reps)
        r->view.entry = value != 0 ? (const int32_t *) (r->slab + w->hdr.off_entry) : nullptr;
    }
  else GSI_ERROR("unknown parameter or value out of range", GSL_SCATTERED_EINVAL);
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_shuffle_mt19937(unsigned long seed, size_t n, size_t order[])
{
  if (n && !order) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  std::vector<int64_t> o(n);
  gsi_shuffle_mt19937(seed, n, o.data());
  for (size_t i = 0; i < n; i++) order[i] = (size_t) o[i];
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_build(gsl_scattered_cuda *w, const double xa[], const double ya[],
                         size_t n_nodes, int init_flags, const size_t order[])
{
  if (!w) GSI_ERROR("null workspace", GSL_SCATTERED_EINVAL);
  if (n_nodes && (!xa || !ya)) GSI_ERROR("null coordinate array", GSL_SCATTERED_EINVAL);
  /* simplex_tree_init returns GSL_FAILURE when the tree has no room, linear_simplex.c:274-278 */
  if (n_nodes > w->n_cap) GSI_ERROR("more nodes than the workspace was allocated for", GSL_SCATTERED_EBADLEN);
  GSI_GUARD_BEGIN
  {
    device_guard restore;
    drop_replicas(w);
  }
  w->built = false;
  w->xa.assign(xa, xa + n_nodes);
  w->ya.assign(ya, ya + n_nodes);
  std::vector<int64_t> ord;
  if (order)
    {
      ord.resize(n_nodes);
      for (size_t i = 0; i < n_nodes; i++) ord[i] = (int64_t) order[i];
    }
  std::string err;
  /* with a CUDA device the builder's own point locations run there in
     batches (same tree, node for node); without one, or when it cannot be set up, on the host */
  cuda_build_locator *loc = nullptr;
  if (w->build_device && n_nodes > 4096)
    {
      int ndev = 0;
      if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0)
        {
          int dev = 0;
          if (!w->devs.empty()) dev = w->devs[0];
          else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
          loc = new (std::nothrow) cuda_build_locator(dev);
        }
      cudaGetLastError();
    }
  int s;
  try
    {
      s = gsi_build_tree(w->xa.data(), w->ya.data(), n_nodes, init_flags, order ? ord.data() : nullptr, &w->tree, &err, loc);
    }
  catch (...)
    {
      delete loc;
      throw;
    }
  delete loc;
  if (s) GSI_ERROR(err.empty() ? "triangulation build failed" : err.c_str(), s);
  w->built = true;
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_upload(gsl_scattered_cuda *w, const double za[], size_t n_fields)
{
  if (!w) GSI_ERROR("null workspace", GSL_SCATTERED_EINVAL);
  if (!w->built) GSI_ERROR("upload before build", GSL_SCATTERED_EINVAL);
  if (n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  if (w->tree.n_points && !za) GSI_ERROR("null value array", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0)
    GSI_ERROR("no CUDA device available: libgsl_scattered_cuda has no CPU evaluation path",
              GSL_SCATTERED_EFAILED);
  GSI_GUARD_BEGIN
  double t0 = now_s();
  std::vector<char> image;
  std::string err;
  int s = gsi_flatten(w->tree, w->xa.data(), w->ya.data(), za, n_fields, w->fprm, &image, &err);
  if (s) GSI_ERROR(err.empty() ? "flatten failed" : err.c_str(), s);
  w->flatten_seconds = now_s() - t0;
  t0 = now_s();
  s = install_image(w, image);
  w->upload_seconds = now_s() - t0;
  return s;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_init(gsl_scattered_cuda *w, const double xa[], const double ya[],
                        const double za[], size_t n_nodes, int init_flags, const size_t order[])
{
  const int s = gsl_scattered_cuda_build(w, xa, ya, n_nodes, init_flags, order);
  if (s) return s;
  return gsl_scattered_cuda_upload(w, za, w->n_fields);
}

int
gsl_scattered_cuda_init_from_tree(gsl_scattered_cuda *w, size_t n_simplexes, const int pts3[],
                                  const int lnk3[], const int type[], const size_t shuffle[],
                                  const double shift[2], const double scale[2],
                                  const double seed_points[6], const double data[],
                                  size_t data_tda, size_t n_nodes, const double za[],
                                  size_t n_fields)
{
  if (!w || !pts3 || !lnk3 || !type || !shift || !scale || !seed_points || n_simplexes == 0)
    GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (n_nodes && (!shuffle || !data || data_tda < 2))
    GSI_ERROR("null data/shuffle or bad stride", GSL_SCATTERED_EINVAL);
  if (n_nodes > w->n_cap) GSI_ERROR("more nodes than the workspace was allocated for", GSL_SCATTERED_EBADLEN);
  if (n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  if (n_simplexes > (size_t) 0x7fffffff) GSI_ERROR("too many simplexes", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  /* validate everything before touching the workspace: a failed call leaves it as it was */
  for (size_t i = 0; i < n_simplexes; i++)
    {
      if (type[i] < 0 || type[i] > 3) GSI_ERROR("bad node type", GSL_SCATTERED_EINVAL);
      for (int k = 0; k < 3; k++)
        {
          const int p = pts3[3 * i + k], l = lnk3[3 * i + k];
          if (p < -3 || (p >= 0 && (size_t) p >= n_nodes)) GSI_ERROR("vertex id out of range", GSL_SCATTERED_EINVAL);
          if (l < 0 || (size_t) l >= n_simplexes) GSI_ERROR("link out of range", GSL_SCATTERED_EINVAL);
        }
    }
  {
    std::vector<bool> seen(n_nodes, false);
    for (size_t i = 0; i < n_nodes; i++)
      {
        if (shuffle[i] >= n_nodes || seen[shuffle[i]]) GSI_ERROR("shuffle is not a permutation", GSL_SCATTERED_EINVAL);
        seen[shuffle[i]] = true;
      }
  }
  {
    device_guard restore;
    drop_replicas(w);
  }
  w->built = false;
  gsi_host_tree &t = w->tree;
  t.n_points = n_nodes;
  t.pts.assign(pts3, pts3 + 3 * n_simplexes);
  t.lnk.assign(lnk3, lnk3 + 3 * n_simplexes);
  t.type.resize(n_simplexes);
  for (size_t i = 0; i < n_simplexes; i++) t.type[i] = (uint8_t) type[i];
  t.shuffle.resize(n_nodes);
  w->xa.resize(n_nodes);
  w->ya.resize(n_nodes);
  for (size_t i = 0; i < n_nodes; i++)
    {
      t.shuffle[i] = (int64_t) shuffle[i];
      w->xa[i] = data[i * data_tda];
      w->ya[i] = data[i * data_tda + 1];
    }
  t.xf.sx = scale[0]; t.xf.sy = scale[1];
  t.xf.cx = shift[0]; t.xf.cy = shift[1];
  std::memcpy(t.seed, seed_points, sizeof t.seed);
  t.build_seconds = 0;
  w->built = true;
  return gsl_scattered_cuda_upload(w, za, n_fields);
  GSI_GUARD_END
}

int
gsl_scattered_cuda_eval_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                          size_t nq, double zq[], int leaf[])
{
  if (nq && (!xq || !yq)) GSI_ERROR("null query array", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  return eval_host(w, xq, yq, nullptr, nq, zq, leaf);
  GSI_GUARD_END
}

double
gsl_scattered_cuda_eval(const gsl_scattered_cuda *w, double x, double y)
{
  try
    {
      std::vector<double> z(w ? w->n_fields : 1, std::numeric_limits<double>::quiet_NaN());
      const int s = eval_host(w, &x, &y, nullptr, 1, z.data(), nullptr);
      if (s != GSL_SCATTERED_SUCCESS && s != GSL_SCATTERED_EDOM)
        return std::numeric_limits<double>::quiet_NaN();
      return z[0];
    }
  catch (...)
    {
      report("internal error (C++ exception)", __FILE__, __LINE__, GSL_SCATTERED_EFAILED);
      return std::numeric_limits<double>::quiet_NaN();
    }
}

int
gsl_scattered_cuda_eval_raster_e(const gsl_scattered_cuda *w, double x0, double dx, size_t nx,
                                 double y0, double dy, size_t ny, double zq[], int leaf[])
{
  if (nx && ny && nx > std::numeric_limits<size_t>::max() / ny)
    GSI_ERROR("raster too large", GSL_SCATTERED_EINVAL);
  if (nx == 0 || ny == 0) return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_BEGIN
  raster_spec rs{x0, dx, y0, dy, ny, 0, nx};
  return eval_host(w, nullptr, nullptr, &rs, nx * ny, zq, leaf);
  GSI_GUARD_END
}

int
gsl_scattered_cuda_eval_device_e(const gsl_scattered_cuda *w, int dev, const double *d_xq,
                                 const double *d_yq, size_t nq, double *d_zq, size_t ld,
                                 int *d_leaf, void *stream)
{
  if (!w || !w->uploaded) GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  if (nq && (!d_xq || !d_yq)) GSI_ERROR("null query array", GSL_SCATTERED_EINVAL);
  if (d_zq && ld < nq) GSI_ERROR("ld < nq", GSL_SCATTERED_EBADLEN);
  if (nq == 0) return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  ctx_lease lease(r, (cudaStream_t) stream);
  if (lease.status) return lease.status;
  eval_ctx *c = lease.c;
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  int s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(lease.st));
      if ((s = read_counters(c, c0))) return s;
    }
  s = eval_on_device(w, r, c, d_xq, d_yq, nullptr, nq, d_zq, ld, d_leaf, lease.st);
  if (s) return s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(lease.st));
      if ((s = read_counters(c, c1))) return s;
      if (c1[GSI_CNT_OUTSIDE] != c0[GSI_CNT_OUTSIDE])
        GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
    }
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_eval_raster_device_e(const gsl_scattered_cuda *w, int dev, double x0, double dx,
                                        size_t nx, double y0, double dy, size_t ny, double *d_zq,
                                        int *d_leaf, void *stream)
{
  if (!w || !w->uploaded) GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  if (nx && ny && nx > std::numeric_limits<size_t>::max() / ny)
    GSI_ERROR("raster too large", GSL_SCATTERED_EINVAL);
  if (nx == 0 || ny == 0) return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  ctx_lease lease(r, (cudaStream_t) stream);
  if (lease.status) return lease.status;
  eval_ctx *c = lease.c;
  raster_spec rs{x0, dx, y0, dy, ny, 0, nx};
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  int s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(lease.st));
      if ((s = read_counters(c, c0))) return s;
    }
  s = eval_on_device(w, r, c, nullptr, nullptr, &rs, nx * ny, d_zq, nx * ny, d_leaf, lease.st);
  if (s) return s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(lease.st));
      if ((s = read_counters(c, c1))) return s;
      if (c1[GSI_CNT_OUTSIDE] != c0[GSI_CNT_OUTSIDE])
        GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
    }
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_get_info(const gsl_scattered_cuda *w, gsl_scattered_cuda_info *info)
{
  if (!w || !info) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  std::memset(info, 0, sizeof *info);
  info->n_nodes = w->built ? w->tree.n_points : (size_t) w->hdr.n_points;
  info->n_fields = w->n_fields;
  info->n_simplexes = w->built ? w->tree.n_nodes() : (size_t) w->hdr.n_dag;
  info->n_leaves = (size_t) w->hdr.n_leaves;
  info->grid_g = w->hdr.grid_g;
  info->slab_bytes = (size_t) w->hdr.total_bytes;
  info->n_devices = w->reps.size();
  info->build_seconds = w->tree.build_seconds;
  info->flatten_seconds = w->flatten_seconds;
  info->upload_seconds = w->upload_seconds;
  device_guard restore;
  for (replica *r : w->reps)
    {
      CU(cudaSetDevice(r->dev));
      CU(cudaDeviceSynchronize()); /* every context's counters and profiling events are final */
      std::lock_guard<std::mutex> lock(r->mu);
      info->n_contexts += r->ctxs.size();
      for (eval_ctx *c : r->ctxs)
        {
          unsigned cn[GSI_CNT_WORDS];
          const int s = read_counters(c, cn);
          if (s) return s;
          unsigned long long u;
          std::memcpy(&u, cn + GSI_CNT_DEFERRED_TOTAL, 8);
          info->n_deferred += u;
          std::memcpy(&u, cn + GSI_CNT_STEPS, 8);
          info->n_steps += u;
          info->n_outside += cn[GSI_CNT_OUTSIDE];
          /* fold finished profiling events into the running totals */
          for (auto &p : c->prof_locate)
            {
              float ms = 0;
              if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) c->locate_seconds += 1e-3 * ms;
              cudaEventDestroy(p.first); cudaEventDestroy(p.second);
            }
          for (auto &p : c->prof_exact)
            {
              float ms = 0;
              if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) c->exact_seconds += 1e-3 * ms;
              cudaEventDestroy(p.first); cudaEventDestroy(p.second);
            }
          c->prof_locate.clear();
          c->prof_exact.clear();
          cudaGetLastError();
          info->locate_seconds += c->locate_seconds;
          info->exact_seconds += c->exact_seconds;
          info->n_kernel_launches += c->n_launches;
          info->n_reordered += c->n_reordered;
        }
    }
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_reset_counters(const gsl_scattered_cuda *w)
{
  if (!w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  device_guard restore;
  for (replica *r : w->reps)
    {
      CU(cudaSetDevice(r->dev));
      CU(cudaDeviceSynchronize());
      std::lock_guard<std::mutex> lock(r->mu);
      for (eval_ctx *c : r->ctxs)
        {
          CU(cudaMemset(c->counters, 0, GSI_CNT_WORDS * sizeof(unsigned)));
          for (auto &p : c->prof_locate) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
          for (auto &p : c->prof_exact) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
          c->prof_locate.clear();
          c->prof_exact.clear();
          c->locate_seconds = c->exact_seconds = 0;
          c->n_launches = 0;
          c->n_reordered = 0;
        }
    }
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_get_tree(const gsl_scattered_cuda *w, int pts3[], int lnk3[], int type[],
                            size_t shuffle[], double shift[2], double scale[2],
                            double seed_points[6])
{
  if (!w || !w->built) GSI_ERROR("no host tree in this workspace", GSL_SCATTERED_EINVAL);
  const gsi_host_tree &t = w->tree;
  const size_t N = t.n_nodes();
  if (pts3) std::copy(t.pts.begin(), t.pts.end(), pts3);
  if (lnk3) std::copy(t.lnk.begin(), t.lnk.end(), lnk3);
  if (type)
    for (size_t i = 0; i < N; i++) type[i] = t.type[i];
  if (shuffle)
    for (size_t i = 0; i < t.n_points; i++) shuffle[i] = (size_t) t.shuffle[i];
  if (shift) { shift[0] = t.xf.cx; shift[1] = t.xf.cy; }
  if (scale) { scale[0] = t.xf.sx; scale[1] = t.xf.sy; }
  if (seed_points) std::memcpy(seed_points, t.seed, sizeof t.seed);
  return GSL_SCATTERED_SUCCESS;
}

size_t
gsl_scattered_cuda_slab_bytes(const gsl_scattered_cuda *w)
{
  return (w && w->uploaded) ? (size_t) w->hdr.total_bytes : 0;
}

int
gsl_scattered_cuda_slab_export_device(const gsl_scattered_cuda *w, int dev, void *d_dst,
                                      void *stream)
{
  if (!w || !w->uploaded || !d_dst) GSI_ERROR("nothing to export", GSL_SCATTERED_EINVAL);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  device_guard restore;
  CU(cudaSetDevice(dev));
  /* the slab is immutable after init: any stream may read it */
  CU(cudaMemcpyAsync(d_dst, r->slab, w->hdr.total_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
  if (!stream) CU(cudaStreamSynchronize(nullptr));
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_init_from_slab_device(gsl_scattered_cuda *w, const void *d_src, size_t bytes)
{
  if (!w || !d_src || bytes < sizeof(gsi_slab_header)) GSI_ERROR("bad slab", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0)
    GSI_ERROR("no CUDA device available", GSL_SCATTERED_EFAILED);
  GSI_GUARD_BEGIN
  gsi_slab_header h;
  CU(cudaMemcpy(&h, d_src, sizeof h, cudaMemcpyDefault));
  return adopt_slab(w, h, d_src, bytes);
  GSI_GUARD_END
}

int
gsl_scattered_cuda_fwrite(FILE *stream, const gsl_scattered_cuda *w)
{
  if (!stream || !w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (!w->uploaded || w->reps.empty()) GSI_ERROR("nothing to write: workspace not initialised", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  replica *r = w->reps[0];
  std::vector<char> image((size_t) w->hdr.total_bytes);
  {
    device_guard restore;
    CU(cudaSetDevice(r->dev));
    CU(cudaMemcpy(image.data(), r->slab, image.size(), cudaMemcpyDeviceToHost));
  }
  gsi_slab_header h;
  std::memcpy(&h, image.data(), sizeof h);
  h.reserved[0] = slab_checksum(image.data(), image.size());
  std::memcpy(image.data(), &h, sizeof h);
  if (fwrite(image.data(), 1, image.size(), stream) != image.size())
    GSI_ERROR("fwrite failed", GSL_SCATTERED_EFAILED);
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_fread(FILE *stream, gsl_scattered_cuda *w)
{
  if (!stream || !w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  gsi_slab_header h;
  if (fread(&h, 1, sizeof h, stream) != sizeof h) GSI_ERROR("fread failed", GSL_SCATTERED_EFAILED);
  if (h.magic != GSI_SLAB_MAGIC || h.version != GSI_SLAB_VERSION || h.total_bytes < sizeof h
      || h.total_bytes > ((uint64_t) 1 << 40))
    GSI_ERROR("not a slab image of this library version", GSL_SCATTERED_EINVAL);
  if (h.n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  std::vector<char> image((size_t) h.total_bytes);
  std::memcpy(image.data(), &h, sizeof h);
  const size_t rest = image.size() - sizeof h;
  if (fread(image.data() + sizeof h, 1, rest, stream) != rest) GSI_ERROR("fread failed", GSL_SCATTERED_EFAILED);
  if (h.reserved[0] != slab_checksum(image.data(), image.size()))
    GSI_ERROR("slab image checksum mismatch", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0)
    GSI_ERROR("no CUDA device available: libgsl_scattered_cuda has no CPU evaluation path",
              GSL_SCATTERED_EFAILED);
  return adopt_slab(w, h, image.data(), image.size());
  GSI_GUARD_END
}

int
gsl_scattered_cuda_save(const gsl_scattered_cuda *w, const char *path)
{
  if (!w || !path) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (!w->uploaded || w->reps.empty()) GSI_ERROR("nothing to write: workspace not initialised", GSL_SCATTERED_EINVAL);
  FILE *f = fopen(path, "wb");
  if (!f) GSI_ERROR("cannot open file for writing", GSL_SCATTERED_EFAILED);
  const int s = gsl_scattered_cuda_fwrite(f, w);
  if (fclose(f) != 0 && s == GSL_SCATTERED_SUCCESS) GSI_ERROR("fclose failed", GSL_SCATTERED_EFAILED);
  return s;
}

int
gsl_scattered_cuda_load(gsl_scattered_cuda *w, const char *path)
{
  if (!w || !path) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  FILE *f = fopen(path, "rb");
  if (!f) GSI_ERROR("cannot open file for reading", GSL_SCATTERED_EFAILED);
  const int s = gsl_scattered_cuda_fread(f, w);
  fclose(f);
  return s;
}

void *
gsl_scattered_cuda_host_alloc(size_t bytes)
{
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess)
    {
      cudaGetLastError();
      return nullptr;
    }
  return p;
}

void
gsl_scattered_cuda_host_free(void *p)
{
  if (p) cudaFreeHost(p);
}

int
gsl_scattered_cuda_generate_uniform_device(const gsl_scattered_cuda *w, int dev,
                                           unsigned long long seed, unsigned long long first,
                                           size_t nq, double xlo, double xhi, double ylo,
                                           double yhi, double *d_xq, double *d_yq, void *stream)
{
  if (!w || !w->uploaded) GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  device_guard restore;
  CU(cudaSetDevice(dev));
  cudaStream_t st = (cudaStream_t) stream;
  CU(gsi_launch_gen_uniform(seed, first, nq, xlo, xhi, ylo, yhi, d_xq, d_yq, r->sm_count, st));
  if (!stream) CU(cudaStreamSynchronize(st));
  return GSL_SCATTERED_SUCCESS;
}

void
gsl_scattered_cuda_generate_uniform_host(unsigned long long seed, unsigned long long first,
                                         size_t nq, double xlo, double xhi, double ylo, double yhi,
                                         double xq[], double yq[])
{
  const double xw = xhi - xlo, yw = yhi - ylo;
  for (size_t q = 0; q < nq; q++)
    {
      const uint64_t k = first + q;
      const double ux = xw * gsi_uniform(seed, 2 * k), uy = yw * gsi_uniform(seed, 2 * k + 1);
      xq[q] = xlo + ux;
      yq[q] = ylo + uy;
    }
}

int
gsl_scattered_cuda_checksum_device(const gsl_scattered_cuda *w, int dev, const int *d_leaf,
                                   const double *d_val, size_t n, unsigned long long out[2],
                                   void *stream)
{
  if (!w || !w->uploaded || !out) GSI_ERROR("bad argument", GSL_SCATTERED_EINVAL);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  ctx_lease lease(r, (cudaStream_t) stream);
  if (lease.status) return lease.status;
  CU(gsi_launch_checksum(d_leaf, d_val, n, lease.c->scratch2, r->sm_count, lease.st));
  CU(cudaMemcpyAsync(out, lease.c->scratch2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, lease.st));
  CU(cudaStreamSynchronize(lease.st));
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

int
gsl_scattered_cuda_selftest_division(int dev, gsl_scattered_cuda_u64 seed, size_t n, gsl_scattered_cuda_u64 out[2])
{
  if (!out) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0) GSI_ERROR("no CUDA device available", GSL_SCATTERED_EFAILED);
  device_guard restore;
  CU(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  unsigned long long *d = nullptr;
  CU(cudaMalloc(&d, 2 * sizeof *d));
  cudaError_t e = gsi_launch_selftest_division(seed, n, d, prop.multiProcessorCount, nullptr);
  unsigned long long h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CU(e);
  out[0] = h[0];
  out[1] = h[1];
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_compare_device(const gsl_scattered_cuda *w, int dev, const int *d_leaf_a,
                                  const int *d_leaf_b, const double *d_val_a,
                                  const double *d_val_b, size_t n, size_t n_val,
                                  unsigned long long out[2], void *stream)
{
  if (!w || !w->uploaded || !out) GSI_ERROR("bad argument", GSL_SCATTERED_EINVAL);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  ctx_lease lease(r, (cudaStream_t) stream);
  if (lease.status) return lease.status;
  CU(gsi_launch_compare(d_leaf_a, d_leaf_b, d_val_a, d_val_b, n, n_val, lease.c->scratch2, r->sm_count, lease.st));
  CU(cudaMemcpyAsync(out, lease.c->scratch2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, lease.st));
  CU(cudaStreamSynchronize(lease.st));
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

/* how much of the last build went through the batched device locate */
int
gsl_scattered_cuda_build_stats(const gsl_scattered_cuda *w, double *batched_seconds, size_t *batched_points, size_t *rounds)
{
  if (!w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (batched_seconds) *batched_seconds = w->tree.batched_seconds;
  if (batched_points) *batched_points = w->tree.batched_points;
  if (rounds) *rounds = w->tree.batched_rounds;
  return GSL_SCATTERED_SUCCESS;
}

/* diagnostic (param "walk_trace"): per-CTA {smid, start ns, end ns, queries handed out} of the most
 * recent k_walkp launch of the replica's first evaluation context; *n_words = words written */
int
gsl_scattered_cuda_debug_walk_trace(const gsl_scattered_cuda *w, int dev, unsigned long long *out, size_t cap_words,
                                    size_t *n_words)
{
  if (!w || !w->uploaded || !out || !n_words) GSI_ERROR("bad argument", GSL_SCATTERED_EINVAL);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  CU(cudaDeviceSynchronize());
  *n_words = 0;
  std::lock_guard<std::mutex> lock(r->mu);
  for (eval_ctx *c : r->ctxs)
    if (c->trace && c->trace_ctas)
      {
        const size_t n = std::min(cap_words, c->trace_ctas * 4);
        CU(cudaMemcpy(out, c->trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        *n_words = n;
        break;
      }
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

} /* extern "C" */

/* ---- derivatives and the type descriptor --------------------------------------------------- */
namespace
{

/* locate (leaf ids only) + gradient kernel on device data, all on stream st; d_leaf is scratch of
 * nq ints (caller's or the context's) */
int
deriv_on_device(const gsl_scattered_cuda *w, const replica *r, eval_ctx *c, const double *d_x, const double *d_y,
                size_t nq, double *d_dzdx, double *d_dzdy, size_t ld, int *d_leaf, cudaStream_t st)
{
  int s = eval_on_device(w, r, c, d_x, d_y, nullptr, nq, nullptr, 0, d_leaf, st);
  if (s) return s;
  CU(gsi_launch_deriv(r->view, d_leaf, nq, d_dzdx, d_dzdy, ld, r->sm_count, st));
  c->n_launches += 1;
  return 0;
}

/* host arrays through the context's device staging, chunk by chunk (not the hot path: plain
 * synchronous copies).  order: 1 = first derivatives, 2 = second derivatives (zero inside the
 * cage, NaN outside) written to dzdx */
int
deriv_host(const gsl_scattered_cuda *w, const double *xq, const double *yq, size_t nq, double *dzdx, double *dzdy,
           int order)
{
  if (!w || !w->uploaded || w->reps.empty())
    GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  if (nq && (!xq || !yq)) GSI_ERROR("null query array", GSL_SCATTERED_EINVAL);
  if (nq == 0) return GSL_SCATTERED_SUCCESS;
  device_guard restore;
  replica *r = w->reps[0];
  CU(cudaSetDevice(r->dev));
  ctx_lease lease(r, nullptr);
  if (lease.status) return lease.status;
  eval_ctx *c = lease.c;
  const size_t F = w->n_fields;
  const size_t C = std::max<size_t>(1, std::min(w->host_chunk, nq));
  int s = ensure_stage(c, C, F);
  if (s) return s;
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  CU(cudaStreamSynchronize(c->s_k));
  if ((s = read_counters(c, c0))) return s;
  std::vector<int> hl;
  if (order == 2) hl.resize(C);
  for (size_t off = 0; off < nq; off += C)
    {
      const size_t n = std::min(C, nq - off);
      stage_buf &a = c->st[0], &b = c->st[1];
      CU(cudaMemcpyAsync(a.x, xq + off, n * sizeof(double), cudaMemcpyHostToDevice, c->s_k));
      CU(cudaMemcpyAsync(a.y, yq + off, n * sizeof(double), cudaMemcpyHostToDevice, c->s_k));
      if (order == 1)
        {
          s = deriv_on_device(w, r, c, a.x, a.y, n, dzdx ? a.z : nullptr, dzdy ? b.z : nullptr, C, a.leaf, c->s_k);
          if (s) return s;
          for (size_t f = 0; f < F; f++)
            {
              if (dzdx) CU(cudaMemcpyAsync(dzdx + f * nq + off, a.z + f * C, n * sizeof(double), cudaMemcpyDeviceToHost, c->s_k));
              if (dzdy) CU(cudaMemcpyAsync(dzdy + f * nq + off, b.z + f * C, n * sizeof(double), cudaMemcpyDeviceToHost, c->s_k));
            }
          CU(cudaStreamSynchronize(c->s_k));
        }
      else
        {
          s = eval_on_device(w, r, c, a.x, a.y, nullptr, n, nullptr, 0, a.leaf, c->s_k);
          if (s) return s;
          CU(cudaMemcpyAsync(hl.data(), a.leaf, n * sizeof(int), cudaMemcpyDeviceToHost, c->s_k));
          CU(cudaStreamSynchronize(c->s_k));
          if (dzdx)
            for (size_t f = 0; f < F; f++)
              for (size_t q = 0; q < n; q++)
                dzdx[f * nq + off + q] = hl[q] >= 0 ? 0.0 : std::numeric_limits<double>::quiet_NaN();
        }
    }
  if ((s = read_counters(c, c1))) return s;
  if (c1[GSI_CNT_OUTSIDE] != c0[GSI_CNT_OUTSIDE])
    GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
  return GSL_SCATTERED_SUCCESS;
}

const gsl_scattered_cuda_type linear_simplex_type = {
    "linear-simplex-cuda",
    3,
    &gsl_scattered_cuda_alloc,
    &gsl_scattered_cuda_init,
    &gsl_scattered_cuda_eval_e,
    &gsl_scattered_cuda_eval_deriv_x_e,
    &gsl_scattered_cuda_eval_deriv_y_e,
    &gsl_scattered_cuda_eval_deriv_xx_e,
    &gsl_scattered_cuda_eval_deriv_xy_e,
    &gsl_scattered_cuda_eval_deriv_yy_e,
    &gsl_scattered_cuda_free,
};

} // namespace

extern "C" {

const gsl_scattered_cuda_type *gsl_scattered_cuda_linear_simplex = &linear_simplex_type;

const gsl_scattered_cuda_type *
gsl_scattered_cuda_type_linear_simplex(void)
{
  return &linear_simplex_type;
}

const char *
gsl_scattered_cuda_name(const gsl_scattered_cuda *w)
{
  (void) w;
  return linear_simplex_type.name;
}

unsigned int
gsl_scattered_cuda_min_size(const gsl_scattered_cuda *w)
{
  (void) w;
  return linear_simplex_type.min_size;
}

unsigned int
gsl_scattered_cuda_type_min_size(const gsl_scattered_cuda_type *T)
{
  return T ? T->min_size : 0;
}

int
gsl_scattered_cuda_eval_deriv_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq,
                                double dzdx[], double dzdy[])
{
  GSI_GUARD_BEGIN
  return deriv_host(w, xq, yq, nq, dzdx, dzdy, 1);
  GSI_GUARD_END
}

int
gsl_scattered_cuda_eval_deriv_x_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[])
{
  return gsl_scattered_cuda_eval_deriv_e(w, xq, yq, nq, d, nullptr);
}

int
gsl_scattered_cuda_eval_deriv_y_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[])
{
  return gsl_scattered_cuda_eval_deriv_e(w, xq, yq, nq, nullptr, d);
}

int
gsl_scattered_cuda_eval_deriv_xx_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[])
{
  GSI_GUARD_BEGIN
  return deriv_host(w, xq, yq, nq, d, nullptr, 2);
  GSI_GUARD_END
}

int
gsl_scattered_cuda_eval_deriv_yy_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[])
{
  return gsl_scattered_cuda_eval_deriv_xx_e(w, xq, yq, nq, d);
}

int
gsl_scattered_cuda_eval_deriv_xy_e(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[])
{
  return gsl_scattered_cuda_eval_deriv_xx_e(w, xq, yq, nq, d);
}

double
gsl_scattered_cuda_eval_deriv_x(const gsl_scattered_cuda *w, double x, double y)
{
  try
    {
      std::vector<double> d(w ? w->n_fields : 1, std::numeric_limits<double>::quiet_NaN());
      const int s = deriv_host(w, &x, &y, 1, d.data(), nullptr, 1);
      return (s == GSL_SCATTERED_SUCCESS || s == GSL_SCATTERED_EDOM) ? d[0] : std::numeric_limits<double>::quiet_NaN();
    }
  catch (...)
    {
      report("internal error (C++ exception)", __FILE__, __LINE__, GSL_SCATTERED_EFAILED);
      return std::numeric_limits<double>::quiet_NaN();
    }
}

double
gsl_scattered_cuda_eval_deriv_y(const gsl_scattered_cuda *w, double x, double y)
{
  try
    {
      std::vector<double> d(w ? w->n_fields : 1, std::numeric_limits<double>::quiet_NaN());
      const int s = deriv_host(w, &x, &y, 1, nullptr, d.data(), 1);
      return (s == GSL_SCATTERED_SUCCESS || s == GSL_SCATTERED_EDOM) ? d[0] : std::numeric_limits<double>::quiet_NaN();
    }
  catch (...)
    {
      report("internal error (C++ exception)", __FILE__, __LINE__, GSL_SCATTERED_EFAILED);
      return std::numeric_limits<double>::quiet_NaN();
    }
}

int
gsl_scattered_cuda_eval_deriv_device_e(const gsl_scattered_cuda *w, int dev, const double *d_xq, const double *d_yq,
                                       size_t nq, double *d_dzdx, double *d_dzdy, size_t ld, int *d_leaf, void *stream)
{
  if (!w || !w->uploaded) GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  if (nq && (!d_xq || !d_yq)) GSI_ERROR("null query array", GSL_SCATTERED_EINVAL);
  if ((d_dzdx || d_dzdy) && ld < nq) GSI_ERROR("ld < nq", GSL_SCATTERED_EBADLEN);
  if (nq == 0) return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_BEGIN
  device_guard restore;
  CU(cudaSetDevice(dev));
  ctx_lease lease(r, (cudaStream_t) stream);
  if (lease.status) return lease.status;
  eval_ctx *c = lease.c;
  int s;
  if (!d_leaf)
    {
      /* the context's staging holds the leaf ids: fixed-size pieces */
      const size_t C = std::max<size_t>(1, std::min(w->host_chunk, nq));
      if ((s = ensure_stage(c, C, w->n_fields))) return s;
      for (size_t off = 0; off < nq; off += C)
        {
          const size_t n = std::min(C, nq - off);
          s = deriv_on_device(w, r, c, d_xq + off, d_yq + off, n, d_dzdx ? d_dzdx + off : nullptr,
                              d_dzdy ? d_dzdy + off : nullptr, ld, c->st[0].leaf, lease.st);
          if (s) return s;
        }
    }
  else if ((s = deriv_on_device(w, r, c, d_xq, d_yq, nq, d_dzdx, d_dzdy, ld, d_leaf, lease.st)))
    return s;
  if (!stream) CU(cudaStreamSynchronize(lease.st));
  return GSL_SCATTERED_SUCCESS;
  GSI_GUARD_END
}

} /* extern "C" */
