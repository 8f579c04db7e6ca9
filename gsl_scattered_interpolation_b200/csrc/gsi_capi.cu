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
  cudaStream_t s_k = nullptr, s_in = nullptr, s_out = nullptr;
  unsigned *list = nullptr;
  size_t list_cap = 0;
  unsigned *ovf_mask = nullptr; /* overflow bitmask, one bit per query of a launch */
  size_t ovf_words = 0;
  unsigned *counters = nullptr;
  unsigned long long *scratch2 = nullptr; /* 2 x u64 for checksum / compare */
  stage_buf st[2];
  size_t st_cap = 0, st_fields = 0;
  gsi_reorder_scratch ro; /* reorder pipeline scratch (lazily sized) */
  size_t ro_fields = 0, ro_off_entries = 0;
  /* profiling (param "profile"): timed event pairs around every kernel launch */
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_locate, prof_exact;
  double locate_seconds = 0, exact_seconds = 0;
  unsigned long long n_launches = 0, n_reordered = 0;
  std::mutex mu;
};

} // namespace

struct gsl_scattered_cuda
{
  size_t n_cap = 0;
  size_t n_fields = 1;
  gsi_host_tree tree;
  std::vector<double> xa, ya;
  bool built = false, uploaded = false;
  gsi_flatten_params fprm;
  int mode = GSI_MODE_FAST, stats = 0, max_steps = 256, ctas_per_sm = 8, profile = 0;
  int kernel = GSI_KERNEL_AUTO, walk_ctas_per_sm = 4;
  int l2_fetch = 32; /* cudaLimitMaxL2FetchGranularity: random 32-byte sector gathers */
  /* reorder pipeline (gsi_kernels.cu): 0 = off (default: measured slower than the direct walk on
     B200), 1 = always for query arrays, 2 = auto (large batches of few
     fields on a triangulation whose fast-path structure exceeds the L2) */
  int reorder = 0;
  size_t reorder_window = (size_t) 1 << 27, reorder_min = (size_t) 1 << 22;
  int reorder_chunk = GSI_BIN_CHUNK_MAX, tile_ranges = 4, tile_smem = 1;
  size_t reorder_max_fields = 2;
  int exact_entry = 1; /* exact replay starts at the cell's certified entry node (0: at the root) */
  int vmm = 1;         /* slab through cuMemCreate/cuMemMap with the recommended granularity (0: cudaMalloc) */
  size_t walk_chunk = 0;
  size_t chunk = (size_t) 0xfffffff0u;
  double list_fraction = 1.0 / 8;
  size_t host_chunk = (size_t) 1 << 22;
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
free_reorder(replica *r)
{
  gsi_reorder_scratch &o = r->ro;
  if (o.bx) cudaFree(o.bx);
  if (o.by) cudaFree(o.by);
  if (o.bi) cudaFree(o.bi);
  if (o.off) cudaFree(o.off);
  if (o.zb) cudaFree(o.zb);
  if (o.lb) cudaFree(o.lb);
  o = gsi_reorder_scratch();
  r->ro_fields = r->ro_off_entries = 0;
}

void
free_replica(replica *r)
{
  if (!r) return;
  cudaSetDevice(r->dev);
  for (auto &p : r->prof_locate) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto &p : r->prof_exact) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  slab_free(r);
  if (r->list) cudaFree(r->list);
  if (r->ovf_mask) cudaFree(r->ovf_mask);
  if (r->counters) cudaFree(r->counters);
  if (r->scratch2) cudaFree(r->scratch2);
  free_reorder(r);
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = r->st[b];
      if (s.x) cudaFree(s.x);
      if (s.y) cudaFree(s.y);
      if (s.z) cudaFree(s.z);
      if (s.leaf) cudaFree(s.leaf);
      if (s.ev_in) cudaEventDestroy(s.ev_in);
      if (s.ev_k) cudaEventDestroy(s.ev_k);
      if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
  if (r->s_k) cudaStreamDestroy(r->s_k);
  if (r->s_in) cudaStreamDestroy(r->s_in);
  if (r->s_out) cudaStreamDestroy(r->s_out);
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
  CU(cudaStreamCreateWithFlags(&r->s_k, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&r->s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&r->s_out, cudaStreamNonBlocking));
  CU(cudaMalloc(&r->counters, GSI_CNT_WORDS * sizeof(unsigned)));
  CU(cudaMemset(r->counters, 0, GSI_CNT_WORDS * sizeof(unsigned)));
  CU(cudaMalloc(&r->scratch2, 2 * sizeof(unsigned long long)));
  for (int b = 0; b < 2; b++)
    {
      CU(cudaEventCreateWithFlags(&r->st[b].ev_in, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&r->st[b].ev_k, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&r->st[b].ev_out, cudaEventDisableTiming));
    }
  return 0;
}

/* The eval path gathers isolated 32-byte sectors from a structure that exceeds the L2 at
 * 10^6 nodes; the default 64-byte L2 fetch granularity would double the DRAM traffic of every
 * miss.  Device-wide hint (cudaLimitMaxL2FetchGranularity); 0 = leave the driver default. */
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
ensure_list(replica *r, size_t cap, size_t per_launch)
{
  if (r->list_cap < cap)
    {
      if (r->list) cudaFree(r->list);
      r->list = nullptr;
      r->list_cap = 0;
      CU(cudaMalloc(&r->list, cap * sizeof(unsigned)));
      r->list_cap = cap;
    }
  const size_t words = (per_launch + 31) / 32;
  if (r->ovf_words < words)
    {
      if (r->ovf_mask) cudaFree(r->ovf_mask);
      r->ovf_mask = nullptr;
      r->ovf_words = 0;
      CU(cudaMalloc(&r->ovf_mask, words * sizeof(unsigned)));
      CU(cudaMemset(r->ovf_mask, 0, words * sizeof(unsigned))); /* kernels leave it zero */
      r->ovf_words = words;
    }
  return 0;
}

gsi_launch_cfg
cfg_of(const gsl_scattered_cuda *w, const replica *r)
{
  gsi_launch_cfg c;
  c.mode = w->mode;
  c.stats = w->stats;
  c.max_steps = w->max_steps;
  c.sm_count = r->sm_count;
  c.locate_ctas_per_sm = w->ctas_per_sm;
  c.kernel = w->kernel;
  c.walk_ctas_per_sm = w->walk_ctas_per_sm;
  c.walk_chunk = w->walk_chunk;
  c.list = r->list;
  c.list_capacity = r->list_cap;
  c.ovf_mask = r->ovf_mask;
  c.counters = r->counters;
  return c;
}

/* Does a batch of nq incoherent queries go through the reorder pipeline? */
bool
want_reorder(const gsl_scattered_cuda *w, size_t nq)
{
  if (w->mode != GSI_MODE_FAST || w->reorder == 0 || nq == 0) return false;
  if (w->reorder == 1) return true;
  return nq >= w->reorder_min && w->n_fields <= w->reorder_max_fields && w->hdr.n_leaves >= 32768;
}

int
ensure_reorder(const gsl_scattered_cuda *w, replica *r, size_t window, bool want_z, bool want_leaf)
{
  gsi_reorder_scratch &o = r->ro;
  const int bits = (int) w->hdr.tile_bits;
  const int S = std::max(32, std::min(w->reorder_chunk, GSI_BIN_CHUNK_MAX)) & ~31;
  const size_t nt = (size_t) 1 << bits;
  const size_t n_chunks = (window + (size_t) S - 1) / (size_t) S;
  const size_t off_entries = n_chunks * (nt + 2);
  const size_t F = w->n_fields;
  if (o.cap < window || o.chunk != S)
    {
      free_reorder(r);
      CU(cudaMalloc(&o.bx, window * sizeof(double)));
      CU(cudaMalloc(&o.by, window * sizeof(double)));
      CU(cudaMalloc(&o.bi, window * sizeof(unsigned short)));
      o.cap = window;
      o.chunk = S;
    }
  if (r->ro_off_entries < off_entries)
    {
      if (o.off) cudaFree(o.off);
      o.off = nullptr;
      r->ro_off_entries = 0;
      CU(cudaMalloc(&o.off, off_entries * sizeof(unsigned short)));
      r->ro_off_entries = off_entries;
    }
  if (want_z && (!o.zb || r->ro_fields < F))
    {
      if (o.zb) cudaFree(o.zb);
      o.zb = nullptr;
      CU(cudaMalloc(&o.zb, o.cap * F * sizeof(double)));
      r->ro_fields = F;
    }
  if (want_leaf && !o.lb) CU(cudaMalloc(&o.lb, o.cap * sizeof(int)));
  o.tile_max_leaves = (int) w->hdr.tile_max_leaves;
  o.tile_smem = w->tile_smem;
  o.ranges = std::max(1, w->tile_ranges);
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
prepare_launch(const gsl_scattered_cuda *w, replica *r, size_t per_launch, cudaEvent_t ev[4],
               gsi_launch_cfg *cfg)
{
  *cfg = cfg_of(w, r);
  if (w->profile)
    {
      for (int i = 0; i < 4; i++) CU(cudaEventCreate(&ev[i]));
      if (w->mode == GSI_MODE_FAST)
        {
          cfg->ev_locate0 = ev[0]; cfg->ev_locate1 = ev[1];
          r->prof_locate.emplace_back(ev[0], ev[1]);
        }
      else
        {
          cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        }
      cfg->ev_exact0 = ev[2]; cfg->ev_exact1 = ev[3];
      r->prof_exact.emplace_back(ev[2], ev[3]);
    }
  r->n_launches += (w->mode == GSI_MODE_FAST) ? 4 : 1;
  (void) per_launch;
  return 0;
}

/* chunked launch of the locate+blend kernels on device data: either query arrays (ras == NULL)
 * or a block of raster columns.  d_z / d_leaf point at the first query / pixel of the block. */
int
eval_on_device(const gsl_scattered_cuda *w, replica *r, const double *d_x, const double *d_y,
               const raster_spec *ras, size_t nq, double *d_z, size_t ld, int *d_leaf,
               cudaStream_t st)
{
  if (nq == 0) return 0;
  size_t chunk = std::max<size_t>(1, std::min(w->chunk, (size_t) 0xfffffff0u));
  const bool reorder = !ras && want_reorder(w, nq);
  if (reorder)
    {
      chunk = std::min(chunk, std::max<size_t>(w->reorder_window, 32));
      const int s = ensure_reorder(w, r, std::min(chunk, nq), d_z != nullptr, d_leaf != nullptr);
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
      /* deferred list: a fraction of the launch size (overflow goes to a bitmask) */
      const size_t per_launch = std::min(chunk, nq);
      const size_t want = std::max<size_t>(std::min<size_t>(per_launch, 65536),
                                           (size_t) (w->list_fraction * (double) per_launch));
      const int s = ensure_list(r, want, per_launch);
      if (s) return s;
    }
  for (size_t off = 0; off < nq; off += chunk)
    {
      const size_t n = std::min(chunk, nq - off);
      gsi_launch_cfg cfg;
      cudaEvent_t ev[4];
      const int s = prepare_launch(w, r, n, ev, &cfg);
      if (s) return s;
      if (reorder)
        {
          cfg.reorder = &r->ro;
          r->n_reordered += n;
          r->n_launches += 2; /* k_bin + k_tile_locate + k_unpermute instead of one locate kernel */
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
  return 0;
}

int
read_counters(replica *r, unsigned out[GSI_CNT_WORDS])
{
  CU(cudaSetDevice(r->dev));
  CU(cudaMemcpy(out, r->counters, GSI_CNT_WORDS * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return 0;
}

int
ensure_stage(replica *r, size_t cap, size_t fields)
{
  if (r->st_cap >= cap && r->st_fields >= fields) return 0;
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = r->st[b];
      if (s.x) cudaFree(s.x);
      if (s.y) cudaFree(s.y);
      if (s.z) cudaFree(s.z);
      if (s.leaf) cudaFree(s.leaf);
      s.x = s.y = s.z = nullptr;
      s.leaf = nullptr;
    }
  r->st_cap = r->st_fields = 0;
  for (int b = 0; b < 2; b++)
    {
      stage_buf &s = r->st[b];
      CU(cudaMalloc(&s.x, cap * sizeof(double)));
      CU(cudaMalloc(&s.y, cap * sizeof(double)));
      CU(cudaMalloc(&s.z, cap * fields * sizeof(double)));
      CU(cudaMalloc(&s.leaf, cap * sizeof(int)));
    }
  r->st_cap = cap;
  r->st_fields = fields;
  return 0;
}

/* host-buffer pipeline on one replica: H2D (s_in) | kernels (s_k) | D2H (s_out), two chunk
 * buffers in flight.  With pinned user buffers (gsl_scattered_cuda_host_alloc) all three
 * stages overlap; pageable buffers work but the copies serialise on the host. */
int
eval_host_slice(const gsl_scattered_cuda *w, replica *r, const double *xq, const double *yq,
                const raster_spec *ras, size_t nq, double *zq, size_t ld, int *leaf,
                unsigned *outside_before_after)
{
  std::lock_guard<std::mutex> lock(r->mu);
  CU(cudaSetDevice(r->dev));
  const size_t F = w->n_fields;
  size_t C = std::max<size_t>(1, std::min(w->host_chunk, nq));
  if (ras) C = std::max<size_t>(1, w->host_chunk / ras->ny) * ras->ny; /* whole columns */
  if (C > nq) C = nq;
  int s = ensure_stage(r, C, F);
  if (s) return s;
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  if ((s = read_counters(r, c0))) return s;
  size_t k = 0;
  for (size_t off = 0; off < nq; off += C, k++)
    {
      const size_t n = std::min(C, nq - off);
      stage_buf &b = r->st[k & 1];
      if (!ras)
        {
          if (k >= 2) CU(cudaStreamWaitEvent(r->s_in, b.ev_k, 0));
          CU(cudaMemcpyAsync(b.x, xq + off, n * sizeof(double), cudaMemcpyHostToDevice, r->s_in));
          CU(cudaMemcpyAsync(b.y, yq + off, n * sizeof(double), cudaMemcpyHostToDevice, r->s_in));
          CU(cudaEventRecord(b.ev_in, r->s_in));
          CU(cudaStreamWaitEvent(r->s_k, b.ev_in, 0));
        }
      if (k >= 2) CU(cudaStreamWaitEvent(r->s_k, b.ev_out, 0));
      raster_spec rs;
      if (ras)
        {
          rs = *ras;
          rs.i_first += off / ras->ny;
          rs.n_cols = n / ras->ny;
        }
      s = eval_on_device(w, r, b.x, b.y, ras ? &rs : nullptr, n, zq ? b.z : nullptr, C,
                         leaf ? b.leaf : nullptr, r->s_k);
      if (s) return s;
      CU(cudaEventRecord(b.ev_k, r->s_k));
      CU(cudaStreamWaitEvent(r->s_out, b.ev_k, 0));
      if (leaf)
        CU(cudaMemcpyAsync(leaf + off, b.leaf, n * sizeof(int), cudaMemcpyDeviceToHost, r->s_out));
      if (zq)
        for (size_t f = 0; f < F; f++)
          CU(cudaMemcpyAsync(zq + f * ld + off, b.z + f * C, n * sizeof(double),
                             cudaMemcpyDeviceToHost, r->s_out));
      CU(cudaEventRecord(b.ev_out, r->s_out));
    }
  CU(cudaStreamSynchronize(r->s_out));
  CU(cudaStreamSynchronize(r->s_k));
  if ((s = read_counters(r, c1))) return s;
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
  const size_t R = w->reps.size();
  std::vector<int> status(R, 0);
  std::vector<unsigned> outside(R, 0);
  auto run = [&](size_t i) {
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
  };
  if (R == 1)
    run(0);
  else
    {
      std::vector<std::thread> th;
      for (size_t i = 0; i < R; i++) th.emplace_back(run, i);
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
  drop_replicas(w);
  std::memcpy(&w->hdr, image.data(), sizeof w->hdr);
  std::vector<int> devs = w->devs;
  if (devs.empty()) devs.push_back(0);
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
  CU(cudaMemcpyAsync(r0->slab, image.data(), image.size(), cudaMemcpyHostToDevice, r0->s_k));
  CU(gsi_launch_entry_cells(r0->view, (int32_t *) (r0->slab + w->hdr.off_entry), r0->sm_count, r0->s_k));
  CU(gsi_launch_seed_cells(r0->view, (int32_t *) (r0->slab + w->hdr.off_seeds), r0->sm_count,
                           r0->s_k));
  CU(cudaStreamSynchronize(r0->s_k));
  for (size_t i = 1; i < w->reps.size(); i++)
    {
      replica *r = w->reps[i];
      CU(cudaMemcpyPeerAsync(r->slab, r->dev, r0->slab, r0->dev, w->hdr.total_bytes, r0->s_k));
    }
  CU(cudaStreamSynchronize(r0->s_k));
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

/* take over a complete slab image (device or host memory; the seed grid is part of it) */
int
adopt_slab(gsl_scattered_cuda *w, const gsi_slab_header &h, const void *src, size_t bytes)
{
  if (h.magic != GSI_SLAB_MAGIC || h.version != GSI_SLAB_VERSION || h.total_bytes != bytes)
    GSI_ERROR("slab header mismatch", GSL_SCATTERED_EINVAL);
  if (h.n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  if (h.off_dag >= bytes || h.off_leaf >= bytes || h.off_z >= bytes || h.off_seeds >= bytes
      || h.n_dag == 0 || h.grid_g == 0
      || h.off_dag + h.n_dag * sizeof(gsi_rec) > bytes
      || h.off_leaf + h.n_leaves * (uint64_t) h.leaf_stride > bytes
      || h.off_seeds + (uint64_t) h.grid_g * h.grid_g * sizeof(int32_t) > bytes
      || h.off_entry >= bytes || h.off_entry + (uint64_t) h.grid_g * h.grid_g * sizeof(int32_t) > bytes
      || h.tile_bits < 1 || h.tile_bits > 12 || h.off_tiles >= bytes
      || h.off_tiles + (((uint64_t) 1 << h.tile_bits) + 1) * sizeof(int32_t) > bytes)
    GSI_ERROR("slab sections out of range", GSL_SCATTERED_EINVAL);
  drop_replicas(w);
  w->built = false;
  w->hdr = h;
  std::vector<int> devs = w->devs;
  if (devs.empty())
    {
      int cur = 0;
      CU(cudaGetDevice(&cur));
      devs.push_back(cur);
    }
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
  return "gsl_scattered_cuda 0.1 (sm_100a)";
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
  gsl_scattered_cuda *w = new (std::nothrow) gsl_scattered_cuda();
  if (!w) GSI_ERROR_NULL("failed to allocate workspace", GSL_SCATTERED_ENOMEM);
  w->n_cap = n_nodes;
  w->n_fields = n_fields;
  std::memset(&w->hdr, 0, sizeof w->hdr);
  return w;
}

void
gsl_scattered_cuda_free(gsl_scattered_cuda *w)
{
  if (!w) return;
  drop_replicas(w);
  delete w;
}

int
gsl_scattered_cuda_set_devices(gsl_scattered_cuda *w, const int dev[], size_t ndev)
{
  if (!w || (ndev && !dev)) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (w->uploaded) GSI_ERROR("set_devices must precede init/upload", GSL_SCATTERED_EINVAL);
  std::vector<int> d(dev, dev + ndev);
  for (size_t i = 0; i < d.size(); i++)
    for (size_t j = i + 1; j < d.size(); j++)
      if (d[i] == d[j]) GSI_ERROR("duplicate device id", GSL_SCATTERED_EINVAL);
  w->devs = d;
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_set_param(gsl_scattered_cuda *w, const char *name, double value)
{
  if (!w || !name) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  const std::string k(name);
  if (k == "cells_per_leaf" && value > 0) w->fprm.cells_per_leaf = value;
  else if (k == "grid_g" && value >= 0 && value <= 32768) w->fprm.grid_g = (uint32_t) value;
  else if (k == "max_steps" && value >= 1) w->max_steps = (int) value;
  else if (k == "delta" && value > 0) w->fprm.delta = value;
  else if (k == "leaf_stride" && (value == 96 || value == 128)) w->fprm.leaf_stride = (uint32_t) value;
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
  else if (k == "reorder_chunk" && value >= 32 && value <= GSI_BIN_CHUNK_MAX) w->reorder_chunk = (int) value;
  else if (k == "reorder_max_fields" && value >= 0) w->reorder_max_fields = (size_t) value;
  else if (k == "tile_ranges" && value >= 1 && value <= 64) w->tile_ranges = (int) value;
  else if (k == "tile_leaves" && value >= 1) w->fprm.tile_leaves = (uint32_t) value;
  else if (k == "tile_smem" && (value == 0 || value == 1)) w->tile_smem = (int) value;
  else if (k == "vmm" && (value == 0 || value == 1)) w->vmm = (int) value;
  else if (k == "exact_entry" && (value == 0 || value == 1))
    {
      w->exact_entry = (int) value;
      for (replica *r : w->reps)
        r->view.entry = value != 0 ? (const int32_t *) (r->slab + w->hdr.off_entry) : nullptr;
    }
  else GSI_ERROR("unknown parameter or value out of range", GSL_SCATTERED_EINVAL);
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_shuffle_mt19937(unsigned long seed, size_t n, size_t order[])
{
  if (n && !order) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  std::vector<int64_t> o(n);
  gsi_shuffle_mt19937(seed, n, o.data());
  for (size_t i = 0; i < n; i++) order[i] = (size_t) o[i];
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_build(gsl_scattered_cuda *w, const double xa[], const double ya[],
                         size_t n_nodes, int init_flags, const size_t order[])
{
  if (!w) GSI_ERROR("null workspace", GSL_SCATTERED_EINVAL);
  if (n_nodes && (!xa || !ya)) GSI_ERROR("null coordinate array", GSL_SCATTERED_EINVAL);
  /* simplex_tree_init returns GSL_FAILURE when the tree has no room, linear_simplex.c:274-278 */
  if (n_nodes > w->n_cap) GSI_ERROR("more nodes than the workspace was allocated for", GSL_SCATTERED_EBADLEN);
  drop_replicas(w);
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
  const int s = gsi_build_tree(w->xa.data(), w->ya.data(), n_nodes, init_flags,
                               order ? ord.data() : nullptr, &w->tree, &err);
  if (s) GSI_ERROR(err.empty() ? "triangulation build failed" : err.c_str(), s);
  w->built = true;
  return GSL_SCATTERED_SUCCESS;
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
  drop_replicas(w);
  gsi_host_tree &t = w->tree;
  t.n_points = n_nodes;
  t.pts.assign(pts3, pts3 + 3 * n_simplexes);
  t.lnk.assign(lnk3, lnk3 + 3 * n_simplexes);
  t.type.resize(n_simplexes);
  for (size_t i = 0; i < n_simplexes; i++)
    {
      if (type[i] < 0 || type[i] > 3) GSI_ERROR("bad node type", GSL_SCATTERED_EINVAL);
      t.type[i] = (uint8_t) type[i];
      for (int k = 0; k < 3; k++)
        {
          const int p = pts3[3 * i + k], l = lnk3[3 * i + k];
          if (p < -3 || (p >= 0 && (size_t) p >= n_nodes)) GSI_ERROR("vertex id out of range", GSL_SCATTERED_EINVAL);
          if (l < 0 || (size_t) l >= n_simplexes) GSI_ERROR("link out of range", GSL_SCATTERED_EINVAL);
        }
    }
  t.shuffle.resize(n_nodes);
  w->xa.resize(n_nodes);
  w->ya.resize(n_nodes);
  for (size_t i = 0; i < n_nodes; i++)
    {
      if (shuffle[i] >= n_nodes) GSI_ERROR("shuffle is not a permutation", GSL_SCATTERED_EINVAL);
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
}

int
gsl_scattered_cuda_eval_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                          size_t nq, double zq[], int leaf[])
{
  if (nq && (!xq || !yq)) GSI_ERROR("null query array", GSL_SCATTERED_EINVAL);
  return eval_host(w, xq, yq, nullptr, nq, zq, leaf);
}

double
gsl_scattered_cuda_eval(const gsl_scattered_cuda *w, double x, double y)
{
  std::vector<double> z(w ? w->n_fields : 1, std::numeric_limits<double>::quiet_NaN());
  const int s = eval_host(w, &x, &y, nullptr, 1, z.data(), nullptr);
  if (s != GSL_SCATTERED_SUCCESS && s != GSL_SCATTERED_EDOM)
    return std::numeric_limits<double>::quiet_NaN();
  return z[0];
}

int
gsl_scattered_cuda_eval_raster_e(const gsl_scattered_cuda *w, double x0, double dx, size_t nx,
                                 double y0, double dy, size_t ny, double zq[], int leaf[])
{
  if (nx && ny && nx > std::numeric_limits<size_t>::max() / ny)
    GSI_ERROR("raster too large", GSL_SCATTERED_EINVAL);
  raster_spec rs{x0, dx, y0, dy, ny, 0, nx};
  return eval_host(w, nullptr, nullptr, &rs, nx * ny, zq, leaf);
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
  std::lock_guard<std::mutex> lock(r->mu);
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  int s;
  if (!stream && (s = read_counters(r, c0))) return s;
  s = eval_on_device(w, r, d_xq, d_yq, nullptr, nq, d_zq, ld, d_leaf, st);
  if (s) return s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(st));
      if ((s = read_counters(r, c1))) return s;
      if (c1[GSI_CNT_OUTSIDE] != c0[GSI_CNT_OUTSIDE])
        GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
    }
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_eval_raster_device_e(const gsl_scattered_cuda *w, int dev, double x0, double dx,
                                        size_t nx, double y0, double dy, size_t ny, double *d_zq,
                                        int *d_leaf, void *stream)
{
  if (!w || !w->uploaded) GSI_ERROR("workspace not initialised on a GPU", GSL_SCATTERED_EFAILED);
  replica *r = find_replica(w, dev);
  if (!r) GSI_ERROR("device holds no replica", GSL_SCATTERED_EINVAL);
  std::lock_guard<std::mutex> lock(r->mu);
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
  raster_spec rs{x0, dx, y0, dy, ny, 0, nx};
  unsigned c0[GSI_CNT_WORDS], c1[GSI_CNT_WORDS];
  int s;
  if (!stream && (s = read_counters(r, c0))) return s;
  s = eval_on_device(w, r, nullptr, nullptr, &rs, nx * ny, d_zq, nx * ny, d_leaf, st);
  if (s) return s;
  if (!stream)
    {
      CU(cudaStreamSynchronize(st));
      if ((s = read_counters(r, c1))) return s;
      if (c1[GSI_CNT_OUTSIDE] != c0[GSI_CNT_OUTSIDE])
        GSI_ERROR("query point outside the domain (cage)", GSL_SCATTERED_EDOM);
    }
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_get_info(const gsl_scattered_cuda *w, gsl_scattered_cuda_info *info)
{
  if (!w || !info) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
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
  for (replica *r : w->reps)
    {
      unsigned c[GSI_CNT_WORDS];
      std::lock_guard<std::mutex> lock(r->mu);
      CU(cudaSetDevice(r->dev));
      CU(cudaStreamSynchronize(r->s_k));
      const int s = read_counters(r, c);
      if (s) return s;
      unsigned long long u;
      std::memcpy(&u, c + GSI_CNT_DEFERRED_TOTAL, 8);
      info->n_deferred += u;
      std::memcpy(&u, c + GSI_CNT_STEPS, 8);
      info->n_steps += u;
      info->n_outside += c[GSI_CNT_OUTSIDE];
      /* fold finished profiling events into the running totals */
      for (auto &p : r->prof_locate)
        {
          float ms = 0;
          if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) r->locate_seconds += 1e-3 * ms;
          cudaEventDestroy(p.first); cudaEventDestroy(p.second);
        }
      for (auto &p : r->prof_exact)
        {
          float ms = 0;
          if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) r->exact_seconds += 1e-3 * ms;
          cudaEventDestroy(p.first); cudaEventDestroy(p.second);
        }
      r->prof_locate.clear();
      r->prof_exact.clear();
      cudaGetLastError();
      info->locate_seconds += r->locate_seconds;
      info->exact_seconds += r->exact_seconds;
      info->n_kernel_launches += r->n_launches;
      info->n_reordered += r->n_reordered;
    }
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_reset_counters(const gsl_scattered_cuda *w)
{
  if (!w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  for (replica *r : w->reps)
    {
      std::lock_guard<std::mutex> lock(r->mu);
      CU(cudaSetDevice(r->dev));
      CU(cudaMemsetAsync(r->counters, 0, GSI_CNT_WORDS * sizeof(unsigned), r->s_k));
      CU(cudaStreamSynchronize(r->s_k));
      for (auto &p : r->prof_locate) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
      for (auto &p : r->prof_exact) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
      r->prof_locate.clear();
      r->prof_exact.clear();
      r->locate_seconds = r->exact_seconds = 0;
      r->n_launches = 0;
      r->n_reordered = 0;
    }
  return GSL_SCATTERED_SUCCESS;
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
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
  CU(cudaMemcpyAsync(d_dst, r->slab, w->hdr.total_bytes, cudaMemcpyDeviceToDevice, st));
  if (!stream) CU(cudaStreamSynchronize(st));
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_init_from_slab_device(gsl_scattered_cuda *w, const void *d_src, size_t bytes)
{
  if (!w || !d_src || bytes < sizeof(gsi_slab_header)) GSI_ERROR("bad slab", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0)
    GSI_ERROR("no CUDA device available", GSL_SCATTERED_EFAILED);
  gsi_slab_header h;
  CU(cudaMemcpy(&h, d_src, sizeof h, cudaMemcpyDefault));
  return adopt_slab(w, h, d_src, bytes);
}

int
gsl_scattered_cuda_fwrite(FILE *stream, const gsl_scattered_cuda *w)
{
  if (!stream || !w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  if (!w->uploaded || w->reps.empty()) GSI_ERROR("nothing to write: workspace not initialised", GSL_SCATTERED_EINVAL);
  replica *r = w->reps[0];
  std::vector<char> image;
  try
    {
      image.resize((size_t) w->hdr.total_bytes);
    }
  catch (const std::bad_alloc &)
    {
      GSI_ERROR("out of host memory for the slab image", GSL_SCATTERED_ENOMEM);
    }
  {
    std::lock_guard<std::mutex> lock(r->mu);
    CU(cudaSetDevice(r->dev));
    CU(cudaStreamSynchronize(r->s_k));
    CU(cudaMemcpy(image.data(), r->slab, image.size(), cudaMemcpyDeviceToHost));
  }
  gsi_slab_header h;
  std::memcpy(&h, image.data(), sizeof h);
  h.reserved[0] = slab_checksum(image.data(), image.size());
  std::memcpy(image.data(), &h, sizeof h);
  if (fwrite(image.data(), 1, image.size(), stream) != image.size())
    GSI_ERROR("fwrite failed", GSL_SCATTERED_EFAILED);
  return GSL_SCATTERED_SUCCESS;
}

int
gsl_scattered_cuda_fread(FILE *stream, gsl_scattered_cuda *w)
{
  if (!stream || !w) GSI_ERROR("null argument", GSL_SCATTERED_EINVAL);
  gsi_slab_header h;
  if (fread(&h, 1, sizeof h, stream) != sizeof h) GSI_ERROR("fread failed", GSL_SCATTERED_EFAILED);
  if (h.magic != GSI_SLAB_MAGIC || h.version != GSI_SLAB_VERSION || h.total_bytes < sizeof h
      || h.total_bytes > ((uint64_t) 1 << 40))
    GSI_ERROR("not a slab image of this library version", GSL_SCATTERED_EINVAL);
  if (h.n_fields != w->n_fields) GSI_ERROR("n_fields differs from alloc", GSL_SCATTERED_EBADLEN);
  std::vector<char> image;
  try
    {
      image.resize((size_t) h.total_bytes);
    }
  catch (const std::bad_alloc &)
    {
      GSI_ERROR("out of host memory for the slab image", GSL_SCATTERED_ENOMEM);
    }
  std::memcpy(image.data(), &h, sizeof h);
  const size_t rest = image.size() - sizeof h;
  if (fread(image.data() + sizeof h, 1, rest, stream) != rest) GSI_ERROR("fread failed", GSL_SCATTERED_EFAILED);
  if (h.reserved[0] != slab_checksum(image.data(), image.size()))
    GSI_ERROR("slab image checksum mismatch", GSL_SCATTERED_EINVAL);
  if (gsl_scattered_cuda_device_count() <= 0)
    GSI_ERROR("no CUDA device available: libgsl_scattered_cuda has no CPU evaluation path",
              GSL_SCATTERED_EFAILED);
  return adopt_slab(w, h, image.data(), image.size());
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
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
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
  std::lock_guard<std::mutex> lock(r->mu);
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
  CU(gsi_launch_checksum(d_leaf, d_val, n, r->scratch2, r->sm_count, st));
  CU(cudaMemcpyAsync(out, r->scratch2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
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
  std::lock_guard<std::mutex> lock(r->mu);
  CU(cudaSetDevice(dev));
  cudaStream_t st = stream ? (cudaStream_t) stream : r->s_k;
  CU(gsi_launch_compare(d_leaf_a, d_leaf_b, d_val_a, d_val_b, n, n_val, r->scratch2, r->sm_count, st));
  CU(cudaMemcpyAsync(out, r->scratch2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GSL_SCATTERED_SUCCESS;
}

} /* extern "C" */
