// lgmi.cu -- host side of liblgmi.so: the C ABI declared in include/lgmi.h.
//
// Owns the CUDA context state (stream, ln table, pinned staging), builds the
// per-batch work tables and launches the kernels in lgmi_kernels.cuh.  There is
// no CPU implementation of the MI step in this library: every entry point that
// computes does so on the device, and lgmi_create() fails without one.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <queue>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/lgmi.h"
#include "lgmi_kernels.cuh"

extern "C" void lgmi_build_lntab(double* hi_lo_pairs, uint64_t k_begin, uint64_t k_end);

using namespace lgmi;

static thread_local std::string g_create_error;

// Size-bucketed cache of device and pinned-host blocks owned by a context.  Batches and pipelines are created and
// destroyed per submit by the drop-in seam (one chunk of footprints = one submit); cudaMalloc / cudaHostAlloc /
// cudaFree cost tens of microseconds to milliseconds each and cudaFree synchronises the device, so their buffers
// are recycled instead.  A block is only returned to the cache after the owning batch's stream has been
// synchronised (lgmi_batch_destroy), i.e. nothing in flight can still touch it.
struct MemPool {
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> live;
  size_t cached_bytes = 0, cap_bytes = 0;
  uint64_t hits = 0, misses = 0;
  bool host = false;
  static size_t bucket(size_t n) {  // sixteenth-of-an-octave steps: at most 1/8 over-allocation
    if (n <= 4096) return 4096;
    size_t p = 1;
    while (p < n) p <<= 1;
    const size_t step = p >> 4;
    return (n + step - 1) / step * step;
  }
  cudaError_t raw_alloc(void** out, size_t sz) {
    return host ? cudaHostAlloc(out, sz, cudaHostAllocDefault) : cudaMalloc(out, sz);
  }
  void raw_free(void* p) {
    if (host) cudaFreeHost(p); else cudaFree(p);
  }
  void trim() {
    for (auto& kv : free_blocks) raw_free(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
  cudaError_t alloc(void** out, size_t bytes) {
    const size_t sz = bucket(bytes ? bytes : 1);
    auto it = free_blocks.lower_bound(sz);
    if (it != free_blocks.end() && it->first <= sz + sz / 4) {
      *out = it->second;
      live[*out] = it->first;
      cached_bytes -= it->first;
      free_blocks.erase(it);
      ++hits;
      return cudaSuccess;
    }
    ++misses;
    cudaError_t e = raw_alloc(out, sz);
    if (e == cudaErrorMemoryAllocation) {  // give the cache back and try once more
      cudaGetLastError();
      trim();
      e = raw_alloc(out, sz);
    }
    if (e == cudaSuccess) live[*out] = sz;
    return e;
  }
  void release(void* p) {
    if (!p) return;
    auto it = live.find(p);
    if (it == live.end()) {  // not ours (never happens): hand it to the runtime
      raw_free(p);
      return;
    }
    const size_t sz = it->second;
    live.erase(it);
    if (cached_bytes + sz > cap_bytes) {
      raw_free(p);
      return;
    }
    free_blocks.emplace(sz, p);
    cached_bytes += sz;
  }
  void destroy() {
    trim();
    for (auto& kv : live) raw_free(kv.first);
    live.clear();
  }
};

struct lgmi_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // ln table
  lg_dd* d_lntab = nullptr;
  uint64_t ln_cap = 0;  // entries 0..ln_cap-1 valid
  // one-shot submit state
  lgmi_batch* oneshot = nullptr;
  // scratch for ecdf / csr
  void* d_scratch = nullptr;
  size_t scratch_cap = 0;
  uint16_t* d_ij_tab = nullptr;  // triangular pair tables for S <= 64 (lg_ij_tab_off)
  // persistent launch shape of k_pairs
  int num_sms = 0;
  int pairs_ctas_per_sm = 0;
  int pre_ctas_per_sm = 0;
  int dense_path = 4;  // deep units: 4 = four Gram blocks + sparse "other" cells where the "other" reads are rare (decided on
                       // the device per unit and run), 9 = always nine blocks
  int small_path = 0;  // 0: popcount (k_pairs_fast, default: faster); 1: counted on the tensor cores
  int tile_path = 2;   // mid-depth units: 0 popcount (k_tile_mi), 1 tensor cores (k_tile_gram),
                       // 2 tensor cores, warp-specialised (k_tile_gram_ws, default)
  // tensor-core path: units at least this large build their tables with k_gram_i8
  uint32_t dense_min_sites = 48, dense_min_reads = 8192;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved through the runtime (no libcuda link)
  MemPool dev_pool, host_pool;   // recycled buffers of batches / pipelines
  uint64_t lntab_gen = 0;        // bumped when the ln table moves: captured graphs hold its address
  int pipeline_graphs = 1;       // groups of a pipelined step replay their launch chain as a CUDA graph (LGMI_GRAPHS=0: off)
};

template <class T>
static cudaError_t pool_malloc(lgmi_ctx* ctx, T** out, size_t bytes) {
  return ctx->dev_pool.alloc(reinterpret_cast<void**>(out), bytes);
}
template <class T>
static cudaError_t pool_host_alloc(lgmi_ctx* ctx, T** out, size_t bytes) {
  return ctx->host_pool.alloc(reinterpret_cast<void**>(out), bytes);
}

// one unit of the tensor-core path
struct DenseTileList {
  uint64_t off = 0, macs = 0;
  uint32_t n = 0;
  uint32_t n_whole = 0;  // the first n_whole work items are whole tiles; the rest share tiles (last wave, split in K)
};
struct DensePlan {
  uint32_t unit = 0, S = 0, W = 0, S_pad = 0, k_blocks = 0;
  uint32_t oth_cap = 0;       // listed "other" reads per site (four-block form)
  bool four = false;          // the four-block form may be chosen (else the run is told "nine" up front)
  DenseTileList list4, list9; // P, M x P, M  /  C, P, M x C, P, M
  uint64_t plane_off = 0, gram_off = 0;
  CUtensorMap tmap;
};

struct lgmi_batch {
  lgmi_ctx* ctx = nullptr;
  cudaStream_t own_stream = nullptr;  // pipeline chunks run on their own stream; otherwise the context's
  uint32_t unit_base = 0;             // added to the unit field of every record (pipeline chunks)
  cudaEvent_t ev[8] = {};             // 0-1 whole run, 2-3 k_pairs_fast, 4-5 tensor-core path, 6-7 last k_gram_i8
  uint32_t n_units = 0, n_items = 0, n_mean_items = 0, n_fast = 0;
  uint64_t plane_words = 0, n_sites = 0, n_candidates = 0, n_dense = 0;
  uint32_t max_reads = 0;
  std::vector<lgmi_unit_desc> h_units;
  std::vector<Item> h_items;
  // device
  DevUnit* d_units = nullptr;
  Item* d_items = nullptr;
  FastItem* d_fast_items = nullptr;
  uint8_t* d_fast_empty = nullptr;           // per fast item: the pre-pass counted no emitted pair
  uint8_t* d_item_dense = nullptr;
  uint32_t* d_n_generic = nullptr;
  MeanItem* d_mean_items = nullptr;     // 32-site blocks of the multi-item units
  uint32_t* d_planes = nullptr;
  uint8_t* d_flags = nullptr;
  unsigned long long* d_item_cnt = nullptr;  // n_items + 1 (last stays 0)
  unsigned long long* d_item_off = nullptr;  // n_items + 1
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  Header* d_header = nullptr;
  lgmi_pair_rec* d_records = nullptr;
  uint64_t rec_cap = 0;
  uint32_t* d_counts = nullptr;
  uint64_t counts_cap = 0;
  double* d_site_mean = nullptr;
  uint32_t* d_site_cnt = nullptr;
  double* d_dense = nullptr;
  unsigned long long* d_unit_rec_off = nullptr;
  // tensor-core path
  std::vector<DensePlan> dense_plans;
  uint8_t* d_x = nullptr;       // indicator matrix of the unit being contracted (reused unit after unit)
  uint32_t* d_gram = nullptr;   // nine count matrices per dense unit
  uint32_t* d_xt = nullptr;     // four-block form: transposed label planes of the unit being contracted
  uint32_t* d_oth_cnt = nullptr, *d_oth_list = nullptr;  // ... and its listed "other" reads per site
  uint32_t* d_unit_mode = nullptr;  // per unit: the form of this run (0 four blocks, else nine)
  // graph replays / pipeline groups: the small units' kernels (k_count_fast before the scan, k_pairs_fast after it) run on
  // a second stream beside the kernels of the larger units (forked and joined by events inside the captured chain)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_ev[4] = {};
  cudaStream_t fix_stream = nullptr;             // k_other_fix runs beside k_dense_x (forked after k_dense_prep,
  cudaEvent_t fix_fork = nullptr, fix_join = nullptr;  // joined before the tables are read)
  DenseTile* d_tiles = nullptr;
  uint64_t dense_macs = 0;      // multiply-accumulates the tile lists amount to
  // small units counted on the tensor cores
  FastItem* d_pre_items = nullptr;
  uint32_t n_pre = 0;
  unsigned long long* d_val = nullptr;  // packed counts, 8 bytes per candidate pair of those units
  // tiled popcount path
  TileItem* d_tile_items = nullptr;
  uint32_t n_tile_items = 0;
  TgTile* d_gram_tiles = nullptr;      // k_tile_gram's work items (128 x 48 site blocks)
  uint32_t n_gram_tiles = 0;
  int gram_kernel = 1;                 // 1 k_tile_gram, 2 k_tile_gram_ws (the context's tile path when the batch was made)
  uint2* d_tile_cnt = nullptr;         // its output: 24 bytes of counts per pair slot
  uint32_t* d_tile_next = nullptr;     // its work counter
  uint32_t n_tiled_work_items = 0;     // work items (Item) of the tiled and of the deep units: k_pairs_generic<1>'s and <2>'s
  TiledDesc* d_tiled_desc = nullptr;   // k_pairs_generic<1>'s items, one self-contained record each
  uint32_t n_tiled_desc = 0;
  uint32_t* d_tile_counts = nullptr;  // EMIT_COUNTS only, allocated on first use
  uint64_t n_tiled_slots = 0;
  // host (pinned) mirrors
  Header* h_header = nullptr;
  lgmi_pair_rec* h_records = nullptr;
  uint64_t h_rec_cap = 0;
  uint32_t* h_counts = nullptr;
  uint64_t h_counts_cap = 0;
  double* h_site_mean = nullptr;
  uint32_t* h_site_cnt = nullptr;
  unsigned long long* h_unit_rec_off = nullptr;
  // launch chains captured as CUDA graphs, one per (min_common, mode)
  struct Graph {
    int min_common;
    uint32_t mode;
    uint64_t lntab_gen;
    uint64_t launches;  // kernels in the chain
    cudaGraphExec_t exec;
  };
  std::vector<Graph> graphs;
  // run state
  bool uploaded = false, ran = false, last_timed = false;
  uint32_t last_mode = 0;
  // candidates with a het_snp partner (what SKIP_NONHET evaluates); known once host flags have been seen
  uint64_t n_het_candidates = 0;
  bool het_known = false;
};

// candidate pairs next to a het_snp site, from the host copy of the flag bytes
static uint64_t het_candidates(const std::vector<lgmi_unit_desc>& units, const uint8_t* site_flags) {
  uint64_t tot = 0;
  for (const lgmi_unit_desc& u : units) {
    uint64_t h = 0;
    for (uint32_t s = 0; s < u.n_sites; ++s)
      h += (site_flags[u.site_off + s] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP ? 1u : 0u;
    const uint64_t S = u.n_sites, o = S - h;
    tot += (S >= 2 ? S * (S - 1) / 2 : 0) - (o >= 2 ? o * (o - 1) / 2 : 0);
  }
  return tot;
}

// --------------------------------------------------------------------------- helpers
static inline cudaStream_t bstream(const lgmi_batch* b) { return b->own_stream ? b->own_stream : b->ctx->stream; }

static int fail(lgmi_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

#define CU(ctx, call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail((ctx), (e__ == cudaErrorMemoryAllocation) ? LGMI_ERR_NOMEM : LGMI_ERR_CUDA, \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static int ensure_lntab(lgmi_ctx* ctx, uint64_t k_max) {
  if (k_max + 1 <= ctx->ln_cap) return LGMI_OK;
  uint64_t cap = std::max<uint64_t>(4096, ctx->ln_cap);
  while (cap < k_max + 1) cap *= 2;
  std::vector<double> host(2 * cap);
  lgmi_build_lntab(host.data(), 0, cap);
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_lntab) CU(ctx, cudaFree(ctx->d_lntab));
  ctx->d_lntab = nullptr;
  ctx->ln_cap = 0;
  CU(ctx, cudaMalloc(&ctx->d_lntab, cap * sizeof(lg_dd)));
  CU(ctx, cudaMemcpy(ctx->d_lntab, host.data(), cap * sizeof(lg_dd), cudaMemcpyHostToDevice));
  ctx->ln_cap = cap;
  ++ctx->lntab_gen;
  return LGMI_OK;
}

static int ensure_scratch(lgmi_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_cap) return LGMI_OK;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_scratch) CU(ctx, cudaFree(ctx->d_scratch));
  ctx->d_scratch = nullptr;
  ctx->scratch_cap = 0;
  size_t cap = std::max<size_t>(bytes, 1 << 20);
  CU(ctx, cudaMalloc(&ctx->d_scratch, cap));
  ctx->scratch_cap = cap;
  return LGMI_OK;
}

// --------------------------------------------------------------------------- context
extern "C" int lgmi_version(void) { return LGMI_VERSION; }

extern "C" int lgmi_create(int device, lgmi_t** out) {
  if (!out) return fail(nullptr, LGMI_ERR_ARG, "lgmi_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, LGMI_ERR_NODEVICE,
                "lgmi_create: no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev)
    return fail(nullptr, LGMI_ERR_ARG, "lgmi_create: device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, LGMI_ERR_NODEVICE,
                "lgmi_create: device %d is sm_%d%d; this build contains sm_100a code only", device,
                prop.major, prop.minor);
  lgmi_ctx* ctx = new (std::nothrow) lgmi_ctx();
  if (!ctx) return fail(nullptr, LGMI_ERR_NOMEM, "lgmi_create: out of host memory");
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return rc;
  }
  ctx->stream = ctx->own_stream;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->host_pool.host = true;
  ctx->dev_pool.cap_bytes = 16ull << 30;   // cached (idle) bytes kept at most; LGMI_POOL_MB overrides both
  ctx->host_pool.cap_bytes = 4ull << 30;
  if (const char* e = getenv("LGMI_POOL_MB")) ctx->dev_pool.cap_bytes = ctx->host_pool.cap_bytes = (size_t)atoll(e) << 20;
  if (cudaFuncSetAttribute(k_pairs_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FastSmem)) !=
          cudaSuccess ||
      cudaFuncSetAttribute(k_pairs_fast_het, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FastSmem)) !=
          cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->pairs_ctas_per_sm, k_pairs_fast, kFastThreads,
                                                    sizeof(FastSmem)) != cudaSuccess ||
      ctx->pairs_ctas_per_sm < 1) {
    int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: k_pairs_fast cannot be resident (%s)",
                  cudaGetErrorString(cudaGetLastError()));
    lgmi_destroy(ctx);
    return rc;
  }
  if (cudaFuncSetAttribute(k_small_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(SgSmem) + 1024)) !=
          cudaSuccess ||
      cudaFuncSetAttribute(k_pairs_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PreSmem)) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->pre_ctas_per_sm, k_pairs_pre, kFastThreads, sizeof(PreSmem)) !=
          cudaSuccess ||
      ctx->pre_ctas_per_sm < 1) {
    int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: small-unit tensor path cannot be resident (%s)",
                  cudaGetErrorString(cudaGetLastError()));
    lgmi_destroy(ctx);
    return rc;
  }
  if (const char* e = getenv("LGMI_SMALL_PATH")) ctx->small_path = atoi(e) ? 1 : 0;
  if (const char* e = getenv("LGMI_DENSE_PATH")) ctx->dense_path = atoi(e) == 9 ? 9 : 4;
  if (const char* e = getenv("LGMI_TILE_PATH")) ctx->tile_path = std::min(2, std::max(0, atoi(e)));
  if (const char* e = getenv("LGMI_GRAPHS")) ctx->pipeline_graphs = atoi(e) ? 1 : 0;
  if (cudaFuncSetAttribute(k_tile_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTgSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_tile_gram_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTwSmemBytes) != cudaSuccess) {
    int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: k_tile_gram cannot be resident (%s)",
                  cudaGetErrorString(cudaGetLastError()));
    lgmi_destroy(ctx);
    return rc;
  }
  {
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !ctx->encode_tiled ||
        cudaFuncSetAttribute(k_gram_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseSmemBytes) !=
            cudaSuccess) {
      int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: tensor-core path unavailable (%s)",
                    cudaGetErrorString(cudaGetLastError()));
      lgmi_destroy(ctx);
      return rc;
    }
  }
  {
    std::vector<uint16_t> tab(kIjTabEntries);
    for (uint32_t S = 2; S <= (uint32_t)kFastMaxS; ++S) {
      uint32_t p = lg_ij_tab_off(S);
      for (uint32_t i = 0; i + 1 < S; ++i)
        for (uint32_t j = i + 1; j < S; ++j) tab[p++] = (uint16_t)(i * 64u + j);
    }
    if (cudaMalloc(&ctx->d_ij_tab, tab.size() * sizeof(uint16_t)) != cudaSuccess ||
        cudaMemcpy(ctx->d_ij_tab, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) {
      int rc = fail(nullptr, LGMI_ERR_CUDA, "lgmi_create: pair table upload failed (%s)",
                    cudaGetErrorString(cudaGetLastError()));
      lgmi_destroy(ctx);
      return rc;
    }
  }
  *out = ctx;
  return LGMI_OK;
}

extern "C" void lgmi_destroy(lgmi_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->oneshot) lgmi_batch_destroy(ctx->oneshot);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->d_lntab) cudaFree(ctx->d_lntab);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_ij_tab) cudaFree(ctx->d_ij_tab);
  ctx->dev_pool.destroy();
  ctx->host_pool.destroy();
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char* lgmi_last_error(const lgmi_t* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int lgmi_set_stream(lgmi_t* ctx, void* cuda_stream) {
  if (!ctx) return LGMI_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return LGMI_OK;
}

extern "C" int lgmi_pinned_alloc(lgmi_t* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return LGMI_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return LGMI_OK;
}

extern "C" int lgmi_pinned_free(lgmi_t* ctx, void* ptr) {
  if (!ctx) return LGMI_ERR_ARG;
  if (ptr) CU(ctx, cudaFreeHost(ptr));
  return LGMI_OK;
}

extern "C" uint64_t lgmi_launch_count(const lgmi_t* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lgmi_set_small_path(lgmi_t* ctx, int tensor_cores) {
  if (!ctx) return LGMI_ERR_ARG;
  ctx->small_path = tensor_cores ? 1 : 0;
  return LGMI_OK;
}

extern "C" int lgmi_set_dense_path(lgmi_t* ctx, int blocks) {
  if (!ctx) return LGMI_ERR_ARG;
  if (blocks != 4 && blocks != 9) return fail(ctx, LGMI_ERR_ARG, "lgmi_set_dense_path: blocks must be 4 or 9");
  ctx->dense_path = blocks;
  return LGMI_OK;
}

extern "C" int lgmi_set_tile_path(lgmi_t* ctx, int tensor_cores) {
  if (!ctx) return LGMI_ERR_ARG;
  ctx->tile_path = tensor_cores < 0 ? 0 : (tensor_cores > 2 ? 2 : tensor_cores);
  return LGMI_OK;
}

extern "C" int lgmi_set_dense_threshold(lgmi_t* ctx, uint32_t min_sites, uint32_t min_reads) {
  if (!ctx) return LGMI_ERR_ARG;
  if (min_sites < 2) return fail(ctx, LGMI_ERR_ARG, "lgmi_set_dense_threshold: min_sites must be >= 2");
  ctx->dense_min_sites = min_sites;
  ctx->dense_min_reads = min_reads;
  return LGMI_OK;
}

// --------------------------------------------------------------------------- batch
extern "C" void lgmi_batch_destroy(lgmi_batch_t* b) {
  if (!b) return;
  lgmi_ctx* ctx = b->ctx;  // (always set: a batch is born with its context)
  if (b->ctx) {
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(bstream(b));
    if (b->ctx->oneshot == b) b->ctx->oneshot = nullptr;
  }
  for (cudaEvent_t e : b->ev)
    if (e) cudaEventDestroy(e);
  for (lgmi_batch::Graph& g : b->graphs) cudaGraphExecDestroy(g.exec);
  ctx->dev_pool.release(b->d_units);
  ctx->dev_pool.release(b->d_items);
  ctx->dev_pool.release(b->d_fast_items);
  ctx->dev_pool.release(b->d_fast_empty);
  ctx->dev_pool.release(b->d_item_dense);
  ctx->dev_pool.release(b->d_n_generic);
  ctx->dev_pool.release(b->d_mean_items);
  ctx->dev_pool.release(b->d_planes);
  ctx->dev_pool.release(b->d_flags);
  ctx->dev_pool.release(b->d_item_cnt);
  ctx->dev_pool.release(b->d_item_off);
  ctx->dev_pool.release(b->d_scan_tmp);
  ctx->dev_pool.release(b->d_header);
  ctx->dev_pool.release(b->d_records);
  ctx->dev_pool.release(b->d_counts);
  ctx->dev_pool.release(b->d_site_mean);
  ctx->dev_pool.release(b->d_site_cnt);
  ctx->dev_pool.release(b->d_dense);
  ctx->dev_pool.release(b->d_unit_rec_off);
  ctx->dev_pool.release(b->d_x);
  ctx->dev_pool.release(b->d_gram);
  ctx->dev_pool.release(b->d_xt);
  ctx->dev_pool.release(b->d_oth_cnt);
  ctx->dev_pool.release(b->d_oth_list);
  ctx->dev_pool.release(b->d_unit_mode);
  if (b->side_stream) cudaStreamDestroy(b->side_stream);
  for (cudaEvent_t e : b->side_ev)
    if (e) cudaEventDestroy(e);
  if (b->fix_stream) cudaStreamDestroy(b->fix_stream);
  if (b->fix_fork) cudaEventDestroy(b->fix_fork);
  if (b->fix_join) cudaEventDestroy(b->fix_join);
  ctx->dev_pool.release(b->d_tiles);
  ctx->dev_pool.release(b->d_tile_items);
  ctx->dev_pool.release(b->d_gram_tiles);
  ctx->dev_pool.release(b->d_tiled_desc);
  ctx->dev_pool.release(b->d_tile_cnt);
  ctx->dev_pool.release(b->d_tile_next);
  ctx->dev_pool.release(b->d_tile_counts);
  ctx->dev_pool.release(b->d_pre_items);
  ctx->dev_pool.release(b->d_val);
  ctx->host_pool.release(b->h_header);
  ctx->host_pool.release(b->h_records);
  ctx->host_pool.release(b->h_counts);
  ctx->host_pool.release(b->h_site_mean);
  ctx->host_pool.release(b->h_site_cnt);
  ctx->host_pool.release(b->h_unit_rec_off);
  delete b;
}

extern "C" int lgmi_batch_create(lgmi_t* ctx, const lgmi_unit_desc* units, uint32_t n_units,
                                 uint64_t plane_words, uint64_t n_sites, lgmi_batch_t** out) {
  if (!ctx || !out || (!units && n_units)) return fail(ctx, LGMI_ERR_ARG, "lgmi_batch_create: NULL argument");
  *out = nullptr;
  CU(ctx, cudaSetDevice(ctx->device));
  lgmi_batch* b = new (std::nothrow) lgmi_batch();
  if (!b) return fail(ctx, LGMI_ERR_NOMEM, "lgmi_batch_create: out of host memory");
  b->ctx = ctx;
  b->n_units = n_units;
  b->plane_words = plane_words;
  b->n_sites = n_sites;
  b->h_units.assign(units, units + n_units);

  std::vector<DevUnit> du(n_units);
  std::vector<FastItem> fast_items, pre_items;
  uint64_t val_slots = 0;
  std::vector<MeanItem> mean_items;
  std::vector<DenseTile> dense_tiles;
  std::vector<TileItem> tile_items;
  std::vector<TiledDesc> tiled_desc;  // k_pairs_generic<1>'s items
  std::vector<TgTile> gram_tiles;
  std::vector<uint32_t> tile_words, gram_words_of;  // W of each tile's unit (launch order: longest first)
  bool any_gram_tiled = false;
  uint64_t dense = 0, gram_words = 0, x_bytes = 0, xt_words = 0, oth_words = 0;
  uint32_t oth_sites = 0;
  for (uint32_t k = 0; k < n_units; ++k) {
    const lgmi_unit_desc& u = units[k];
    if (u.n_sites > 65535u) {
      delete b;
      return fail(ctx, LGMI_ERR_UNSUPPORTED, "unit %u: %u sites > 65535", k, u.n_sites);
    }
    if (u.row_words != 4u * ((u.n_reads + 127u) / 128u) || (u.plane_off & 3ull) ||
        u.plane_off + 3ull * u.n_sites * u.row_words > plane_words ||
        (uint64_t)u.site_off + u.n_sites > n_sites) {
      delete b;
      return fail(ctx, LGMI_ERR_ARG,
                  "unit %u: inconsistent descriptor (S=%u R=%u W=%u plane_off=%llu site_off=%u)", k,
                  u.n_sites, u.n_reads, u.row_words, (unsigned long long)u.plane_off, u.site_off);
    }
    const size_t first_gram_tile = gram_tiles.size();  // this unit's k_tile_gram work items start here
    DevUnit& d = du[k];
    d.plane_off = u.plane_off;
    d.S = u.n_sites;
    d.R = u.n_reads;
    d.W = u.row_words;
    d.site_off = u.site_off;
    d.first_item = (uint32_t)b->h_items.size();
    const uint64_t np = u.n_sites >= 2 ? (uint64_t)u.n_sites * (u.n_sites - 1) / 2 : 0;
    b->n_candidates += np;
    b->max_reads = std::max(b->max_reads, u.n_reads);
    const uint32_t nit = np ? (uint32_t)((np + kPairsMax - 1) / kPairsMax) : 1u;
    d.n_items = nit;
    d.dense_off = ~0ull;
    d.gram_off = kNoGram;
    d.S_pad = 0;
    d.tiled = 0;
    if (np && u.n_sites >= ctx->dense_min_sites && u.n_reads >= ctx->dense_min_reads) {
      // tensor-core path: the unit's nine count matrices come from k_gram_i8
      DensePlan pl;
      pl.unit = k;
      pl.S = u.n_sites;
      pl.W = u.row_words;
      pl.S_pad = (u.n_sites + (uint32_t)kDenseBN - 1u) / (uint32_t)kDenseBN * (uint32_t)kDenseBN;
      pl.k_blocks = u.row_words * 32u / (uint32_t)kDenseBK;  // W is a multiple of 4: whole 128-read blocks
      pl.plane_off = u.plane_off;
      pl.gram_off = gram_words;
      // (k_other_fix addresses a label's transposed plane with 32-bit word offsets)
      pl.four = ctx->dense_path == 4 && (uint64_t)pl.k_blocks * kDenseBK * (pl.S_pad / 32u) < (1ull << 32);
      pl.oth_cap = std::max<uint32_t>(kDenseOthMin, u.n_reads / kDenseOthDiv);
      auto build_list = [&](DenseTileList& L, uint32_t g0) {  // row groups g0 .. 2 of X against each other
        L.off = dense_tiles.size();
        for (uint32_t a = g0; a < 3; ++a)
          for (uint32_t bb = g0; bb < 3; ++bb)
            for (uint32_t J = 0; J < pl.S_pad / (uint32_t)kDenseBN; ++J)
              for (uint32_t I = 0; I < pl.S_pad / (uint32_t)kDenseBM; ++I) {
                // some pair i < j inside the tile, both real sites
                const uint32_t i_min = I * (uint32_t)kDenseBM;
                const uint32_t j_max = std::min(J * (uint32_t)kDenseBN + (uint32_t)kDenseBN, pl.S) - 1u;
                if (J * (uint32_t)kDenseBN >= pl.S || i_min >= j_max) continue;
                dense_tiles.push_back(DenseTile{(uint8_t)a, (uint8_t)bb, (uint16_t)I, (uint16_t)J, 0, 0u, pl.k_blocks});
              }
        // one CTA per SM walks the list with stride #SMs: a last, partly filled wave leaves SMs idle for a whole
        // tile time, so its tiles are cut along K into as many parts as fill it (added atomically into zeroed tiles)
        const uint32_t n = (uint32_t)(dense_tiles.size() - L.off), grid = (uint32_t)ctx->num_sms;
        const uint32_t rem = n % grid;
        uint32_t parts = rem ? std::min<uint32_t>(8u, grid / rem) : 1u;
        while (parts > 1u && pl.k_blocks < 8u * parts) --parts;  // at least 8 k-blocks (1 024 reads) per part
        L.n_whole = n;
        if (parts >= 2u) {
          L.n_whole = n - rem;
          std::vector<DenseTile> tail(dense_tiles.end() - rem, dense_tiles.end());
          dense_tiles.resize(dense_tiles.size() - rem);
          for (uint32_t q = 0; q < parts; ++q)
            for (DenseTile t : tail) {
              t.kb0 = (uint32_t)((uint64_t)pl.k_blocks * q / parts);
              t.kb1 = (uint32_t)((uint64_t)pl.k_blocks * (q + 1) / parts);
              t.partial = 1;
              dense_tiles.push_back(t);
            }
        }
        L.n = (uint32_t)(dense_tiles.size() - L.off);
        for (uint32_t t = 0; t < L.n; ++t) {
          const DenseTile& dt = dense_tiles[L.off + t];
          L.macs += (uint64_t)kDenseBM * kDenseBN * (dt.kb1 - dt.kb0) * kDenseBK;
        }
      };
      if (pl.four) build_list(pl.list4, 1u);
      build_list(pl.list9, 0u);
      if (pl.four) {
        xt_words = std::max<uint64_t>(xt_words, 3ull * pl.k_blocks * kDenseBK * (pl.S_pad / 32u));
        oth_words = std::max<uint64_t>(oth_words, (uint64_t)pl.S * pl.oth_cap);
        oth_sites = std::max<uint32_t>(oth_sites, pl.S);
      }
      gram_words += 9ull * pl.S_pad * pl.S_pad;
      x_bytes = std::max<uint64_t>(x_bytes, 3ull * pl.S_pad * pl.k_blocks * kDenseBK);
      d.gram_off = pl.gram_off;
      d.S_pad = pl.S_pad;
      b->dense_plans.push_back(pl);
    } else if (np && !(nit == 1 && u.n_sites <= (uint32_t)kFastMaxS && u.n_reads <= (uint32_t)kFastMaxR)) {
      // tiled popcount path: MI of every pair precomputed by k_tile_mi into the dense scratch
      // (or, by default, counted on the tensor cores by k_tile_gram and finished by k_tile_finish)
      const bool gram = ctx->tile_path && u.n_reads >= 1u && u.n_reads <= kTgMaxReads;
      d.tiled = gram ? 2 : 1;
      if (nit == 1) {
        d.dense_off = dense;
        dense += np;
      }
      if (gram) {
        any_gram_tiled = true;
        const uint32_t nI = (u.n_sites + (uint32_t)kTgSitesI - 1u) / (uint32_t)kTgSitesI;
        const uint32_t nJ = (u.n_sites + (uint32_t)kTgSitesJ - 1u) / (uint32_t)kTgSitesJ;
        for (uint32_t I = 0; I < nI; ++I)
          for (uint32_t J = 0; J < nJ; ++J) {
            // some pair i < j with i in row block I and j in column block J
            const uint32_t j_max = std::min((J + 1u) * (uint32_t)kTgSitesJ, u.n_sites) - 1u;
            if (I * (uint32_t)kTgSitesI >= j_max) continue;
            gram_tiles.push_back(TgTile{u.plane_off, 0ull /* set below */, u.n_sites, u.row_words, u.site_off, (uint16_t)I, (uint16_t)J});
            gram_words_of.push_back(u.row_words);
          }
      } else {
        const uint32_t nb = (u.n_sites + (uint32_t)kTileSites - 1u) / (uint32_t)kTileSites;
        for (uint32_t I = 0; I < nb; ++I)
          for (uint32_t J = I; J < nb; ++J) {
            tile_items.push_back(TileItem{k, (uint16_t)I, (uint16_t)J});
            tile_words.push_back(u.row_words);
          }
      }
    }
    if (nit > 1) {
      d.dense_off = dense;
      dense += np;
      for (uint32_t s0 = 0; s0 < u.n_sites; s0 += (uint32_t)kMeanSites) mean_items.push_back(MeanItem{k, s0});
    }
    for (size_t t = first_gram_tile; t < gram_tiles.size(); ++t) gram_tiles[t].dense_off = d.dense_off;  // known now
    for (uint32_t t = 0; t < nit; ++t) {
      Item it;
      it.unit = k;
      it.pair_begin = t * (uint32_t)kPairsMax;
      it.pair_cnt = (uint32_t)std::min<uint64_t>(kPairsMax, np - (uint64_t)t * kPairsMax);
      it.flags = (t == 0 ? ITEM_FIRST : 0u) | (nit == 1 ? ITEM_SINGLE : 0u) | (d.tiled ? ITEM_TILED : 0u) |
                 (d.tiled == 2 ? ITEM_TILED_GRAM : 0u) | (d.gram_off != kNoGram ? ITEM_GRAM : 0u);
      if (d.tiled || d.gram_off != kNoGram) ++b->n_tiled_work_items;  // (not k_pairs_generic<0>'s)
      if (d.tiled) {
        TiledDesc td{};
        td.item_idx = (uint32_t)b->h_items.size();
        td.unit = k, td.pair_begin = it.pair_begin, td.pair_cnt = it.pair_cnt, td.flags = it.flags;
        td.S = u.n_sites, td.site_off = u.site_off, td.dense_off = d.dense_off;
        tiled_desc.push_back(td);
      }
      if (nit == 1 && np >= 1 && d.gram_off == kNoGram && u.n_sites <= (uint32_t)kFastMaxS && u.n_reads <= (uint32_t)kFastMaxR) {
        it.flags |= ITEM_FAST;  // (a unit sent to the tensor-core path by a lowered threshold is not also k_pairs_fast's)
        FastItem f;
        f.plane_off = u.plane_off;
        f.site_off = u.site_off;
        f.unit = k;
        f.item = (uint32_t)b->h_items.size();
        f.S = (uint16_t)u.n_sites;
        f.R = (uint16_t)u.n_reads;
        f.W = u.row_words;
        f.pad = 0;
        if (ctx->small_path && u.n_sites <= (uint32_t)kSgMaxS && val_slots + np + 1 < 0xffffffffull) {
          it.flags |= ITEM_PRE;
          f.pad = (uint32_t)val_slots;          // first slot of the unit's packed counts (even: 16-byte aligned)
          val_slots += (np + 1) & ~1ull;
          pre_items.push_back(f);
        } else {
          fast_items.push_back(f);
        }
      }
      b->h_items.push_back(it);
    }
  }
  auto deepest_first = [](auto& items, const std::vector<uint32_t>& words) {
    // the deepest tiles first: CTAs are handed out in order, so the long ones do not end up in the tail
    std::vector<uint32_t> order(items.size());
    for (uint32_t t = 0; t < order.size(); ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return words[x] > words[y]; });
    typename std::remove_reference<decltype(items)>::type sorted(items.size());
    for (uint32_t t = 0; t < order.size(); ++t) sorted[t] = items[order[t]];
    items.swap(sorted);
  };
  deepest_first(tile_items, tile_words);
  deepest_first(gram_tiles, gram_words_of);
  // per-site sums: a CTA's time grows with its unit's number of sites (one serial chain of S steps per site), so the
  // CTAs of the largest units go first and the short ones fill the tail
  std::stable_sort(mean_items.begin(), mean_items.end(), [&](const MeanItem& x, const MeanItem& y) {
    return units[x.unit].n_sites > units[y.unit].n_sites;
  });
  b->n_tile_items = (uint32_t)tile_items.size();
  b->n_gram_tiles = (uint32_t)gram_tiles.size();
  b->gram_kernel = ctx->tile_path == 2 ? 2 : 1;

  b->n_items = (uint32_t)b->h_items.size();
  b->n_fast = (uint32_t)fast_items.size();
  b->n_pre = (uint32_t)pre_items.size();
  b->n_mean_items = (uint32_t)mean_items.size();
  b->n_dense = dense;

#define BCU(call)                                                                    \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      int rc__ = fail(ctx, (e__ == cudaErrorMemoryAllocation) ? LGMI_ERR_NOMEM : LGMI_ERR_CUDA, \
                      "%s failed: %s", #call, cudaGetErrorString(e__));              \
      lgmi_batch_destroy(b);                                                         \
      return rc__;                                                                   \
    }                                                                                \
  } while (0)

  for (cudaEvent_t& e : b->ev) BCU(cudaEventCreate(&e));
  if ((b->n_fast || b->n_pre) && (!b->dense_plans.empty() || b->n_tile_items || b->n_gram_tiles)) {
    BCU(cudaStreamCreateWithFlags(&b->side_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : b->side_ev) BCU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  BCU(pool_malloc(ctx, &b->d_units, std::max<size_t>(1, n_units) * sizeof(DevUnit)));
  BCU(pool_malloc(ctx, &b->d_items, std::max<size_t>(1, b->n_items) * sizeof(Item)));
  BCU(pool_malloc(ctx, &b->d_fast_items, std::max<size_t>(1, fast_items.size()) * sizeof(FastItem)));
  BCU(pool_malloc(ctx, &b->d_fast_empty, std::max<size_t>(1, fast_items.size())));
  BCU(pool_malloc(ctx, &b->d_item_dense, std::max<size_t>(1, b->n_items)));
  BCU(pool_malloc(ctx, &b->d_n_generic, sizeof(uint32_t)));
  BCU(pool_malloc(ctx, &b->d_mean_items, std::max<size_t>(1, mean_items.size()) * sizeof(MeanItem)));
  BCU(pool_malloc(ctx, &b->d_planes, std::max<uint64_t>(4, plane_words) * sizeof(uint32_t)));
  BCU(pool_malloc(ctx, &b->d_flags, std::max<uint64_t>(1, n_sites)));
  BCU(pool_malloc(ctx, &b->d_item_cnt, ((size_t)b->n_items + 1) * sizeof(unsigned long long)));
  BCU(pool_malloc(ctx, &b->d_item_off, ((size_t)b->n_items + 1) * sizeof(unsigned long long)));
  BCU(cub::DeviceScan::ExclusiveSum(nullptr, b->scan_tmp_bytes, b->d_item_cnt, b->d_item_off, (int)b->n_items + 1,
                                    bstream(b)));
  BCU(pool_malloc(ctx, &b->d_scan_tmp, std::max<size_t>(b->scan_tmp_bytes, 16)));
  BCU(cudaMemsetAsync(b->d_item_cnt, 0, ((size_t)b->n_items + 1) * sizeof(unsigned long long), bstream(b)));
  BCU(pool_malloc(ctx, &b->d_header, sizeof(Header)));
  BCU(pool_malloc(ctx, &b->d_site_mean, std::max<uint64_t>(1, n_sites) * sizeof(double)));
  BCU(pool_malloc(ctx, &b->d_site_cnt, std::max<uint64_t>(1, n_sites) * sizeof(uint32_t)));
  BCU(pool_malloc(ctx, &b->d_dense, std::max<uint64_t>(1, dense) * sizeof(double)));
  BCU(pool_malloc(ctx, &b->d_unit_rec_off, ((size_t)n_units + 1) * sizeof(unsigned long long)));
  if (!pre_items.empty()) {
    BCU(pool_malloc(ctx, &b->d_pre_items, pre_items.size() * sizeof(FastItem)));
    BCU(cudaMemcpyAsync(b->d_pre_items, pre_items.data(), pre_items.size() * sizeof(FastItem), cudaMemcpyHostToDevice,
                        bstream(b)));
    BCU(pool_malloc(ctx, &b->d_val, val_slots * sizeof(unsigned long long)));
  }
  if (!tile_items.empty()) {
    BCU(pool_malloc(ctx, &b->d_tile_items, tile_items.size() * sizeof(TileItem)));
    BCU(cudaMemcpyAsync(b->d_tile_items, tile_items.data(), tile_items.size() * sizeof(TileItem), cudaMemcpyHostToDevice,
                        bstream(b)));
  }
  BCU(pool_malloc(ctx, &b->d_tile_next, sizeof(uint32_t)));
  if (!gram_tiles.empty()) {
    BCU(pool_malloc(ctx, &b->d_gram_tiles, gram_tiles.size() * sizeof(TgTile)));
    BCU(cudaMemcpyAsync(b->d_gram_tiles, gram_tiles.data(), gram_tiles.size() * sizeof(TgTile), cudaMemcpyHostToDevice,
                        bstream(b)));
  }
  b->n_tiled_desc = (uint32_t)tiled_desc.size();
  if (!tiled_desc.empty()) {
    BCU(pool_malloc(ctx, &b->d_tiled_desc, tiled_desc.size() * sizeof(TiledDesc)));
    BCU(cudaMemcpyAsync(b->d_tiled_desc, tiled_desc.data(), tiled_desc.size() * sizeof(TiledDesc), cudaMemcpyHostToDevice,
                        bstream(b)));
  }
  if (any_gram_tiled) BCU(pool_malloc(ctx, &b->d_tile_cnt, std::max<uint64_t>(1, dense) * 3ull * sizeof(uint2)));
  if (!b->dense_plans.empty()) {
    BCU(pool_malloc(ctx, &b->d_x, x_bytes));
    BCU(cudaMemsetAsync(b->d_x, 0, x_bytes, bstream(b)));  // rows of the site padding stay zero
    BCU(pool_malloc(ctx, &b->d_gram, gram_words * sizeof(uint32_t)));
    BCU(pool_malloc(ctx, &b->d_unit_mode, (size_t)n_units * sizeof(uint32_t)));
    if (xt_words) {
      BCU(pool_malloc(ctx, &b->d_xt, xt_words * sizeof(uint32_t)));
      BCU(pool_malloc(ctx, &b->d_oth_list, oth_words * sizeof(uint32_t)));
      BCU(pool_malloc(ctx, &b->d_oth_cnt, (size_t)oth_sites * sizeof(uint32_t)));
      BCU(cudaStreamCreateWithFlags(&b->fix_stream, cudaStreamNonBlocking));
      BCU(cudaEventCreateWithFlags(&b->fix_fork, cudaEventDisableTiming));
      BCU(cudaEventCreateWithFlags(&b->fix_join, cudaEventDisableTiming));
    }
    BCU(pool_malloc(ctx, &b->d_tiles, dense_tiles.size() * sizeof(DenseTile)));
    BCU(cudaMemcpyAsync(b->d_tiles, dense_tiles.data(), dense_tiles.size() * sizeof(DenseTile), cudaMemcpyHostToDevice,
                        bstream(b)));
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiled encode = reinterpret_cast<EncodeTiled>(ctx->encode_tiled);
    for (DensePlan& pl : b->dense_plans) {
      // X of this unit: [3 * S_pad rows][K_pad bytes], K contiguous; box = 128 rows x 128 bytes, 128B swizzle
      const cuuint64_t dims[2] = {(cuuint64_t)pl.k_blocks * kDenseBK, 3ull * pl.S_pad};
      const cuuint64_t strides[1] = {(cuuint64_t)pl.k_blocks * kDenseBK};
      const cuuint32_t box[2] = {(cuuint32_t)kDenseBK, 128u};
      const cuuint32_t estr[2] = {1u, 1u};
      const CUresult r = encode(&pl.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b->d_x, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        int rc__ = fail(ctx, LGMI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for unit %u", (int)r, pl.unit);
        lgmi_batch_destroy(b);
        return rc__;
      }
    }
  }
  b->rec_cap = std::max<uint64_t>(1, b->n_candidates);
  BCU(pool_malloc(ctx, &b->d_records, b->rec_cap * sizeof(lgmi_pair_rec)));
  BCU(pool_host_alloc(ctx, &b->h_header, sizeof(Header)));
  BCU(pool_host_alloc(ctx, &b->h_site_mean, std::max<uint64_t>(1, n_sites) * sizeof(double)));
  BCU(pool_host_alloc(ctx, &b->h_site_cnt, std::max<uint64_t>(1, n_sites) * sizeof(uint32_t)));
  BCU(pool_host_alloc(ctx, &b->h_unit_rec_off, ((size_t)n_units + 1) * sizeof(unsigned long long)));
  if (n_units) BCU(cudaMemcpyAsync(b->d_units, du.data(), n_units * sizeof(DevUnit), cudaMemcpyHostToDevice, bstream(b)));
  if (b->n_items) BCU(cudaMemcpyAsync(b->d_items, b->h_items.data(), b->n_items * sizeof(Item), cudaMemcpyHostToDevice, bstream(b)));
  if (!fast_items.empty())
    BCU(cudaMemcpyAsync(b->d_fast_items, fast_items.data(), fast_items.size() * sizeof(FastItem), cudaMemcpyHostToDevice,
                        bstream(b)));
  if (!mean_items.empty())
    BCU(cudaMemcpyAsync(b->d_mean_items, mean_items.data(), mean_items.size() * sizeof(MeanItem), cudaMemcpyHostToDevice, bstream(b)));
  BCU(cudaStreamSynchronize(bstream(b)));  // the host vectors above go out of scope
#undef BCU
  int rc = ensure_lntab(ctx, std::max<uint32_t>(b->max_reads, 1u));
  if (rc) {
    lgmi_batch_destroy(b);
    return rc;
  }
  *out = b;
  return LGMI_OK;
}

extern "C" int lgmi_batch_upload(lgmi_batch_t* b, const uint32_t* planes, const uint8_t* site_flags) {
  if (!b) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = b->ctx;
  if ((!planes && b->plane_words) || (!site_flags && b->n_sites))
    return fail(ctx, LGMI_ERR_ARG, "lgmi_batch_upload: NULL input");
  CU(ctx, cudaSetDevice(ctx->device));
  if (b->plane_words)
    CU(ctx, cudaMemcpyAsync(b->d_planes, planes, b->plane_words * sizeof(uint32_t), cudaMemcpyHostToDevice, bstream(b)));
  if (b->n_sites)
    CU(ctx, cudaMemcpyAsync(b->d_flags, site_flags, b->n_sites, cudaMemcpyHostToDevice, bstream(b)));
  b->n_het_candidates = b->n_sites ? het_candidates(b->h_units, site_flags) : 0;
  b->het_known = true;
  b->uploaded = true;
  return LGMI_OK;
}

static void drop_graphs(lgmi_batch* b) {
  for (lgmi_batch::Graph& g : b->graphs) cudaGraphExecDestroy(g.exec);
  b->graphs.clear();
}

static int run_chain(lgmi_batch* b, int min_common, uint32_t mode, bool timing);

extern "C" int lgmi_batch_run(lgmi_batch_t* b, int min_common, uint32_t mode) {
  if (!b) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = b->ctx;
  if (!b->uploaded) return fail(ctx, LGMI_ERR_STATE, "lgmi_batch_run: no input uploaded");
  if ((mode & LGMI_MODE_SKIP_NONHET) && !(mode & LGMI_MODE_HET_ONLY))
    return fail(ctx, LGMI_ERR_ARG, "LGMI_MODE_SKIP_NONHET requires LGMI_MODE_HET_ONLY");
  CU(ctx, cudaSetDevice(ctx->device));
  // LGMI_MODE_GRAPH (and every group of a pipelined step): the whole launch chain as ONE cudaGraphLaunch
  const bool as_graph = (mode & LGMI_MODE_GRAPH) != 0u || (b->own_stream != nullptr && ctx->pipeline_graphs);
  mode &= ~LGMI_MODE_GRAPH;
  // buffers the chain needs exist before anything is launched or captured
  if ((mode & LGMI_MODE_EMIT_COUNTS) && b->counts_cap < b->rec_cap) {
    CU(ctx, cudaStreamSynchronize(bstream(b)));
    if (b->d_counts) ctx->dev_pool.release(b->d_counts);
    b->d_counts = nullptr;
    b->counts_cap = 0;
    drop_graphs(b);
    CU(ctx, pool_malloc(ctx, &b->d_counts, b->rec_cap * 9ull * sizeof(uint32_t)));
    b->counts_cap = b->rec_cap;
  }
  if ((mode & LGMI_MODE_EMIT_COUNTS) && (b->n_tile_items || b->n_gram_tiles) && !b->d_tile_counts) {
    CU(ctx, cudaStreamSynchronize(bstream(b)));
    drop_graphs(b);
    CU(ctx, pool_malloc(ctx, &b->d_tile_counts, std::max<uint64_t>(1, b->n_dense) * 9ull * sizeof(uint32_t)));
  }
  if (!as_graph) return run_chain(b, min_common, mode, b->own_stream == nullptr);
  lgmi_batch::Graph* hit = nullptr;
  for (lgmi_batch::Graph& g : b->graphs)
    if (g.min_common == min_common && g.mode == mode) hit = &g;
  if (hit && hit->lntab_gen != ctx->lntab_gen) {  // the ln table has moved since the capture
    drop_graphs(b);
    hit = nullptr;
  }
  if (!hit) {
    cudaStream_t st = bstream(b);
    cudaGraph_t graph = nullptr;
    const uint64_t launches0 = ctx->launches;
    CU(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = run_chain(b, min_common, mode, false);
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const uint64_t n_launch = ctx->launches - launches0;
    ctx->launches = launches0;
    if (rc != LGMI_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc != LGMI_OK ? rc : fail(ctx, LGMI_ERR_CUDA, "graph capture of the launch chain failed: %s", cudaGetErrorString(e));
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(ctx, LGMI_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
    b->graphs.push_back(lgmi_batch::Graph{min_common, mode, ctx->lntab_gen, n_launch, exec});
    hit = &b->graphs.back();
  }
  CU(ctx, cudaGraphLaunch(hit->exec, bstream(b)));
  ctx->launches += hit->launches;
  b->ran = true;
  b->last_mode = mode;
  b->last_timed = false;
  return LGMI_OK;
}

// the launch chain of one run on the batch's stream (also what a graph capture records)
static int run_chain(lgmi_batch* b, int min_common, uint32_t mode, bool timing) {
  lgmi_ctx* ctx = b->ctx;
  RunParams P;
  P.units = b->d_units;
  P.items = b->d_items;
  P.n_items = b->n_items;
  P.tiled_desc = b->d_tiled_desc;
  P.n_tiled_desc = b->n_tiled_desc;
  P.n_units = b->n_units;
  P.planes = b->d_planes;
  P.site_flags = b->d_flags;
  P.lntab = ctx->d_lntab;
  P.ln_cap = (uint32_t)std::min<uint64_t>(ctx->ln_cap, 0xffffffffull);
  P.min_common = min_common;
  P.mode = mode;
  P.item_cnt = b->d_item_cnt;
  P.item_off = b->d_item_off;
  P.item_dense = b->d_item_dense;
  P.n_generic = b->d_n_generic;
  P.header = b->d_header;
  P.records = b->d_records;
  P.counts = b->d_counts;
  P.site_mean = b->d_site_mean;
  P.site_cnt = b->d_site_cnt;
  P.dense = b->d_dense;
  P.unit_rec_off = b->d_unit_rec_off;
  P.gram = b->d_gram;
  P.unit_mode = b->d_unit_mode;
  P.unit_base = b->unit_base;
  P.tile_counts = b->d_tile_counts;

  // (pipeline groups and graph replays skip the per-kernel events)
  if (timing) CU(ctx, cudaEventRecord(b->ev[0], bstream(b)));
  {
    // one launch resets the run state: header, offsets, flags, NaN means (sites of pair-less units keep them)
    const uint64_t n = std::max<uint64_t>(std::max<uint64_t>(b->n_sites, b->n_items), (uint64_t)b->n_units + 1);
    k_run_init<<<(unsigned)((n + 255) / 256), 256, 0, bstream(b)>>>(b->d_header, b->own_stream ? b->h_header : nullptr,
                                                                  b->d_unit_rec_off, b->n_units + 1,
                                                                  b->d_item_dense, b->n_items, b->d_n_generic,
                                                                  b->n_items - b->n_fast - b->n_pre - b->n_tiled_work_items, b->d_site_mean,
                                                                  b->d_site_cnt,
                                                                  b->n_sites, b->d_tile_next);
    ++ctx->launches;
  }
  if (timing) CU(ctx, cudaEventRecord(b->ev[4], bstream(b)));
  // untimed runs (graph captures, pipeline groups) of mixed batches: the small units' kernels on the side stream
  static const bool fork_ok = !(getenv("LGMI_FORK") && atoi(getenv("LGMI_FORK")) == 0);  // LGMI_FORK=0: one stream
  const bool forked = !timing && b->side_stream != nullptr && fork_ok;
  cudaStream_t small = forked ? b->side_stream : bstream(b);
  if (forked) {
    CU(ctx, cudaEventRecord(b->side_ev[0], bstream(b)));
    CU(ctx, cudaStreamWaitEvent(small, b->side_ev[0], 0));
  }
  if (!b->dense_plans.empty())  // every dense unit starts a run in the four-block form
    CU(ctx, cudaMemsetAsync(b->d_unit_mode, 0, (size_t)b->n_units * sizeof(uint32_t), bstream(b)));
  for (const DensePlan& pl : b->dense_plans) {
    // K3: bit-planes -> 0/1 bytes -> count matrices on the tensor cores (lgmi_dense.cuh)
    uint32_t* mode = b->d_unit_mode + pl.unit;
    const uint32_t* planes = b->d_planes + pl.plane_off;
    const uint64_t K_pad = (uint64_t)pl.k_blocks * kDenseBK;
    if (pl.four) {
      CU(ctx, cudaMemsetAsync(b->d_oth_cnt, 0, (size_t)pl.S * sizeof(uint32_t), bstream(b)));
      k_dense_prep<<<dim3((pl.W + 7u) / 8u, pl.S_pad / 256u), 256, 0, bstream(b)>>>(planes, pl.S, pl.W, pl.S_pad, b->d_xt,
                                                                                   b->d_oth_cnt, b->d_oth_list, pl.oth_cap, mode);
      ++ctx->launches;
      // the "other" cells need the lists and the transposed planes only: counted on a second stream, beside the
      // X rows and the first tiles of the GEMM; joined below, before the next unit's k_dense_prep reuses the lists and
      // before anything reads the tables.  (Measured: its three CTAs per SM hold the whole register file, so
      // k_dense_x mostly runs after it and the step costs what one fused pass cost; capping it at two CTAs per SM,
      // or k_dense_x first on a smaller grid, was slower -- tools/experiments/README.md.  The nine-block form no
      // longer pays for lists and transposes it does not use: 2.01 -> 1.65 ms.)
      CU(ctx, cudaEventRecord(b->fix_fork, bstream(b)));
      CU(ctx, cudaStreamWaitEvent(b->fix_stream, b->fix_fork, 0));
      k_other_fix<<<pl.S, kFixThreads, 0, b->fix_stream>>>(b->d_xt, K_pad, pl.S, pl.S_pad, b->d_oth_cnt, b->d_oth_list, pl.oth_cap,
                                                        b->d_gram + pl.gram_off, mode);
      ++ctx->launches;
      CU(ctx, cudaEventRecord(b->fix_join, b->fix_stream));
    } else {
      CU(ctx, cudaMemsetAsync(mode, 1, sizeof(uint32_t), bstream(b)));  // (any non-zero value: nine blocks)
    }
    {
      const uint64_t warps = (uint64_t)pl.S * ((pl.W + 7u) / 8u);
      const unsigned xgrid = (unsigned)std::min<uint64_t>((warps + 7) / 8, (uint64_t)ctx->num_sms * 16u);
      k_dense_x<<<xgrid, 256, 0, bstream(b)>>>(planes, pl.S, pl.W, pl.S_pad, b->d_x, mode);
      ++ctx->launches;
    }
    const bool last = &pl == &b->dense_plans.back();
    if (last && timing) CU(ctx, cudaEventRecord(b->ev[6], bstream(b)));
    for (int form = pl.four ? 0 : 1; form < 2; ++form) {  // the launch of the form the unit is not in returns at once
      const DenseTileList& L = form ? pl.list9 : pl.list4;
      DenseParams D;
      D.tiles = b->d_tiles + L.off;
      D.n_tiles = L.n;
      D.k_blocks = pl.k_blocks;
      D.S_pad = pl.S_pad;
      D.gram = b->d_gram + pl.gram_off;
      D.error = reinterpret_cast<uint32_t*>(&b->d_header->pad);
      D.mode = mode;
      D.nine = (uint32_t)form;
      const unsigned ggrid = (unsigned)std::min<uint32_t>(L.n, (uint32_t)ctx->num_sms);
      if (L.n > L.n_whole) {
        k_zero_partial_tiles<<<L.n - L.n_whole, 256, 0, bstream(b)>>>(D.tiles + L.n_whole, L.n - L.n_whole, pl.S_pad, D.gram,
                                                                     mode, (uint32_t)form);
        ++ctx->launches;
      }
      k_gram_i8<<<ggrid, kDenseThreads, kDenseSmemBytes, bstream(b)>>>(pl.tmap, D);
      ++ctx->launches;
    }
    if (last && timing) CU(ctx, cudaEventRecord(b->ev[7], bstream(b)));
    if (pl.four) CU(ctx, cudaStreamWaitEvent(bstream(b), b->fix_join, 0));
  }
  if (timing) CU(ctx, cudaEventRecord(b->ev[5], bstream(b)));
  if (b->n_tile_items) {
    // K1 + K2 of the medium units: counts + MI per 16 x 16 block of site pairs
    const unsigned tgrid = (unsigned)std::min<uint64_t>(b->n_tile_items, (uint64_t)ctx->num_sms * 3u);
    k_tile_mi<<<tgrid, kThreads, 0, bstream(b)>>>(P, b->d_tile_items, b->n_tile_items);
    ++ctx->launches;
  }
  if (b->n_gram_tiles) {
    // K1 of the medium units on the tensor cores (batched, bits expanded in the kernel), then K2 per pair
    TileGramParams T;
    T.planes = b->d_planes;
    T.site_flags = b->d_flags;
    T.mode = mode;
    T.tiles = b->d_gram_tiles;
    T.n_tiles = b->n_gram_tiles;
    T.cnt = b->d_tile_cnt;
    T.next = b->d_tile_next;
    T.error = reinterpret_cast<uint32_t*>(&b->d_header->pad);
    const unsigned ggrid = (unsigned)std::min<uint32_t>(b->n_gram_tiles, (uint32_t)ctx->num_sms);
    if (b->gram_kernel == 2) k_tile_gram_ws<<<ggrid, kTwThreads, kTwSmemBytes, bstream(b)>>>(T);
    else k_tile_gram<<<ggrid, kTgThreads, kTgSmemBytes, bstream(b)>>>(T);
    ++ctx->launches;
    const unsigned fgrid = (unsigned)std::min<uint64_t>(b->n_items, (uint64_t)ctx->num_sms * 8u);
    k_tile_finish<<<fgrid, kThreads, 0, bstream(b)>>>(P, b->d_tile_cnt);
    ++ctx->launches;
  }
  if (b->n_pre) {
    // small units: counts of every pair on the tensor cores -> packed values + emitted pairs per unit
    SgParams G;
    G.items = b->d_pre_items;
    G.n_items = b->n_pre;
    G.planes = b->d_planes;
    G.site_flags = b->d_flags;
    G.min_common = min_common;
    G.mode = mode;
    G.val = b->d_val;
    G.item_cnt = b->d_item_cnt;
    G.item_dense = b->d_item_dense;
    G.n_generic = b->d_n_generic;
    G.error = reinterpret_cast<uint32_t*>(&b->d_header->pad);
    const unsigned grid = (unsigned)std::min<uint64_t>(b->n_pre, (uint64_t)ctx->num_sms * 2u);
    k_small_gram<<<grid, kSgThreads, sizeof(SgSmem) + 1024, small>>>(G);
    ++ctx->launches;
  }
  if (b->n_fast) {
    // K0 of the small units: emitted pairs per unit from the C planes alone
    CountFastParams C;
    C.items = b->d_fast_items;
    C.n_items = b->n_fast;
    C.planes = b->d_planes;
    C.site_flags = b->d_flags;
    C.ij_tab = ctx->d_ij_tab;
    C.min_common = min_common;
    C.mode = mode;
    C.item_cnt = b->d_item_cnt;
    C.fast_empty = b->d_fast_empty;
    const unsigned grid = (unsigned)std::min<uint64_t>(b->n_fast, (uint64_t)ctx->num_sms * 8u);
    k_count_fast<<<grid, kThreads, 0, small>>>(C);
    ++ctx->launches;
  }
  if (forked) {  // the scan needs every item's count
    CU(ctx, cudaEventRecord(b->side_ev[1], small));
    CU(ctx, cudaStreamWaitEvent(bstream(b), b->side_ev[1], 0));
  }
  if (b->n_items) {
    // K0 of everything else + scan: every item's place in the ordered output
    if (b->n_items > b->n_fast + b->n_pre) {
      const unsigned cgrid = (unsigned)std::min<uint64_t>(b->n_items, (uint64_t)ctx->num_sms * 8u);
      k_count<<<cgrid, kThreads, 0, bstream(b)>>>(P);
      ++ctx->launches;
    }
    CU(ctx, cub::DeviceScan::ExclusiveSum(b->d_scan_tmp, b->scan_tmp_bytes, b->d_item_cnt, b->d_item_off,
                                          (int)b->n_items + 1, bstream(b)));
    ++ctx->launches;
    k_scan_finish<<<1, 32, 0, bstream(b)>>>(b->d_item_off + b->n_items, b->d_header,
                                            b->own_stream ? b->h_header : nullptr, b->d_unit_rec_off + b->n_units);
    ++ctx->launches;
  }
  if (timing) CU(ctx, cudaEventRecord(b->ev[2], bstream(b)));
  if (forked) {
    CU(ctx, cudaEventRecord(b->side_ev[2], bstream(b)));
    CU(ctx, cudaStreamWaitEvent(small, b->side_ev[2], 0));
  }
  if (b->n_fast || b->n_pre) {
    FastParams F;
    F.items = b->d_fast_items;
    F.n_items = b->n_fast;
    F.planes = b->d_planes;
    F.site_flags = b->d_flags;
    F.lntab = ctx->d_lntab;
    F.ln_cap = P.ln_cap;
    F.ij_tab = ctx->d_ij_tab;
    F.min_common = min_common;
    F.mode = mode;
    F.item_off = b->d_item_off;
    F.fast_empty = b->d_fast_empty;
    F.item_dense = b->d_item_dense;
    F.n_generic = b->d_n_generic;
    F.records = b->d_records;
    F.counts = b->d_counts;
    F.site_mean = b->d_site_mean;
    F.site_cnt = b->d_site_cnt;
    F.unit_rec_off = b->d_unit_rec_off;
    F.unit_base = b->unit_base;
    if (b->n_pre) {
      PreParams Q;
      Q.F = F;
      Q.F.items = b->d_pre_items;
      Q.F.n_items = b->n_pre;
      Q.val = b->d_val;
      const unsigned grid = (unsigned)std::min<uint64_t>(b->n_pre, (uint64_t)ctx->num_sms * ctx->pre_ctas_per_sm);
      k_pairs_pre<<<grid, kFastThreads, sizeof(PreSmem), small>>>(Q);
      ++ctx->launches;
    }
    if (b->n_fast) {
      // persistent CTAs: a whole number of CTAs per SM, never more than there are items
      const unsigned grid = (unsigned)std::min<uint64_t>(b->n_fast, (uint64_t)ctx->num_sms * ctx->pairs_ctas_per_sm);
      if ((mode & (LGMI_MODE_HET_ONLY | LGMI_MODE_SKIP_NONHET)) == (LGMI_MODE_HET_ONLY | LGMI_MODE_SKIP_NONHET))
        k_pairs_fast_het<<<grid, kFastThreads, sizeof(FastSmem), small>>>(F);
      else
        k_pairs_fast<<<grid, kFastThreads, sizeof(FastSmem), small>>>(F);
      ++ctx->launches;
    }
  }
  if (timing) CU(ctx, cudaEventRecord(b->ev[3], bstream(b)));
  if (forked) CU(ctx, cudaEventRecord(b->side_ev[3], small));
  if (b->n_items) {
    // everything the small-unit kernel does not take
    if (b->n_tiled_desc) {  // ordering + emission of what k_tile_mi / k_tile_finish computed
      const unsigned grid = (unsigned)std::min<uint64_t>(b->n_tiled_desc, (uint64_t)ctx->num_sms * 4u);
      k_pairs_generic<1><<<grid, kThreads, 0, bstream(b)>>>(P);
      ++ctx->launches;
    }
    if (!b->dense_plans.empty()) {  // the deep units: tables from the count matrices
      k_pairs_generic<2><<<(unsigned)std::min<uint64_t>(b->n_items, (uint64_t)ctx->num_sms * 3u), kThreads, 0, bstream(b)>>>(P);
      ++ctx->launches;
    }
  }
  if (b->n_mean_items) {  // (units of several items: none of them is k_pairs_fast's or k_pairs_generic<0>'s)
    k_site_mean_dense<<<b->n_mean_items, kMeanThreads, 0, bstream(b)>>>(b->d_units, b->d_mean_items, b->d_flags, b->d_dense,
                                                                       b->d_site_mean, b->d_site_cnt);
    ++ctx->launches;
  }
  if (forked) CU(ctx, cudaStreamWaitEvent(bstream(b), b->side_ev[3], 0));
  if (b->n_items) {
    // small units k_pairs_fast handed on (a site with too many "other" reads) and pair-less units; exits at once
    // when there is nothing
    const unsigned grid = (unsigned)std::min<uint64_t>(b->n_items, (uint64_t)ctx->num_sms * 4u);
    k_pairs_generic<0><<<grid, kThreads, 0, bstream(b)>>>(P);
    ++ctx->launches;
  }
  if (timing) CU(ctx, cudaEventRecord(b->ev[1], bstream(b)));
  CU(ctx, cudaGetLastError());
  b->ran = true;
  b->last_mode = mode;
  b->last_timed = timing;
  return LGMI_OK;
}

static int fill_scalars(lgmi_batch* b, lgmi_result* out) {
  lgmi_ctx* ctx = b->ctx;
  memset(out, 0, sizeof *out);
  out->n_candidates = b->n_candidates;
  // SKIP_NONHET evaluates only the candidates next to a het_snp (counted from the host flags; a caller that
  // filled the device buffers itself gets n_candidates)
  out->n_evaluated = ((b->last_mode & LGMI_MODE_SKIP_NONHET) && b->het_known) ? b->n_het_candidates : b->n_candidates;
  out->n_records = b->h_header->n_records;
  out->n_sites = b->n_sites;
  float ms = 0.f;
  const bool timing = b->last_timed;
  if (timing) {
    CU(ctx, cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]));
    out->kernel_ms = ms;
    CU(ctx, cudaEventElapsedTime(&ms, b->ev[2], b->ev[3]));
    out->pairs_kernel_ms = ms;
    CU(ctx, cudaEventElapsedTime(&ms, b->ev[4], b->ev[5]));
    out->dense_kernel_ms = ms;
  }
  if (!b->dense_plans.empty()) {
    // which form each deep unit took is decided on the device: read back (the stream has been synchronised)
    std::vector<uint32_t> modes(b->n_units);
    CU(ctx, cudaMemcpy(modes.data(), b->d_unit_mode, (size_t)b->n_units * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    b->dense_macs = 0;
    out->n_dense_four = 0;
    for (const DensePlan& pl : b->dense_plans) {
      const bool nine = modes[pl.unit] != 0u;
      b->dense_macs += nine ? pl.list9.macs : pl.list4.macs;
      out->n_dense_four += nine ? 0u : 1u;
    }
    if (timing) {
      CU(ctx, cudaEventElapsedTime(&ms, b->ev[6], b->ev[7]));
      out->gram_kernel_ms = ms;
      const DensePlan& pl = b->dense_plans.back();
      out->gram_macs = modes[pl.unit] ? pl.list9.macs : pl.list4.macs;
    }
  }
  out->n_dense_units = (uint32_t)b->dense_plans.size();
  out->dense_macs = b->dense_macs;
  if (b->h_header->pad)
    return fail(ctx, LGMI_ERR_CUDA, "k_gram_i8 / k_tile_gram: a pipeline barrier timed out (tensor-core path)");
  return LGMI_OK;
}

extern "C" int lgmi_batch_sync(lgmi_batch_t* b, lgmi_result* out) {
  if (!b || !out) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = b->ctx;
  if (!b->ran) return fail(ctx, LGMI_ERR_STATE, "lgmi_batch_sync: batch has not been run");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemcpyAsync(b->h_header, b->d_header, sizeof(Header), cudaMemcpyDeviceToHost, bstream(b)));
  CU(ctx, cudaStreamSynchronize(bstream(b)));
  return fill_scalars(b, out);
}

extern "C" int lgmi_batch_download(lgmi_batch_t* b, lgmi_result* out) {
  if (!b || !out) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = b->ctx;
  if (!b->ran) return fail(ctx, LGMI_ERR_STATE, "lgmi_batch_download: batch has not been run");
  CU(ctx, cudaSetDevice(ctx->device));
  // the small per-site / per-unit outputs can go while we learn the record count
  CU(ctx, cudaMemcpyAsync(b->h_header, b->d_header, sizeof(Header), cudaMemcpyDeviceToHost, bstream(b)));
  if (b->n_sites) {
    CU(ctx, cudaMemcpyAsync(b->h_site_mean, b->d_site_mean, b->n_sites * sizeof(double), cudaMemcpyDeviceToHost, bstream(b)));
    CU(ctx, cudaMemcpyAsync(b->h_site_cnt, b->d_site_cnt, b->n_sites * sizeof(uint32_t), cudaMemcpyDeviceToHost, bstream(b)));
  }
  CU(ctx, cudaMemcpyAsync(b->h_unit_rec_off, b->d_unit_rec_off, ((size_t)b->n_units + 1) * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, bstream(b)));
  CU(ctx, cudaStreamSynchronize(bstream(b)));
  const uint64_t nrec = b->h_header->n_records;
  if (nrec > b->rec_cap) return fail(ctx, LGMI_ERR_STATE, "record count %llu exceeds capacity", (unsigned long long)nrec);
  if (nrec > b->h_rec_cap) {
    if (b->h_records) ctx->host_pool.release(b->h_records);
    b->h_records = nullptr;
    b->h_rec_cap = 0;
    const uint64_t cap = std::max<uint64_t>(nrec + nrec / 8, 1024);
    CU(ctx, pool_host_alloc(ctx, &b->h_records, cap * sizeof(lgmi_pair_rec)));
    b->h_rec_cap = cap;
  }
  const bool want_counts = (b->last_mode & LGMI_MODE_EMIT_COUNTS) != 0u;
  if (want_counts && nrec > b->h_counts_cap) {
    if (b->h_counts) ctx->host_pool.release(b->h_counts);
    b->h_counts = nullptr;
    b->h_counts_cap = 0;
    const uint64_t cap = std::max<uint64_t>(nrec + nrec / 8, 1024);
    CU(ctx, pool_host_alloc(ctx, &b->h_counts, cap * 9ull * sizeof(uint32_t)));
    b->h_counts_cap = cap;
  }
  if (nrec) {
    CU(ctx, cudaMemcpyAsync(b->h_records, b->d_records, nrec * sizeof(lgmi_pair_rec), cudaMemcpyDeviceToHost, bstream(b)));
    if (want_counts)
      CU(ctx, cudaMemcpyAsync(b->h_counts, b->d_counts, nrec * 9ull * sizeof(uint32_t), cudaMemcpyDeviceToHost, bstream(b)));
    CU(ctx, cudaStreamSynchronize(bstream(b)));
  }
  int rc = fill_scalars(b, out);
  if (rc) return rc;
  out->records = b->h_records;
  out->counts = want_counts ? b->h_counts : nullptr;
  out->site_mean = b->h_site_mean;
  out->site_cnt = b->h_site_cnt;
  out->unit_rec_off = reinterpret_cast<const uint64_t*>(b->h_unit_rec_off);
  return LGMI_OK;
}

extern "C" int lgmi_batch_device_ptrs(lgmi_batch_t* b, void** d_planes, void** d_site_flags, void** d_records,
                                      void** d_site_mean) {
  if (!b) return LGMI_ERR_ARG;
  if (d_planes) *d_planes = b->d_planes;
  if (d_site_flags) *d_site_flags = b->d_flags;
  if (d_records) *d_records = b->d_records;
  if (d_site_mean) *d_site_mean = b->d_site_mean;
  b->uploaded = true;  // a device-resident caller fills the buffers itself
  b->het_known = false;
  return LGMI_OK;
}

extern "C" int lgmi_batch_algorithmic_bytes(lgmi_batch_t* b, uint64_t* bytes) {
  if (!b || !bytes) return LGMI_ERR_ARG;
  if (!b->ran) return fail(b->ctx, LGMI_ERR_STATE, "lgmi_batch_algorithmic_bytes: batch has not been run");
  uint64_t in = 0;
  for (const lgmi_unit_desc& u : b->h_units) in += 3ull * u.n_sites * ((u.n_reads + 7ull) / 8ull) + u.n_sites;
  *bytes = in + 16ull * b->h_header->n_records + 12ull * b->n_sites;
  return LGMI_OK;
}

// --------------------------------------------------------------------------- pipelined step
// The batch cut into consecutive groups of units, each a batch of its own on its own
// stream: H2D of group k+1, kernels of group k and D2H of group k-1 overlap (two copy
// engines + the SMs).  Outputs are merged into one set of pinned host arrays in the
// reference's row order, identical to upload + run + download of the whole batch.
struct lgmi_pipeline {
  lgmi_ctx* ctx = nullptr;
  uint32_t n_units = 0;
  uint64_t plane_words = 0, n_sites = 0, n_candidates = 0;
  struct Chunk {
    lgmi_batch* b = nullptr;
    uint32_t unit0 = 0, n_units = 0;
    uint64_t plane0 = 0, site0 = 0;
    cudaEvent_t done = nullptr;     // the group's kernels have finished
    cudaEvent_t landed = nullptr;   // the group's input is in device memory
    uint32_t* d_packed = nullptr;   // two-plane input staging (lgmi_pipeline_step_packed)
    unsigned long long* d_tight_off = nullptr;  // LGMI_MODE_TIGHT_INPUT: first word of each unit, relative to the group
    uint64_t tight0 = 0, tight_words = 0;       // the group's slice of the tight two-plane input
    double* d_rec_mi = nullptr;     // split output (LGMI_MODE_SPLIT_RECORDS)
    uint32_t* d_rec_ij = nullptr;
    uint64_t base = 0;              // first row of the group in the merged output of the current step
  };
  std::vector<Chunk> chunks;
  // one stream per copy direction: copies of a direction run one after the other anyway, and a stream that
  // copies both ways can land on one engine with another stream's uploads in front of its downloads
  cudaStream_t h2d = nullptr, d2h = nullptr;
  lgmi_pair_rec* h_records = nullptr;
  uint64_t h_rec_cap = 0;
  uint32_t* h_counts = nullptr;
  uint64_t h_counts_cap = 0;
  double* h_rec_mi = nullptr;
  uint64_t h_rec_mi_cap = 0;
  uint32_t* h_rec_ij = nullptr;   // (2-byte entries under LGMI_MODE_COMPACT_OUTPUT when every unit has <= 256 sites)
  uint64_t h_rec_ij_cap = 0;
  uint32_t max_sites = 0;
  std::vector<uint64_t> tight_off;  // per unit: first word in the tight two-plane input
  double* h_site_mean = nullptr;
  uint32_t* h_site_cnt = nullptr;
  unsigned long long* h_unit_rec_off = nullptr;
  // a step between lgmi_pipeline_begin* and lgmi_pipeline_finish
  bool in_flight = false;   // begun, not finished
  bool collected = false;   // ... and its downloads queued (lgmi_pipeline_collect)
  uint32_t pend_mode = 0;
  uint64_t pend_rows = 0;
};

extern "C" void lgmi_pipeline_destroy(lgmi_pipeline_t* p) {
  if (!p) return;
  lgmi_ctx* ctx = p->ctx;
  if (p->ctx) cudaSetDevice(p->ctx->device);
  // copies into the pinned output arrays must have landed before the arrays go back to the cache
  if (p->d2h) cudaStreamSynchronize(p->d2h);
  if (p->h2d) cudaStreamSynchronize(p->h2d);
  for (lgmi_pipeline::Chunk& c : p->chunks) {
    if (c.b) {
      cudaStream_t st = c.b->own_stream;
      lgmi_batch_destroy(c.b);
      if (st) cudaStreamDestroy(st);
    }
    if (c.done) cudaEventDestroy(c.done);
    if (c.landed) cudaEventDestroy(c.landed);
    ctx->dev_pool.release(c.d_packed);
    ctx->dev_pool.release(c.d_tight_off);
    ctx->dev_pool.release(c.d_rec_mi);
    ctx->dev_pool.release(c.d_rec_ij);
  }
  if (p->h2d) cudaStreamDestroy(p->h2d);
  if (p->d2h) cudaStreamDestroy(p->d2h);
  ctx->host_pool.release(p->h_rec_mi);
  ctx->host_pool.release(p->h_rec_ij);
  ctx->host_pool.release(p->h_records);
  ctx->host_pool.release(p->h_counts);
  ctx->host_pool.release(p->h_site_mean);
  ctx->host_pool.release(p->h_site_cnt);
  ctx->host_pool.release(p->h_unit_rec_off);
  delete p;
}

extern "C" int lgmi_pipeline_create(lgmi_t* ctx, const lgmi_unit_desc* units, uint32_t n_units, uint64_t plane_words,
                                    uint64_t n_sites, uint32_t n_chunks, lgmi_pipeline_t** out) {
  if (!ctx || !out || (!units && n_units)) return fail(ctx, LGMI_ERR_ARG, "lgmi_pipeline_create: NULL argument");
  *out = nullptr;
  if (n_chunks == 0) return fail(ctx, LGMI_ERR_ARG, "lgmi_pipeline_create: n_chunks must be >= 1");
  CU(ctx, cudaSetDevice(ctx->device));
  // groups are slices of the input buffers: units must be laid out back to back, in order
  uint64_t total = 0;
  for (uint32_t k = 0; k < n_units; ++k) {
    const lgmi_unit_desc& u = units[k];
    const uint64_t want_plane = k ? units[k - 1].plane_off + 3ull * units[k - 1].n_sites * units[k - 1].row_words : u.plane_off;
    const uint64_t want_site = k ? (uint64_t)units[k - 1].site_off + units[k - 1].n_sites : u.site_off;
    if (u.plane_off != want_plane || u.site_off != want_site)
      return fail(ctx, LGMI_ERR_UNSUPPORTED, "lgmi_pipeline_create: unit %u is not laid out right after unit %u", k, k - 1);
    total += u.n_sites >= 2 ? (uint64_t)u.n_sites * (u.n_sites - 1) / 2 : 0;
  }
  lgmi_pipeline* p = new (std::nothrow) lgmi_pipeline();
  if (!p) return fail(ctx, LGMI_ERR_NOMEM, "lgmi_pipeline_create: out of host memory");
  p->ctx = ctx;
  p->n_units = n_units;
  p->plane_words = plane_words;
  p->n_sites = n_sites;
  p->n_candidates = total;
#define PCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      int rc__ = fail(ctx, (e__ == cudaErrorMemoryAllocation) ? LGMI_ERR_NOMEM : LGMI_ERR_CUDA, \
                      "%s failed: %s", #call, cudaGetErrorString(e__));                        \
      lgmi_pipeline_destroy(p);                                                                \
      return rc__;                                                                             \
    }                                                                                          \
  } while (0)
  // the tight two-plane form (LGMI_MODE_TIGHT_INPUT): rows of ceil(R/32) words, units back to back
  p->tight_off.resize((size_t)n_units + 1);
  {
    uint64_t off = 0;
    for (uint32_t k = 0; k < n_units; ++k) {
      p->tight_off[k] = off;
      off += 2ull * units[k].n_sites * ((units[k].n_reads + 31u) / 32u);
      p->max_sites = std::max(p->max_sites, units[k].n_sites);
    }
    p->tight_off[n_units] = off;
  }
  // equal shares of the work: candidate pairs plus a term for the bytes to move
  std::vector<uint64_t> cost(n_units);
  uint64_t cost_total = 0;
  for (uint32_t k = 0; k < n_units; ++k) {
    const lgmi_unit_desc& u = units[k];
    cost[k] = (u.n_sites >= 2 ? (uint64_t)u.n_sites * (u.n_sites - 1) / 2 : 0) + (uint64_t)u.n_sites * u.row_words / 4 + 1;
    cost_total += cost[k];
  }
  n_chunks = std::min<uint32_t>(n_chunks, std::max<uint32_t>(1, n_units));
  uint32_t k = 0;
  uint64_t done_cost = 0;
  for (uint32_t c = 0; c < n_chunks && (k < n_units || c == 0); ++c) {
    lgmi_pipeline::Chunk ch;
    ch.unit0 = k;
    const uint64_t target = cost_total * (c + 1) / n_chunks;
    while (k < n_units && (done_cost < target || c + 1 == n_chunks)) done_cost += cost[k++];
    ch.n_units = k - ch.unit0;
    if (ch.n_units == 0 && n_units) continue;  // one unit outweighs a whole share
    ch.plane0 = ch.n_units ? units[ch.unit0].plane_off : 0;
    ch.site0 = ch.n_units ? units[ch.unit0].site_off : 0;
    std::vector<lgmi_unit_desc> local(units + ch.unit0, units + k);
    uint64_t pw = 0, ns = 0;
    for (lgmi_unit_desc& u : local) {
      u.plane_off -= ch.plane0;
      u.site_off -= (uint32_t)ch.site0;
      pw = u.plane_off + 3ull * u.n_sites * u.row_words;
      ns = (uint64_t)u.site_off + u.n_sites;
    }
    int rc = lgmi_batch_create(ctx, local.data(), ch.n_units, pw, ns, &ch.b);
    if (rc) {
      lgmi_pipeline_destroy(p);
      return rc;
    }
    p->chunks.push_back(ch);
    lgmi_pipeline::Chunk& ref = p->chunks.back();
    ref.b->unit_base = ref.unit0;  // records carry the unit's index in the whole batch
    ref.tight0 = p->tight_off[ref.unit0];
    ref.tight_words = p->tight_off[ref.unit0 + ref.n_units] - ref.tight0;
    if (ref.n_units) {
      std::vector<unsigned long long> rel(ref.n_units);
      for (uint32_t q = 0; q < ref.n_units; ++q) rel[q] = p->tight_off[ref.unit0 + q] - ref.tight0;
      PCU(pool_malloc(ctx, &ref.d_tight_off, rel.size() * sizeof(unsigned long long)));
      PCU(cudaMemcpy(ref.d_tight_off, rel.data(), rel.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    PCU(cudaStreamCreateWithFlags(&ref.b->own_stream, cudaStreamNonBlocking));
    PCU(cudaEventCreateWithFlags(&ref.done, cudaEventDisableTiming));
    PCU(cudaEventCreateWithFlags(&ref.landed, cudaEventDisableTiming));
  }
  PCU(cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking));
  PCU(cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking));
  PCU(pool_host_alloc(ctx, &p->h_site_mean, std::max<uint64_t>(1, n_sites) * sizeof(double)));
  PCU(pool_host_alloc(ctx, &p->h_site_cnt, std::max<uint64_t>(1, n_sites) * sizeof(uint32_t)));
  PCU(pool_host_alloc(ctx, &p->h_unit_rec_off, ((size_t)n_units + 1) * sizeof(unsigned long long)));
#undef PCU
  *out = p;
  return LGMI_OK;
}

// grows a pinned output array, keeping what has already landed in it
template <class T>
static int grow_pinned(lgmi_pipeline* p, T*& buf, uint64_t& cap, uint64_t need, uint64_t keep, size_t elem) {
  if (need <= cap) return LGMI_OK;
  lgmi_ctx* ctx = p->ctx;
  CU(ctx, cudaStreamSynchronize(p->d2h));  // copies into the old array
  const uint64_t new_cap = std::max<uint64_t>(need + need / 4, 1024);
  T* fresh = nullptr;
  CU(ctx, pool_host_alloc(ctx, &fresh, new_cap * elem));
  if (keep) memcpy(fresh, buf, keep * elem);
  if (buf) ctx->host_pool.release(buf);
  buf = fresh;
  cap = new_cap;
  return LGMI_OK;
}

// First half of a step: everything that does not depend on a count -- the uploads, the kernels and the row split of
// every group -- is queued, group after group, and the call returns.
static int pipeline_begin(lgmi_pipeline* p, const uint32_t* planes, bool packed, const uint8_t* site_flags, int min_common,
                          uint32_t mode) {
  if (!p) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = p->ctx;
  if (p->in_flight) return fail(ctx, LGMI_ERR_STATE, "lgmi_pipeline_begin: the previous step has not been finished");
  if ((!planes && p->plane_words) || (!site_flags && p->n_sites))
    return fail(ctx, LGMI_ERR_ARG, "lgmi_pipeline_step: NULL input");
  CU(ctx, cudaSetDevice(ctx->device));
  const bool compact = (mode & LGMI_MODE_COMPACT_OUTPUT) != 0u;  // split rows with 2-byte (i, j) where possible, no site_cnt
  const bool split = (mode & LGMI_MODE_SPLIT_RECORDS) != 0u || compact;
  const bool ij16 = compact && p->max_sites <= 256u;
  const bool tight = (mode & LGMI_MODE_TIGHT_INPUT) != 0u;
  if (tight && !packed) return fail(ctx, LGMI_ERR_ARG, "LGMI_MODE_TIGHT_INPUT is a form of the two-plane input (lgmi_pipeline_step_packed)");
  p->pend_mode = mode;
  mode &= ~(LGMI_MODE_SPLIT_RECORDS | LGMI_MODE_COMPACT_OUTPUT | LGMI_MODE_TIGHT_INPUT);  // a matter of the copies, not of the kernels
  for (lgmi_pipeline::Chunk& c : p->chunks) {
    lgmi_batch* b = c.b;
    cudaStream_t st = b->own_stream;
    int rc = LGMI_OK;
    // input on the upload stream, in group order; the group's own stream picks up when it has landed
    if (b->n_sites) CU(ctx, cudaMemcpyAsync(b->d_flags, site_flags + c.site0, b->n_sites, cudaMemcpyHostToDevice, p->h2d));
    const uint64_t words2 = b->plane_words / 3u * 2u;           // staging sized for the padded form: either fits
    const uint64_t words_in = tight ? c.tight_words : words2;
    if (packed) {  // two planes per site over PCIe, expanded on the device
      if (!c.d_packed && words2) CU(ctx, pool_malloc(ctx, &c.d_packed, words2 * sizeof(uint32_t)));
      if (words_in)
        CU(ctx, cudaMemcpyAsync(c.d_packed, planes + (tight ? c.tight0 : c.plane0 / 3u * 2u), words_in * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, p->h2d));
    } else if (b->plane_words) {
      CU(ctx, cudaMemcpyAsync(b->d_planes, planes + c.plane0, b->plane_words * sizeof(uint32_t), cudaMemcpyHostToDevice,
                              p->h2d));
    }
    CU(ctx, cudaEventRecord(c.landed, p->h2d));
    CU(ctx, cudaStreamWaitEvent(st, c.landed, 0));
    if (packed && words2) {
      k_unpack2<<<std::min<uint32_t>(b->n_units, (uint32_t)ctx->num_sms * 16u), 256, 0, st>>>(
          b->d_units, b->n_units, c.d_packed, b->d_planes, tight ? c.d_tight_off : nullptr);
      ++ctx->launches;
    }
    b->uploaded = true;
    if (mode & LGMI_MODE_SKIP_NONHET) {
      b->n_het_candidates = b->n_sites ? het_candidates(b->h_units, site_flags + c.site0) : 0;
      b->het_known = true;
    }
    if (!rc) rc = lgmi_batch_run(b, min_common, mode);
    if (rc) return rc;
    if (split) {
      if (!c.d_rec_mi) {
        CU(ctx, pool_malloc(ctx, &c.d_rec_mi, b->rec_cap * sizeof(double)));
        CU(ctx, pool_malloc(ctx, &c.d_rec_ij, b->rec_cap * sizeof(uint32_t)));
      }
      k_split_records<<<(unsigned)ctx->num_sms * 8u, 256, 0, st>>>(b->d_header, b->d_records, c.d_rec_mi, c.d_rec_ij,
                                                                   ij16 ? reinterpret_cast<uint16_t*>(c.d_rec_ij) : nullptr);
      ++ctx->launches;
    }
    CU(ctx, cudaEventRecord(c.done, st));  // the group's record count is in its (host-resident) header by then
  }
  p->in_flight = true;
  return LGMI_OK;
}

// Second half: waits for the groups in order and copies their rows to the host; a group's place in the merged
// arrays is known once the groups before it have been counted.
static int pipeline_collect(lgmi_pipeline* p) {
  if (!p) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = p->ctx;
  if (!p->in_flight) return fail(ctx, LGMI_ERR_STATE, "lgmi_pipeline_collect / finish: no step has been begun");
  if (p->collected) return LGMI_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  const uint32_t mode = p->pend_mode;
  const bool want_counts = (mode & LGMI_MODE_EMIT_COUNTS) != 0u;
  const bool compact = (mode & LGMI_MODE_COMPACT_OUTPUT) != 0u;
  const bool split = (mode & LGMI_MODE_SPLIT_RECORDS) != 0u || compact;
  const bool ij16 = compact && p->max_sites <= 256u;
  const size_t ij_bytes = ij16 ? sizeof(uint16_t) : sizeof(uint32_t);
  static const bool debug = getenv("LGMI_PIPE_DEBUG") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::vector<double> t_done, t_issue;
  uint64_t base = 0;
  for (lgmi_pipeline::Chunk& c : p->chunks) {
    lgmi_batch* b = c.b;
    CU(ctx, cudaEventSynchronize(c.done));
    if (debug) t_done.push_back(since());
    const uint64_t nrec = b->h_header->n_records;
    if (nrec > b->rec_cap) return fail(ctx, LGMI_ERR_STATE, "record count %llu exceeds capacity", (unsigned long long)nrec);
    int rc = LGMI_OK;
    if (split) {
      rc = grow_pinned(p, p->h_rec_mi, p->h_rec_mi_cap, base + nrec, base, sizeof(double));
      if (!rc) rc = grow_pinned(p, p->h_rec_ij, p->h_rec_ij_cap, base + nrec, base, sizeof(uint32_t));
    } else {
      rc = grow_pinned(p, p->h_records, p->h_rec_cap, base + nrec, base, sizeof(lgmi_pair_rec));
    }
    if (!rc && want_counts) rc = grow_pinned(p, p->h_counts, p->h_counts_cap, base + nrec, base, 9 * sizeof(uint32_t));
    if (rc) return rc;
    cudaStream_t st = p->d2h;  // the group is complete (event above): nothing here waits on a kernel
    if (nrec) {
      if (split) {
        CU(ctx, cudaMemcpyAsync(p->h_rec_mi + base, c.d_rec_mi, nrec * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(p->h_rec_ij) + base * ij_bytes, c.d_rec_ij, nrec * ij_bytes,
                                cudaMemcpyDeviceToHost, st));
      } else {
        CU(ctx, cudaMemcpyAsync(p->h_records + base, b->d_records, nrec * sizeof(lgmi_pair_rec), cudaMemcpyDeviceToHost, st));
      }
      if (want_counts)
        CU(ctx, cudaMemcpyAsync(p->h_counts + base * 9ull, b->d_counts, nrec * 9ull * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaMemcpyAsync(b->h_unit_rec_off, b->d_unit_rec_off, ((size_t)b->n_units + 1) * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    if (b->n_sites) {
      CU(ctx, cudaMemcpyAsync(p->h_site_mean + c.site0, b->d_site_mean, b->n_sites * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (!compact)
        CU(ctx, cudaMemcpyAsync(p->h_site_cnt + c.site0, b->d_site_cnt, b->n_sites * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    c.base = base;
    base += nrec;
    if (debug) t_issue.push_back(since());
  }
  p->pend_rows = base;
  p->collected = true;
  if (debug) {
    fprintf(stderr, "[lgmi pipeline] (ms since lgmi_pipeline_collect) groups counted at");
    for (double t : t_done) fprintf(stderr, " %.3f", t);
    fprintf(stderr, "; D2H issued by");
    for (double t : t_issue) fprintf(stderr, " %.3f", t);
    fprintf(stderr, " ms\n");
  }
  return LGMI_OK;
}

// Third part: the copies have landed; per-unit row offsets of the merged arrays, scalars.
static int pipeline_finish(lgmi_pipeline* p, lgmi_result* out) {
  if (!p || !out) return LGMI_ERR_ARG;
  lgmi_ctx* ctx = p->ctx;
  int rc_collect = pipeline_collect(p);
  if (rc_collect) {
    p->in_flight = p->collected = false;
    return rc_collect;
  }
  p->in_flight = p->collected = false;
  const uint32_t mode = p->pend_mode;
  const bool want_counts = (mode & LGMI_MODE_EMIT_COUNTS) != 0u;
  const bool compact = (mode & LGMI_MODE_COMPACT_OUTPUT) != 0u;
  const bool split = (mode & LGMI_MODE_SPLIT_RECORDS) != 0u || compact;
  const size_t ij_bytes = (compact && p->max_sites <= 256u) ? sizeof(uint16_t) : sizeof(uint32_t);
  const uint64_t base = p->pend_rows;
  CU(ctx, cudaStreamSynchronize(p->d2h));
  for (lgmi_pipeline::Chunk& c : p->chunks) {
    for (uint32_t u = 0; u < c.n_units; ++u) p->h_unit_rec_off[c.unit0 + u] = c.base + c.b->h_unit_rec_off[u];
  }
  p->h_unit_rec_off[p->n_units] = base;
  memset(out, 0, sizeof *out);
  out->n_candidates = p->n_candidates;
  out->n_evaluated = 0;
  out->n_records = base;
  out->records = split ? nullptr : p->h_records;
  out->rec_mi = split ? p->h_rec_mi : nullptr;
  out->rec_ij = split ? p->h_rec_ij : nullptr;
  out->rec_ij_bytes = split ? (uint32_t)ij_bytes : 0u;
  out->counts = want_counts ? p->h_counts : nullptr;
  out->n_sites = p->n_sites;
  out->site_mean = p->h_site_mean;
  out->site_cnt = compact ? nullptr : p->h_site_cnt;
  out->unit_rec_off = reinterpret_cast<const uint64_t*>(p->h_unit_rec_off);
  for (lgmi_pipeline::Chunk& c : p->chunks) {
    lgmi_result r;
    int rc = fill_scalars(c.b, &r);
    if (rc) return rc;
    out->n_evaluated += r.n_evaluated;
    out->kernel_ms += r.kernel_ms;
    out->pairs_kernel_ms += r.pairs_kernel_ms;
    out->dense_kernel_ms += r.dense_kernel_ms;
    out->n_dense_units += r.n_dense_units;
    out->dense_macs += r.dense_macs;
  }
  return LGMI_OK;
}

extern "C" int lgmi_pipeline_step(lgmi_pipeline_t* p, const uint32_t* planes, const uint8_t* site_flags, int min_common,
                                  uint32_t mode, lgmi_result* out) {
  if (!p || !out) return LGMI_ERR_ARG;
  int rc = pipeline_begin(p, planes, false, site_flags, min_common, mode);
  return rc ? rc : pipeline_finish(p, out);
}

extern "C" int lgmi_pipeline_step_packed(lgmi_pipeline_t* p, const uint32_t* planes2, const uint8_t* site_flags,
                                         int min_common, uint32_t mode, lgmi_result* out) {
  if (!p || !out) return LGMI_ERR_ARG;
  int rc = pipeline_begin(p, planes2, true, site_flags, min_common, mode);
  return rc ? rc : pipeline_finish(p, out);
}

extern "C" int lgmi_pipeline_begin(lgmi_pipeline_t* p, const uint32_t* planes, const uint8_t* site_flags, int min_common,
                                   uint32_t mode) {
  return pipeline_begin(p, planes, false, site_flags, min_common, mode);
}

extern "C" int lgmi_pipeline_begin_packed(lgmi_pipeline_t* p, const uint32_t* planes2, const uint8_t* site_flags,
                                          int min_common, uint32_t mode) {
  return pipeline_begin(p, planes2, true, site_flags, min_common, mode);
}

extern "C" int lgmi_pipeline_collect(lgmi_pipeline_t* p) { return pipeline_collect(p); }

extern "C" int lgmi_pipeline_finish(lgmi_pipeline_t* p, lgmi_result* out) { return pipeline_finish(p, out); }

// --------------------------------------------------------------------------- one-shot
extern "C" int lgmi_submit(lgmi_t* ctx, const lgmi_unit_desc* units, uint32_t n_units, const uint32_t* planes,
                           uint64_t plane_words, const uint8_t* site_flags, uint64_t n_sites, int min_common,
                           uint32_t mode) {
  if (!ctx) return LGMI_ERR_ARG;
  if (ctx->oneshot) lgmi_batch_destroy(ctx->oneshot);
  ctx->oneshot = nullptr;
  lgmi_batch* b = nullptr;
  int rc = lgmi_batch_create(ctx, units, n_units, plane_words, n_sites, &b);
  if (rc) return rc;
  ctx->oneshot = b;
  rc = lgmi_batch_upload(b, planes, site_flags);
  if (rc) return rc;
  return lgmi_batch_run(b, min_common, mode);
}

extern "C" int lgmi_wait(lgmi_t* ctx, lgmi_result* out) {
  if (!ctx || !out) return LGMI_ERR_ARG;
  if (!ctx->oneshot) return fail(ctx, LGMI_ERR_STATE, "lgmi_wait: nothing submitted");
  return lgmi_batch_download(ctx->oneshot, out);
}

// --------------------------------------------------------------------------- mean of rows
extern "C" int lgmi_site_mean_csr(lgmi_t* ctx, const uint64_t* offsets, const double* values, uint64_t n_sites,
                                  double* mean_out) {
  if (!ctx || !offsets || !mean_out) return fail(ctx, LGMI_ERR_ARG, "lgmi_site_mean_csr: NULL argument");
  if (n_sites == 0) return LGMI_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  const uint64_t nval = offsets[n_sites];
  if (nval && !values) return fail(ctx, LGMI_ERR_ARG, "lgmi_site_mean_csr: NULL values");
  const size_t off_bytes = (n_sites + 1) * sizeof(uint64_t);
  const size_t val_bytes = std::max<uint64_t>(1, nval) * sizeof(double);
  const size_t out_bytes = n_sites * sizeof(double);
  int rc = ensure_scratch(ctx, off_bytes + val_bytes + out_bytes + 64);
  if (rc) return rc;
  char* p = static_cast<char*>(ctx->d_scratch);
  unsigned long long* d_off = reinterpret_cast<unsigned long long*>(p);
  double* d_val = reinterpret_cast<double*>(p + off_bytes);
  double* d_out = reinterpret_cast<double*>(p + off_bytes + val_bytes);
  CU(ctx, cudaMemcpyAsync(d_off, offsets, off_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (nval) CU(ctx, cudaMemcpyAsync(d_val, values, nval * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_site_mean_csr<<<(unsigned)((n_sites + 127) / 128), 128, 0, ctx->stream>>>(d_off, d_val, n_sites, d_out);
  ++ctx->launches;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(mean_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LGMI_OK;
}

// --------------------------------------------------------------------------- ecdf
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int lgmi_ecdf(lgmi_t* ctx, const double* mean, const uint8_t* site_flags, uint64_t n, double threshold,
                         double* mip, uint8_t* call) {
  if (!ctx || (n && (!mean || !site_flags || !mip))) return fail(ctx, LGMI_ERR_ARG, "lgmi_ecdf: NULL argument");
  if (n == 0) return LGMI_OK;
  if (n > 0x7fffffffull) return fail(ctx, LGMI_ERR_UNSUPPORTED, "lgmi_ecdf: more than 2^31-1 sites");
  CU(ctx, cudaSetDevice(ctx->device));
  size_t sort_bytes = 0;
  CU(ctx, cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const double*)nullptr, (double*)nullptr, (int)n, 0, 64,
                                        ctx->stream));
  const size_t a_mean = 0;
  const size_t a_keys = align_up(a_mean + n * 8, 256);
  const size_t a_sorted = align_up(a_keys + n * 8, 256);
  const size_t a_mip = align_up(a_sorted + n * 8, 256);
  const size_t a_flags = align_up(a_mip + n * 8, 256);
  const size_t a_call = align_up(a_flags + n, 256);
  const size_t a_cnt = align_up(a_call + n, 256);
  const size_t a_tmp = align_up(a_cnt + 8, 256);
  int rc = ensure_scratch(ctx, a_tmp + sort_bytes);
  if (rc) return rc;
  char* p = static_cast<char*>(ctx->d_scratch);
  double* d_mean = reinterpret_cast<double*>(p + a_mean);
  double* d_keys = reinterpret_cast<double*>(p + a_keys);
  double* d_sorted = reinterpret_cast<double*>(p + a_sorted);
  double* d_mip = reinterpret_cast<double*>(p + a_mip);
  uint8_t* d_flags = reinterpret_cast<uint8_t*>(p + a_flags);
  uint8_t* d_call = reinterpret_cast<uint8_t*>(p + a_call);
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(p + a_cnt);
  CU(ctx, cudaMemcpyAsync(d_mean, mean, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d_flags, site_flags, n, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, ctx->stream));
  const unsigned grid = (unsigned)((n + 255) / 256);
  k_ecdf_keys<<<grid, 256, 0, ctx->stream>>>(d_mean, d_flags, n, d_keys, d_cnt);
  ++ctx->launches;
  CU(ctx, cub::DeviceRadixSort::SortKeys(p + a_tmp, sort_bytes, d_keys, d_sorted, (int)n, 0, 64, ctx->stream));
  ++ctx->launches;
  k_ecdf_mip<<<grid, 256, 0, ctx->stream>>>(d_mean, d_flags, n, d_sorted, d_cnt, threshold, d_mip, d_call);
  ++ctx->launches;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(mip, d_mip, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (call) CU(ctx, cudaMemcpyAsync(call, d_call, n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LGMI_OK;
}

// sorts x on the device (NaNs canonicalised: last, as numpy.sort leaves them); d_sorted / d_nan are in the scratch
static int ecdf_sort(lgmi_ctx* ctx, const double* x, uint64_t n, size_t extra_bytes, char** scratch, double** d_sorted,
                     unsigned long long** d_nan) {
  size_t sort_bytes = 0;
  CU(ctx, cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const double*)nullptr, (double*)nullptr, (int)n, 0, 64,
                                        ctx->stream));
  const size_t a_x = 0;
  const size_t a_sorted = align_up(a_x + n * 8, 256);
  const size_t a_nan = align_up(a_sorted + n * 8, 256);
  const size_t a_tmp = align_up(a_nan + 8, 256);
  const size_t a_extra = align_up(a_tmp + sort_bytes, 256);
  int rc = ensure_scratch(ctx, a_extra + extra_bytes);
  if (rc) return rc;
  char* p = static_cast<char*>(ctx->d_scratch);
  double* d_x = reinterpret_cast<double*>(p + a_x);
  *d_sorted = reinterpret_cast<double*>(p + a_sorted);
  *d_nan = reinterpret_cast<unsigned long long*>(p + a_nan);
  *scratch = p + a_extra;
  CU(ctx, cudaMemcpyAsync(d_x, x, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemsetAsync(*d_nan, 0, 8, ctx->stream));
  k_canon_nan<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_x, n, *d_nan);
  ++ctx->launches;
  CU(ctx, cub::DeviceRadixSort::SortKeys(p + a_tmp, sort_bytes, d_x, *d_sorted, (int)n, 0, 64, ctx->stream));
  ++ctx->launches;
  return LGMI_OK;
}

extern "C" int lgmi_ecdf_eval(lgmi_t* ctx, const double* x, uint64_t n, const double* samples, uint64_t m,
                              double* out) {
  if (!ctx || (n && !x) || (m && (!samples || !out))) return fail(ctx, LGMI_ERR_ARG, "lgmi_ecdf_eval: NULL argument");
  if (n == 0) return fail(ctx, LGMI_ERR_ARG, "lgmi_ecdf_eval: empty x (the reference divides by len(x))");
  if (m == 0) return LGMI_OK;
  if (n > 0x7fffffffull) return fail(ctx, LGMI_ERR_UNSUPPORTED, "lgmi_ecdf_eval: more than 2^31-1 values");
  CU(ctx, cudaSetDevice(ctx->device));
  char* extra = nullptr;
  double* d_sorted = nullptr;
  unsigned long long* d_nan = nullptr;
  const size_t a_o = align_up(m * 8, 256);
  int rc = ecdf_sort(ctx, x, n, a_o + m * 8, &extra, &d_sorted, &d_nan);
  if (rc) return rc;
  double* d_s = reinterpret_cast<double*>(extra);
  double* d_o = reinterpret_cast<double*>(extra + a_o);
  CU(ctx, cudaMemcpyAsync(d_s, samples, m * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_ecdf_eval<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(d_sorted, n, d_nan, d_s, m, d_o);
  ++ctx->launches;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(out, d_o, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LGMI_OK;
}

extern "C" int lgmi_ecdf_table(lgmi_t* ctx, const double* x, uint64_t n, double* sorted_out, double* y_out) {
  if (!ctx || (n && (!x || !sorted_out)) || !y_out) return fail(ctx, LGMI_ERR_ARG, "lgmi_ecdf_table: NULL argument");
  if (n == 0) return fail(ctx, LGMI_ERR_ARG, "lgmi_ecdf_table: empty x (the reference divides by len(x))");
  if (n > 0x7fffffffull) return fail(ctx, LGMI_ERR_UNSUPPORTED, "lgmi_ecdf_table: more than 2^31-1 values");
  CU(ctx, cudaSetDevice(ctx->device));
  char* extra = nullptr;
  double* d_sorted = nullptr;
  unsigned long long* d_nan = nullptr;
  int rc = ecdf_sort(ctx, x, n, (n + 1) * 8, &extra, &d_sorted, &d_nan);
  if (rc) return rc;
  double* d_y = reinterpret_cast<double*>(extra);
  k_ecdf_ordinates<<<(unsigned)((n + 256) / 256), 256, 0, ctx->stream>>>(n, d_y);
  ++ctx->launches;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(sorted_out, d_sorted, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(y_out, d_y, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LGMI_OK;
}

extern "C" int lgmi_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

// --------------------------------------------------------------------------- partitioning
extern "C" uint64_t lgmi_unit_cost(uint32_t n_sites, uint32_t n_reads) {
  if (n_sites < 2) return 0;
  return (uint64_t)n_sites * (n_sites - 1) / 2 * ((n_reads + 63ull) / 64ull);
}

extern "C" int lgmi_partition_lpt(const uint64_t* cost, uint32_t n_units, uint32_t n_bins, uint32_t* bin_of,
                                  uint64_t* bin_load) {
  if ((!cost && n_units) || !bin_of || n_bins == 0) return LGMI_ERR_ARG;
  std::vector<uint32_t> order(n_units);
  for (uint32_t k = 0; k < n_units; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });
  typedef std::pair<uint64_t, uint32_t> Bin;  // (load, bin); min-heap, ties -> lower bin index
  std::priority_queue<Bin, std::vector<Bin>, std::greater<Bin>> heap;
  for (uint32_t k = 0; k < n_bins; ++k) heap.push(Bin(0, k));
  std::vector<uint64_t> load(n_bins, 0);
  for (uint32_t k : order) {
    Bin top = heap.top();
    heap.pop();
    bin_of[k] = top.second;
    top.first += cost[k];
    load[top.second] = top.first;
    heap.push(top);
  }
  if (bin_load) memcpy(bin_load, load.data(), n_bins * sizeof(uint64_t));
  return LGMI_OK;
}

#include "lgmi_host.inl"
