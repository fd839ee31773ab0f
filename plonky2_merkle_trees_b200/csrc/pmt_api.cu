// pmt_api.cu -- C ABI of libpmt.so (include/pmt.h): context, host<->device staging and the launch plans of the
// tree / MMR builders.  Every entry point returns a status code and never throws across the boundary.
#include "../../include/pmt.h"
#include "merkle_kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is loaded with dlopen (pmt_comm_init), libpmt does not link against NCCL
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace pmt;

// A few persistent host threads that copy one buffer in parallel slices: the staging copies of the host-buffer builders
// when the caller's buffers are PAGEABLE (a plain Vec / malloc).  One memcpy thread moves ~10 GB/s, a PCIe 5 x16 link 55.
struct HostCopyPool {
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  char* dst = nullptr; const char* src = nullptr; size_t bytes = 0;
  unsigned generation = 0, pending = 0;
  bool stop = false;
  explicit HostCopyPool(unsigned n_threads) {
    for (unsigned t = 0; t < n_threads; t++) {
      try { workers.emplace_back([this, t] { run(t); }); } catch (...) { break; }
    }
  }
  ~HostCopyPool() {
    { std::lock_guard<std::mutex> g(m); stop = true; generation++; }
    cv_work.notify_all();
    for (auto& w : workers) w.join();
  }
  void slice(unsigned t, unsigned parts) const {   // part t of `parts`, 4 KiB-aligned cuts
    const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095, lo = (size_t)t * per;
    if (lo >= bytes) return;
    const size_t len = lo + per > bytes || t + 1 == parts ? bytes - lo : per;
    memcpy(dst + lo, src + lo, len);
  }
  void run(unsigned t) {
    unsigned seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m);
      cv_work.wait(lk, [&] { return generation != seen; });
      seen = generation;
      if (stop) return;
      lk.unlock();
      slice(t + 1, (unsigned)workers.size() + 1);
      lk.lock();
      if (--pending == 0) cv_done.notify_one();
    }
  }
  void copy(void* d, const void* s_, size_t n) {   // the caller copies slice 0 itself
    if (n < ((size_t)1 << 20) || workers.empty()) { memcpy(d, s_, n); return; }
    { std::lock_guard<std::mutex> g(m); dst = (char*)d; src = (const char*)s_; bytes = n; pending = (unsigned)workers.size(); generation++; }
    cv_work.notify_all();
    slice(0, (unsigned)workers.size() + 1);
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

struct pmt_ctx {
  int device = 0;
  int sms = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  char err[512] = {0};
  // grow-only staging arenas for the host-buffer entry points
  void* arena[3] = {nullptr, nullptr, nullptr};
  size_t arena_bytes[3] = {0, 0, 0};
  std::vector<void*> user_allocs;
  // optional per-launch timing (pmt_profile_enable): CUDA events on the launching stream around every kernel
  struct Rec { const char* name; double units; cudaEvent_t a, b; };
  bool profiling = false;
  std::vector<Rec> recs;
  bool rec_open = false;
  // copy streams + events of the pipelined host-buffer tree build (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> ev;
  // levels with at most this many nodes run the cooperative kernel (COOP_MAX; PMT_COOP_MAX_LOG2 is a tuning knob)
  size_t coop_max = (size_t)1 << 13;
  // the latency-bound levels of a perfect (sub)tree run as ONE launch (k_tree_coop: block-local subtrees, then the last
  // block to finish takes the levels above); PMT_FUSE_SUBTREES=0 restores one cooperative launch per level (A/B knob)
  bool fuse_subtrees = true;
  // tickets of k_tree_coop: a ring of zero-initialised device counters, one per launch in flight (the last block of a
  // launch resets its counter); rotating them keeps launches that overlap on different streams apart
  unsigned* tickets = nullptr;
  unsigned ticket_next = 0;
  // k_levels_wave: one flag per block of a launch (grow-only; a flag holds the epoch of the launch that set it), the
  // epoch counter, and PMT_WAVE=0 to go back to one launch per level (A/B knob)
  unsigned* wave_flags = nullptr;
  size_t wave_flags_n = 0;
  unsigned wave_epoch = 0;
  bool wave = true;
  size_t wave_min = (size_t)1 << 14;      // levels with fewer nodes than this stay one launch each (PMT_WAVE_MIN_LOG2)
  // one process per GPU: the NCCL communicator of pmt_comm_init (pmt_merkle_tree_build_sharded_dev)
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_world = 0;
  // one process, several GPUs: the event that publishes this ctx's subtree root to ctxs[0] (pmt_merkle_tree_build_multi_dev)
  cudaEvent_t root_ready = nullptr;
  std::vector<int> peers;      // devices this ctx's device has been given peer access to
  // the peer-memory exchange of the sharded builds (k_exchange_top): this ctx's mailbox, the table of every rank's mailbox as
  // this device addresses it, and who the table belongs to (MAIL_COMM: the ranks of pmt_comm_init, mapped with CUDA IPC;
  // MAIL_LOCAL: the contexts of one process, mail_group = their mailboxes in rank order)
  uint64_t* mail = nullptr;
  bool mail_exported = false;
  uint64_t** d_peers = nullptr;
  std::vector<void*> ipc_open;
  int mail_mode = 0, mail_rank = 0, mail_world = 0;
  std::vector<uint64_t*> mail_group;
  unsigned xchg_seq = 0;
  unsigned* fault_h = nullptr;
  unsigned* fault_d = nullptr;
  long long xchg_timeout_cycles = 0;
  // pageable caller buffers: library-owned page-locked staging slots (2 in, 2 out) and the copy threads
  void* stage[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t stage_bytes[4] = {0, 0, 0, 0};
  HostCopyPool* pool = nullptr;
};
constexpr unsigned TICKET_RING = 4096;
enum { MAIL_NONE = 0, MAIL_COMM = 1, MAIL_LOCAL = 2 };

static inline void prof_begin(pmt_ctx* c, const char* name, double units) {
  if (!c->profiling) return;
  pmt_ctx::Rec r{name, units, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c->stream);
  c->recs.push_back(r);
  c->rec_open = true;
}
static inline void prof_end(pmt_ctx* c) {
  if (!c->profiling || !c->rec_open) return;
  cudaEventRecord(c->recs.back().b, c->stream);
  c->rec_open = false;
}

namespace {

int fail(pmt_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(c, call)                                                                                        \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return fail((c), e_ == cudaErrorMemoryAllocation ? PMT_E_OOM : PMT_E_CUDA, "%s: %s (%s:%d)", #call,  \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                             \
  } while (0)

// TAG(c, "kernel", units) goes right before a launch, CHECK_LAUNCH(c) right after it
#define TAG(c, nm, u) prof_begin((c), (nm), (double)(u))

#define CHECK_LAUNCH(c)                                                                                    \
  do {                                                                                                     \
    (c)->launches++;                                                                                       \
    prof_end(c);                                                                                           \
    cudaError_t e_ = cudaGetLastError();                                                                   \
    if (e_ != cudaSuccess) return fail((c), PMT_E_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int bind(pmt_ctx* c) {
  if (!c) return PMT_E_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  return PMT_OK;
}

int log2_strict(size_t n) {  // plonky2_util::log2_strict: -1 unless n is a power of two
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

// grid for `count` independent permutations: ONE node per thread.  Measured on B200 (tools/ab_level.cu, 2^22 nodes):
// a persistent grid of resident blocks striding over the level is 7 % slower than count / 128 short-lived blocks, because
// blocks that start together stay in phase (all warps in the ALU-heavy S-box layer, then all in the DFMA layer), while
// blocks retiring at different times keep the warps of an SM in different phases so the two issue pipes overlap.
unsigned grid_for(const pmt_ctx* c, size_t count) {
  (void)c;
  size_t blocks = (count + BLOCK - 1) / BLOCK;
  if (blocks > 0x7fffffffu) blocks = 0x7fffffffu;   // threads stride over the rest
  return (unsigned)(blocks ? blocks : 1);
}

// grid for the narrow-leaf copy (w <= 4: hash_or_noop is a canonicalising copy, HBM-bound): a few waves of resident
// blocks striding over the rows (6.5 TB/s measured, against 4.4 TB/s with one row per thread)
unsigned grid_copy(const pmt_ctx* c, size_t count) {
  const size_t wave = (size_t)c->sms * 8;
  size_t blocks = (count + BLOCK - 1) / BLOCK;
  if (blocks > 8 * wave) blocks = 8 * wave;
  return (unsigned)(blocks ? blocks : 1);
}

int arena_get(pmt_ctx* c, int slot, size_t bytes, void** out) {
  if (bytes > c->arena_bytes[slot]) {
    if (c->arena[slot]) { CU(c, cudaFree(c->arena[slot])); c->arena[slot] = nullptr; c->arena_bytes[slot] = 0; }
    size_t want = bytes + bytes / 8 + 256;
    CU(c, cudaMalloc(&c->arena[slot], want));
    c->arena_bytes[slot] = want;
  }
  *out = c->arena[slot];
  return PMT_OK;
}

// One tree level of `count` nodes starting at node k0.  Big levels: one thread per node.  Levels that cannot fill the
// GPU that way (<= COOP_MAX nodes) are latency-bound: they run the cooperative 16-lanes-per-permutation kernel.
// Measured on B200 (tools/perm_bench.cu lat): a lone warp needs 38 us per thread-per-state permutation but 6.6 us per
// cooperative one; per permutation the cooperative form issues 2.4x more instructions, so it only wins while the level
// is latency-bound: <= 2^13 nodes.
constexpr size_t COOP_MAX = (size_t)1 << 13;
// proof batches up to this size are verified by quads (the GPU holds 148 x 3 x 64 = 28 416 such states at once; beyond
// that the thread-per-proof kernel's lower instruction count wins)
constexpr size_t COOP_VERIFY_MAX = (size_t)1 << 14;
// Hasher batches (rows, permutations) up to this size run by quads: latency instead of throughput
constexpr size_t COOP_ROWS_MAX = (size_t)1 << 12;

// n independent chains of dependent permutations (proof paths, sponge rows) by groups of threads: while one block per SM
// holds them all only latency counts -- 16-lane groups, 8 per block (one warp per sub-partition) or 16 per block; beyond
// that quads (4 threads per chain, 64 per block), the throughput form
struct CoopPlan { bool wide; unsigned threads, blocks; };
CoopPlan coop_plan(const pmt_ctx* c, size_t n) {
  const size_t sms = (size_t)c->sms;
  if (n <= sms * 8) return {true, 128, (unsigned)((n + 7) / 8)};
  if (n <= sms * WIDE_NODES) return {true, COOP_BLOCK, (unsigned)((n + WIDE_NODES - 1) / WIDE_NODES)};
  return {false, COOP_BLOCK, (unsigned)((n + COOP_NODES - 1) / COOP_NODES)};
}

int log2_floor(size_t x) { int l = 0; while (((size_t)2 << l) <= x) l++; return l; }
int ctz_cap(size_t x, int cap) { int z = 0; while (z < cap && !((x >> z) & 1)) z++; return z; }

// `n` consecutive zeroed counters for one k_tree_coop launch (1 for the final round + 1 per group of 8 blocks)
unsigned* next_ticket(pmt_ctx* c, unsigned n) {
  if (c->ticket_next + n > TICKET_RING) c->ticket_next = 0;
  unsigned* t = c->tickets + c->ticket_next;
  c->ticket_next += n;
  return t;
}

// The cooperative plan for the nodes k0 + [0, count) of level l (count <= coop_max) and up to `room` levels from there:
// returns how many levels the ONE launch covers.  Block-local subtrees need k0 and count aligned to their size; the
// ticket phase on top needs a perfect subtree (count a power of two, k0 a multiple of it).
template <class Layout>
int launch_coop(pmt_ctx* c, const Layout& lay, int l, int room, size_t k0, size_t count, unsigned sets = 1) {
  int local = 1, top = 0;
  if (c->fuse_subtrees && room > 1) {
    const int align = k0 ? (ctz_cap(k0, 6) < ctz_cap(count, 6) ? ctz_cap(k0, 6) : ctz_cap(count, 6)) : ctz_cap(count, 6);
    local = align + 1 < room ? align + 1 : room;
    if (local > COOP_LOCAL_LEVELS) local = COOP_LOCAL_LEVELS;
    const bool perfect = (count & (count - 1)) == 0 && k0 % count == 0;
    if (perfect && local == COOP_LOCAL_LEVELS && count > (size_t)COOP_NODES && sets == 1 &&
        count / COOP_NODES / TICKET_GROUP + 1 <= TICKET_RING / 4) {
      const int height = log2_floor(count) + 1;           // levels of the subtree, its root included
      top = (height < room ? height : room) - local;
    }
  }
  const size_t blocks = (count + COOP_NODES - 1) / COOP_NODES;
  size_t units = 0;
  for (int j = 0; j < local + top; j++) units += count >> j;
  TAG(c, top ? "k_tree_coop" : (local > 1 ? "k_subtree_coop" : "k_level_coop"), units * sets);
  k_tree_coop<Layout><<<dim3((unsigned)blocks, sets), COOP_BLOCK, 0, c->stream>>>(
      lay, l, k0, count, local, top, top ? next_ticket(c, 1 + (unsigned)(blocks / TICKET_GROUP)) : nullptr);
  CHECK_LAUNCH(c);
  return local + top;
}

// One tree level of `count` nodes starting at node k0, one thread per node: levels that fill the GPU.
template <class Layout>
int launch_level(pmt_ctx* c, const Layout& lay, int l, size_t k0, size_t count) {
  if (count == 0) return PMT_OK;
  TAG(c, "k_level", count);
  k_level<Layout><<<grid_for(c, count), BLOCK, 0, c->stream>>>(lay, l, k0, count);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// levels la .. la + n_levels - 1 of the nodes the leaves [n0, n1) complete, all big, in ONE launch (k_levels_wave)
template <class Layout>
int launch_wave(pmt_ctx* c, const Layout& lay, int la, int n_levels, size_t n0, size_t n1, size_t blocks,
                const uint64_t* d_rows = nullptr, size_t w = 0) {
  if (blocks > c->wave_flags_n) {               // grow-only; a fresh buffer is zero = no epoch
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->wave_flags) { CU(c, cudaFree(c->wave_flags)); c->wave_flags = nullptr; c->wave_flags_n = 0; }
    const size_t want = blocks + blocks / 4 + 1024;
    CU(c, cudaMalloc((void**)&c->wave_flags, want * sizeof(unsigned)));
    CU(c, cudaMemset(c->wave_flags, 0, want * sizeof(unsigned)));
    c->wave_flags_n = want;
  }
  if (++c->wave_epoch == 0) {                   // 2^32 launches later: start the epochs over on clean flags
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemset(c->wave_flags, 0, c->wave_flags_n * sizeof(unsigned)));
    c->wave_epoch = 1;
  }
  size_t units = 0;
  for (int j = 0; j < n_levels; j++) units += (n1 >> (la + j)) - (n0 >> (la + j));
  TAG(c, "k_level", units);
  k_levels_wave<Layout><<<(unsigned)blocks, BLOCK, 0, c->stream>>>(lay, la, n_levels, n0, n1, c->wave_flags, c->wave_epoch, next_ticket(c, 1),
                                                                   c->xchg_timeout_cycles, c->fault_d, d_rows, w);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// levels l_first .. l_last over the nodes that the leaves [n0, n1) complete: level l has the nodes k in [n0 >> l, n1 >> l)
// (a whole tree, a chunk of the pipelined tree build, a batch append to an MMR).  Big levels: one thread per node, one
// launch per level (the level has to be written anyway -- proofs need it -- and re-reading it costs 1/50 of the time its
// permutations take).  Levels that cannot fill the GPU that way (<= coop_max nodes) are latency-bound (a lone warp needs
// 38 us per thread-per-state permutation, ~6 us per cooperative one): cooperative launches, fused as far as the range's
// alignment allows -- for a perfect subtree all the way to its root in one launch.
template <class Layout>
int launch_level_span(pmt_ctx* c, const Layout& lay, int l_first, int l_last, size_t n0, size_t n1, const uint64_t* d_rows = nullptr,
                      size_t w = 0) {
  // d_rows: l_first == 1 and level 1 is computed from the narrow leaf rows of [n0, n1) themselves (n0 even, w <= 4, more than
  // coop_max level-1 nodes; the caller handles an unpaired last leaf): the leaf copy rides along with level 1
  for (int l = l_first; l <= l_last;) {
    const size_t k0 = n0 >> l, k1 = n1 >> l;
    if (k1 == 0) break;
    if (k1 <= k0) { l++; continue; }
    const size_t count = k1 - k0;
    const bool from_rows = d_rows != nullptr && l == 1;
    if (count > c->coop_max) {
      // the run of big levels from here that can go into one wavefront launch: level l' + 1 joins while it is big too and
      // level l' starts at an even node (then the children of its block b are exactly the blocks 2 b, 2 b + 1 of level l').
      // A wave only starts on a level that more than fills the GPU's resident blocks: below that the blocks of SEVERAL levels
      // become resident at once and take their numbers in arbitrary order, which spreads the small levels unevenly over the
      // SMs (measured: 2^16 + 2^15 + 2^14 nodes as a wave 155 us, as three launches 121 us)
      int l2 = l;
      size_t blocks = (count + BLOCK - 1) / BLOCK;
      const bool fills = count > (size_t)c->sms * PMT_MINB * BLOCK;
      while (c->wave && fills && l2 < l_last && !((n0 >> l2) & 1) && (n1 >> (l2 + 1)) - (n0 >> (l2 + 1)) > c->coop_max &&
             (n1 >> (l2 + 1)) - (n0 >> (l2 + 1)) >= c->wave_min) {
        l2++;
        blocks += ((n1 >> l2) - (n0 >> l2) + BLOCK - 1) / BLOCK;
      }
      if (l2 > l && blocks <= 0x7fffffffu) {
        if (int rc = launch_wave(c, lay, l, l2 - l + 1, n0, n1, blocks, from_rows ? d_rows : nullptr, w)) return rc;
        l = l2 + 1;
        continue;
      }
      if (from_rows) {
        TAG(c, "k_level", count);
        k_leaves_level1<Layout><<<grid_for(c, count), BLOCK, 0, c->stream>>>(lay, d_rows, w, k0, count);
        CHECK_LAUNCH(c);
      } else if (int rc = launch_level(c, lay, l, k0, count)) return rc;
      l++;
      continue;
    }
    const int done = launch_coop(c, lay, l, l_last - l + 1, k0, count);
    if (done < 0) return done;
    l += done;
  }
  return PMT_OK;
}

// levels l0 .. top of a perfect tree whose level l0 has `count` nodes starting at node 0
template <class Layout>
int run_levels(pmt_ctx* c, const Layout& lay, int l0, int top, size_t count) {
  return launch_level_span(c, lay, l0, top, 0, count << l0);
}

// level 0: digest(0, k0 + i) = hash_or_noop(row i).  Narrow rows (<= 4 felts) are a canonicalising copy (HBM-bound); wide
// rows are sponge-hashed, one row per thread when there are enough rows to fill the GPU, else one row per quad (the
// 17 permutations of a 135-column row are sequential: 0.65 ms per row-per-thread pass, ~0.1 ms by quads).
template <class Layout>
int launch_leaves(pmt_ctx* c, const Layout& lay, const uint64_t* d_rows, size_t w, size_t k0, size_t count) {
  if (count == 0) return PMT_OK;
  if (w > 4 && count <= c->coop_max) {
    const CoopPlan p = coop_plan(c, count);
    TAG(c, "k_rows_coop", count * ((w + 7) / 8));
    if (p.wide) k_rows_coop<Layout, true, Wide><<<p.blocks, p.threads, 0, c->stream>>>(lay, d_rows, w, k0, count);
    else k_rows_coop<Layout, true, Quad><<<p.blocks, p.threads, 0, c->stream>>>(lay, d_rows, w, k0, count);
  } else {
    TAG(c, "k_leaves", w <= 4 ? 0 : count * ((w + 7) / 8));
    k_leaves<Layout><<<w <= 4 ? grid_copy(c, count) : grid_for(c, count), BLOCK, 0, c->stream>>>(lay, d_rows, w, k0, count);
  }
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// Layout = Mmr (the whole post-order array is on the device) or MmrAppend (only the old peaks and the new elements are).
template <class Layout>
int mmr_extend_plan(pmt_ctx* c, const Layout& lay, size_t n0, const uint64_t* d_new, size_t m) {
  if ((n0 & 1) == 0 && m / 2 > c->coop_max) {   // every level-1 node has two NEW leaves: copy them and hash in one pass
    if (m & 1) {                             // the unpaired last leaf
      TAG(c, "k_leaves", 0);
      k_leaves<Layout><<<1, BLOCK, 0, c->stream>>>(lay, d_new + (m - 1), 1, n0 + m - 1, 1);
      CHECK_LAUNCH(c);
    }
    return launch_level_span(c, lay, 1, 40, n0, n0 + m, d_new, 1);
  }
  TAG(c, "k_leaves", 0);
  k_leaves<Layout><<<grid_copy(c, m), BLOCK, 0, c->stream>>>(lay, d_new, 1, n0, m);
  CHECK_LAUNCH(c);
  return launch_level_span(c, lay, 1, 40, n0, n0 + m);
}

// work(0 .. n-1), one host thread per index (index 0 on the calling thread), all joined before it returns.  Never throws:
// an exception inside work(r) becomes rcs[r] = PMT_E_OOM with a message on ctxs[r], and if no thread can be had that index
// runs on the caller.
template <class F>
void run_per_ctx(pmt_ctx* const* ctxs, std::vector<int>& rcs, F&& work) {
  const size_t n = rcs.size();
  auto guarded = [&](size_t r) {
    try { work(r); } catch (const std::exception& e) {
      rcs[r] = fail(ctxs[r], PMT_E_OOM, "worker of ctx %zu threw: %s", r, e.what());
    } catch (...) {
      rcs[r] = fail(ctxs[r], PMT_E_OOM, "worker of ctx %zu threw", r);
    }
  };
  std::vector<std::thread> pool;
  try { pool.reserve(n - 1); } catch (...) {}
  for (size_t r = 1; r < n; r++) {
    try { pool.emplace_back(guarded, r); } catch (...) { guarded(r); }
  }
  guarded(0);
  for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {

// ---- context -------------------------------------------------------------------------------------------------------
int pmt_init(pmt_ctx** out, int device_id) {
  if (!out) return PMT_E_INVALID_ARG;
  *out = nullptr;
  pmt_ctx* c = new (std::nothrow) pmt_ctx();
  if (!c) return PMT_E_OOM;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device_id < 0 || device_id >= n) {
    // no CPU fallback: without a usable device the engine refuses to exist
    delete c;
    return PMT_E_CUDA;
  }
  c->device = device_id;
  cudaDeviceProp prop;
  if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return PMT_E_CUDA;
  }
  c->sms = prop.multiProcessorCount;
  c->stream = c->own_stream;
  c->coop_max = COOP_MAX;
  if (const char* e2 = getenv("PMT_COOP_MAX_LOG2")) {   // tuning knob (tools/bench_configs.py), like PMT_PIPELINE_LOG2_CHUNKS
    const int lg = atoi(e2);
    if (lg >= 4 && lg <= 24) c->coop_max = (size_t)1 << lg;
  }
  if (const char* e3 = getenv("PMT_FUSE_SUBTREES")) c->fuse_subtrees = atoi(e3) != 0;
  if (const char* e4 = getenv("PMT_WAVE")) c->wave = atoi(e4) != 0;
  if (const char* e5 = getenv("PMT_WAVE_MIN_LOG2")) { const int lg = atoi(e5); if (lg >= 8 && lg <= 40) c->wave_min = (size_t)1 << lg; }
  if (cudaMalloc(&c->tickets, TICKET_RING * sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(c->tickets, 0, TICKET_RING * sizeof(unsigned)) != cudaSuccess ||
      // the word a kernel raises when a bounded wait expires (k_exchange_top: a peer; k_levels_wave: a block of the level
      // below), read by pmt_sync; host-mapped so that it survives whatever the stream does next
      cudaHostAlloc((void**)&c->fault_h, sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&c->fault_d, c->fault_h, 0) != cudaSuccess) {
    pmt_destroy(c);
    return PMT_E_OOM;
  }
  *c->fault_h = 0;
  {
    long long ms = 20000;                          // a peer that has not arrived after 20 s is reported, not waited for
    if (const char* e = getenv("PMT_EXCHANGE_TIMEOUT_MS")) { const long long v = atoll(e); if (v > 0) ms = v; }
    c->xchg_timeout_cycles = ms * (long long)(prop.clockRate > 0 ? prop.clockRate : 1965000);
  }
  // the cooperative kernels stage their round constants from global memory (poseidon_coop.cuh): fill the tables of this device
  poseidon::coop::k_coop_tables_init<<<1, 256, 0, c->stream>>>();
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
    pmt_destroy(c);
    return PMT_E_CUDA;
  }
  *out = c;
  return PMT_OK;
}

void pmt_destroy(pmt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (void* p : c->arena) if (p) cudaFree(p);
  for (void* p : c->user_allocs) cudaFree(p);
  if (c->tickets) cudaFree(c->tickets);
  if (c->wave_flags) cudaFree(c->wave_flags);
  if (c->comm) pmt_comm_destroy(c);
  if (c->root_ready) cudaEventDestroy(c->root_ready);
  for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
  // a mailbox that was exported to other processes may still be mapped there: it is left to process teardown
  if (c->mail && !c->mail_exported) cudaFree(c->mail);
  if (c->d_peers) cudaFree(c->d_peers);
  if (c->fault_h) cudaFreeHost(c->fault_h);
  for (void* p : c->stage) if (p) cudaFreeHost(p);
  delete c->pool;
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
}

const char* pmt_last_error(const pmt_ctx* c) { return c ? c->err : "null ctx"; }
const char* pmt_version(void) { return "pmt 0.1 (sm_100a)"; }
int pmt_device_id(const pmt_ctx* c) { return c ? c->device : -1; }
int pmt_set_stream(pmt_ctx* c, void* s) {
  if (int rc = bind(c)) return rc;
  cudaStream_t next = s ? (cudaStream_t)s : c->own_stream;
  // per-ctx device state (ticket counters, wavefront flags, mailbox slots) assumes that the ctx's launches run in order: drain
  // the stream that is being left before work goes to another one
  if (next != c->stream) CU(c, cudaStreamSynchronize(c->stream));
  c->stream = next;
  return PMT_OK;
}
void* pmt_get_stream(const pmt_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pmt_sync(pmt_ctx* c) {
  if (int rc = bind(c)) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->fault_h && *c->fault_h) {            // a kernel gave up a bounded wait: its outputs are garbage
    const unsigned v = *c->fault_h;
    *c->fault_h = 0;
    if (v & 0x80000000u) return fail(c, PMT_E_CUDA, "k_levels_wave: a block waited longer than the timeout for the level below (PMT_EXCHANGE_TIMEOUT_MS)");
    return fail(c, PMT_E_NCCL, "sharded build: rank %u did not deliver its subtree root within the exchange timeout (PMT_EXCHANGE_TIMEOUT_MS)", v - 1);
  }
  return PMT_OK;
}
uint64_t pmt_kernel_launches(const pmt_ctx* c) { return c ? c->launches : 0; }

// per-launch timing with CUDA events on the launching stream; read() synchronises, aggregates by kernel and resets.
// Output: one line per kernel "name launches total_ms total_units" (units = permutations the launches computed).
int pmt_profile_enable(pmt_ctx* c, int on) {
  if (!c) return PMT_E_INVALID_ARG;
  c->profiling = on != 0;
  return PMT_OK;
}
int pmt_profile_read(pmt_ctx* c, char* buf, size_t cap) {
  if (int rc = bind(c)) return rc;
  if (!buf || cap == 0) return fail(c, PMT_E_INVALID_ARG, "pmt_profile_read: null buffer");
  CU(c, cudaStreamSynchronize(c->stream));
  struct Agg { const char* name; int launches; double ms, units; };
  std::vector<Agg> agg;
  for (auto& r : c->recs) {
    float ms = 0;
    if (r.a && r.b && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      Agg* hit = nullptr;
      for (auto& a : agg) if (!strcmp(a.name, r.name)) hit = &a;
      if (!hit) { agg.push_back(Agg{r.name, 0, 0, 0}); hit = &agg.back(); }
      hit->launches++; hit->ms += ms; hit->units += r.units;
    }
    if (r.a) cudaEventDestroy(r.a);
    if (r.b) cudaEventDestroy(r.b);
  }
  c->recs.clear();
  c->rec_open = false;
  size_t off = 0;
  buf[0] = 0;
  for (auto& a : agg) {
    int k = snprintf(buf + off, cap - off, "%s %d %.6f %.0f\n", a.name, a.launches, a.ms, a.units);
    if (k < 0 || (size_t)k >= cap - off) break;
    off += (size_t)k;
  }
  return PMT_OK;
}

int pmt_malloc(pmt_ctx* c, size_t bytes, void** out) {
  if (int rc = bind(c)) return rc;
  if (!out) return fail(c, PMT_E_INVALID_ARG, "pmt_malloc: null out");
  void* p = nullptr;
  CU(c, cudaMalloc(&p, bytes ? bytes : 1));
  c->user_allocs.push_back(p);
  *out = p;
  return PMT_OK;
}
int pmt_free(pmt_ctx* c, void* p) {
  if (int rc = bind(c)) return rc;
  for (size_t i = 0; i < c->user_allocs.size(); i++)
    if (c->user_allocs[i] == p) {
      c->user_allocs.erase(c->user_allocs.begin() + i);
      CU(c, cudaFree(p));
      return PMT_OK;
    }
  return fail(c, PMT_E_INVALID_ARG, "pmt_free: pointer not owned by this ctx");
}
int pmt_memcpy_h2d(pmt_ctx* c, void* dst, const void* src, size_t bytes) {
  if (int rc = bind(c)) return rc;
  CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return PMT_OK;
}
int pmt_memcpy_d2h(pmt_ctx* c, void* dst, const void* src, size_t bytes) {
  if (int rc = bind(c)) return rc;
  CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return PMT_OK;
}

// ---- page-locked host memory for the pipelined host-buffer entry points ---------------------------------------------------
int pmt_host_register(pmt_ctx* c, void* ptr, size_t bytes) {
  if (int rc = bind(c)) return rc;
  if (!ptr || bytes == 0) return fail(c, PMT_E_INVALID_ARG, "pmt_host_register: null pointer / zero size");
  CU(c, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return PMT_OK;
}
int pmt_host_unregister(pmt_ctx* c, void* ptr) {
  if (int rc = bind(c)) return rc;
  if (!ptr) return fail(c, PMT_E_INVALID_ARG, "pmt_host_unregister: null pointer");
  CU(c, cudaHostUnregister(ptr));
  return PMT_OK;
}
int pmt_host_alloc(pmt_ctx* c, size_t bytes, void** out) {
  if (int rc = bind(c)) return rc;
  if (!out) return fail(c, PMT_E_INVALID_ARG, "pmt_host_alloc: null out");
  void* p = nullptr;
  CU(c, cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable));
  *out = p;
  return PMT_OK;
}
int pmt_host_free(pmt_ctx* c, void* ptr) {
  if (int rc = bind(c)) return rc;
  if (!ptr) return PMT_OK;
  CU(c, cudaFreeHost(ptr));
  return PMT_OK;
}

// ---- device-pointer entry points ---------------------------------------------------------------------------------------
int pmt_permute_dev(pmt_ctx* c, const uint64_t* d_in, size_t n, uint64_t* d_out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!d_in || !d_out) return fail(c, PMT_E_INVALID_ARG, "pmt_permute: null pointer");
  if (n <= COOP_ROWS_MAX) {
    TAG(c, "k_permute_coop", n);
    k_permute_coop<<<(unsigned)((n + COOP_NODES - 1) / COOP_NODES), COOP_BLOCK, 0, c->stream>>>(d_in, d_out, n);
  } else {
    TAG(c, "k_permute", n);
    k_permute<<<grid_for(c, n), BLOCK, 0, c->stream>>>(d_in, d_out, n);
  }
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_hash_two_to_one_dev(pmt_ctx* c, const uint64_t* d_l, const uint64_t* d_r, size_t n, uint64_t* d_out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!d_l || !d_r || !d_out) return fail(c, PMT_E_INVALID_ARG, "pmt_hash_two_to_one: null pointer");
  TAG(c, "k_two_to_one", n);
  k_two_to_one<<<grid_for(c, n), BLOCK, 0, c->stream>>>(d_l, d_r, d_out, n, 4);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_hash_rows_dev(pmt_ctx* c, const uint64_t* d_rows, size_t n, size_t w, int noop_rule, uint64_t* d_out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!d_out || (!d_rows && w)) return fail(c, PMT_E_INVALID_ARG, "pmt_hash_rows: null pointer");
  const bool permuted = !(w <= 4 && noop_rule);
  if (permuted && n <= COOP_ROWS_MAX) {   // few rows: the sponge's sequential permutations by groups of threads
    const CoopPlan p = coop_plan(c, n);
    TAG(c, "k_rows_coop", n * ((w + 7) / 8));
    if (noop_rule && p.wide) k_rows_coop<Flat, true, Wide><<<p.blocks, p.threads, 0, c->stream>>>(Flat{d_out}, d_rows, w, 0, n);
    else if (noop_rule) k_rows_coop<Flat, true, Quad><<<p.blocks, p.threads, 0, c->stream>>>(Flat{d_out}, d_rows, w, 0, n);
    else if (p.wide) k_rows_coop<Flat, false, Wide><<<p.blocks, p.threads, 0, c->stream>>>(Flat{d_out}, d_rows, w, 0, n);
    else k_rows_coop<Flat, false, Quad><<<p.blocks, p.threads, 0, c->stream>>>(Flat{d_out}, d_rows, w, 0, n);
  } else {
    TAG(c, "k_hash_rows", permuted ? n * ((w + 7) / 8) : 0);
    if (noop_rule) k_hash_rows<true><<<grid_for(c, n), BLOCK, 0, c->stream>>>(d_rows, n, w, d_out);
    else k_hash_rows<false><<<grid_for(c, n), BLOCK, 0, c->stream>>>(d_rows, n, w, d_out);
  }
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_simple_tree_build_dev(pmt_ctx* c, const uint64_t* d_leaves, size_t n, uint64_t* d_levels, uint64_t* d_root) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "simple tree: %zu leaves is not a power of two (log2_strict, simple_merkle_tree.rs:30)", n);
  if (lg < 1) return fail(c, PMT_E_INVALID_ARG, "simple tree: needs at least 2 leaves (simple_merkle_tree.rs:38)");
  if (!d_leaves || !d_levels || !d_root) return fail(c, PMT_E_INVALID_ARG, "simple tree: null pointer");
  LevelMajor lay{d_levels, d_root, n, lg};
  if (n / 2 > c->coop_max) return launch_level_span(c, lay, 1, lg, 0, n, d_leaves, 1);   // big tree: the leaf copy rides along with level 1
  TAG(c, "k_leaves", 0);
  k_leaves<LevelMajor><<<grid_copy(c, n), BLOCK, 0, c->stream>>>(lay, d_leaves, 1, 0, n);
  CHECK_LAUNCH(c);
  return run_levels(c, lay, 1, lg, n / 2);
}

int pmt_simple_tree_prove_dev(pmt_ctx* c, const uint64_t* d_levels, size_t n, const uint64_t* d_idx, size_t n_idx,
                              uint64_t* d_out) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 1) return fail(c, PMT_E_NOT_POW2, "simple tree prove: bad leaf count %zu", n);
  if (n_idx == 0) return PMT_OK;
  if (!d_levels || !d_idx || !d_out) return fail(c, PMT_E_INVALID_ARG, "simple tree prove: null pointer");
  LevelMajor lay{const_cast<uint64_t*>(d_levels), nullptr, n, lg};
  const size_t total = n_idx * (size_t)lg;
  k_simple_prove<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(lay, d_idx, n_idx, d_out);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// idx_mask: see k_verify_to_cap (all ones for upstream's verifier, the low path_len bits for the simple tree's)
static int verify_to_cap_dev(pmt_ctx* c, const uint64_t* d_rows, size_t w, const uint64_t* d_idx, size_t idx_mask, size_t n_idx,
                             const uint64_t* d_cap, uint32_t cap_height, const uint64_t* d_proofs, size_t path_len,
                             uint8_t* d_ok) {
  if (int rc = bind(c)) return rc;
  if (n_idx == 0) return PMT_OK;
  if (!d_rows || !d_idx || !d_cap || !d_ok || (!d_proofs && path_len)) return fail(c, PMT_E_INVALID_ARG, "verify: null pointer");
  if (w == 0 || cap_height > 40 || path_len > 63) return fail(c, PMT_E_INVALID_ARG, "verify: bad width / cap_height / path_len");
  if (n_idx <= COOP_VERIFY_MAX) {   // latency-bound batch: a group of threads per proof
    const CoopPlan p = coop_plan(c, n_idx);
    TAG(c, "k_verify_to_cap_coop", n_idx * (path_len + (w <= 4 ? 0 : (w + 7) / 8)));
    if (p.wide) k_verify_to_cap_coop<Wide><<<p.blocks, p.threads, 0, c->stream>>>(d_rows, w, d_idx, idx_mask, n_idx, d_cap, cap_height, d_proofs, path_len, d_ok);
    else k_verify_to_cap_coop<Quad><<<p.blocks, p.threads, 0, c->stream>>>(d_rows, w, d_idx, idx_mask, n_idx, d_cap, cap_height, d_proofs, path_len, d_ok);
  } else {
    TAG(c, "k_verify_to_cap", n_idx * (path_len + (w <= 4 ? 0 : (w + 7) / 8)));
    k_verify_to_cap<<<(unsigned)((n_idx + BLOCK - 1) / BLOCK), BLOCK, 0, c->stream>>>(d_rows, w, d_idx, idx_mask, n_idx, d_cap,
                                                                                     cap_height, d_proofs, path_len, d_ok);
  }
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_merkle_verify_dev(pmt_ctx* c, const uint64_t* d_rows, size_t w, const uint64_t* d_idx, size_t n_idx,
                          const uint64_t* d_cap, uint32_t cap_height, const uint64_t* d_proofs, size_t path_len,
                          uint8_t* d_ok) {
  return verify_to_cap_dev(c, d_rows, w, d_idx, ~(size_t)0, n_idx, d_cap, cap_height, d_proofs, path_len, d_ok);
}

int pmt_simple_tree_verify_dev(pmt_ctx* c, const uint64_t* d_leaves, const uint64_t* d_idx, size_t n_idx,
                               const uint64_t* d_root, const uint64_t* d_proofs, size_t path_len, uint8_t* d_ok) {
  // verify_merkle_proof == verify_to_cap with width 1 and a one-entry cap.  The reference folds by the parities of
  // leaf_index >> i for i < path_len and never looks at the bits above (simple_merkle_tree.rs:97-105): an index of
  // leaf_index + k 2^path_len verifies there, so it does here -- the index is masked to its low path_len bits.
  if (path_len > 63) return fail(c, PMT_E_INVALID_ARG, "verify: bad path_len");
  return verify_to_cap_dev(c, d_leaves, 1, d_idx, ((size_t)1 << path_len) - 1, n_idx, d_root, 0, d_proofs, path_len, d_ok);
}

int pmt_merkle_tree_build_dev(pmt_ctx* c, const uint64_t* d_leaves, size_t n, size_t w, uint32_t cap_height,
                              uint64_t* d_digests, uint64_t* d_cap) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "MerkleTree::new: %zu leaves is not a power of two (log2_strict)", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "MerkleTree::new: cap_height=%u should be at most log2(leaves.len())=%d", cap_height, lg);
  if (w == 0) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: zero-width leaves");
  if (!d_leaves || !d_cap || (!d_digests && (size_t)lg > cap_height)) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: null pointer");
  const int L = lg - (int)cap_height;
  Plonky2 lay{d_digests, d_cap, L};
  if (w <= 4 && L >= 1 && n / 2 > c->coop_max) {   // narrow leaves, big tree: the leaf copy rides along with level 1
    return launch_level_span(c, lay, 1, L, 0, n, d_leaves, w);
  }
  if (int rc = launch_leaves(c, lay, d_leaves, w, 0, n)) return rc;
  // levels 1 .. L over all subtrees at once: level l has n >> l nodes (2^h subtrees x 2^(L-l))
  if (L == 0) return PMT_OK;
  return run_levels(c, lay, 1, L, n / 2);
}

// MerkleTree::new fed by the prover's column-major LDE values (transpose + reverse_index_bits fused into the leaf kernel)
int pmt_merkle_tree_build_from_columns_dev(pmt_ctx* c, const uint64_t* d_columns, size_t n, size_t w, int bit_reverse,
                                           uint32_t cap_height, uint64_t* d_leaves_out, uint64_t* d_digests, uint64_t* d_cap) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "MerkleTree::new: %zu leaves is not a power of two (log2_strict)", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "MerkleTree::new: cap_height=%u should be at most log2(leaves.len())=%d", cap_height, lg);
  if (w == 0) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: zero-width leaves");
  if (!d_columns || !d_cap || (!d_digests && (size_t)lg > cap_height)) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: null pointer");
  const int L = lg - (int)cap_height;
  Plonky2 lay{d_digests, d_cap, L};
  TAG(c, "k_leaves_columns", w <= 4 ? 0 : n * ((w + 7) / 8));
  k_leaves_columns<Plonky2><<<w <= 4 ? grid_copy(c, n) : grid_for(c, n), BLOCK, 0, c->stream>>>(lay, d_columns, n, w, lg, bit_reverse != 0,
                                                                                             d_leaves_out);
  CHECK_LAUNCH(c);
  if (L == 0) return PMT_OK;
  return run_levels(c, lay, 1, L, n / 2);
}

int pmt_merkle_prove_dev(pmt_ctx* c, const uint64_t* d_digests, size_t n, uint32_t cap_height, const uint64_t* d_idx,
                         size_t n_idx, uint64_t* d_out) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "prove: %zu leaves is not a power of two", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "prove: cap_height %u > log2 n", cap_height);
  const int L = lg - (int)cap_height;
  if (n_idx == 0 || L == 0) return PMT_OK;
  if (!d_digests || !d_idx || !d_out) return fail(c, PMT_E_INVALID_ARG, "prove: null pointer");
  Plonky2 lay{const_cast<uint64_t*>(d_digests), nullptr, L};
  const size_t total = n_idx * (size_t)L;
  k_plonky2_prove<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(lay, n, d_idx, n_idx, d_out);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// top of a sharded tree: the g - h levels above the gathered subtree roots.  d_top_out is level-major: n_roots/2,
// n_roots/4, ..., 2^h digests (n_roots - 2^h in total); its last 2^h digests are the cap.
int pmt_top_levels_dev(pmt_ctx* c, const uint64_t* d_roots, size_t n_roots, uint32_t cap_height, uint64_t* d_top_out) {
  return pmt_top_levels_batch_dev(c, d_roots, 1, n_roots, cap_height, d_top_out);
}

// `batch` independent finishes of the same shape (the rounds of a sharded MMR): one launch for all sets when a set fits
// one block (<= 64 roots: always, with one root per rank), else one launch per set
int pmt_top_levels_batch_dev(pmt_ctx* c, const uint64_t* d_roots, size_t batch, size_t n_roots, uint32_t cap_height, uint64_t* d_top_out) {
  if (int rc = bind(c)) return rc;
  const int g = log2_strict(n_roots);
  if (g < 0) return fail(c, PMT_E_NOT_POW2, "top levels: %zu roots is not a power of two", n_roots);
  if ((int)cap_height > g) return fail(c, PMT_E_RANGE, "top levels: cap_height %u > log2(roots)", cap_height);
  if (batch > 65535) return fail(c, PMT_E_RANGE, "top levels batch: at most 65535 sets");
  if ((int)cap_height == g || batch == 0) return PMT_OK;   // the roots are the cap
  if (!d_roots || !d_top_out) return fail(c, PMT_E_INVALID_ARG, "top levels: null pointer");
  const size_t n_cap = (size_t)1 << cap_height;
  const int levels = g - (int)cap_height;                  // levels 1 .. levels of TopRoots
  if (n_roots / 2 <= (size_t)COOP_NODES) {
    TopRoots lay{d_roots, d_top_out, n_roots, n_roots, n_roots - n_cap, 1};
    const int done = launch_coop(c, lay, 1, levels, 0, n_roots / 2, (unsigned)batch);
    return done < 0 ? done : PMT_OK;
  }
  for (size_t b = 0; b < batch; b++) {
    TopRoots lay{d_roots + 4 * n_roots * b, d_top_out + 4 * (n_roots - n_cap) * b, n_roots, 0, 0, 1};
    if (int rc = launch_level_span(c, lay, 1, levels, 0, n_roots)) return rc;
  }
  return PMT_OK;
}

// ---- the peer-memory exchange (k_exchange_top): mailboxes, peer tables -------------------------------------------------------
// PMT_EXCHANGE=nccl keeps the sharded builds on ncclAllGather / peer copies + a separate finish launch (the A/B knob and the
// fallback when peers cannot map each other's memory); anything else: mailboxes where they can be set up.
static bool exchange_wanted() {
  const char* e = getenv("PMT_EXCHANGE");
  return !(e && (!strcmp(e, "nccl") || !strcmp(e, "copy") || !strcmp(e, "0")));
}
static int mail_alloc(pmt_ctx* c) {
  if (!c->mail) {
    CU(c, cudaMalloc((void**)&c->mail, MAIL_BYTES));
    CU(c, cudaMalloc((void**)&c->d_peers, MAIL_MAX_WORLD * sizeof(uint64_t*)));
  }
  // stream first: a kernel of an earlier group may still be reading the mailbox
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaMemset(c->mail, 0, MAIL_BYTES));
  c->xchg_seq = 0;
  return PMT_OK;
}
static void mail_close_peers(pmt_ctx* c) {
  for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
  c->ipc_open.clear();
  c->mail_mode = MAIL_NONE;
  c->mail_group.clear();
}
// one launch: push `mine`, wait for every peer, gather, finish `sets` sets of `world` roots (see k_exchange_top)
static int launch_exchange(pmt_ctx* c, const uint64_t* d_mine, size_t n_mine, uint64_t* d_gathered, size_t sets, uint64_t* d_tops,
                           int top_levels, uint64_t* d_finals) {
  const Exchange x{c->d_peers, (unsigned)c->mail_world, (unsigned)c->mail_rank, c->xchg_seq++, c->xchg_timeout_cycles, c->fault_d};
  TAG(c, "k_exchange_top", sets * ((size_t)c->mail_world - ((size_t)c->mail_world >> top_levels)));
  k_exchange_top<<<dim3(1, (unsigned)(sets ? sets : 1)), COOP_BLOCK, 0, c->stream>>>(x, d_mine, (unsigned)n_mine, d_gathered, (unsigned)sets,
                                                                                    d_tops, top_levels, d_finals);
  CHECK_LAUNCH(c);
  return PMT_OK;
}
// the contexts of ONE process as an exchange group (rank = position in ctxs): every device must be able to write every other
// device's memory.  Returns false (and leaves the contexts usable for the copy path) when that cannot be arranged.
static bool mail_local_group(pmt_ctx* const* ctxs, size_t n) {
  if (!exchange_wanted() || n < 2 || n > MAIL_MAX_WORLD) return false;
  bool cached = true;
  for (size_t i = 0; i < n && cached; i++) {
    pmt_ctx* c = ctxs[i];
    cached = c->mail_mode == MAIL_LOCAL && c->mail_rank == (int)i && c->mail_group.size() == n;
    for (size_t j = 0; j < n && cached; j++) cached = c->mail_group[j] == ctxs[j]->mail && ctxs[j]->mail;
  }
  if (cached) return true;
  for (size_t i = 0; i < n; i++) if (ctxs[i]->mail_mode == MAIL_COMM) return false;   // owned by a communicator
  for (size_t i = 0; i < n; i++)                  // a kernel of an earlier group may still be writing a mailbox that is re-zeroed below
    if (cudaSetDevice(ctxs[i]->device) != cudaSuccess || cudaStreamSynchronize(ctxs[i]->stream) != cudaSuccess) return false;
  // Contexts that share a device keep the copy path: their streams may share a hardware queue, and a kernel that spins on a
  // flag would then block the very kernel that sets it (PMT_EXCHANGE_SAME_DEVICE=1: the test hook for one-GPU boxes).
  const char* same = getenv("PMT_EXCHANGE_SAME_DEVICE");
  const bool same_ok = same && !strcmp(same, "1");
  for (size_t a = 0; a < n; a++)
    for (size_t b = 0; b < n; b++) {
      if (ctxs[a]->device == ctxs[b]->device) {
        if (a != b && !same_ok) return false;
        continue;
      }
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device) != cudaSuccess || !can) { cudaGetLastError(); return false; }
    }
  for (size_t i = 0; i < n; i++) {
    pmt_ctx* c = ctxs[i];
    if (cudaSetDevice(c->device) != cudaSuccess) return false;
    mail_close_peers(c);
    if (mail_alloc(c) != PMT_OK) return false;
    for (size_t b = 0; b < n; b++) {
      const int dev = ctxs[b]->device;
      if (dev == c->device || std::find(c->peers.begin(), c->peers.end(), dev) != c->peers.end()) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
      cudaGetLastError();
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return false;
      c->peers.push_back(dev);
    }
  }
  std::vector<uint64_t*> table(MAIL_MAX_WORLD, nullptr);
  for (size_t i = 0; i < n; i++) table[i] = ctxs[i]->mail;
  for (size_t i = 0; i < n; i++) {
    pmt_ctx* c = ctxs[i];
    if (cudaSetDevice(c->device) != cudaSuccess) return false;
    if (cudaMemcpy(c->d_peers, table.data(), MAIL_MAX_WORLD * sizeof(uint64_t*), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    c->mail_mode = MAIL_LOCAL;
    c->mail_rank = (int)i;
    c->mail_world = (int)n;
    c->mail_group.assign(table.begin(), table.begin() + n);
  }
  return true;
}

// ---- subtree-sharded MerkleTree::new, device resident ---------------------------------------------------------------------
static int check_ctxs(pmt_ctx* const* ctxs, size_t n_ctx, const char* who);
// the validation + local build shared by the two sharded forms: rank r of G = 2^g builds its n / G rows with the local cap
// height max(h - g, 0); cap entries / the root go to `d_cap_or_root`
static int sharded_local_build(pmt_ctx* c, const uint64_t* d_local_leaves, size_t n, size_t w, uint32_t cap_height, int g,
                               uint64_t* d_local_digests, uint64_t* d_cap_or_root) {
  const size_t per = n >> g;
  const uint32_t hl = (int)cap_height >= g ? cap_height - (uint32_t)g : 0;
  return pmt_merkle_tree_build_dev(c, d_local_leaves, per, w, hl, d_local_digests, d_cap_or_root);
}
static int sharded_check(pmt_ctx* c, size_t G, size_t n, size_t w, uint32_t cap_height, int* g_out) {
  const int g = log2_strict(G), lg = log2_strict(n);
  if (g < 0) return fail(c, PMT_E_NOT_POW2, "sharded build: %zu ranks / contexts is not a power of two", G);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "MerkleTree::new: %zu leaves is not a power of two (log2_strict)", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "MerkleTree::new: cap_height=%u should be at most log2(leaves.len())=%d", cap_height, lg);
  if (w == 0) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: zero-width leaves");
  if (g > lg) return fail(c, PMT_E_RANGE, "sharded build: more ranks (%zu) than leaves (%zu)", G, n);
  *g_out = g;
  return PMT_OK;
}

int pmt_merkle_tree_build_multi_dev(pmt_ctx* const* ctxs, size_t n_ctx, const uint64_t* const* d_leaves, size_t n, size_t w,
                                    uint32_t cap_height, uint64_t* const* d_digests, uint64_t* d_roots, uint64_t* d_top, uint64_t* d_cap) {
  if (int rc = check_ctxs(ctxs, n_ctx, "multi build")) return rc;
  pmt_ctx* c0 = ctxs[0];
  int g = 0;
  if (int rc = sharded_check(c0, n_ctx, n, w, cap_height, &g)) return rc;
  if (!d_leaves || !d_digests || !d_cap) return fail(c0, PMT_E_INVALID_ARG, "multi build: null pointer");
  const bool gather = (int)cap_height < g;
  if (gather && (!d_roots || (!d_top && n_ctx > 1))) return fail(c0, PMT_E_INVALID_ARG, "multi build: null d_roots / d_top");
  const size_t cap_l = gather ? 1 : (size_t)1 << (cap_height - (uint32_t)g);
  // Mailbox form: every context pushes its root / cap entries into every other context's mailbox and finishes the top itself
  // (k_exchange_top: one launch per context, no events, no copies); ctxs[0] writes the caller's d_roots / d_top / d_cap, the
  // others write scratch.  pmt_sync(ctxs[0]) still orders after every context's subtree: ctxs[0]'s kernel has seen their flags.
  const bool mailbox = n_ctx > 1 && cap_l <= MAIL_DIGESTS && mail_local_group(ctxs, n_ctx);
  const size_t n_cap = (size_t)1 << cap_height;
  // arena 2 of a context in the mailbox form: [64 digests + 64 B: other calls' scratch][mine: cap_l][gathered: G cap_l][tops: G][finals: n_cap].
  // Allocated for ALL contexts before anything is launched: cudaFree / cudaMalloc may wait for the device, and a context's
  // exchange kernel waits for its peers -- whose launches would then never be enqueued.
  const size_t mail_words = 64 * 4 + 8 + 4 * (cap_l + n_ctx * cap_l + n_ctx + n_cap);
  if (mailbox)
    for (size_t r = 0; r < n_ctx; r++) {
      void* stage = nullptr;
      if (int rc = bind(ctxs[r])) return rc;
      if (int rc = arena_get(ctxs[r], 2, mail_words * 8, &stage)) {
        char msg[sizeof ctxs[r]->err];
        memcpy(msg, ctxs[r]->err, sizeof msg);
        msg[sizeof msg - 1] = 0;
        return fail(c0, rc, "multi build: ctx %zu (device %d): %.400s", r, ctxs[r]->device, msg);
      }
    }
  // what context r enqueues on its own device: its subtree, then the exchange (mailbox) or the peer copy of its root / cap
  // entries to device 0 and an event
  auto enqueue = [&](size_t r) -> int {
    pmt_ctx* c = ctxs[r];
    if (int rc = bind(c)) return rc;
    if (mailbox) {
      uint64_t* d_mine = (uint64_t*)c->arena[2] + 64 * 4 + 8;
      uint64_t* sc_gathered = d_mine + 4 * cap_l, *sc_tops = sc_gathered + 4 * n_ctx * cap_l, *sc_finals = sc_tops + 4 * n_ctx;
      if (int rc = sharded_local_build(c, d_leaves[r], n, w, cap_height, g, d_digests[r], d_mine)) return rc;
      if (gather) return launch_exchange(c, d_mine, 1, r ? sc_gathered : d_roots, 1, r ? sc_tops : d_top, g - (int)cap_height, r ? sc_finals : d_cap);
      return launch_exchange(c, d_mine, cap_l, r ? sc_gathered : d_cap, 0, nullptr, 0, nullptr);
    }
    if (c->device != c0->device && std::find(c->peers.begin(), c->peers.end(), c0->device) == c->peers.end()) {
      int can = 0;                                   // let device r write device 0's memory (NVLink peer access), once
      CU(c, cudaDeviceCanAccessPeer(&can, c->device, c0->device));
      if (can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(c0->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, PMT_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", c->device, c0->device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      c->peers.push_back(c0->device);
    }
    // the subtree of ctx r, in its own device memory; its root / cap entries land in a local staging digest first
    void* stage = nullptr;
    if (int rc = arena_get(c, 2, 64 * 32 + 64 + cap_l * 32, &stage)) return rc;
    uint64_t* d_mine = (uint64_t*)stage + 64 * 4 + 8;
    if (int rc = sharded_local_build(c, d_leaves[r], n, w, cap_height, g, d_digests[r], d_mine)) return rc;
    uint64_t* dst = gather ? d_roots + 4 * r : d_cap + 4 * r * cap_l;
    CU(c, cudaMemcpyPeerAsync(dst, c0->device, d_mine, c->device, cap_l * 32, c->stream));
    if (!c->root_ready) CU(c, cudaEventCreateWithFlags(&c->root_ready, cudaEventDisableTiming));
    CU(c, cudaEventRecord(c->root_ready, c->stream));
    return PMT_OK;
  };
  // a dozen launches per device: from one host thread the last device would start ~0.4 ms after the first on an 8-GPU box
  // (measured: 2^24 leaves on 8 devices 2.08 ms enqueued serially), so from 4 contexts on every context is fed by its own thread
  std::vector<int> rcs(n_ctx, PMT_OK);
  if (n_ctx >= 4) run_per_ctx(ctxs, rcs, [&](size_t r) { rcs[r] = enqueue(r); });
  else for (size_t r = 0; r < n_ctx; r++) rcs[r] = enqueue(r);
  for (size_t r = 0; r < n_ctx; r++)
    if (rcs[r] != PMT_OK) {
      char msg[sizeof ctxs[r]->err];
      memcpy(msg, ctxs[r]->err, sizeof msg);
      msg[sizeof msg - 1] = 0;
      return fail(c0, rcs[r], "multi build: ctx %zu (device %d): %.400s", r, ctxs[r]->device, msg);
    }
  if (int rc = bind(c0)) return rc;
  if (mailbox) return PMT_OK;
  for (size_t r = 1; r < n_ctx; r++) CU(c0, cudaStreamWaitEvent(c0->stream, ctxs[r]->root_ready, 0));
  if (!gather) return PMT_OK;
  if (int rc = pmt_top_levels_dev(c0, d_roots, n_ctx, cap_height, d_top)) return rc;
  CU(c0, cudaMemcpyAsync(d_cap, d_top + 4 * (n_ctx - 2 * n_cap), n_cap * 32, cudaMemcpyDeviceToDevice, c0->stream));
  return PMT_OK;
}

// NCCL, loaded at run time
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    }
  }
  return &api;
}
}  // namespace
#define NC(c, call)                                                                                         \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess) return fail((c), PMT_E_NCCL, "%s: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

// The ranks of a communicator as an exchange group: every rank exports its mailbox (CUDA IPC), the handles travel by
// ncclAllGather, every rank maps its peers' mailboxes, and a second all-gather makes the decision unanimous -- either all
// ranks use k_exchange_top or all stay on ncclAllGather (ranks inside one process, or devices without peer access).
static int comm_setup_mail(pmt_ctx* c) {
  const int G = c->comm_world, r = c->comm_rank;
  if (G < 2 || G > (int)MAIL_MAX_WORLD || !exchange_wanted()) return PMT_OK;
  NcclApi* a = nccl_api();
  struct Info { cudaIpcMemHandle_t h; int ok; int pad; };
  static_assert(sizeof(Info) == 72, "Info is gathered as bytes");
  if (int rc = mail_alloc(c)) return rc;
  Info mine{};
  mine.ok = cudaIpcGetMemHandle(&mine.h, c->mail) == cudaSuccess;
  cudaGetLastError();
  c->mail_exported = c->mail_exported || mine.ok;
  void* buf = nullptr;
  if (int rc = arena_get(c, 2, (size_t)G * sizeof(Info), &buf)) return rc;
  std::vector<Info> all((size_t)G);
  auto gather = [&](const Info& v) -> int {
    CU(c, cudaMemcpyAsync((char*)buf + (size_t)r * sizeof(Info), &v, sizeof(Info), cudaMemcpyHostToDevice, c->stream));
    NC(c, a->AllGather((char*)buf + (size_t)r * sizeof(Info), buf, sizeof(Info), ncclChar, c->comm, c->stream));
    CU(c, cudaMemcpyAsync(all.data(), buf, (size_t)G * sizeof(Info), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return PMT_OK;
  };
  if (int rc = gather(mine)) return rc;
  std::vector<uint64_t*> table(MAIL_MAX_WORLD, nullptr);
  bool ok = true;
  for (int p = 0; p < G; p++) ok = ok && all[(size_t)p].ok;
  for (int p = 0; p < G && ok; p++) {
    if (p == r) { table[(size_t)p] = c->mail; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    c->ipc_open.push_back(ptr);
    table[(size_t)p] = (uint64_t*)ptr;
  }
  Info vote{};
  vote.ok = ok;
  if (int rc = gather(vote)) return rc;
  for (int p = 0; p < G; p++) ok = ok && all[(size_t)p].ok;
  if (!ok) { mail_close_peers(c); return PMT_OK; }
  CU(c, cudaMemcpy(c->d_peers, table.data(), MAIL_MAX_WORLD * sizeof(uint64_t*), cudaMemcpyHostToDevice));
  c->mail_mode = MAIL_COMM;
  c->mail_rank = r;
  c->mail_world = G;
  return PMT_OK;
}

int pmt_nccl_unique_id(pmt_ctx* c, void* id_out) {
  if (!c || !id_out) return PMT_E_INVALID_ARG;
  NcclApi* a = nccl_api();
  if (!a->ok) return fail(c, PMT_E_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
  ncclUniqueId id;
  NC(c, a->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof id);
  return PMT_OK;
}
int pmt_comm_init(pmt_ctx* c, const void* unique_id, int rank, int world) {
  if (int rc = bind(c)) return rc;
  if (!unique_id || world < 1 || rank < 0 || rank >= world) return fail(c, PMT_E_INVALID_ARG, "pmt_comm_init: bad arguments");
  if (log2_strict((size_t)world) < 0) return fail(c, PMT_E_NOT_POW2, "pmt_comm_init: world size %d is not a power of two", world);
  NcclApi* a = nccl_api();
  if (!a->ok) return fail(c, PMT_E_NCCL, "libnccl.so.2 could not be loaded");
  if (c->comm) pmt_comm_destroy(c);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof id);
  NC(c, a->CommInitRank(&c->comm, world, id, rank));
  c->comm_rank = rank;
  c->comm_world = world;
  if (c->mail_mode != MAIL_NONE) mail_close_peers(c);
  return comm_setup_mail(c);
}
int pmt_comm_uses_peer_memory(const pmt_ctx* c) { return c && c->comm && c->mail_mode == MAIL_COMM ? 1 : 0; }
int pmt_comm_destroy(pmt_ctx* c) {
  if (!c) return PMT_E_INVALID_ARG;
  if (c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->mail_mode == MAIL_COMM) mail_close_peers(c);
    nccl_api()->CommDestroy(c->comm);
    c->comm = nullptr;
    c->comm_world = 0;
  }
  return PMT_OK;
}

int pmt_merkle_tree_build_sharded_dev(pmt_ctx* c, const uint64_t* d_local_leaves, size_t n, size_t w, uint32_t cap_height,
                                      uint64_t* d_local_digests, uint64_t* d_roots, uint64_t* d_top, uint64_t* d_cap) {
  if (int rc = bind(c)) return rc;
  if (!c->comm) return fail(c, PMT_E_INVALID_ARG, "sharded build: pmt_comm_init has not been called on this ctx");
  const size_t G = (size_t)c->comm_world, r = (size_t)c->comm_rank;
  int g = 0;
  if (int rc = sharded_check(c, G, n, w, cap_height, &g)) return rc;
  if (!d_local_leaves || !d_cap) return fail(c, PMT_E_INVALID_ARG, "sharded build: null pointer");
  NcclApi* a = nccl_api();
  const bool mailbox = G > 1 && c->mail_mode == MAIL_COMM;     // unanimous among the ranks (comm_setup_mail)
  if ((int)cap_height >= g) {            // every rank yields 2^(h-g) cap entries: gathered in place into d_cap
    const size_t cap_l = (size_t)1 << (cap_height - (uint32_t)g);
    if (int rc = sharded_local_build(c, d_local_leaves, n, w, cap_height, g, d_local_digests, d_cap + 4 * r * cap_l)) return rc;
    if (mailbox && cap_l <= MAIL_DIGESTS) return launch_exchange(c, d_cap + 4 * r * cap_l, cap_l, d_cap, 0, nullptr, 0, nullptr);
    if (G > 1) NC(c, a->AllGather(d_cap + 4 * r * cap_l, d_cap, 4 * cap_l, ncclUint64, c->comm, c->stream));
    return PMT_OK;
  }
  if (!d_roots || !d_top) return fail(c, PMT_E_INVALID_ARG, "sharded build: null d_roots / d_top");
  if (int rc = sharded_local_build(c, d_local_leaves, n, w, cap_height, g, d_local_digests, d_roots + 4 * r)) return rc;
  // exchange + the g - h levels above the roots + the cap: ONE launch over peer memory
  if (mailbox) return launch_exchange(c, d_roots + 4 * r, 1, d_roots, 1, d_top, g - (int)cap_height, d_cap);
  NC(c, a->AllGather(d_roots + 4 * r, d_roots, 4, ncclUint64, c->comm, c->stream));
  if (int rc = pmt_top_levels_dev(c, d_roots, G, cap_height, d_top)) return rc;
  const size_t n_cap = (size_t)1 << cap_height;
  CU(c, cudaMemcpyAsync(d_cap, d_top + 4 * (G - 2 * n_cap), n_cap * 32, cudaMemcpyDeviceToDevice, c->stream));
  return PMT_OK;
}

// ---- MMR ---------------------------------------------------------------------------------------------------------------
size_t pmt_mmr_size(size_t n) { return 2 * n - (size_t)__builtin_popcountll((unsigned long long)n); }
size_t pmt_mmr_index(size_t i) { return 2 * i - (size_t)__builtin_popcountll((unsigned long long)i); }

// batch add_leaf (merkle_mountain_ranges.rs:89-120): the nodes created by appending leaves [n0, n0 + m) are, per
// height l, the k with n0 < (k + 1) 2^l <= n0 + m, i.e. k in [n0 >> l, (n0 + m) >> l).  Level l only reads level l - 1
// (new or already present in `elements`), so one launch per height is a correct schedule.
int pmt_mmr_extend_dev(pmt_ctx* c, uint64_t* d_elements, size_t n0, const uint64_t* d_new, size_t m) {
  if (int rc = bind(c)) return rc;
  if (m == 0) return PMT_OK;
  if (!d_elements || !d_new) return fail(c, PMT_E_INVALID_ARG, "mmr extend: null pointer");
  if (n0 + m > ((size_t)1 << 30)) return fail(c, PMT_E_RANGE, "mmr extend: more than 2^30 leaves (get_mmr_index is i32, merkle_mountain_ranges.rs:264)");
  return mmr_extend_plan(c, Mmr{d_elements}, n0, d_new, m);
}

int pmt_mmr_peaks_dev(pmt_ctx* c, const uint64_t* d_elements, size_t n_leaves, uint64_t* d_peaks, uint32_t* n_peaks_out) {
  if (int rc = bind(c)) return rc;
  const uint32_t k = (uint32_t)__builtin_popcountll((unsigned long long)n_leaves);
  if (n_peaks_out) *n_peaks_out = k;
  if (k == 0) return PMT_OK;
  if (!d_elements || !d_peaks) return fail(c, PMT_E_INVALID_ARG, "mmr peaks: null pointer");
  if (n_leaves >> 32) return fail(c, PMT_E_RANGE, "mmr peaks: size does not fit u32 (merkle_mountain_ranges.rs:184)");
  k_mmr_peaks<<<1, 64, 0, c->stream>>>(Mmr{const_cast<uint64_t*>(d_elements)}, n_leaves, d_peaks);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// bagging_the_peaks (:122-127) = hash_or_noop over the flattened peaks: ONE row of 4 k felts, hashed by one quad (the
// sponge's permutations are sequential)
static int bag_peaks(pmt_ctx* c, const uint64_t* d_peaks, uint32_t k, uint64_t* d_root) {
  TAG(c, "k_rows_coop", k <= 1 ? 0 : (4 * k + 7) / 8);
  k_rows_coop<Flat, true, Wide><<<1, 32, 0, c->stream>>>(Flat{d_root}, d_peaks, (size_t)4 * k, 0, 1);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_mmr_bag_dev(pmt_ctx* c, const uint64_t* d_elements, size_t n_leaves, uint64_t* d_root) {
  if (int rc = bind(c)) return rc;
  if (n_leaves == 0) return fail(c, PMT_E_INVALID_ARG, "mmr bag: empty MMR");
  if (!d_root) return fail(c, PMT_E_INVALID_ARG, "mmr bag: null pointer");
  void* peaks = nullptr;
  if (int rc = arena_get(c, 2, 64 * 32 + 64, &peaks)) return rc;
  uint32_t k = 0;
  if (int rc = pmt_mmr_peaks_dev(c, d_elements, n_leaves, (uint64_t*)peaks, &k)) return rc;
  return bag_peaks(c, (const uint64_t*)peaks, k, d_root);
}

int pmt_mmr_prove_dev(pmt_ctx* c, const uint64_t* d_elements, size_t n_leaves, const uint64_t* d_idx, size_t n_idx,
                      uint64_t* d_sib, uint8_t* d_left, uint32_t* d_len) {
  if (int rc = bind(c)) return rc;
  if (n_idx == 0) return PMT_OK;
  if (!d_elements || !d_idx || !d_sib || !d_left || !d_len) return fail(c, PMT_E_INVALID_ARG, "mmr prove: null pointer");
  if (n_leaves == 0 || n_leaves > ((size_t)1 << 30)) return fail(c, PMT_E_RANGE, "mmr prove: bad leaf count");
  const size_t total = n_idx * 32;
  k_mmr_prove<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(Mmr{const_cast<uint64_t*>(d_elements)}, n_leaves, d_idx,
                                                                      n_idx, d_sib, d_left, d_len);
  CHECK_LAUNCH(c);
  return PMT_OK;
}

int pmt_mmr_verify_dev(pmt_ctx* c, const uint64_t* d_leaves, size_t n_idx, const uint64_t* d_sib, const uint8_t* d_left,
                       const uint32_t* d_len, const uint64_t* d_peaks, uint32_t n_peaks, const uint64_t* d_root,
                       int8_t* d_status) {
  if (int rc = bind(c)) return rc;
  if (n_idx == 0) return PMT_OK;
  if (!d_leaves || !d_sib || !d_left || !d_len || !d_peaks || !d_root || !d_status) return fail(c, PMT_E_INVALID_ARG, "mmr verify: null pointer");
  if (n_peaks == 0 || n_peaks > 64) return fail(c, PMT_E_INVALID_ARG, "mmr verify: bad peak count");
  void* bag = nullptr;
  if (int rc = arena_get(c, 2, 64 * 32 + 64, &bag)) return rc;
  uint64_t* d_bag = (uint64_t*)bag + 64 * 4;
  if (int rc = bag_peaks(c, d_peaks, n_peaks, d_bag)) return rc;
  if (n_idx <= COOP_VERIFY_MAX) {
    const CoopPlan p = coop_plan(c, n_idx);
    TAG(c, "k_mmr_verify_coop", n_idx);
    if (p.wide) k_mmr_verify_coop<Wide><<<p.blocks, p.threads, 0, c->stream>>>(d_leaves, n_idx, d_sib, d_left, d_len, d_peaks, n_peaks, d_bag, d_root, d_status);
    else k_mmr_verify_coop<Quad><<<p.blocks, p.threads, 0, c->stream>>>(d_leaves, n_idx, d_sib, d_left, d_len, d_peaks, n_peaks, d_bag, d_root, d_status);
  } else {
    TAG(c, "k_mmr_verify", n_idx);
    k_mmr_verify<<<(unsigned)((n_idx + BLOCK - 1) / BLOCK), BLOCK, 0, c->stream>>>(d_leaves, n_idx, d_sib, d_left, d_len, d_peaks,
                                                                                  n_peaks, d_bag, d_root, d_status);
  }
  CHECK_LAUNCH(c);
  return PMT_OK;
}

// ---- host-buffer entry points: stage through the ctx arenas, run the *_dev plan, copy back, synchronise ------------------
#define H2D(c, dst, src, bytes) CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (c)->stream))
#define D2H(c, dst, src, bytes) CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (c)->stream))
#define FINISH(c) CU(c, cudaStreamSynchronize((c)->stream))

// Every host-buffer entry point returns through this: on an error nothing of the ctx may still be reading or writing the
// caller's buffers (DMA queued on the copy streams before the failing call would otherwise outlive the function, and a
// caller that frees its buffers after an error would race the copy engines).  Errors of the drain itself are ignored.
static int drained(pmt_ctx* c, int rc) {
  if (rc != PMT_OK) {
    if (c->copy_in) cudaStreamSynchronize(c->copy_in);
    if (c->copy_out) cudaStreamSynchronize(c->copy_out);
    cudaStreamSynchronize(c->stream);
    cudaGetLastError();
    return rc;
  }
  // a host-buffer call has synchronised already: a kernel that gave up a bounded wait (k_levels_wave) has left garbage in the
  // caller's buffers, and this is the call that must say so
  if (c->fault_h && *c->fault_h) return pmt_sync(c);
  return rc;
}

// the copy streams + events of the pipelined builders (created on first use); the copy streams are ordered behind the
// work already queued on the compute stream (earlier users of the arenas)
// true for plain pageable host memory (malloc / a Rust Vec / numpy): not page-locked, not registered, not managed
static bool is_pageable(const void* p) {
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}
// page-locked staging slot `slot` of at least `bytes`, and the copy threads (PMT_COPY_THREADS, default min(8, cores / 2))
static int stage_get(pmt_ctx* c, int slot, size_t bytes, void** out) {
  if (bytes > c->stage_bytes[slot]) {
    if (c->stage[slot]) { CU(c, cudaFreeHost(c->stage[slot])); c->stage[slot] = nullptr; c->stage_bytes[slot] = 0; }
    CU(c, cudaHostAlloc(&c->stage[slot], bytes, cudaHostAllocDefault));
    c->stage_bytes[slot] = bytes;
  }
  *out = c->stage[slot];
  if (!c->pool) {
    unsigned n = std::thread::hardware_concurrency() / 2;
    if (n > 8) n = 8;
    if (const char* e = getenv("PMT_COPY_THREADS")) n = (unsigned)atoi(e);
    if (n < 1) n = 1;
    c->pool = new (std::nothrow) HostCopyPool(n - 1);
    if (!c->pool) return fail(c, PMT_E_OOM, "host copy pool");
  }
  return PMT_OK;
}

static int pipeline_streams(pmt_ctx* c, size_t chunks) {
  if (!c->copy_in) CU(c, cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  if (!c->copy_out) CU(c, cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  while (c->ev.size() < 3 * chunks + 2) {      // per chunk: H2D done, hashed, D2H done (staged path); + the fence
    cudaEvent_t e;
    CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev.push_back(e);
  }
  CU(c, cudaEventRecord(c->ev[2 * chunks], c->stream));
  CU(c, cudaStreamWaitEvent(c->copy_in, c->ev[2 * chunks], 0));
  CU(c, cudaStreamWaitEvent(c->copy_out, c->ev[2 * chunks], 0));
  return PMT_OK;
}

static int pmt_permute_impl(pmt_ctx* c, const uint64_t* in, size_t n, uint64_t* out);
int pmt_permute(pmt_ctx* c, const uint64_t* in, size_t n, uint64_t* out) { return c ? drained(c, pmt_permute_impl(c, in, n, out)) : PMT_E_INVALID_ARG; }
static int pmt_permute_impl(pmt_ctx* c, const uint64_t* in, size_t n, uint64_t* out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!in || !out) return fail(c, PMT_E_INVALID_ARG, "pmt_permute: null pointer");
  void *a, *b;
  if (int rc = arena_get(c, 0, n * 96, &a)) return rc;
  if (int rc = arena_get(c, 1, n * 96, &b)) return rc;
  H2D(c, a, in, n * 96);
  if (int rc = pmt_permute_dev(c, (uint64_t*)a, n, (uint64_t*)b)) return rc;
  D2H(c, out, b, n * 96);
  FINISH(c);
  return PMT_OK;
}

static int pmt_hash_two_to_one_impl(pmt_ctx* c, const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out);
int pmt_hash_two_to_one(pmt_ctx* c, const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out) { return c ? drained(c, pmt_hash_two_to_one_impl(c, l, r, n, out)) : PMT_E_INVALID_ARG; }
static int pmt_hash_two_to_one_impl(pmt_ctx* c, const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!l || !r || !out) return fail(c, PMT_E_INVALID_ARG, "pmt_hash_two_to_one: null pointer");
  void *a, *b;
  if (int rc = arena_get(c, 0, n * 64, &a)) return rc;
  if (int rc = arena_get(c, 1, n * 32, &b)) return rc;
  uint64_t* dl = (uint64_t*)a; uint64_t* dr = dl + 4 * n;
  H2D(c, dl, l, n * 32);
  H2D(c, dr, r, n * 32);
  if (int rc = pmt_hash_two_to_one_dev(c, dl, dr, n, (uint64_t*)b)) return rc;
  D2H(c, out, b, n * 32);
  FINISH(c);
  return PMT_OK;
}

static int hash_rows_host(pmt_ctx* c, const uint64_t* rows, size_t n, size_t w, int noop, uint64_t* out) {
  if (int rc = bind(c)) return rc;
  if (n == 0) return PMT_OK;
  if (!out || (!rows && w)) return fail(c, PMT_E_INVALID_ARG, "pmt_hash_rows: null pointer");
  void *a, *b;
  if (int rc = arena_get(c, 0, n * w * 8 + 8, &a)) return rc;
  if (int rc = arena_get(c, 1, n * 32, &b)) return rc;
  if (w) H2D(c, a, rows, n * w * 8);
  if (int rc = pmt_hash_rows_dev(c, (uint64_t*)a, n, w, noop, (uint64_t*)b)) return rc;
  D2H(c, out, b, n * 32);
  FINISH(c);
  return PMT_OK;
}
int pmt_hash_or_noop(pmt_ctx* c, const uint64_t* rows, size_t n, size_t w, uint64_t* out) {
  return c ? drained(c, hash_rows_host(c, rows, n, w, 1, out)) : PMT_E_INVALID_ARG;
}
int pmt_hash_no_pad(pmt_ctx* c, const uint64_t* rows, size_t n, size_t w, uint64_t* out) {
  return c ? drained(c, hash_rows_host(c, rows, n, w, 0, out)) : PMT_E_INVALID_ARG;
}

static int pmt_simple_tree_build_impl(pmt_ctx* c, const uint64_t* leaves, size_t n, uint64_t* levels_out, uint64_t* root_out);
int pmt_simple_tree_build(pmt_ctx* c, const uint64_t* leaves, size_t n, uint64_t* levels_out, uint64_t* root_out) { return c ? drained(c, pmt_simple_tree_build_impl(c, leaves, n, levels_out, root_out)) : PMT_E_INVALID_ARG; }
static int pmt_simple_tree_build_impl(pmt_ctx* c, const uint64_t* leaves, size_t n, uint64_t* levels_out, uint64_t* root_out) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "simple tree: %zu leaves is not a power of two (log2_strict, simple_merkle_tree.rs:30)", n);
  if (lg < 1) return fail(c, PMT_E_INVALID_ARG, "simple tree: needs at least 2 leaves (simple_merkle_tree.rs:38)");
  if (!leaves || !levels_out || !root_out) return fail(c, PMT_E_INVALID_ARG, "simple tree: null pointer");
  void *a, *b;
  if (int rc = arena_get(c, 0, n * 8, &a)) return rc;
  if (int rc = arena_get(c, 1, (2 * n - 1) * 32, &b)) return rc;
  uint64_t* d_levels = (uint64_t*)b; uint64_t* d_root = d_levels + 4 * (2 * n - 2);
  H2D(c, a, leaves, n * 8);
  if (int rc = pmt_simple_tree_build_dev(c, (uint64_t*)a, n, d_levels, d_root)) return rc;
  D2H(c, levels_out, d_levels, (2 * n - 2) * 32);
  D2H(c, root_out, d_root, 32);
  FINISH(c);
  return PMT_OK;
}

// host index of node (l, k) in upstream's `digests` (same formula as Plonky2::at on the device)
static size_t plonky2_index(int L, int l, size_t k) {
  const int per = L - l;
  const size_t c = k >> per, kk = k & (((size_t)1 << per) - 1);
  return c * (((size_t)2 << L) - 2) + 2 * (((kk >> 1) << (l + 1)) + ((size_t)1 << l) - 1) + (kk & 1);
}

// Host-buffer MerkleTree::new.  Large trees are built as a 3-stage pipeline over 16 leaf chunks: chunk i's leaves go up
// on the copy-in stream while chunk i-1's subtree is hashed on the compute stream and chunk i-2's digests (one
// contiguous slice of upstream's layout) go down on the copy-out stream, so PCIe traffic in both directions hides
// behind the permutations.  The few digests above the chunk roots are finished and downloaded at the end.
static int pmt_merkle_tree_build_impl(pmt_ctx* c, const uint64_t* leaves, size_t n, size_t w, uint32_t cap_height,
                          uint64_t* digests_out, uint64_t* cap_out);
int pmt_merkle_tree_build(pmt_ctx* c, const uint64_t* leaves, size_t n, size_t w, uint32_t cap_height,
                          uint64_t* digests_out, uint64_t* cap_out) {
  return c ? drained(c, pmt_merkle_tree_build_impl(c, leaves, n, w, cap_height, digests_out, cap_out)) : PMT_E_INVALID_ARG;
}
static int pmt_merkle_tree_build_impl(pmt_ctx* c, const uint64_t* leaves, size_t n, size_t w, uint32_t cap_height,
                          uint64_t* digests_out, uint64_t* cap_out) {
  if (int rc = bind(c)) return rc;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "MerkleTree::new: %zu leaves is not a power of two (log2_strict)", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "MerkleTree::new: cap_height=%u should be at most log2(leaves.len())=%d", cap_height, lg);
  if (w == 0) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: zero-width leaves");
  const size_t n_cap = (size_t)1 << cap_height, n_dig = 2 * (n - n_cap);
  if (!leaves || !cap_out || (!digests_out && n_dig)) return fail(c, PMT_E_INVALID_ARG, "MerkleTree::new: null pointer");
  void *a, *b;
  if (int rc = arena_get(c, 0, n * w * 8, &a)) return rc;
  if (int rc = arena_get(c, 1, (n_dig + n_cap) * 32, &b)) return rc;
  uint64_t* d_leaves = (uint64_t*)a;
  uint64_t* d_dig = (uint64_t*)b; uint64_t* d_cap = d_dig + 4 * n_dig;
  const int L = lg - (int)cap_height;
  // log2(leaves per chunk): 16 chunks.  The D2H of every digest (1 GiB at 2^24 leaves, 18.9 ms) is the bound; measured
  // with tools/e2e_bench.py: 8 / 16 / 32 / 64 chunks = 25.0 / 22.1 / 25.9 / 34.3 ms (every chunk adds one latency-bound
  // subtree tail of ~13 tiny launches); alternating chunks between two compute streams did not help (23.1 ms at 16).
  int log2_chunks = 4;
  if (const char* e = getenv("PMT_PIPELINE_LOG2_CHUNKS")) log2_chunks = atoi(e);   // tuning knob (tools/e2e_bench.py)
  if (log2_chunks < 1) log2_chunks = 1;
  if (log2_chunks > 10) log2_chunks = 10;
  int cb = lg - log2_chunks;
  if (cb > L) cb = L;
  if (cb < 12 || n * w * 8 < ((size_t)8 << 20)) {   // small: one shot
    H2D(c, d_leaves, leaves, n * w * 8);
    if (int rc = pmt_merkle_tree_build_dev(c, d_leaves, n, w, cap_height, d_dig, d_cap)) return rc;
    if (n_dig) D2H(c, digests_out, d_dig, n_dig * 32);
    D2H(c, cap_out, d_cap, n_cap * 32);
    FINISH(c);
    return PMT_OK;
  }
  const size_t chunks = n >> cb, chunk = (size_t)1 << cb;
  if (int rc = pipeline_streams(c, chunks)) return rc;
  // PAGEABLE caller buffers (a plain Vec): cudaMemcpyAsync would stage every copy through the driver's bounce buffer on the
  // calling thread and the three stages would run one after the other (measured: 114.6 ms against 21.6 ms from pinned
  // buffers for 2^24 x 4 leaves).  Instead the chunks go through the ctx's own page-locked slots, filled / drained by the
  // copy threads while the DMA engines and the SMs work on the neighbouring chunks.
  const size_t in_bytes = chunk * w * 8, out_len = cb >= 1 ? 2 * chunk - 2 : 0;
  const bool in_staged = is_pageable(leaves), out_staged = out_len && is_pageable(digests_out);
  void *si[2] = {nullptr, nullptr}, *so[2] = {nullptr, nullptr};
  if (in_staged) for (int k = 0; k < 2; k++) if (int rc = stage_get(c, k, in_bytes, &si[k])) return rc;
  if (out_staged) for (int k = 0; k < 2; k++) if (int rc = stage_get(c, 2 + k, out_len * 32, &so[k])) return rc;
  cudaEvent_t* ev_out = c->ev.data() + 2 * chunks + 2;          // D2H of chunk i into its slot is complete
  Plonky2 lay{d_dig, d_cap, L};
  for (size_t i = 0; i < chunks; i++) {
    const uint64_t* src = leaves + i * chunk * w;
    if (in_staged) {
      if (i >= 2) CU(c, cudaEventSynchronize(c->ev[2 * (i - 2)]));     // the slot's previous chunk has left for the device
      c->pool->copy(si[i & 1], src, in_bytes);
      src = (const uint64_t*)si[i & 1];
    }
    CU(c, cudaMemcpyAsync(d_leaves + i * chunk * w, src, in_bytes, cudaMemcpyHostToDevice, c->copy_in));
    CU(c, cudaEventRecord(c->ev[2 * i], c->copy_in));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev[2 * i], 0));
    if (int rc = launch_leaves(c, lay, d_leaves + i * chunk * w, w, i * chunk, chunk)) return rc;
    if (int rc = launch_level_span(c, lay, 1, cb, i * chunk, (i + 1) * chunk)) return rc;
    CU(c, cudaEventRecord(c->ev[2 * i + 1], c->stream));
    CU(c, cudaStreamWaitEvent(c->copy_out, c->ev[2 * i + 1], 0));
    if (out_len) {
      const size_t start = plonky2_index(L, 0, i * chunk);      // the chunk's subtree is one contiguous slice of `digests`
      if (out_staged) {      // slot i & 1 is free: the host drained chunk i - 2 out of it during iteration i - 1
        CU(c, cudaMemcpyAsync(so[i & 1], d_dig + 4 * start, out_len * 32, cudaMemcpyDeviceToHost, c->copy_out));
        CU(c, cudaEventRecord(ev_out[i], c->copy_out));
        if (i >= 1) {
          CU(c, cudaEventSynchronize(ev_out[i - 1]));
          c->pool->copy(digests_out + 4 * plonky2_index(L, 0, (i - 1) * chunk), so[(i - 1) & 1], out_len * 32);
        }
      } else {
        CU(c, cudaMemcpyAsync(digests_out + 4 * start, d_dig + 4 * start, out_len * 32, cudaMemcpyDeviceToHost, c->copy_out));
      }
    }
  }
  if (out_staged) {
    CU(c, cudaEventSynchronize(ev_out[chunks - 1]));
    c->pool->copy(digests_out + 4 * plonky2_index(L, 0, (chunks - 1) * chunk), so[(chunks - 1) & 1], out_len * 32);
  }
  if (cb < L)
    if (int rc = run_levels(c, lay, cb + 1, L, n >> (cb + 1))) return rc;
  // digests of levels cb .. L-1 (the chunk roots and everything above them, below the cap): sibling pairs are adjacent
  for (int l = cb; l < L; l++)
    for (size_t k = 0; k < (n >> l); k += 2) {
      const size_t pos = plonky2_index(L, l, k);
      D2H(c, digests_out + 4 * pos, d_dig + 4 * pos, 64);
    }
  D2H(c, cap_out, d_cap, n_cap * 32);
  CU(c, cudaStreamSynchronize(c->copy_out));
  FINISH(c);
  return PMT_OK;
}

// Single-process multi-GPU MerkleTree::new for a compiled host (the reference is one process): ctx r owns
// leaves [r n/G, (r+1) n/G) = one subtree and runs the pipelined host-buffer build above on its own device from its own
// host thread, straight into its contiguous slice of upstream's `digests`.  cap_height >= log2 G: the slices and cap
// entries tile the outputs, nothing else to do.  Otherwise the G subtree roots (32 B each) come back through the host,
// ctx 0 finishes the log2 G - h levels above them in one cooperative launch and the 2G - 2^(h+1) digests are placed at
// their closed-form positions between the slices.  Nothing but roots ever leaves a device.
int pmt_merkle_tree_build_multi(pmt_ctx* const* ctxs, size_t n_ctx, const uint64_t* leaves, size_t n, size_t w,
                                uint32_t cap_height, uint64_t* digests_out, uint64_t* cap_out) {
  if (!ctxs || n_ctx == 0 || !ctxs[0]) return PMT_E_INVALID_ARG;
  pmt_ctx* c0 = ctxs[0];
  for (size_t i = 0; i < n_ctx; i++) {
    if (!ctxs[i]) return fail(c0, PMT_E_INVALID_ARG, "multi build: null ctx %zu", i);
    for (size_t j = 0; j < i; j++)
      if (ctxs[j] == ctxs[i]) return fail(c0, PMT_E_INVALID_ARG, "multi build: ctx %zu given twice (a ctx is not thread-safe)", i);
  }
  const int g = log2_strict(n_ctx), lg = log2_strict(n);
  if (g < 0) return fail(c0, PMT_E_NOT_POW2, "multi build: %zu contexts is not a power of two", n_ctx);
  if (lg < 0) return fail(c0, PMT_E_NOT_POW2, "MerkleTree::new: %zu leaves is not a power of two (log2_strict)", n);
  if ((int)cap_height > lg) return fail(c0, PMT_E_RANGE, "MerkleTree::new: cap_height=%u should be at most log2(leaves.len())=%d", cap_height, lg);
  if (w == 0) return fail(c0, PMT_E_INVALID_ARG, "MerkleTree::new: zero-width leaves");
  if (g > lg) return fail(c0, PMT_E_RANGE, "multi build: more contexts (%zu) than leaves (%zu)", n_ctx, n);
  const size_t n_cap = (size_t)1 << cap_height, n_dig = 2 * (n - n_cap);
  if (!leaves || !cap_out || (!digests_out && n_dig)) return fail(c0, PMT_E_INVALID_ARG, "MerkleTree::new: null pointer");
  if (n_ctx == 1) return pmt_merkle_tree_build(c0, leaves, n, w, cap_height, digests_out, cap_out);
  try {
  const size_t per = n >> g;
  const int L = lg - (int)cap_height, Lr = lg - g;
  const bool gather = (int)cap_height < g;
  std::vector<uint64_t> roots(gather ? 4 * n_ctx : 0);
  std::vector<int> rcs(n_ctx, PMT_OK);
  auto work = [&](size_t r) {
    const uint64_t* my = leaves + r * per * w;
    if (!gather) {   // 2^(h-g) cap subtrees per ctx
      const uint32_t hl = cap_height - (uint32_t)g;
      const size_t cap_l = (size_t)1 << hl, dig_l = 2 * (per - cap_l);
      rcs[r] = pmt_merkle_tree_build(ctxs[r], my, per, w, hl, digests_out ? digests_out + 4 * r * dig_l : nullptr, cap_out + 4 * r * cap_l);
    } else {         // one subtree of height Lr inside a cap subtree of height L: 2 per - 2 contiguous digests
      rcs[r] = pmt_merkle_tree_build(ctxs[r], my, per, w, 0, digests_out + 4 * plonky2_index(L, 0, r * per), roots.data() + 4 * r);
    }
  };
  run_per_ctx(ctxs, rcs, work);
  for (size_t r = 0; r < n_ctx; r++)
    if (rcs[r] != PMT_OK) {
      char msg[sizeof ctxs[r]->err];
      memcpy(msg, ctxs[r]->err, sizeof msg);
      msg[sizeof msg - 1] = 0;
      return fail(c0, rcs[r], "multi build: ctx %zu (device %d): %.400s", r, ctxs[r]->device, msg);
    }
  if (!gather) return PMT_OK;
  if (int rc = bind(c0)) return rc;
  void* t = nullptr;
  if (int rc = arena_get(c0, 2, 2 * n_ctx * 32, &t)) return rc;
  uint64_t* d_roots = (uint64_t*)t; uint64_t* d_top = d_roots + 4 * n_ctx;
  std::vector<uint64_t> top(4 * (n_ctx - n_cap));
  auto finish = [&]() -> int {
    H2D(c0, d_roots, roots.data(), n_ctx * 32);
    if (int rc = pmt_top_levels_dev(c0, d_roots, n_ctx, cap_height, d_top)) return rc;
    D2H(c0, top.data(), d_top, (n_ctx - n_cap) * 32);
    FINISH(c0);
    return PMT_OK;
  };
  if (int rc = drained(c0, finish())) return rc;
  for (size_t k = 0; k < n_ctx; k++) memcpy(digests_out + 4 * plonky2_index(L, Lr, k), roots.data() + 4 * k, 32);
  const uint64_t* lvl = top.data();
  int level = Lr + 1;
  for (size_t cnt = n_ctx / 2; cnt >= n_cap; cnt >>= 1, level++) {
    for (size_t k = 0; k < cnt; k++)
      memcpy(level < L ? digests_out + 4 * plonky2_index(L, level, k) : cap_out + 4 * k, lvl + 4 * k, 32);
    lvl += 4 * cnt;
    if (cnt == 1) break;
  }
  return PMT_OK;
  } catch (const std::bad_alloc&) {
    return fail(c0, PMT_E_OOM, "multi build: out of host memory");
  }
}

// Host-buffer batch append.  The device holds only the old PEAKS (a new node's left child is either new or an old peak)
// and the new elements (MmrAppend: O(m + log n) device memory, whatever the size of the MMR appended to), and large batches
// run as a 3-stage pipeline over aligned power-of-two chunks: appending chunk i to the MMR of the leaves before it creates
// exactly the elements [mmr_size(n0 + i*chunk), mmr_size(n0 + (i+1)*chunk)) -- the chunk's perfect sub-mountain followed by
// every ancestor it completes -- ONE contiguous slice of the post-order array, so it is downloaded while the next chunk is
// hashed and the one after that is uploaded.
static int mmr_extend_host(pmt_ctx* c, uint64_t* elements, size_t n0, const uint64_t* new_leaves, size_t m) {
  const size_t s0 = pmt_mmr_size(n0), s1 = pmt_mmr_size(n0 + m);
  const uint32_t n_peaks = (uint32_t)__builtin_popcountll((unsigned long long)n0);
  void *a, *b;
  if (int rc = arena_get(c, 0, m * 8, &a)) return rc;
  if (int rc = arena_get(c, 1, (n_peaks + (s1 - s0)) * 32, &b)) return rc;
  uint64_t* d_leaves = (uint64_t*)a;
  const MmrAppend lay{(uint64_t*)b, n0, s0, n_peaks};
  uint64_t* d_new = lay.buf + 4 * (size_t)n_peaks;       // element at post-order position p >= s0: d_new + 4 (p - s0)
  // old peaks: one per set bit of n0, at the last position of its mountain (get_peaks, :179-200), largest first
  {
    size_t base = 0;
    uint32_t slot = 0;
    for (int bit = 63; bit >= 0; bit--)
      if ((n0 >> bit) & 1) {
        base += (size_t)1 << bit;
        H2D(c, lay.buf + 4 * (size_t)slot++, elements + 4 * (pmt_mmr_size(base) - 1), 32);
      }
  }
  const int lgm = log2_floor(m);
  const size_t chunk = lgm >= 20 ? (size_t)1 << (lgm - 4) : m;   // 16 .. 31 chunks for big batches, else one shot
  if (chunk >= m) {
    H2D(c, d_leaves, new_leaves, m * 8);
    if (int rc = mmr_extend_plan(c, lay, n0, d_leaves, m)) return rc;
    D2H(c, elements + 4 * s0, d_new, (s1 - s0) * 32);
    FINISH(c);
    return PMT_OK;
  }
  // first piece: up to the next multiple of `chunk` so that every later piece is an aligned perfect sub-mountain
  const size_t chunks = (m + chunk - 1) / chunk + 1;
  if (int rc = pipeline_streams(c, chunks)) return rc;
  // pageable caller buffers go through the ctx's page-locked slots and copy threads, as in pmt_merkle_tree_build
  const size_t out_slot_bytes = (2 * chunk + 64) * 32;       // a piece of <= chunk leaves creates < 2 chunk + 64 elements
  const bool in_staged = is_pageable(new_leaves), out_staged = is_pageable(elements);
  void *si[2] = {nullptr, nullptr}, *so[2] = {nullptr, nullptr};
  if (in_staged) for (int k = 0; k < 2; k++) if (int rc = stage_get(c, k, chunk * 8, &si[k])) return rc;
  if (out_staged) for (int k = 0; k < 2; k++) if (int rc = stage_get(c, 2 + k, out_slot_bytes, &so[k])) return rc;
  cudaEvent_t* ev_out = c->ev.data() + 2 * chunks + 2;
  size_t done = 0, i = 0, prev_p0 = 0, prev_bytes = 0;
  while (done < m) {
    size_t len = chunk - ((n0 + done) & (chunk - 1));     // to the next chunk boundary
    if (len > m - done) len = m - done;
    const uint64_t* src = new_leaves + done;
    if (in_staged) {
      if (i >= 2) CU(c, cudaEventSynchronize(c->ev[2 * (i - 2)]));
      c->pool->copy(si[i & 1], src, len * 8);
      src = (const uint64_t*)si[i & 1];
    }
    CU(c, cudaMemcpyAsync(d_leaves + done, src, len * 8, cudaMemcpyHostToDevice, c->copy_in));
    CU(c, cudaEventRecord(c->ev[2 * i], c->copy_in));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev[2 * i], 0));
    if (int rc = mmr_extend_plan(c, lay, n0 + done, d_leaves + done, len)) return rc;
    CU(c, cudaEventRecord(c->ev[2 * i + 1], c->stream));
    CU(c, cudaStreamWaitEvent(c->copy_out, c->ev[2 * i + 1], 0));
    const size_t p0 = pmt_mmr_size(n0 + done), p1 = pmt_mmr_size(n0 + done + len);
    if (out_staged) {
      CU(c, cudaMemcpyAsync(so[i & 1], d_new + 4 * (p0 - s0), (p1 - p0) * 32, cudaMemcpyDeviceToHost, c->copy_out));
      CU(c, cudaEventRecord(ev_out[i], c->copy_out));
      if (i >= 1) {      // drain the previous piece while this one is hashed
        CU(c, cudaEventSynchronize(ev_out[i - 1]));
        c->pool->copy(elements + 4 * prev_p0, so[(i - 1) & 1], prev_bytes);
      }
      prev_p0 = p0; prev_bytes = (p1 - p0) * 32;
    } else {
      CU(c, cudaMemcpyAsync(elements + 4 * p0, d_new + 4 * (p0 - s0), (p1 - p0) * 32, cudaMemcpyDeviceToHost, c->copy_out));
    }
    done += len;
    i++;
  }
  if (out_staged && i >= 1) {
    CU(c, cudaEventSynchronize(ev_out[i - 1]));
    c->pool->copy(elements + 4 * prev_p0, so[(i - 1) & 1], prev_bytes);
  }
  CU(c, cudaStreamSynchronize(c->copy_out));
  FINISH(c);
  return PMT_OK;
}

int pmt_mmr_extend(pmt_ctx* c, uint64_t* elements, size_t n0, const uint64_t* new_leaves, size_t m) {
  if (int rc = bind(c)) return rc;
  if (m == 0) return PMT_OK;
  if (!elements || !new_leaves) return fail(c, PMT_E_INVALID_ARG, "mmr extend: null pointer");
  if (n0 + m > ((size_t)1 << 30)) return fail(c, PMT_E_RANGE, "mmr extend: more than 2^30 leaves (get_mmr_index is i32, merkle_mountain_ranges.rs:264)");
  return drained(c, mmr_extend_host(c, elements, n0, new_leaves, m));
}

// ---- single-process multi-GPU batch append ---------------------------------------------------------------------------------
// post-order position of the node at height l covering leaves [k 2^l, (k + 1) 2^l)
static size_t mmr_pos(int l, size_t k) {
  const size_t last = ((k + 1) << l) - 1;
  return 2 * last - (size_t)__builtin_popcountll((unsigned long long)last) + (size_t)l;
}

static int check_ctxs(pmt_ctx* const* ctxs, size_t n_ctx, const char* who) {
  if (!ctxs || n_ctx == 0 || !ctxs[0]) return PMT_E_INVALID_ARG;
  for (size_t i = 0; i < n_ctx; i++) {
    if (!ctxs[i]) return fail(ctxs[0], PMT_E_INVALID_ARG, "%s: null ctx %zu", who, i);
    for (size_t j = 0; j < i; j++)
      if (ctxs[j] == ctxs[i]) return fail(ctxs[0], PMT_E_INVALID_ARG, "%s: ctx %zu given twice (a ctx is not thread-safe)", who, i);
  }
  return PMT_OK;
}

// The plan of pmt_mmr_extend_multi, pure index math (no device): the appended leaves [n0, n0 + m) are cut into a head up to
// the first multiple of B = 2^b, aligned blocks of B leaves, and a tail; b = floor(log2(m / n_ctx)) - 2, at least 12, so that
// the contexts get 4 .. 8 blocks each.  Returns 1 and the plan if there is at least one block and more than one context,
// 0 if one context should do the whole append.
int pmt_mmr_multi_plan(size_t n0, size_t m, size_t n_ctx, uint32_t* log2_block, size_t* first_aligned, size_t* last_aligned) {
  if (n_ctx <= 1 || m / n_ctx < 4096) return 0;
  int b = 0;
  while (((size_t)2 << b) <= m / n_ctx) b++;   // floor(log2(m / n_ctx)) >= 12
  b = b - 2 < 12 ? 12 : b - 2;
  const size_t A = ((n0 + ((size_t)1 << b) - 1) >> b) << b, Z = ((n0 + m) >> b) << b;
  if (Z <= A) return 0;
  if (log2_block) *log2_block = (uint32_t)b;
  if (first_aligned) *first_aligned = A;
  if (last_aligned) *last_aligned = Z;
  return 1;
}

// Batch append over several GPUs from one process.  An aligned block of B leaves is a perfect sub-mountain whose 2B - 1
// elements are one contiguous slice of the post-order array at mmr_size(start), identical to a fresh MMR of those leaves:
// context j mod G builds block j from empty straight into its slice (pmt_mmr_extend, pipelined for big blocks), context 0
// also appends the head.  The nodes above the block roots are the MMR over the block roots ("coarse" MMR: node (l, k) of
// it is node (l + b, k) of the fine one): its old peaks (= the fine peaks at the first block boundary) and the block roots
// go up to context 0, one launch_level_span computes the rest, and the few new nodes come back to their fine positions.
// The tail (< B leaves) is appended last.  Output identical to pmt_mmr_extend.
int pmt_mmr_extend_multi(pmt_ctx* const* ctxs, size_t n_ctx, uint64_t* elements, size_t n0, const uint64_t* new_leaves, size_t m) {
  if (int rc = check_ctxs(ctxs, n_ctx, "mmr extend multi")) return rc;
  pmt_ctx* c0 = ctxs[0];
  if (m == 0) return PMT_OK;
  if (!elements || !new_leaves) return fail(c0, PMT_E_INVALID_ARG, "mmr extend: null pointer");
  if (n0 + m > ((size_t)1 << 30)) return fail(c0, PMT_E_RANGE, "mmr extend: more than 2^30 leaves (get_mmr_index is i32, merkle_mountain_ranges.rs:264)");
  uint32_t b = 0;
  size_t A = 0, Z = 0;
  if (!pmt_mmr_multi_plan(n0, m, n_ctx, &b, &A, &Z)) return pmt_mmr_extend(c0, elements, n0, new_leaves, m);
  try {
  const size_t B = (size_t)1 << b, C = (Z - A) >> b;
  std::vector<int> rcs(n_ctx, PMT_OK);
  auto work = [&](size_t r) {
    if (r == 0 && A > n0) {
      rcs[0] = pmt_mmr_extend(ctxs[0], elements, n0, new_leaves, A - n0);
      if (rcs[0] != PMT_OK) return;
    }
    for (size_t j = r; j < C; j += n_ctx) {
      const size_t start = A + j * B;
      rcs[r] = pmt_mmr_extend(ctxs[r], elements + 4 * pmt_mmr_size(start), 0, new_leaves + (start - n0), B);
      if (rcs[r] != PMT_OK) return;
    }
  };
  run_per_ctx(ctxs, rcs, work);
  for (size_t r = 0; r < n_ctx; r++)
    if (rcs[r] != PMT_OK) {
      char msg[sizeof ctxs[r]->err];
      memcpy(msg, ctxs[r]->err, sizeof msg);
      msg[sizeof msg - 1] = 0;
      return fail(c0, rcs[r], "mmr extend multi: ctx %zu (device %d): %.400s", r, ctxs[r]->device, msg);
    }
  // the coarse MMR over the block roots, on context 0: its old peaks, the block roots and what the append creates
  if (int rc = bind(c0)) return rc;
  const size_t s0 = A >> b, s1 = s0 + C;
  const size_t cs0 = pmt_mmr_size(s0), cs1 = pmt_mmr_size(s1);
  const uint32_t n_peaks = (uint32_t)__builtin_popcountll((unsigned long long)s0);
  void* t = nullptr;
  if (int rc = arena_get(c0, 1, (n_peaks + (cs1 - cs0)) * 32, &t)) return rc;
  const MmrAppend sup{(uint64_t*)t, s0, cs0, n_peaks};
  uint64_t* d_new = sup.buf + 4 * (size_t)n_peaks;     // coarse element at post-order position p >= cs0: d_new + 4 (p - cs0)
  auto coarse = [&]() -> int {
    size_t base = 0;
    uint32_t slot = 0;
    for (int bit = 63; bit >= 0; bit--)
      if ((s0 >> bit) & 1) {
        base += (size_t)1 << bit;
        H2D(c0, sup.buf + 4 * (size_t)slot++, elements + 4 * (pmt_mmr_size(base << b) - 1), 32);
      }
    for (size_t s = s0; s < s1; s++) H2D(c0, d_new + 4 * (mmr_pos(0, s) - cs0), elements + 4 * mmr_pos((int)b, s), 32);
    if (int rc = launch_level_span(c0, sup, 1, 40, s0, s1)) return rc;
    for (int l = 1; l < 40; l++) {
      const size_t k0 = s0 >> l, k1 = s1 >> l;
      if (k1 == 0) break;
      for (size_t k = k0; k < k1; k++) D2H(c0, elements + 4 * mmr_pos(l + (int)b, k), d_new + 4 * (mmr_pos(l, k) - cs0), 32);
    }
    FINISH(c0);
    return PMT_OK;
  };
  if (int rc = drained(c0, coarse())) return rc;
  if (n0 + m > Z) return pmt_mmr_extend(c0, elements, Z, new_leaves + (Z - n0), n0 + m - Z);
  return PMT_OK;
  } catch (const std::bad_alloc&) {
    return fail(c0, PMT_E_OOM, "mmr extend multi: out of host memory");
  }
}

// ---- subtree-sharded MMR, one process per GPU, NCCL inside the library ------------------------------------------------------
// The plan: every set bit 2^b >= G of n is one mountain cut into G equal perfect sub-mountains of m_i = 2^b / G
// leaves, one per rank ("round" i); the bits below G (t < G leaves) are the tail, kept by the last rank.  Pure index math.
int pmt_mmr_shard_plan(size_t n_total, size_t world, uint32_t* n_rounds, size_t* m_out, size_t* tail) {
  if (world == 0 || (world & (world - 1))) return PMT_E_NOT_POW2;
  uint32_t k = 0;
  size_t rest = n_total;
  while (rest >= world) {
    const size_t m = (size_t)1 << log2_floor(rest / world);
    if (m_out) m_out[k] = m;
    k++;
    rest -= world * m;
  }
  if (n_rounds) *n_rounds = k;
  if (tail) *tail = rest;
  return PMT_OK;
}

// Collective.  This rank's leaves: its m_i leaves of every round i, in round order, then -- last rank only -- the t tail
// leaves.  ONE batch append builds all of the rank's sub-mountains (they are exactly the mountains of its local MMR),
// k_mmr_peaks writes their roots and the tail's peaks into this rank's row of d_gathered, ONE ncclAllGather exchanges the rows
// (<= 62 digests per rank), one batched launch finishes the log2 G levels above every round's G roots, and the global peaks
// (one per round, then the tail's) are collected.  One stream, no host synchronisation.
//   d_local_elements  pmt_mmr_size(sum m_i) digests: the rank's local MMR; sub-mountain i is the slice of the global
//                     `elements` that starts at pmt_mmr_size(S_i + rank m_i), S_i = world (m_0 + .. + m_(i-1))
//   d_tail_elements   pmt_mmr_size(t) digests (last rank; ignored elsewhere)
//   d_gathered        world x slots digests, slots = rounds + popcount(t): row r = [rank r's sub-mountain roots | tail peaks]
//   d_tops            rounds x (world - 1) digests: the levels above each round's roots, level-major; last one = the round's peak
//   d_peaks           slots digests: get_peaks() of the whole MMR, on every rank
int pmt_mmr_build_sharded_dev(pmt_ctx* c, const uint64_t* d_local_leaves, size_t n_total, uint64_t* d_local_elements,
                              uint64_t* d_tail_elements, uint64_t* d_gathered, uint64_t* d_tops, uint64_t* d_peaks) {
  if (int rc = bind(c)) return rc;
  if (!c->comm) return fail(c, PMT_E_INVALID_ARG, "sharded mmr: pmt_comm_init has not been called on this ctx");
  const size_t G = (size_t)c->comm_world, r = (size_t)c->comm_rank;
  if (n_total == 0 || n_total > ((size_t)1 << 30)) return fail(c, PMT_E_RANGE, "sharded mmr: bad leaf count (get_mmr_index is i32, merkle_mountain_ranges.rs:264)");
  uint32_t k = 0;
  size_t ms[64], t = 0;
  pmt_mmr_shard_plan(n_total, G, &k, ms, &t);
  size_t n_main = 0;
  for (uint32_t i = 0; i < k; i++) n_main += ms[i];
  const uint32_t tp = (uint32_t)__builtin_popcountll((unsigned long long)t);
  const size_t slots = (size_t)k + tp;
  const bool last = r + 1 == G;
  if (!d_gathered || !d_peaks || (n_main && (!d_local_leaves || !d_local_elements)) || (k && G > 1 && !d_tops) || (t && last && !d_tail_elements))
    return fail(c, PMT_E_INVALID_ARG, "sharded mmr: null pointer");
  uint64_t* mine = d_gathered + 4 * slots * r;
  CU(c, cudaMemsetAsync(mine, 0, slots * 32, c->stream));
  if (n_main) {
    if (int rc = mmr_extend_plan(c, Mmr{d_local_elements}, 0, d_local_leaves, n_main)) return rc;
    k_mmr_peaks<<<1, 64, 0, c->stream>>>(Mmr{d_local_elements}, n_main, mine);      // popcount(n_main) = k peaks: the sub-mountain roots
    CHECK_LAUNCH(c);
  }
  if (t && last) {
    if (int rc = mmr_extend_plan(c, Mmr{d_tail_elements}, 0, d_local_leaves + n_main, t)) return rc;
    k_mmr_peaks<<<1, 64, 0, c->stream>>>(Mmr{d_tail_elements}, t, mine + 4 * k);
    CHECK_LAUNCH(c);
  }
  if (G > 1 && c->mail_mode == MAIL_COMM && slots <= MAIL_DIGESTS)      // exchange + every round's finish + the peaks: ONE launch
    return launch_exchange(c, mine, slots, d_gathered, k, d_tops, log2_strict(G), d_peaks);
  if (G > 1) NC(c, nccl_api()->AllGather(mine, d_gathered, 4 * slots, ncclUint64, c->comm, c->stream));
  if (k && G > 1) {      // round i: roots = column i of the gathered matrix (slots digests apart), finished in set i
    TopRoots lay{d_gathered, d_tops, G, 1, G - 1, slots};
    if (G / 2 <= (size_t)COOP_NODES) {
      const int done = launch_coop(c, lay, 1, log2_strict(G), 0, G / 2, k);
      if (done < 0) return done;
    } else {
      for (uint32_t i = 0; i < k; i++)
        if (int rc = launch_level_span(c, lay.for_set_host(i), 1, log2_strict(G), 0, G)) return rc;
    }
    CU(c, cudaMemcpy2DAsync(d_peaks, 32, d_tops + 4 * (G - 2), (G - 1) * 32, 32, k, cudaMemcpyDeviceToDevice, c->stream));
  } else if (k) {
    CU(c, cudaMemcpyAsync(d_peaks, d_gathered, (size_t)k * 32, cudaMemcpyDeviceToDevice, c->stream));
  }
  if (tp) CU(c, cudaMemcpyAsync(d_peaks + 4 * k, d_gathered + 4 * (slots * (G - 1) + k), (size_t)tp * 32, cudaMemcpyDeviceToDevice, c->stream));
  return PMT_OK;
}

// positions of the peaks in the post-order array, largest mountain first (get_peaks, :179-200)
static uint32_t peak_positions(size_t n_leaves, size_t* pos) {
  uint32_t k = 0;
  size_t base = 0;
  for (int bit = 63; bit >= 0; bit--)
    if ((n_leaves >> bit) & 1) {
      base += (size_t)1 << bit;
      pos[k++] = pmt_mmr_size(base) - 1;
    }
  return k;
}

// get_peaks on a HOST array is a gather of <= 32 digests: done on the host (nothing to compute); canonicalised like
// every output of the library.
int pmt_mmr_peaks(pmt_ctx* c, const uint64_t* elements, size_t n_leaves, uint64_t* peaks_out, uint32_t* n_peaks_out) {
  if (!c) return PMT_E_INVALID_ARG;
  if (n_leaves == 0) { if (n_peaks_out) *n_peaks_out = 0; return PMT_OK; }
  if (!elements || !peaks_out) return fail(c, PMT_E_INVALID_ARG, "mmr peaks: null pointer");
  if (n_leaves >> 32) return fail(c, PMT_E_RANGE, "mmr peaks: size does not fit u32 (merkle_mountain_ranges.rs:184)");
  size_t pos[64];
  const uint32_t k = peak_positions(n_leaves, pos);
  for (uint32_t i = 0; i < k; i++)
    for (int e = 0; e < 4; e++) {
      const uint64_t x = elements[4 * pos[i] + e];
      peaks_out[4 * i + e] = x >= PMT_P ? x - PMT_P : x;
    }
  if (n_peaks_out) *n_peaks_out = k;
  return PMT_OK;
}

// bagging_the_peaks on a HOST array: only the peaks travel to the GPU (the old form uploaded all of `elements`)
static int pmt_mmr_bag_impl(pmt_ctx* c, const uint64_t* elements, size_t n_leaves, uint64_t* root_out);
int pmt_mmr_bag(pmt_ctx* c, const uint64_t* elements, size_t n_leaves, uint64_t* root_out) {
  return c ? drained(c, pmt_mmr_bag_impl(c, elements, n_leaves, root_out)) : PMT_E_INVALID_ARG;
}
static int pmt_mmr_bag_impl(pmt_ctx* c, const uint64_t* elements, size_t n_leaves, uint64_t* root_out) {
  if (int rc = bind(c)) return rc;
  if (n_leaves == 0) return fail(c, PMT_E_INVALID_ARG, "mmr bag: empty MMR");
  if (!elements || !root_out) return fail(c, PMT_E_INVALID_ARG, "mmr bag: null pointer");
  if (n_leaves >> 32) return fail(c, PMT_E_RANGE, "mmr bag: size does not fit u32 (merkle_mountain_ranges.rs:184)");
  size_t pos[64];
  const uint32_t k = peak_positions(n_leaves, pos);
  void* p = nullptr;
  if (int rc = arena_get(c, 2, 64 * 32 + 64, &p)) return rc;
  uint64_t* d_peaks = (uint64_t*)p;
  uint64_t* d_root = d_peaks + 64 * 4;
  for (uint32_t i = 0; i < k; i++) H2D(c, d_peaks + 4 * i, elements + 4 * pos[i], 32);
  if (int rc = bag_peaks(c, d_peaks, k, d_root)) return rc;
  D2H(c, root_out, d_root, 32);
  FINISH(c);
  return PMT_OK;
}

// ---- host-buffer proofs: gathers from the caller's arrays (nothing to compute: done on the host, like get_peaks) ---------
static inline void copy_digest_canonical(uint64_t* dst, const uint64_t* src) {
  for (int e = 0; e < 4; e++) dst[e] = src[e] >= PMT_P ? src[e] - PMT_P : src[e];
}

// get_merkle_proof (simple_merkle_tree.rs:55-74) for a batch, from the level-major host array pmt_simple_tree_build filled
int pmt_simple_tree_prove(pmt_ctx* c, const uint64_t* levels, size_t n, const uint64_t* idx, size_t n_idx, uint64_t* siblings_out) {
  if (!c) return PMT_E_INVALID_ARG;
  const int lg = log2_strict(n);
  if (lg < 1) return fail(c, PMT_E_NOT_POW2, "simple tree prove: bad leaf count %zu", n);
  if (n_idx == 0) return PMT_OK;
  if (!levels || !idx || !siblings_out) return fail(c, PMT_E_INVALID_ARG, "simple tree prove: null pointer");
  for (size_t q = 0; q < n_idx; q++) {
    if (idx[q] >= n) return fail(c, PMT_E_RANGE, "assert!(leaf_index < self.tree[0].len()): %llu >= %zu (simple_merkle_tree.rs:56)", (unsigned long long)idx[q], n);
    for (int l = 0; l < lg; l++) {
      const size_t k = (idx[q] >> l) ^ 1;
      copy_digest_canonical(siblings_out + 4 * (q * lg + l), levels + 4 * ((2 * n - ((2 * n) >> l)) + k));
    }
  }
  return PMT_OK;
}

// [UPSTREAM] MerkleTree::prove for a batch, from the host `digests` array pmt_merkle_tree_build filled
int pmt_merkle_prove(pmt_ctx* c, const uint64_t* digests, size_t n, uint32_t cap_height, const uint64_t* idx, size_t n_idx,
                     uint64_t* siblings_out) {
  if (!c) return PMT_E_INVALID_ARG;
  const int lg = log2_strict(n);
  if (lg < 0) return fail(c, PMT_E_NOT_POW2, "prove: %zu leaves is not a power of two", n);
  if ((int)cap_height > lg) return fail(c, PMT_E_RANGE, "prove: cap_height %u > log2 n", cap_height);
  const int L = lg - (int)cap_height;
  if (n_idx == 0 || L == 0) return PMT_OK;
  if (!digests || !idx || !siblings_out) return fail(c, PMT_E_INVALID_ARG, "prove: null pointer");
  for (size_t q = 0; q < n_idx; q++) {
    if (idx[q] >= n) return fail(c, PMT_E_RANGE, "prove: leaf index %llu >= %zu", (unsigned long long)idx[q], n);
    for (int l = 0; l < L; l++)
      copy_digest_canonical(siblings_out + 4 * (q * L + l), digests + 4 * plonky2_index(L, l, (idx[q] >> l) ^ 1));
  }
  return PMT_OK;
}

// get_proof_normal_index (merkle_mountain_ranges.rs:203-223) for a batch, from the host `elements` array; same output
// layout as pmt_mmr_prove_dev (32 entries per proof)
int pmt_mmr_prove(pmt_ctx* c, const uint64_t* elements, size_t n_leaves, const uint64_t* leaf_idx, size_t n_idx,
                  uint64_t* siblings_out, uint8_t* on_left_out, uint32_t* path_len_out) {
  if (!c) return PMT_E_INVALID_ARG;
  if (n_idx == 0) return PMT_OK;
  if (!elements || !leaf_idx || !siblings_out || !on_left_out || !path_len_out) return fail(c, PMT_E_INVALID_ARG, "mmr prove: null pointer");
  if (n_leaves == 0 || n_leaves > ((size_t)1 << 30)) return fail(c, PMT_E_RANGE, "mmr prove: bad leaf count");
  for (size_t q = 0; q < n_idx; q++) {
    const size_t i = leaf_idx[q];
    if (i >= n_leaves) return fail(c, PMT_E_RANGE, "mmr prove: leaf index %zu >= %zu", i, n_leaves);
    int H = 0;
    size_t base = 0;
    for (int b = 63; b >= 0; b--)
      if ((n_leaves >> b) & 1) {
        if (i < base + ((size_t)1 << b)) { H = b; break; }
        base += (size_t)1 << b;
      }
    path_len_out[q] = (uint32_t)H;
    for (int j = 0; j < H; j++) {
      const size_t k = (i >> j) ^ 1, last = ((k + 1) << j) - 1;                        // sibling (height j, index k)
      const size_t pos = 2 * last - (size_t)__builtin_popcountll((unsigned long long)last) + (size_t)j;
      copy_digest_canonical(siblings_out + 4 * (32 * q + j), elements + 4 * pos);
      on_left_out[32 * q + j] = (uint8_t)((i >> j) & 1);
    }
  }
  return PMT_OK;
}

// ---- host-buffer verification: upload the batch, fold every path on the GPU, download the verdicts ------------------------
static int verify_to_cap_host(pmt_ctx* c, const uint64_t* leaf_rows, size_t w, const uint64_t* idx, size_t idx_mask, size_t n_idx,
                              const uint64_t* cap, uint32_t cap_height, const uint64_t* proofs, size_t path_len, uint8_t* ok_out);
int pmt_merkle_verify(pmt_ctx* c, const uint64_t* leaf_rows, size_t w, const uint64_t* idx, size_t n_idx, const uint64_t* cap,
                      uint32_t cap_height, const uint64_t* proofs, size_t path_len, uint8_t* ok_out) {
  return c ? drained(c, verify_to_cap_host(c, leaf_rows, w, idx, ~(size_t)0, n_idx, cap, cap_height, proofs, path_len, ok_out)) : PMT_E_INVALID_ARG;
}
static int verify_to_cap_host(pmt_ctx* c, const uint64_t* leaf_rows, size_t w, const uint64_t* idx, size_t idx_mask, size_t n_idx,
                              const uint64_t* cap, uint32_t cap_height, const uint64_t* proofs, size_t path_len, uint8_t* ok_out) {
  if (int rc = bind(c)) return rc;
  if (n_idx == 0) return PMT_OK;
  if (!leaf_rows || !idx || !cap || !ok_out || (!proofs && path_len)) return fail(c, PMT_E_INVALID_ARG, "verify: null pointer");
  if (w == 0 || cap_height > 40 || path_len > 63) return fail(c, PMT_E_INVALID_ARG, "verify: bad width / cap_height / path_len");
  const size_t b_rows = n_idx * w * 8, b_idx = n_idx * 8, b_cap = ((size_t)32) << cap_height, b_pr = n_idx * path_len * 32;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  void* a;
  if (int rc = arena_get(c, 0, al(b_rows) + al(b_idx) + al(b_cap) + al(b_pr) + al(n_idx), &a)) return rc;
  char* p = (char*)a;
  uint64_t* d_rows = (uint64_t*)p; p += al(b_rows);
  uint64_t* d_idx = (uint64_t*)p; p += al(b_idx);
  uint64_t* d_cap = (uint64_t*)p; p += al(b_cap);
  uint64_t* d_pr = (uint64_t*)p; p += al(b_pr);
  uint8_t* d_ok = (uint8_t*)p;
  H2D(c, d_rows, leaf_rows, b_rows);
  H2D(c, d_idx, idx, b_idx);
  H2D(c, d_cap, cap, b_cap);
  if (b_pr) H2D(c, d_pr, proofs, b_pr);
  if (int rc = verify_to_cap_dev(c, d_rows, w, d_idx, idx_mask, n_idx, d_cap, cap_height, d_pr, path_len, d_ok)) return rc;
  D2H(c, ok_out, d_ok, n_idx);
  FINISH(c);
  return PMT_OK;
}

int pmt_simple_tree_verify(pmt_ctx* c, const uint64_t* leaves, const uint64_t* idx, size_t n_idx, const uint64_t* root,
                           const uint64_t* proofs, size_t path_len, uint8_t* ok_out) {
  if (!c) return PMT_E_INVALID_ARG;
  if (path_len > 63) return fail(c, PMT_E_INVALID_ARG, "verify: bad path_len");
  // index masked to its low path_len bits: see pmt_simple_tree_verify_dev
  return drained(c, verify_to_cap_host(c, leaves, 1, idx, ((size_t)1 << path_len) - 1, n_idx, root, 0, proofs, path_len, ok_out));
}

static int pmt_mmr_verify_impl(pmt_ctx* c, const uint64_t* leaves, size_t n_idx, const uint64_t* siblings, const uint8_t* on_left,
                   const uint32_t* path_len, const uint64_t* peaks, uint32_t n_peaks, const uint64_t* root, int8_t* status_out);
int pmt_mmr_verify(pmt_ctx* c, const uint64_t* leaves, size_t n_idx, const uint64_t* siblings, const uint8_t* on_left,
                   const uint32_t* path_len, const uint64_t* peaks, uint32_t n_peaks, const uint64_t* root, int8_t* status_out) {
  return c ? drained(c, pmt_mmr_verify_impl(c, leaves, n_idx, siblings, on_left, path_len, peaks, n_peaks, root, status_out)) : PMT_E_INVALID_ARG;
}
static int pmt_mmr_verify_impl(pmt_ctx* c, const uint64_t* leaves, size_t n_idx, const uint64_t* siblings, const uint8_t* on_left,
                   const uint32_t* path_len, const uint64_t* peaks, uint32_t n_peaks, const uint64_t* root, int8_t* status_out) {
  if (int rc = bind(c)) return rc;
  if (n_idx == 0) return PMT_OK;
  if (!leaves || !siblings || !on_left || !path_len || !peaks || !root || !status_out) return fail(c, PMT_E_INVALID_ARG, "mmr verify: null pointer");
  if (n_peaks == 0 || n_peaks > 64) return fail(c, PMT_E_INVALID_ARG, "mmr verify: bad peak count");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_lv = n_idx * 8, b_sib = n_idx * 32 * 32, b_left = n_idx * 32, b_len = n_idx * 4, b_pk = (size_t)n_peaks * 32;
  void* a;
  if (int rc = arena_get(c, 0, al(b_lv) + al(b_sib) + al(b_left) + al(b_len) + al(b_pk) + al(32) + al(n_idx), &a)) return rc;
  char* p = (char*)a;
  uint64_t* d_lv = (uint64_t*)p; p += al(b_lv);
  uint64_t* d_sib = (uint64_t*)p; p += al(b_sib);
  uint8_t* d_left = (uint8_t*)p; p += al(b_left);
  uint32_t* d_len = (uint32_t*)p; p += al(b_len);
  uint64_t* d_pk = (uint64_t*)p; p += al(b_pk);
  uint64_t* d_root = (uint64_t*)p; p += al(32);
  int8_t* d_st = (int8_t*)p;
  H2D(c, d_lv, leaves, b_lv);
  H2D(c, d_sib, siblings, b_sib);
  H2D(c, d_left, on_left, b_left);
  H2D(c, d_len, path_len, b_len);
  H2D(c, d_pk, peaks, b_pk);
  H2D(c, d_root, root, 32);
  if (int rc = pmt_mmr_verify_dev(c, d_lv, n_idx, d_sib, d_left, d_len, d_pk, n_peaks, d_root, d_st)) return rc;
  D2H(c, status_out, d_st, n_idx);
  FINISH(c);
  return PMT_OK;
}

}  // extern "C"
