// backend.cu - the thin C-ABI layer between zfp's host dispatch and the sm_100a kernels.
//
// Replaces the reference's src/cuda_zfp/cuZFP.cu (entry points cuda_compress :357-414 and
// cuda_decompress :416-491) with the same signatures and the same post-conditions on the host
// bitstream, but supports every compression mode, a non-zero stream offset, strided and
// partial-block fields, 64-bit block counts, and keeps everything on the caller's device.
//
// There is no CPU code path in this file: if a CUDA call fails the entry points return failure.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "../../include/zfp_b200_backend.h"
#include "bitstream_impl.h"
#include "types.h"
#include "stream_kernels.cuh"

using namespace zb;

// ------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};

static bool cuda_ok(cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return true;
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
#define CU(call) do { if (!cuda_ok((call), #call)) return ZFP_B200_ECUDA; } while (0)
#define LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); \
                        if (!cuda_ok(cudaGetLastError(), "kernel launch")) return ZFP_B200_ECUDA; } while (0)

extern "C" const char* zfp_b200_last_error(void) { return g_error.c_str(); }
extern "C" uint64 zfp_b200_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// device scratch: grow-only buffer SETS cached per device and leased to one call at a time.
// The reference's contract is "thread-safe as long as threads do not share a zfp_stream"
// (docs/source/faq.rst:1096-1098; its CUDA backend allocates per call).  Here every entry point takes
// a lease for its duration: two host threads, or two CUDA streams, working on the same device get
// different sets.  A call may return with its kernels still in flight (device_only_sync), so a set
// goes back to the pool with an event recorded on the caller's stream and its next user makes its own
// stream wait for that event first.  Growing a buffer frees the old one with cudaFree, which waits for
// the device, so work still using it completes.
// ------------------------------------------------------------------------------------------------
namespace {

enum { SCR_SLOTS = 0, SCR_LENGTHS, SCR_TILES, SCR_OFFSETS, SCR_CURSOR, SCR_STAGE_DATA, SCR_STAGE_WORDS, SCR_COUNT };

struct ScratchBuf { void* p = nullptr; size_t bytes = 0; };
struct ScratchSet {
  ScratchBuf buf[SCR_COUNT];
  cudaEvent_t idle = nullptr;  // recorded when the last lease ended
  bool recorded = false, busy = false;
  // variable-rate encode: the compaction of chunk i runs on a second stream beside the encode of chunk i+1
  cudaStream_t aux = nullptr;
  cudaEvent_t encoded[2] = { nullptr, nullptr }, compacted[2] = { nullptr, nullptr };
  bool aux_ok = false, aux_tried = false;
};
constexpr int kMaxDevices = 64, kMaxSets = 16;
std::mutex g_scratch_mutex;
ScratchSet* g_sets[kMaxDevices][kMaxSets];
thread_local ScratchSet* t_lease = nullptr;

class ScratchLease {
 public:
  explicit ScratchLease(cudaStream_t st) : st_(st)
  {
    if (t_lease) return;  // nested entry point: the outer call's set
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (int i = 0; i < kMaxSets && !set_; i++) {
      if (!g_sets[dev][i]) g_sets[dev][i] = new (std::nothrow) ScratchSet();
      if (g_sets[dev][i] && !g_sets[dev][i]->busy) set_ = g_sets[dev][i];
    }
    if (!set_) return;  // more than kMaxSets concurrent calls on one device: scratch() reports the failure
    set_->busy = true;
    if (set_->recorded) cudaStreamWaitEvent(st_, set_->idle, 0);
    t_lease = set_;
  }
  ~ScratchLease()
  {
    if (!set_) return;
    if (!set_->idle && cudaEventCreateWithFlags(&set_->idle, cudaEventDisableTiming) != cudaSuccess) set_->idle = nullptr;
    set_->recorded = set_->idle && cudaEventRecord(set_->idle, st_) == cudaSuccess;
    if (!set_->recorded) cudaStreamSynchronize(st_);
    t_lease = nullptr;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    set_->busy = false;
  }
  ScratchLease(const ScratchLease&) = delete;
  ScratchLease& operator=(const ScratchLease&) = delete;

 private:
  cudaStream_t st_;
  ScratchSet* set_ = nullptr;
};

void* scratch(int slot, size_t bytes)
{
  if (!t_lease) {
    g_error = "no scratch set available (too many concurrent calls on this device)";
    return nullptr;
  }
  ScratchBuf& b = t_lease->buf[slot];
  if (b.bytes < bytes) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (!cuda_ok(cudaMalloc(&b.p, want), "cudaMalloc(scratch)")) return nullptr;
    b.bytes = want;
  }
  return b.p;
}

// the leased set's second stream and its fork / join events (nullptr when they cannot be made: single-stream order)
ScratchSet* overlap_set()
{
  ScratchSet* s = t_lease;
  // (opt-in, ZFP_B200_OVERLAP=1: measured slower - 1024^3 fp64 accuracy 1e-6 compress 5.97 ms against 5.81 - the
  // encode kernel's CTAs hold every multiprocessor, the helpers only get what is left, and the chunks are halved)
  if (!s || !getenv("ZFP_B200_OVERLAP")) return nullptr;
  if (!s->aux_tried) {
    s->aux_tried = true;
    bool ok = cudaStreamCreateWithFlags(&s->aux, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
      ok = cudaEventCreateWithFlags(&s->encoded[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&s->compacted[i], cudaEventDisableTiming) == cudaSuccess;
    s->aux_ok = ok;
  }
  return s->aux_ok ? s : nullptr;
}

// multiprocessors of the current device (grid sizing of the helper kernels)
int sm_count()
{
  static std::atomic<int> cached[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (!n) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace

extern "C" void zfp_b200_release_scratch(void)
{
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < kMaxDevices; d++)
    for (int i = 0; i < kMaxSets; i++) {
      ScratchSet* set = g_sets[d][i];
      if (!set || set->busy) continue;
      cudaSetDevice(d);
      for (int k = 0; k < SCR_COUNT; k++)
        if (set->buf[k].p) cudaFree(set->buf[k].p);
      if (set->idle) cudaEventDestroy(set->idle);
      delete set;
      g_sets[d][i] = nullptr;
    }
  cudaSetDevice(cur);
}

// ------------------------------------------------------------------------------------------------
// block-offset index
// ------------------------------------------------------------------------------------------------
struct zfp_b200_index {
  uint16_t* d_lengths = nullptr;
  size_t blocks = 0;
  size_t capacity = 0;
  uint64_t total_bits = 0;  // sum of the lengths, recorded by the encode that filled the index
  bool speculative = false; // made by zfp_b200_index_rebuild for whatever stream was decoded last: never reused as is
  // what the lengths describe (filled by zfp_b200_encode): a decode with other parameters or another
  // start phase ignores the index and rebuilds one; a decode that finds a block whose parsed length
  // differs from the recorded one (same shape and parameters, other data) reports it and is redone
  bool keyed = false;
  zfp_b200_desc key_desc;
  uint64_t key_start = 0;
};

static bool index_matches(const zfp_b200_index* ix, const zfp_b200_desc* d, const Geom& g, uint64_t start_bit)
{
  if (!ix || ix->blocks != g.nblocks) return false;
  if (!ix->keyed) return true;  // imported lengths: the caller vouches for them (checked block by block during the decode)
  const zfp_b200_desc& k = ix->key_desc;
  bool same = k.type == d->type && k.dims == d->dims && k.minbits == d->minbits && k.maxbits == d->maxbits &&
              k.maxprec == d->maxprec && k.minexp == d->minexp && (ix->key_start & 63) == (start_bit & 63);
  for (uint32_t i = 0; i < d->dims && same; i++)
    same = k.n[i] == d->n[i];
  return same;
}

extern "C" zfp_b200_index* zfp_b200_index_create(void) { return new (std::nothrow) zfp_b200_index(); }

extern "C" void zfp_b200_index_destroy(zfp_b200_index* ix)
{
  if (!ix) return;
  if (ix->d_lengths) cudaFree(ix->d_lengths);
  delete ix;
}

static bool index_reserve(zfp_b200_index* ix, size_t blocks)
{
  if (ix->capacity < blocks) {
    if (ix->d_lengths) cudaFree(ix->d_lengths);
    ix->d_lengths = nullptr;
    ix->capacity = 0;
    if (!cuda_ok(cudaMalloc(&ix->d_lengths, blocks * sizeof(uint16_t) + 64), "cudaMalloc(index)")) return false;
    ix->capacity = blocks;
  }
  ix->blocks = blocks;
  return true;
}

extern "C" size_t zfp_b200_index_blocks(const zfp_b200_index* ix) { return ix ? ix->blocks : 0; }
extern "C" uint64 zfp_b200_index_bits(const zfp_b200_index* ix) { return ix ? ix->total_bits : 0; }

extern "C" size_t zfp_b200_index_export(const zfp_b200_index* ix, uint16_t* host, size_t capacity)
{
  if (!ix || !host || capacity < ix->blocks) return 0;
  if (!cuda_ok(cudaMemcpy(host, ix->d_lengths, ix->blocks * sizeof(uint16_t), cudaMemcpyDeviceToHost), "index export"))
    return 0;
  return ix->blocks;
}

extern "C" int zfp_b200_index_import(zfp_b200_index* ix, const uint16_t* host, size_t blocks)
{
  if (!ix || !host) return ZFP_B200_EINVAL;
  if (!index_reserve(ix, blocks)) return ZFP_B200_ECUDA;
  CU(cudaMemcpy(ix->d_lengths, host, blocks * sizeof(uint16_t), cudaMemcpyHostToDevice));
  ix->keyed = false;
  ix->speculative = false;
  ix->total_bits = 0;
  return ZFP_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry and parameters
// ------------------------------------------------------------------------------------------------
static size_t scalar_bytes(int type) { return (type == T_INT32 || type == T_FLOAT) ? 4 : 8; }

static bool make_geom(const zfp_b200_desc* d, const void* data, Geom* g)
{
  if (!d || d->dims < 1 || d->dims > 4 || d->type < T_INT32 || d->type > T_DOUBLE) return false;
  int64_t contiguous = 1;
  g->nblocks = 1;
  for (uint32_t i = 0; i < 4; i++) {
    uint64_t n = i < d->dims ? d->n[i] : 1;
    if (i < d->dims && n == 0) return false;
    g->n[i] = n;
    g->s[i] = (i < d->dims && d->s[i]) ? (int64_t)d->s[i] : contiguous;
    g->nb[i] = (n + 3) / 4;
    g->nblocks *= g->nb[i];
    contiguous *= (int64_t)n;
  }
  const size_t row = 4 * scalar_bytes(d->type);
  bool vec = g->s[0] == 1 && (reinterpret_cast<uintptr_t>(data) % row) == 0;
  for (uint32_t i = 1; i < d->dims; i++)
    vec = vec && (g->s[i] % 4) == 0;
  g->vec_rows = vec ? 1 : 0;
  g->box = 0;
  for (int i = 0; i < 4; i++) {
    g->bl[i] = 0;
    g->be[i] = 1;
  }
  return true;
}

// bits a block spends before its coefficients: exponent / reversible-mode prefix and precision
static uint32_t block_header_bits(const zfp_b200_desc* d)
{
  const bool reversible = d->minexp < kMinExp;
  switch (d->type) {
    case T_INT32: return reversible ? 5 : 0;
    case T_INT64: return reversible ? 6 : 0;
    case T_FLOAT: return reversible ? 15 : 9;
    default: return reversible ? 19 : 12;
  }
}

// Besides the reference's own checks (src/zfp.c:813-824): a maxbits below the block header is refused.
// Upstream accepts it through zfp_stream_set_params, but then writes the whole header anyway and
// hands the coefficient coder the wrapped-around budget `maxbits - bits` (src/template/encodef.c:77-82),
// i.e. unbounded blocks that overrun the size zfp_stream_maximum_size promises the caller
// (src/zfp.c:711-742).  The rate setter never produces such parameters (src/zfp.c:760-784).
static bool check_params(const zfp_b200_desc* d)
{
  if (!(d->minbits <= d->maxbits && d->maxprec >= 1 && d->maxprec <= 64 && d->maxbits >= 1))
    return false;
  if (d->maxbits < block_header_bits(d)) {
    g_error = "maxbits is smaller than the block header: parameters outside zfp_stream_maximum_size's contract";
    return false;
  }
  // variable rate: block lengths travel as 16-bit words (the index format).  No block codes more than
  // 16 658 bits of its own (4-D double, reversible), so only a minbits padding beyond 65 535 could
  // overflow them: refused (expert parameters nobody uses) rather than wrapped.
  if (d->minbits != d->maxbits && d->minbits > 65535) {
    g_error = "variable-rate parameters with minbits > 65535 are not supported (16-bit block lengths)";
    return false;
  }
  return true;
}

// worst-case coded size of one block (the per-block term of zfp_stream_maximum_size, src/zfp.c:711-742)
static uint32_t block_capacity_bits(const zfp_b200_desc* d)
{
  const uint32_t values = 1u << (2 * d->dims), prec = (uint32_t)(8 * scalar_bytes(d->type));
  uint32_t bits = block_header_bits(d);
  bits += values - 1 + values * (d->maxprec < prec ? d->maxprec : prec);
  if (bits > d->maxbits) bits = d->maxbits;
  if (bits < d->minbits) bits = d->minbits;
  return bits;
}

extern "C" int zfp_b200_is_fixed_rate(const zfp_b200_desc* d) { return d && d->minbits == d->maxbits; }

extern "C" size_t zfp_b200_blocks(const zfp_b200_desc* d)
{
  Geom g;
  return make_geom(d, nullptr, &g) ? (size_t)g.nblocks : 0;
}

extern "C" size_t zfp_b200_capacity(const zfp_b200_desc* d, uint64 start_bit)
{
  Geom g;
  if (!make_geom(d, nullptr, &g)) return 0;
  uint64_t bits = start_bit + 148 + g.nblocks * (uint64_t)block_capacity_bits(d);
  return (size_t)(((bits + 63) & ~(uint64_t)63) / 8);
}

// ------------------------------------------------------------------------------------------------
// launches: the kernel instances live in inst_{enc,dec}_<type>.cu
// ------------------------------------------------------------------------------------------------
static int encode_any(int out_mode, int type, uint32_t dims, const void* data, const Geom& g, const Params& prm, void* out,
                      uint64_t start_bit, uint32_t slot_words, uint16_t* lengths, uint64_t b0, uint64_t b1, cudaStream_t st)
{
  static const int staged = getenv("ZFP_B200_NO_STAGED") ? 0 : 1;
  const EncodeArgs a = { data, g, prm, out, start_bit, slot_words, lengths, b0, b1, st, staged };
  cudaError_t e;
  switch (type) {
    case T_INT32: e = launch_encode_t<T_INT32>((int)dims, out_mode, a); break;
    case T_INT64: e = launch_encode_t<T_INT64>((int)dims, out_mode, a); break;
    case T_FLOAT: e = launch_encode_t<T_FLOAT>((int)dims, out_mode, a); break;
    case T_DOUBLE: e = launch_encode_t<T_DOUBLE>((int)dims, out_mode, a); break;
    default: return ZFP_B200_EINVAL;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_ok(e, "encode kernel launch") ? ZFP_B200_OK : ZFP_B200_ECUDA;
}

static int decode_any(int offs_mode, int type, uint32_t dims, void* data, const Geom& g, const Params& prm, const void* in,
                      uint64_t start_bit, const uint64_t* offsets, const uint16_t* lengths, cudaStream_t st, uint64_t b0, uint64_t b1,
                      uint32_t* check = nullptr)
{
  static const int staged = getenv("ZFP_B200_NO_STAGED") ? 0 : 1;
  if (b0 >= b1) return ZFP_B200_OK;
  const DecodeArgs a = { data, g, prm, in, start_bit, offsets, lengths, st, staged, b0, b1, check };
  cudaError_t e;
  switch (type) {
    case T_INT32: e = launch_decode_t<T_INT32>((int)dims, offs_mode, a); break;
    case T_INT64: e = launch_decode_t<T_INT64>((int)dims, offs_mode, a); break;
    case T_FLOAT: e = launch_decode_t<T_FLOAT>((int)dims, offs_mode, a); break;
    case T_DOUBLE: e = launch_decode_t<T_DOUBLE>((int)dims, offs_mode, a); break;
    default: return ZFP_B200_EINVAL;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_ok(e, "decode kernel launch") ? ZFP_B200_OK : ZFP_B200_ECUDA;
}

// exclusive scan of `n` block lengths into bit offsets, continuing at cursor[1]
static int scan_lengths(const uint16_t* lengths, uint64_t n, uint64_t* tiles, uint64_t* offsets, uint64_t* cursor,
                        cudaStream_t st, uint32_t cap = 0xffffffffu)
{
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_sums<<<(unsigned)ntiles, kScanThreads, 0, st>>>(lengths, n, tiles, cap);
  LAUNCHED();
  scan_tile_offsets<<<1, 1024, 0, st>>>(tiles, ntiles, cursor);
  LAUNCHED();
  scan_apply<<<(unsigned)ntiles, kScanThreads, 0, st>>>(lengths, n, tiles, offsets, cap);
  LAUNCHED();
  return ZFP_B200_OK;
}

// cursor[0] = first bit, cursor[1] = running end, cursor[2] = decode-time index check (0 = lengths agree)
__global__ void set_cursor(uint64_t* cursor, uint64_t v) { cursor[0] = v; cursor[1] = v; cursor[2] = 0; }

// ------------------------------------------------------------------------------------------------
// raw entry points
// ------------------------------------------------------------------------------------------------
__global__ void store_u64(uint64_t* dst, uint64_t v) { *dst = v; }
__global__ void copy_u64(uint64_t* dst, const uint64_t* src) { *dst = *src; }

// carry = {end position, last 64 bits of the stream so far}: the payload starts at start_bit, whatever a
// header left below it in its word is the tail the first tile completes
__global__ void init_carry(unsigned long long* carry, const uint64_t* words, uint64_t start_bit)
{
  const uint32_t r = (uint32_t)(start_bit & 63);
  carry[0] = start_bit;
  carry[1] = r ? (words[start_bit >> 6] & ((1ull << r) - 1)) << (64 - r) : 0;
}

__global__ void finish_async_var1(uint64_t* d_end, const unsigned long long* carry, const unsigned int* count, unsigned int capacity)
{
  *d_end = *count > capacity ? ~0ull : carry[0];
}

static int launch_var1(int type, const EncodeArgs& a, const Var1Bufs& v)
{
  cudaError_t e;
  switch (type) {
    case T_INT32: e = launch_encode_var1_t<T_INT32>(a, v); break;
    case T_INT64: e = launch_encode_var1_t<T_INT64>(a, v); break;
    case T_FLOAT: e = launch_encode_var1_t<T_FLOAT>(a, v); break;
    case T_DOUBLE: e = launch_encode_var1_t<T_DOUBLE>(a, v); break;
    default: return ZFP_B200_EINVAL;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_ok(e, "single-pass encode launch") ? ZFP_B200_OK : ZFP_B200_ECUDA;
}

// hand over from the single pass to the slot path: the running end becomes the cursor and the partial word the
// single pass keeps in its carry is written out (upper bits zero), as clear_word_tail leaves it
// 3-D variable rate in one pass.  Blocks that outgrow the shared-memory window leave a hole and are coded again
// by the clean-up launch; when there are more of them than the overflow list holds (noise at tight tolerances)
// a synchronous call starts over on the slot path (*resume = 0), a stream-ordered call (d_end_bit) reports an
// end position of ~0.  (A probe of the first 64 Ki blocks with a hand-over to the slot path in mid-stream was
// tried and removed: it lost streams of long blocks at 290 K blocks and the path is slower anyway.)
static int encode_var1(const zfp_b200_desc* d, const Geom& g, const Params& prm, const void* d_data, void* d_words,
                       uint64 start_bit, uint64* end_bit, uint64* d_end_bit, zfp_b200_index* index, cudaStream_t st,
                       uint64_t* resume)
{
  const int type = d->type;
  int tile = 128;
  switch (type) {
    case T_INT32: tile = var1_tile_blocks<T_INT32>(); break;
    case T_INT64: tile = var1_tile_blocks<T_INT64>(); break;
    case T_FLOAT: tile = var1_tile_blocks<T_FLOAT>(); break;
    default: tile = var1_tile_blocks<T_DOUBLE>(); break;
  }
  *resume = g.nblocks;
  const uint64_t tiles = (g.nblocks + tile - 1) / tile + 1;
  uint64_t cap = g.nblocks / 16 + 4096;
  if (cap > ((uint64_t)1 << 24)) cap = (uint64_t)1 << 24;
  uint16_t* lengths;
  if (index) {
    if (!index_reserve(index, g.nblocks)) return ZFP_B200_ECUDA;
    lengths = index->d_lengths;
  }
  else
    lengths = static_cast<uint16_t*>(scratch(SCR_LENGTHS, g.nblocks * sizeof(uint16_t)));
  // {ticket, overflow count} + status in one buffer; overflow list; carry (the cursor scratch)
  char* stat = static_cast<char*>(scratch(SCR_TILES, tiles * 16 + 16));
  void* list = scratch(SCR_OFFSETS, cap * 16);
  unsigned long long* carry = static_cast<unsigned long long*>(scratch(SCR_CURSOR, 64));
  if (!lengths || !stat || !list || !carry) return ZFP_B200_ECUDA;
  init_carry<<<1, 1, 0, st>>>(carry, static_cast<const uint64_t*>(d_words), start_bit);
  LAUNCHED();
  Var1Bufs v;
  v.status = stat + 16;
  v.ticket = reinterpret_cast<unsigned int*>(stat);
  v.overflow_count = reinterpret_cast<unsigned int*>(stat) + 1;
  v.carry = carry;
  v.overflow = list;
  v.overflow_capacity = (unsigned int)cap;
  v.sms = sm_count();
  v.cleanup = 0;
  CU(cudaMemsetAsync(stat, 0, 16, st));
  EncodeArgs a = { d_data, g, prm, d_words, start_bit, 0, lengths, 0, g.nblocks, st, 1 };
  int rc;
  CU(cudaMemsetAsync(stat + 16, 0, ((a.b1 - a.b0 + tile - 1) / tile) * 16, st));
  if ((rc = launch_var1(type, a, v))) return rc;
  v.cleanup = 1;
  if ((rc = launch_var1(type, a, v))) return rc;
  if (index) {
    index->keyed = true;
    index->speculative = false;
    index->key_desc = *d;
    index->key_start = start_bit;
    index->total_bits = 0;
  }
  if (d_end_bit) {  // stream-ordered: the size stays on the device
    finish_async_var1<<<1, 1, 0, st>>>(d_end_bit, carry, v.overflow_count, v.overflow_capacity);
    LAUNCHED();
    return ZFP_B200_OK;
  }
  unsigned long long h_end = 0;
  unsigned int h_cnt[2];
  CU(cudaMemcpyAsync(&h_end, carry, sizeof(h_end), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h_cnt, stat, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (h_cnt[1] > cap) {
    // more long blocks than the list holds: start over on the slot path
    *resume = 0;
    return ZFP_B200_OK;
  }
  if (end_bit) *end_bit = h_end;
  if (index) index->total_bits = h_end - start_bit;
  return ZFP_B200_OK;
}

// d_end_bit != nullptr: the end position is left in DEVICE memory and nothing synchronises the stream
// (variable rate: the host never learns the size; the index's total_bits stays 0)
static int encode_impl(const zfp_b200_desc* d, const void* d_data, void* d_words, uint64 start_bit,
                       uint64* end_bit, uint64* d_end_bit, zfp_b200_index* index, void* cuda_stream);

extern "C" int zfp_b200_encode(const zfp_b200_desc* d, const void* d_data, void* d_words, uint64 start_bit,
                               uint64* end_bit, zfp_b200_index* index, void* cuda_stream)
{
  return encode_impl(d, d_data, d_words, start_bit, end_bit, nullptr, index, cuda_stream);
}

extern "C" int zfp_b200_encode_async(const zfp_b200_desc* d, const void* d_data, void* d_words, uint64 start_bit,
                                     uint64* d_end_bit, zfp_b200_index* index, void* cuda_stream)
{
  if (!d_end_bit) return ZFP_B200_EINVAL;
  return encode_impl(d, d_data, d_words, start_bit, nullptr, d_end_bit, index, cuda_stream);
}

static int encode_impl(const zfp_b200_desc* d, const void* d_data, void* d_words, uint64 start_bit,
                       uint64* end_bit, uint64* d_end_bit, zfp_b200_index* index, void* cuda_stream)
{
  Geom g;
  if (!make_geom(d, d_data, &g) || !d_data || !d_words) {
    g_error = "zfp_b200_encode: invalid descriptor";
    return ZFP_B200_EINVAL;
  }
  g_error = "zfp_b200_encode: invalid compression parameters";
  if (!check_params(d))
    return ZFP_B200_EINVAL;
  g_error.clear();
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ScratchLease lease(st);
  const Params prm = { d->minbits, d->maxbits, d->maxprec, d->minexp };
  const int type = d->type;
  const uint32_t dims = d->dims;
  int rc;

  if (d->minbits == d->maxbits) {
    // fixed rate: block b lives at start + b*maxbits, no communication between blocks
    const uint64_t total = g.nblocks * (uint64_t)d->maxbits, end = start_bit + total;
    // blocks of whole 64-bit words go out with plain stores from any kernel; blocks of whole 32-bit
    // words too, from the column kernels (dims <= 3, budget within their staging limit)
    static const bool staged = getenv("ZFP_B200_NO_STAGED") == nullptr;
    const bool words32 = (d->maxbits & 31) == 0 && staged && dims <= 3 && d->maxbits <= 4096;
    if ((start_bit & 63) == 0 && ((d->maxbits & 63) == 0 || words32)) {
      if (end & 63)  // the stream ends mid-word: the rest of that word reads as zero, as after stream_flush
        CU(cudaMemsetAsync(static_cast<uint64_t*>(d_words) + (end >> 6), 0, 8, st));
      rc = encode_any(0, type, dims, d_data, g, prm, d_words, start_bit, 0, nullptr, 0, g.nblocks, st);
    }
    else {
      uint64_t* w = static_cast<uint64_t*>(d_words);
      const uint64_t w0 = (start_bit + 63) >> 6, w1 = (end + 63) >> 6;
      clear_word_tail<<<1, 1, 0, st>>>(w, start_bit);
      LAUNCHED();
      if (w1 > w0) CU(cudaMemsetAsync(w + w0, 0, (w1 - w0) * 8, st));
      rc = encode_any(1, type, dims, d_data, g, prm, d_words, start_bit, 0, nullptr, 0, g.nblocks, st);
    }
    if (rc) return rc;
    if (end_bit) *end_bit = end;
    if (d_end_bit) {
      store_u64<<<1, 1, 0, st>>>(d_end_bit, end);
      LAUNCHED();
    }
    return ZFP_B200_OK;
  }

  // variable rate, 3-D, on request (ZFP_B200_VAR1=1): one pass (kernels_var1.cuh) - encode, decoupled look-back,
  // shifted plain stores.  Bit-exact and without the slot scratch, but measured SLOWER than the slot path below on
  // the B200 (1024^3 fp64 accuracy 1e-6: 6.39 ms against 5.78 ms), so it is not the default.
  static const bool single_pass = getenv("ZFP_B200_VAR1") != nullptr;
  // (reversible mode on 64-bit types: nearly every block outgrows the window - straight to the slot path)
  if (dims == 3 && single_pass && !(prm.minexp < kMinExp && (type == T_DOUBLE || type == T_INT64))) {
    uint64_t resume = 0;
    rc = encode_var1(d, g, prm, d_data, d_words, start_bit, end_bit, d_end_bit, index, st, &resume);
    if (rc || resume >= g.nblocks) return rc;  // (otherwise: too many long blocks, start over below)
  }

  // variable rate: encode into per-block scratch slots, scan the lengths, compact bit-granularly
  const uint32_t slot_words = (block_capacity_bits(d) + 63) / 64 + 4;  // + room for a plane of overshoot past the budget
  const uint64_t slot_bytes = (uint64_t)slot_words * 8;
  // chunks of half a GiB of slots, two slot buffers: while chunk i is scanned and compacted on the set's second
  // stream (bandwidth-bound helpers), the encode of chunk i+1 (issue-bound) runs on the caller's
  ScratchSet* ov = overlap_set();
  uint64_t chunk = ((uint64_t)1 << (ov ? 29 : 30)) / slot_bytes;
  chunk = chunk / kScanTile * kScanTile;
  if (chunk < (uint64_t)kScanTile) chunk = kScanTile;
  if (chunk >= g.nblocks) {
    chunk = g.nblocks;
    ov = nullptr;  // a single chunk: nothing to overlap
  }

  uint16_t* lengths;
  if (index) {
    if (!index_reserve(index, g.nblocks)) return ZFP_B200_ECUDA;
    lengths = index->d_lengths;
  }
  else
    lengths = static_cast<uint16_t*>(scratch(SCR_LENGTHS, g.nblocks * sizeof(uint16_t)));
  uint64_t* slots = static_cast<uint64_t*>(scratch(SCR_SLOTS, chunk * slot_bytes * (ov ? 2 : 1)));
  uint64_t* tiles = static_cast<uint64_t*>(scratch(SCR_TILES, ((chunk + kScanTile - 1) / kScanTile + 1) * 8));
  uint64_t* offsets = static_cast<uint64_t*>(scratch(SCR_OFFSETS, chunk * 8));
  uint64_t* cursor = static_cast<uint64_t*>(scratch(SCR_CURSOR, 64));
  if (!lengths || !slots || !tiles || !offsets || !cursor) return ZFP_B200_ECUDA;
  set_cursor<<<1, 1, 0, st>>>(cursor, start_bit);
  LAUNCHED();
  clear_word_tail<<<1, 1, 0, st>>>(static_cast<uint64_t*>(d_words), start_bit);
  LAUNCHED();
  cudaStream_t helper = ov ? ov->aux : st;
  int k = 0;
  for (uint64_t b0 = 0; b0 < g.nblocks; b0 += chunk, k++) {
    const uint64_t b1 = b0 + chunk < g.nblocks ? b0 + chunk : g.nblocks, cn = b1 - b0;
    uint64_t* buf = slots + (ov ? (uint64_t)(k & 1) * chunk * slot_words : 0);
    if (ov && k >= 2) CU(cudaStreamWaitEvent(st, ov->compacted[k & 1], 0));  // this slot buffer is free again
    rc = encode_any(2, type, dims, d_data, g, prm, buf, 0, slot_words, lengths, b0, b1, st);
    if (rc) return rc;
    if (ov) {
      CU(cudaEventRecord(ov->encoded[k & 1], st));
      CU(cudaStreamWaitEvent(helper, ov->encoded[k & 1], 0));
    }
    rc = scan_lengths(lengths + b0, cn, tiles, offsets, cursor, helper);
    if (rc) return rc;
    zero_new_words<<<sm_count() * 4, 256, 0, helper>>>(static_cast<uint64_t*>(d_words), cursor);
    LAUNCHED();
    compact_blocks<<<(unsigned)((cn * kCompactLanes + 255) / 256), 256, 0, helper>>>(buf, slot_words, lengths + b0, offsets, cn, d_words);
    LAUNCHED();
    if (ov) CU(cudaEventRecord(ov->compacted[k & 1], helper));
  }
  if (ov && k > 0) CU(cudaStreamWaitEvent(st, ov->compacted[(k - 1) & 1], 0));  // join: the helper stream is in order
  if (index) {
    index->keyed = true;
    index->speculative = false;
    index->key_desc = *d;
    index->key_start = start_bit;
    index->total_bits = 0;
  }
  if (d_end_bit) {  // stream-ordered: the size stays on the device
    copy_u64<<<1, 1, 0, st>>>(d_end_bit, cursor + 1);
    LAUNCHED();
    return ZFP_B200_OK;
  }
  uint64_t h_cursor[2];
  CU(cudaMemcpyAsync(h_cursor, cursor, sizeof(h_cursor), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (end_bit) *end_bit = h_cursor[1];
  if (index) index->total_bits = h_cursor[1] - start_bit;
  return ZFP_B200_OK;
}

static int decode_range(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit, uint64* end_bit,
                        const zfp_b200_index* index, void* cuda_stream, uint64_t block0, uint64_t block1, bool whole,
                        uint32_t* d_status = nullptr, const size_t* box_lo = nullptr, const size_t* box_hi = nullptr);

// Candidate block index of a variable-rate stream that came without one, rebuilt segment-parallel (kernels.cuh
// spec_index_kernel).  The caller says where the buffer ends (words_bytes from d_words): speculative walks start
// anywhere in it and must not read beyond.  The index is marked "imported": the decode verifies it block by block
// and falls back to the sequential rebuild if it is wrong.  Returns ZFP_B200_EINVAL when the case is not covered
// (4-D, few blocks, fixed rate) - the caller then simply decodes without an index.
extern "C" int zfp_b200_index_rebuild(const zfp_b200_desc* d, const void* d_words, uint64 start_bit, size_t words_bytes,
                                      zfp_b200_index* index, void* cuda_stream)
{
  Geom g;
  if (!index || !d_words || !make_geom(d, d_words, &g) || !check_params(d)) return ZFP_B200_EINVAL;
  if (d->minbits == d->maxbits || d->dims > 3 || g.nblocks < 16384) return ZFP_B200_EINVAL;
  const uint64_t avail = (uint64_t)words_bytes * 8;
  const uint32_t cap = block_capacity_bits(d), margin = cap + 128;
  if (avail <= start_bit + margin) return ZFP_B200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ScratchLease lease(st);
  // A walk that starts off a block boundary lands on one with probability ~1 / (bits per block) per garbage block,
  // i.e. after some 10^5 bits; segments are an order of magnitude longer, so that a walk has almost surely met the
  // true chain when it leaves its segment (if not, pass 1 repeats; after 8 rounds the sequential walk takes over).
  uint64_t seg = (avail - start_bit + 65535) / 65536;   // at most 65536 segments ...
  if (seg < (1ull << 20)) seg = 1ull << 20;              // ... of at least 1 Mbit
  if (seg < 256ull * cap) seg = 256ull * cap;
  const uint32_t nseg = (uint32_t)((avail - start_bit + seg - 1) / seg);
  if (nseg < 4) return ZFP_B200_EINVAL;                  // (too short to gain anything)
  if (!index_reserve(index, g.nblocks)) return ZFP_B200_ECUDA;
  uint64_t *d_exit = nullptr, *d_off = nullptr;
  uint32_t* d_cnt = nullptr;
  std::vector<uint32_t> h_cnt(nseg);
  std::vector<uint64_t> h_off(nseg);
  int rc = ZFP_B200_ECUDA;
  auto launch = [&](int pass, const SpecIndexArgs& a) -> cudaError_t {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (d->type) {
      case T_INT32: return launch_spec_index_t<T_INT32>((int)d->dims, pass, a, st);
      case T_INT64: return launch_spec_index_t<T_INT64>((int)d->dims, pass, a, st);
      case T_FLOAT: return launch_spec_index_t<T_FLOAT>((int)d->dims, pass, a, st);
      default: return launch_spec_index_t<T_DOUBLE>((int)d->dims, pass, a, st);
    }
  };
  do {
    if (cudaMalloc(&d_exit, (size_t)nseg * 8) != cudaSuccess || cudaMalloc(&d_off, (size_t)nseg * 8) != cudaSuccess ||
        cudaMalloc(&d_cnt, (size_t)nseg * 4 + 4) != cudaSuccess) break;
    uint32_t* d_changed = d_cnt + nseg;
    const Params prm = { d->minbits, d->maxbits, d->maxprec, d->minexp };
    SpecIndexArgs a = { d_words, start_bit, avail, seg, nseg, margin, prm, d_exit, d_cnt, d_changed, d_off, index->d_lengths, g.nblocks };
    if (!cuda_ok(launch(0, a), "index rebuild, pass 0")) break;
    bool settled = false;
    for (int it = 0; it < 8 && !settled; it++) {
      uint32_t changed = 0;
      if (cudaMemsetAsync(d_changed, 0, 4, st) != cudaSuccess || !cuda_ok(launch(1, a), "index rebuild, pass 1")) break;
      if (cudaMemcpyAsync(&changed, d_changed, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) break;
      settled = !changed;
    }
    if (!settled) { rc = ZFP_B200_EINVAL; break; }
    if (cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)nseg * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) break;
    uint64_t total = 0;
    for (uint32_t t = 0; t < nseg; t++) { h_off[t] = total; total += h_cnt[t]; }
    if (cudaMemcpyAsync(d_off, h_off.data(), (size_t)nseg * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (!cuda_ok(launch(2, a), "index rebuild, pass 2")) break;
    if (total < g.nblocks) {
      // the walks stop a worst-case block short of the buffer's end: the sequential walk finishes from there
      uint64_t last = 0;
      if (cudaMemcpyAsync(&last, d_exit + (nseg - 1), 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) break;
      const DecodeArgs da = { nullptr, g, prm, d_words, last, nullptr, nullptr, st, 0, total, g.nblocks, nullptr };
      cudaError_t e;
      switch (d->type) {
        case T_INT32: e = launch_index_t<T_INT32>((int)d->dims, da, index->d_lengths); break;
        case T_INT64: e = launch_index_t<T_INT64>((int)d->dims, da, index->d_lengths); break;
        case T_FLOAT: e = launch_index_t<T_FLOAT>((int)d->dims, da, index->d_lengths); break;
        default: e = launch_index_t<T_DOUBLE>((int)d->dims, da, index->d_lengths); break;
      }
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (!cuda_ok(e, "index rebuild, tail")) break;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    index->keyed = false;
    index->total_bits = 0;
    index->speculative = true;
    rc = ZFP_B200_OK;
  } while (false);
  cudaFree(d_exit);
  cudaFree(d_off);
  cudaFree(d_cnt);
  if (rc != ZFP_B200_OK) index->blocks = 0;
  return rc;
}

extern "C" int zfp_b200_decode(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit,
                               uint64* end_bit, const zfp_b200_index* index, void* cuda_stream)
{
  return decode_range(d, d_data, d_words, start_bit, end_bit, index, cuda_stream, 0, 0, true);
}

extern "C" int zfp_b200_decode_async(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit,
                                     const zfp_b200_index* index, uint32_t* d_status, void* cuda_stream)
{
  if (!d_status) return ZFP_B200_EINVAL;
  return decode_range(d, d_data, d_words, start_bit, nullptr, index, cuda_stream, 0, 0, true, d_status);
}

extern "C" int zfp_b200_decode_blocks(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit,
                                      uint64 block0, uint64 block1, const zfp_b200_index* index, void* cuda_stream)
{
  return decode_range(d, d_data, d_words, start_bit, nullptr, index, cuda_stream, block0, block1, false);
}

extern "C" int zfp_b200_decode_box(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit,
                                   const size_t* lo, const size_t* hi, const zfp_b200_index* index, void* cuda_stream)
{
  if (!lo || !hi) return ZFP_B200_EINVAL;
  return decode_range(d, d_data, d_words, start_bit, nullptr, index, cuda_stream, 0, 0, false, nullptr, lo, hi);
}

static int decode_range(const zfp_b200_desc* d, void* d_data, const void* d_words, uint64 start_bit, uint64* end_bit,
                        const zfp_b200_index* index, void* cuda_stream, uint64_t block0, uint64_t block1, bool whole,
                        uint32_t* d_status, const size_t* box_lo, const size_t* box_hi)
{
  Geom g;
  if (!make_geom(d, d_data, &g) || !d_data || !d_words) {
    g_error = "zfp_b200_decode: invalid descriptor";
    return ZFP_B200_EINVAL;
  }
  g_error = "zfp_b200_decode: invalid compression parameters";
  if (!check_params(d))
    return ZFP_B200_EINVAL;
  g_error.clear();
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ScratchLease lease(st);
  const Params prm = { d->minbits, d->maxbits, d->maxprec, d->minexp };
  int rc;
  if (whole) {
    block0 = 0;
    block1 = g.nblocks;
  }
  else if (box_lo) {
    // the blocks that intersect lo <= index < hi (x first, like desc->n): one launch over their list
    uint64_t count = 1;
    for (uint32_t i = 0; i < 4; i++) {
      uint64_t l = 0, h = 1;
      if (i < d->dims) {
        const uint64_t hi_i = box_hi[i] < g.n[i] ? box_hi[i] : g.n[i];
        if (box_lo[i] >= hi_i) return ZFP_B200_OK;  // empty box
        l = box_lo[i] / 4;
        h = (hi_i + 3) / 4;
      }
      g.bl[i] = (uint32_t)l;
      g.be[i] = (uint32_t)(h - l);
      count *= h - l;
    }
    g.box = 1;
    block0 = 0;
    block1 = count;
  }
  else if (block0 > block1 || block1 > g.nblocks) {
    g_error = "zfp_b200_decode_blocks: block range outside the field";
    return ZFP_B200_EINVAL;
  }

  if (d->minbits == d->maxbits) {
    rc = decode_any(0, d->type, d->dims, d_data, g, prm, d_words, start_bit, nullptr, nullptr, st, block0, block1);
    if (rc) return rc;
    if (end_bit) *end_bit = start_bit + g.nblocks * (uint64_t)d->maxbits;
    return ZFP_B200_OK;
  }

  // An index that was made for other parameters, another shape or another bit phase is not used; one
  // that passes is still checked block by block while decoding (the lengths it records against the
  // lengths the parse finds): the same zfp_stream may have compressed another field of the same shape
  // since, or the buffer may have been refilled.  A stale index costs one wasted decode, never wrong data.
  const uint16_t* lengths;
  const bool use_index = index_matches(index, d, g, start_bit);
  if (d_status && !use_index) {
    g_error = "zfp_b200_decode_async: variable-rate parameters need the block index of this stream";
    return ZFP_B200_ENOINDEX;
  }
  if (use_index)
    lengths = index->d_lengths;
  else {
    // foreign stream: rebuild the index by parsing the stream sequentially on the device
    uint16_t* rebuilt = static_cast<uint16_t*>(scratch(SCR_LENGTHS, g.nblocks * sizeof(uint16_t)));
    if (!rebuilt) return ZFP_B200_ECUDA;
    const DecodeArgs a = { d_data, g, prm, d_words, start_bit, nullptr, nullptr, st, 0, 0, g.nblocks, nullptr };
    cudaError_t e;
    switch (d->type) {
      case T_INT32: e = launch_index_t<T_INT32>((int)d->dims, a, rebuilt); break;
      case T_INT64: e = launch_index_t<T_INT64>((int)d->dims, a, rebuilt); break;
      case T_FLOAT: e = launch_index_t<T_FLOAT>((int)d->dims, a, rebuilt); break;
      default: e = launch_index_t<T_DOUBLE>((int)d->dims, a, rebuilt); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!cuda_ok(e, "index scan launch")) return ZFP_B200_ECUDA;
    lengths = rebuilt;
  }
  uint64_t* tiles = static_cast<uint64_t*>(scratch(SCR_TILES, ((g.nblocks + kScanTile - 1) / kScanTile + 1) * 8));
  uint64_t* offsets = static_cast<uint64_t*>(scratch(SCR_OFFSETS, g.nblocks * 8));
  uint64_t* cursor = static_cast<uint64_t*>(scratch(SCR_CURSOR, 64));
  if (!tiles || !offsets || !cursor) return ZFP_B200_ECUDA;
  set_cursor<<<1, 1, 0, st>>>(cursor, start_bit);
  LAUNCHED();
  rc = scan_lengths(lengths, g.nblocks, tiles, offsets, cursor, st, block_capacity_bits(d));
  if (rc) return rc;
  rc = decode_any(1, d->type, d->dims, d_data, g, prm, d_words, start_bit, offsets, lengths, st, block0, block1,
                  d_status ? d_status : reinterpret_cast<uint32_t*>(cursor + 2));
  if (rc) return rc;
  if (d_status)  // stream-ordered: the caller looks at *d_status (0 = every block parsed to its indexed length) when it likes
    return ZFP_B200_OK;
  uint64_t h_cursor[3];
  CU(cudaMemcpyAsync(h_cursor, cursor, sizeof(h_cursor), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (h_cursor[2]) {
    if (!use_index) {
      g_error = "zfp_b200_decode: the stream does not parse to the lengths just rebuilt from it (corrupt stream?)";
      return ZFP_B200_EINVAL;
    }
    // stale index: rebuild from the stream itself and decode again
    return decode_range(d, d_data, d_words, start_bit, end_bit, nullptr, cuda_stream, block0, block1, whole);
  }
  if (end_bit) *end_bit = h_cursor[1];
  return ZFP_B200_OK;
}

extern "C" int zfp_b200_bitcopy(void* d_dst_words, uint64 dst_bit, const void* d_src_words, uint64 src_bit, uint64 nbits,
                                void* cuda_stream)
{
  if (!nbits) return ZFP_B200_OK;
  if (!d_dst_words || !d_src_words) return ZFP_B200_EINVAL;
  const uint64_t words = ((dst_bit + nbits + 63) >> 6) - (dst_bit >> 6);
  unsigned ctas = (unsigned)((words + 255) / 256);
  const unsigned cap = (unsigned)sm_count() * 16;
  if (ctas > cap) ctas = cap;
  bitcopy_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<uint64_t*>(d_dst_words), dst_bit,
                                                                           static_cast<const uint64_t*>(d_src_words), src_bit, nbits);
  LAUNCHED();
  return ZFP_B200_OK;
}

// Slab placement with the offsets still on the device: d_end_bits[r] = end bit of rank r's slab stream
// encoded at bit 0 of its own buffer (= its length), as all-gathered on this stream; slab `rank` goes to
// start_bit + sum of the lengths of the lower ranks.  No host round trip between the encode, the
// exchange and the placement.
__global__ void __launch_bounds__(256)
bitcopy_ranked_kernel(uint64_t* __restrict__ dst, uint64_t start_bit, const uint64_t* __restrict__ lens, uint32_t rank,
                      const uint64_t* __restrict__ src, uint64_t* __restrict__ d_base_out)
{
  uint64_t dst_bit = start_bit;
  for (uint32_t r = 0; r < rank; r++)
    dst_bit += lens[r];
  const uint64_t nbits = lens[rank];
  if (d_base_out && blockIdx.x == 0 && threadIdx.x == 0)
    *d_base_out = dst_bit;
  if (!dst)
    return;
  const uint64_t w0 = dst_bit >> 6, w1 = (dst_bit + nbits + 63) >> 6;
  for (uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t lo = w * 64 > dst_bit ? w * 64 : dst_bit;
    const uint64_t hi = (w + 1) * 64 < dst_bit + nbits ? (w + 1) * 64 : dst_bit + nbits;
    const uint32_t len = (uint32_t)(hi - lo);
    const uint64_t sb = lo - dst_bit;
    const uint32_t sh = (uint32_t)(sb & 63);
    uint64_t v = src[sb >> 6] >> sh;
    if (sh && sh + len > 64)
      v |= src[(sb >> 6) + 1] << (64 - sh);
    v &= len >= 64 ? ~0ull : ((1ull << len) - 1);
    v <<= (uint32_t)(lo & 63);
    if (len == 64)
      dst[w] = v;
    else
      atomicOr(reinterpret_cast<unsigned long long*>(dst + w), (unsigned long long)v);
  }
}

extern "C" int zfp_b200_bitcopy_ranked(void* d_dst_words, uint64 start_bit, const uint64* d_lengths, uint rank,
                                       const void* d_src_words, uint64* d_base_out, void* cuda_stream)
{
  if (!d_lengths || (d_dst_words && !d_src_words)) return ZFP_B200_EINVAL;
  const unsigned ctas = d_dst_words ? (unsigned)sm_count() * 8 : 1;
  bitcopy_ranked_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      static_cast<uint64_t*>(d_dst_words), start_bit, d_lengths, rank, static_cast<const uint64_t*>(d_src_words), d_base_out);
  LAUNCHED();
  return ZFP_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// Several GPUs driven by ONE process (SURVEY section 8e "process model"): slab i of the array lives on
// device i; the slabs are encoded concurrently, each on its device's stream, and the only exchange of
// the variable-rate path - one 64-bit slab length per device - is an ncclAllGather enqueued on those
// same streams, followed by the device-side prefix that gives every slab its place in the global
// stream.  Nothing returns to the host between the encode and the placement; one synchronisation at the
// end delivers the lengths.  NCCL is resolved at run time (dlopen of libnccl.so.2, the library torch
// ships or the system one), so libzfp_b200.so itself has no NCCL dependency.
// ------------------------------------------------------------------------------------------------
namespace {

typedef void* nccl_comm_t;
struct NcclApi {
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclUint64 = 5;  // ncclUint64 (nccl.h ncclDataType_t)

NcclApi* nccl_api()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.CommInitAll && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd;
  });
  return &api;
}

}  // namespace

struct zfp_b200_multi {
  int ndev = 0;
  int dev[16];
  nccl_comm_t comm[16];
  cudaStream_t stream[16];
  uint64_t* d_len[16];    // this device's slab length (bits)
  uint64_t* d_all[16];    // all slab lengths, gathered
  uint64_t* d_base[16];   // this device's slab base in the global stream
  uint64_t* h_pinned = nullptr;  // [2 * ndev]: lengths, bases
};

extern "C" void zfp_b200_multi_destroy(zfp_b200_multi* m)
{
  if (!m) return;
  int cur = 0;
  cudaGetDevice(&cur);
  NcclApi* nc = nccl_api();
  for (int i = 0; i < m->ndev; i++) {
    cudaSetDevice(m->dev[i]);
    if (m->stream[i]) cudaStreamSynchronize(m->stream[i]);
    if (m->comm[i] && nc->ok) nc->CommDestroy(m->comm[i]);
    if (m->d_len[i]) cudaFree(m->d_len[i]);
    if (m->stream[i]) cudaStreamDestroy(m->stream[i]);
  }
  if (m->h_pinned) cudaFreeHost(m->h_pinned);
  cudaSetDevice(cur);
  delete m;
}

extern "C" zfp_b200_multi* zfp_b200_multi_create(int ndev, const int* devices)
{
  NcclApi* nc = nccl_api();
  if (ndev < 1 || ndev > 16 || !devices) {
    g_error = "zfp_b200_multi_create: 1..16 devices";
    return nullptr;
  }
  if (!nc->ok) {
    g_error = "zfp_b200_multi_create: libnccl.so.2 not found (dlopen)";
    return nullptr;
  }
  zfp_b200_multi* m = new (std::nothrow) zfp_b200_multi();
  if (!m) return nullptr;
  m->ndev = ndev;
  for (int i = 0; i < 16; i++) {
    m->dev[i] = i < ndev ? devices[i] : -1;
    m->comm[i] = nullptr;
    m->stream[i] = nullptr;
    m->d_len[i] = m->d_all[i] = m->d_base[i] = nullptr;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  bool ok = true;
  const int rc = nc->CommInitAll(m->comm, ndev, m->dev);
  if (rc != 0) {
    g_error = std::string("ncclCommInitAll: ") + (nc->GetErrorString ? nc->GetErrorString(rc) : "failed");
    ok = false;
  }
  for (int i = 0; i < ndev && ok; i++) {
    ok = cuda_ok(cudaSetDevice(m->dev[i]), "cudaSetDevice") &&
         cuda_ok(cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking), "cudaStreamCreate") &&
         cuda_ok(cudaMalloc(&m->d_len[i], (size_t)(ndev + 2) * 8), "cudaMalloc(multi)");
    if (ok) {
      m->d_all[i] = m->d_len[i] + 1;
      m->d_base[i] = m->d_len[i] + 1 + ndev;
    }
  }
  ok = ok && cuda_ok(cudaMallocHost(&m->h_pinned, (size_t)ndev * 16), "cudaMallocHost(multi)");
  cudaSetDevice(cur);
  if (!ok) {
    zfp_b200_multi_destroy(m);
    return nullptr;
  }
  return m;
}

extern "C" int zfp_b200_multi_devices(const zfp_b200_multi* m) { return m ? m->ndev : 0; }
extern "C" void* zfp_b200_multi_stream(const zfp_b200_multi* m, int i) { return m && i >= 0 && i < m->ndev ? m->stream[i] : nullptr; }

extern "C" int zfp_b200_multi_compress(zfp_b200_multi* m, const zfp_b200_desc* descs, const void* const* d_slabs,
                                       void* const* d_words, zfp_b200_index* const* indexes, uint64* slab_bits, uint64* slab_base)
{
  if (!m || !descs || !d_slabs || !d_words) return ZFP_B200_EINVAL;
  NcclApi* nc = nccl_api();
  int cur = 0, rc = ZFP_B200_OK;
  cudaGetDevice(&cur);
  // 1. every slab encoded at bit 0 of its own buffer, length left in device memory
  for (int i = 0; i < m->ndev && rc == ZFP_B200_OK; i++) {
    if (!cuda_ok(cudaSetDevice(m->dev[i]), "cudaSetDevice")) rc = ZFP_B200_ECUDA;
    else rc = zfp_b200_encode_async(&descs[i], d_slabs[i], d_words[i], 0, m->d_len[i], indexes ? indexes[i] : nullptr, m->stream[i]);
  }
  // 2. the exchange, on the same streams
  if (rc == ZFP_B200_OK) {
    int e = nc->GroupStart();
    for (int i = 0; i < m->ndev && e == 0; i++)
      e = nc->AllGather(m->d_len[i], m->d_all[i], 1, kNcclUint64, m->comm[i], m->stream[i]);
    const int e2 = nc->GroupEnd();
    if (e != 0 || e2 != 0) {
      g_error = std::string("ncclAllGather: ") + (nc->GetErrorString ? nc->GetErrorString(e ? e : e2) : "failed");
      rc = ZFP_B200_ECUDA;
    }
  }
  // 3. device-side prefix: where each slab starts in the global stream; results to the host in one go
  for (int i = 0; i < m->ndev && rc == ZFP_B200_OK; i++) {
    if (!cuda_ok(cudaSetDevice(m->dev[i]), "cudaSetDevice")) { rc = ZFP_B200_ECUDA; break; }
    rc = zfp_b200_bitcopy_ranked(nullptr, 0, m->d_all[i], (uint)i, nullptr, m->d_base[i], m->stream[i]);
    if (rc == ZFP_B200_OK && (slab_bits || slab_base)) {
      if (!cuda_ok(cudaMemcpyAsync(m->h_pinned + i, m->d_len[i], 8, cudaMemcpyDeviceToHost, m->stream[i]), "D2H length") ||
          !cuda_ok(cudaMemcpyAsync(m->h_pinned + m->ndev + i, m->d_base[i], 8, cudaMemcpyDeviceToHost, m->stream[i]), "D2H base"))
        rc = ZFP_B200_ECUDA;
    }
  }
  for (int i = 0; i < m->ndev; i++) {
    cudaSetDevice(m->dev[i]);
    if (!cuda_ok(cudaStreamSynchronize(m->stream[i]), "sync") && rc == ZFP_B200_OK) rc = ZFP_B200_ECUDA;
  }
  cudaSetDevice(cur);
  if (rc == ZFP_B200_OK)
    for (int i = 0; i < m->ndev; i++) {
      if (slab_bits) slab_bits[i] = m->h_pinned[i];
      if (slab_base) slab_base[i] = m->h_pinned[m->ndev + i];
    }
  return rc;
}

extern "C" int zfp_b200_multi_decompress(zfp_b200_multi* m, const zfp_b200_desc* descs, void* const* d_slabs,
                                         const void* const* d_words, zfp_b200_index* const* indexes)
{
  if (!m || !descs || !d_slabs || !d_words) return ZFP_B200_EINVAL;
  int cur = 0, rc = ZFP_B200_OK;
  cudaGetDevice(&cur);
  // fixed-rate slabs are enqueued on all devices before anything waits; variable-rate ones synchronise their
  // own stream at the end of each call (index check), so they overlap only across the launches already queued
  for (int i = 0; i < m->ndev && rc == ZFP_B200_OK; i++) {
    if (!cuda_ok(cudaSetDevice(m->dev[i]), "cudaSetDevice")) rc = ZFP_B200_ECUDA;
    else rc = zfp_b200_decode(&descs[i], d_slabs[i], d_words[i], 0, nullptr, indexes ? indexes[i] : nullptr, m->stream[i]);
  }
  for (int i = 0; i < m->ndev; i++) {
    cudaSetDevice(m->dev[i]);
    if (!cuda_ok(cudaStreamSynchronize(m->stream[i]), "sync") && rc == ZFP_B200_OK) rc = ZFP_B200_ECUDA;
  }
  cudaSetDevice(cur);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// zfp_stream / zfp_field entry points (drop-in for src/cuda_zfp/cuZFP.h:9-10)
// ------------------------------------------------------------------------------------------------
static bool on_device(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static bool fill_desc(const zfp_stream* zfp, const zfp_field* f, zfp_b200_desc* d)
{
  memset(d, 0, sizeof(*d));
  d->type = (int)f->type;
  d->dims = f->nx ? f->ny ? f->nz ? f->nw ? 4 : 3 : 2 : 1 : 0;
  if (!d->dims) return false;
  d->n[0] = f->nx; d->n[1] = f->ny; d->n[2] = f->nz; d->n[3] = f->nw;
  d->s[0] = f->sx; d->s[1] = f->sy; d->s[2] = f->sz; d->s[3] = f->sw;
  d->minbits = zfp->minbits; d->maxbits = zfp->maxbits; d->maxprec = zfp->maxprec; d->minexp = zfp->minexp;
  return true;
}

// lowest and highest element index touched by the field (src/zfp.c field_index_span)
static void index_span(const zfp_b200_desc* d, int64_t* lo, int64_t* hi)
{
  Geom g;
  make_geom(d, nullptr, &g);
  *lo = *hi = 0;
  for (uint32_t i = 0; i < d->dims; i++) {
    int64_t reach = g.s[i] * (int64_t)(g.n[i] - 1);
    if (reach < 0) *lo += reach; else *hi += reach;
  }
}

static zfp_exec_params_cuda* get_cuda_params(const zfp_stream* zfp)
{
  if (zfp->exec.policy != zfp_exec_cuda || !zfp->exec.params) return nullptr;
  zfp_exec_params_cuda* p = static_cast<zfp_exec_params_cuda*>(zfp->exec.params);
  return p->magic == ZFP_B200_PARAMS_MAGIC ? p : nullptr;
}

// ------------------------------------------------------------------------------------------------
// Host-resident field AND host-resident stream at a fixed rate: the transfer is the cost (PCIe),
// so the array is cut into slabs of whole block layers along its slowest dimension and the slabs
// are pipelined over a few CUDA streams - H2D of slab i+1, the kernel of slab i and D2H of slab
// i-1 overlap (both copy engines busy, the kernels hidden).  Fixed-rate slabs sit at deterministic
// bit offsets; the path is taken when those are word aligned and the field is contiguous.
// Pinned host buffers give real overlap; pageable ones degrade to the serial order, still correct.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kPipeLanes = 3;
struct PipeStreams { cudaStream_t s[kPipeLanes] = { nullptr, nullptr, nullptr }; bool ok = false; };
PipeStreams g_pipe[64];

PipeStreams* pipe_streams()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  PipeStreams& p = g_pipe[dev];
  if (!p.ok) {
    for (int i = 0; i < kPipeLanes; i++)
      if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    p.ok = true;
  }
  return &p;
}

struct SlabPlan {
  uint32_t slow;            // index of the slowest dimension
  size_t rows_per_slab;     // extent of a slab along it (multiple of 4)
  size_t nslabs;
  size_t elems_per_row;     // values per unit of the slowest dimension
  uint64_t blocks_per_row4; // blocks per 4 rows
};

// decide whether (and how) to pipeline; false = use the monolithic path
bool plan_slabs(const zfp_b200_desc& d, uint64_t start_bit, size_t esize, SlabPlan* plan)
{
  if (d.minbits != d.maxbits || d.dims < 1 || (start_bit & 63)) return false;
  ptrdiff_t expect = 1;
  for (uint32_t i = 0; i < d.dims; i++) {  // default (contiguous) layout only, implied or spelled out
    if (d.s[i] != 0 && d.s[i] != expect) return false;
    expect *= (ptrdiff_t)d.n[i];
  }
  const uint32_t slow = d.dims - 1;
  size_t elems = 1;
  uint64_t blocks = 1;
  for (uint32_t i = 0; i < slow; i++) {
    elems *= d.n[i];
    blocks *= (d.n[i] + 3) / 4;
  }
  const size_t total_bytes = elems * d.n[slow] * esize;
  if (total_bytes < ((size_t)64 << 20)) return false;  // small arrays: not worth the extra launches
  // slab size: about 1/16 of the array, at least 16 MiB, whole block layers, word-aligned stream range
  size_t layers = (d.n[slow] + 3) / 4;
  size_t per = layers / 16 ? layers / 16 : 1;
  while (per < layers && per * 4 * elems * esize < ((size_t)16 << 20)) per++;
  while (per < layers && ((blocks * per * d.maxbits) & 63)) per++;
  if ((blocks * per * d.maxbits) & 63) return false;
  plan->slow = slow;
  plan->rows_per_slab = per * 4;
  plan->nslabs = (layers + per - 1) / per;
  plan->elems_per_row = elems;
  plan->blocks_per_row4 = blocks;
  return plan->nslabs >= 2;
}

// returns the number of stream bits produced / consumed, 0 on failure
uint64_t run_pipelined(const zfp_b200_desc& d, const SlabPlan& plan, size_t esize, char* h_data, uint64_t* h_words, bool encode,
                       cudaStream_t user)
{
  PipeStreams* ps = pipe_streams();
  if (!ps) return 0;
  const size_t total_rows = d.n[plan.slow];
  const size_t field_bytes = plan.elems_per_row * total_rows * esize;
  const uint64_t total_blocks = plan.blocks_per_row4 * ((total_rows + 3) / 4);
  const uint64_t total_bits = total_blocks * d.maxbits;
  const size_t words_bytes = (size_t)((total_bits + 63) >> 6) * 8;
  char* d_data = static_cast<char*>(scratch(SCR_STAGE_DATA, field_bytes));
  uint64_t* d_words = static_cast<uint64_t*>(scratch(SCR_STAGE_WORDS, words_bytes + 64));
  if (!d_data || !d_words) return 0;
  // everything already enqueued on the caller's stream comes first
  cudaEvent_t ready;
  if (!cuda_ok(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming), "event")) return 0;
  cudaEventRecord(ready, user);
  for (int i = 0; i < kPipeLanes; i++)
    cudaStreamWaitEvent(ps->s[i], ready, 0);
  bool ok = true;
  for (size_t i = 0; i < plan.nslabs && ok; i++) {
    cudaStream_t st = ps->s[i % kPipeLanes];
    const size_t row0 = i * plan.rows_per_slab;
    const size_t rows = row0 + plan.rows_per_slab <= total_rows ? plan.rows_per_slab : total_rows - row0;
    const size_t off = row0 * plan.elems_per_row * esize, bytes = rows * plan.elems_per_row * esize;
    const uint64_t bit0 = (uint64_t)(row0 / 4) * plan.blocks_per_row4 * d.maxbits;
    const uint64_t nbits = (uint64_t)((rows + 3) / 4) * plan.blocks_per_row4 * d.maxbits;
    const size_t w0 = (size_t)(bit0 >> 6), nw = (size_t)((nbits + 63) >> 6);
    zfp_b200_desc slab = d;
    slab.n[plan.slow] = rows;
    slab.s[0] = slab.s[1] = slab.s[2] = slab.s[3] = 0;
    uint64_t end = 0;
    if (encode) {
      ok = cuda_ok(cudaMemcpyAsync(d_data + off, h_data + off, bytes, cudaMemcpyHostToDevice, st), "H2D slab") &&
           zfp_b200_encode(&slab, d_data + off, d_words + w0, 0, &end, nullptr, st) == ZFP_B200_OK &&
           cuda_ok(cudaMemcpyAsync(h_words + w0, d_words + w0, nw * 8, cudaMemcpyDeviceToHost, st), "D2H slab stream");
    }
    else {
      ok = cuda_ok(cudaMemcpyAsync(d_words + w0, h_words + w0, nw * 8, cudaMemcpyHostToDevice, st), "H2D slab stream") &&
           zfp_b200_decode(&slab, d_data + off, d_words + w0, 0, &end, nullptr, st) == ZFP_B200_OK &&
           cuda_ok(cudaMemcpyAsync(h_data + off, d_data + off, bytes, cudaMemcpyDeviceToHost, st), "D2H slab");
    }
  }
  for (int i = 0; i < kPipeLanes; i++)
    ok = cuda_ok(cudaStreamSynchronize(ps->s[i]), "sync") && ok;
  cudaEventDestroy(ready);
  return ok ? total_bits : 0;
}

}  // namespace

static size_t compress_impl(zfp_stream* zfp, const zfp_field* field)
{
  zfp_b200_desc d;
  bitstream* s = zfp->stream;
  if (!s || !field->data || !fill_desc(zfp, field, &d)) return 0;
  zfp_exec_params_cuda* xp = get_cuda_params(zfp);
  cudaStream_t st = xp ? static_cast<cudaStream_t>(xp->cuda_stream) : nullptr;
  ScratchLease lease(st);
  const size_t esize = scalar_bytes(d.type);
  const uint64_t start_bit = (uint64_t)(s->ptr - s->begin) * 64 + s->bits;
  const uint64_t first_word = start_bit >> 6;

  // host field, host stream, fixed rate: slab pipeline over the copy engines
  SlabPlan plan;
  if (!on_device(field->data) && !on_device(s->begin) && !s->bits && !getenv("ZFP_B200_NO_PIPELINE") &&
      plan_slabs(d, start_bit, esize, &plan)) {
    const uint64_t bits = run_pipelined(d, plan, esize, static_cast<char*>(const_cast<void*>(field->data)), s->begin + first_word,
                                        true, st);
    if (!bits) return 0;
    s->ptr = s->begin + first_word + ((bits + 63) >> 6);
    s->bits = 0;
    s->buffer = 0;
    return (size_t)(s->ptr - s->begin) * 8;
  }

  // field data: use in place when device resident, else stage the touched span
  int64_t lo, hi;
  index_span(&d, &lo, &hi);
  const void* d_data = field->data;
  if (!on_device(field->data)) {
    const size_t bytes = (size_t)(hi - lo + 1) * esize;
    char* stage = static_cast<char*>(scratch(SCR_STAGE_DATA, bytes));
    if (!stage) return 0;
    if (!cuda_ok(cudaMemcpyAsync(stage, static_cast<const char*>(field->data) + lo * (int64_t)esize, bytes,
                                 cudaMemcpyHostToDevice, st), "H2D field")) return 0;
    d_data = stage - lo * (int64_t)esize;
  }

  // stream buffer: same
  const bool stream_dev = on_device(s->begin);
  const size_t cap_bytes = zfp_b200_capacity(&d, start_bit & 63) + 8;
  uint64_t* d_words;
  if (stream_dev)
    d_words = s->begin + first_word;
  else {
    d_words = static_cast<uint64_t*>(scratch(SCR_STAGE_WORDS, cap_bytes));
    if (!d_words) return 0;
  }
  // bits of a partially filled word (e.g. a header) still sit in the host-side buffer
  if (s->bits) {
    const uint64_t partial = s->buffer & ((1ull << s->bits) - 1);
    if (!cuda_ok(cudaMemcpyAsync(d_words, &partial, 8, cudaMemcpyHostToDevice, st), "H2D partial word")) return 0;
    if (!cuda_ok(cudaStreamSynchronize(st), "sync")) return 0;
  }

  zfp_b200_index* index = nullptr;
  if (d.minbits != d.maxbits && xp) {
    if (!xp->index) xp->index = zfp_b200_index_create();
    index = xp->index;
  }
  uint64_t end_rel = 0;
  if (zfp_b200_encode(&d, d_data, d_words, start_bit & 63, &end_rel, index, st) != ZFP_B200_OK) return 0;
  const uint64_t words_rel = (end_rel + 63) >> 6;

  if (!stream_dev) {
    if (!cuda_ok(cudaMemcpyAsync(s->begin + first_word, d_words, words_rel * 8, cudaMemcpyDeviceToHost, st), "D2H stream"))
      return 0;
  }
  if (!(xp && xp->device_only_sync && stream_dev && d.minbits == d.maxbits))
    if (!cuda_ok(cudaStreamSynchronize(st), "sync")) return 0;

  // leave the host bitstream flushed and positioned after the last word (cuZFP.cu:406-411)
  s->ptr = s->begin + first_word + words_rel;
  s->bits = 0;
  s->buffer = 0;
  return (size_t)(s->ptr - s->begin) * 8;
}

static size_t decompress_impl(zfp_stream* zfp, zfp_field* field)
{
  zfp_b200_desc d;
  bitstream* s = zfp->stream;
  if (!s || !field->data || !fill_desc(zfp, field, &d)) return 0;
  zfp_exec_params_cuda* xp = get_cuda_params(zfp);
  cudaStream_t st = xp ? static_cast<cudaStream_t>(xp->cuda_stream) : nullptr;
  ScratchLease lease(st);
  const size_t esize = scalar_bytes(d.type);
  const uint64_t start_bit = (uint64_t)(s->ptr - s->begin) * 64 - s->bits;
  const uint64_t first_word = start_bit >> 6;

  SlabPlan plan;
  if (!on_device(field->data) && !on_device(s->begin) && !getenv("ZFP_B200_NO_PIPELINE") && plan_slabs(d, start_bit, esize, &plan)) {
    const uint64_t bits = run_pipelined(d, plan, esize, static_cast<char*>(field->data), s->begin + first_word, false, st);
    if (!bits) return 0;
    s->ptr = s->begin + first_word + ((bits + 63) >> 6);
    s->bits = 0;
    s->buffer = 0;
    return (size_t)(s->ptr - s->begin) * 8;
  }

  int64_t lo, hi;
  index_span(&d, &lo, &hi);
  const bool data_dev = on_device(field->data);
  const size_t span_bytes = (size_t)(hi - lo + 1) * esize;
  void* d_data = field->data;
  char* stage = nullptr;
  if (!data_dev) {
    stage = static_cast<char*>(scratch(SCR_STAGE_DATA, span_bytes));
    if (!stage) return 0;
    // gaps of a strided host array must survive the round trip
    if ((uint64_t)(hi - lo + 1) != (uint64_t)(d.n[0] * (d.dims > 1 ? d.n[1] : 1) * (d.dims > 2 ? d.n[2] : 1) * (d.dims > 3 ? d.n[3] : 1)))
      if (!cuda_ok(cudaMemcpyAsync(stage, static_cast<char*>(field->data) + lo * (int64_t)esize, span_bytes,
                                   cudaMemcpyHostToDevice, st), "H2D field")) return 0;
    d_data = stage - lo * (int64_t)esize;
  }

  const bool stream_dev = on_device(s->begin);
  const uint64_t* d_words;
  if (stream_dev)
    d_words = s->begin + first_word;
  else {
    size_t want = zfp_b200_capacity(&d, start_bit & 63);
    const size_t have = (size_t)(s->end - (s->begin + first_word)) * 8;
    if (s->end > s->begin && want > have) want = have;
    uint64_t* w = static_cast<uint64_t*>(scratch(SCR_STAGE_WORDS, want + 16));
    if (!w) return 0;
    if (!cuda_ok(cudaMemcpyAsync(w, s->begin + first_word, want, cudaMemcpyHostToDevice, st), "H2D stream")) return 0;
    d_words = w;
  }

  const zfp_b200_index* index = (d.minbits != d.maxbits && xp) ? xp->index : nullptr;
  zfp_b200_index* own_index = nullptr;  // (a zfp_stream without CUDA parameters has no place to keep one)
  if (d.minbits != d.maxbits && s->end > s->begin + first_word && !getenv("ZFP_B200_SERIAL_INDEX")) {
    // a stream that did not come from this zfp_stream: rebuild a candidate index in parallel (the bit stream says
    // where its buffer ends); the decode verifies it and walks the stream sequentially if it does not hold
    Geom g;
    if (make_geom(&d, d_data, &g) && (!index_matches(index, &d, g, start_bit & 63) || index->speculative)) {
      zfp_b200_index* target;
      if (xp) {
        if (!xp->index) xp->index = zfp_b200_index_create();
        target = xp->index;
      }
      else
        target = own_index = zfp_b200_index_create();
      const size_t have = (size_t)(s->end - (s->begin + first_word)) * 8;
      size_t avail = zfp_b200_capacity(&d, start_bit & 63);
      if (avail > have) avail = have;
      index = (target && zfp_b200_index_rebuild(&d, d_words, start_bit & 63, avail, target, st) == ZFP_B200_OK) ? target : nullptr;
      g_error.clear();
    }
  }
  uint64_t end_rel = 0;
  const int decoded = zfp_b200_decode(&d, d_data, d_words, start_bit & 63, &end_rel, index, st);  // (variable rate: synchronous)
  zfp_b200_index_destroy(own_index);
  if (decoded != ZFP_B200_OK) return 0;

  if (!data_dev)
    if (!cuda_ok(cudaMemcpyAsync(static_cast<char*>(field->data) + lo * (int64_t)esize, stage, span_bytes,
                                 cudaMemcpyDeviceToHost, st), "D2H field")) return 0;
  if (!(xp && xp->device_only_sync && data_dev && d.minbits == d.maxbits))
    if (!cuda_ok(cudaStreamSynchronize(st), "sync")) return 0;

  s->ptr = s->begin + first_word + ((end_rel + 63) >> 6);
  s->bits = 0;
  s->buffer = 0;
  return (size_t)(s->ptr - s->begin) * 8;
}

extern "C" size_t zfp_b200_compress_stream(zfp_stream* stream, const zfp_field* field) { return compress_impl(stream, field); }
extern "C" size_t zfp_b200_decompress_stream(zfp_stream* stream, zfp_field* field) { return decompress_impl(stream, field); }
extern "C" size_t cuda_compress(zfp_stream* stream, const zfp_field* field) { return compress_impl(stream, field); }
extern "C" void cuda_decompress(zfp_stream* stream, zfp_field* field) { decompress_impl(stream, field); }
