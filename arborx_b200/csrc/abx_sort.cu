// abx_sort.cu -- hand-written onesweep LSD radix sort of (key, index) pairs and
// the exclusive scan used by the CRS path.
//
// Replaces thrust::sort_by_key in the reference's sortObjects
// (misc/ArborX_SortUtils.hpp:28-43, kokkos_ext/ArborX_KokkosExtSort.hpp:115-148).
// Stable: equal keys keep their original order, so the tree
// built on top is deterministic and bit-comparable with the oracle.
//
// Structure (one launch per digit + small launches up front):
//   0. prefixSampleKernel     2048 sampled keys -> how clustered the top bits are (sortPairsDB)
//   1. radixHistogramKernel   reads the keys once, builds the histograms of the digits to sort
//   2. radixScanHistKernel    exclusive scan of each 2^BITS-bin histogram
//   3. onesweepPassKernel     per digit: tiles rank their keys with a warp multisplit
//                             (8 ballots per key), obtain their global digit offsets with a
//                             decoupled look-back over per-tile digit counts, stage
//                             the tile in shared memory in sorted order and write
//                             it out in coalesced runs.
//   4. segmentFixKernel       when only the top digits were sorted (keys spread over many
//                             prefixes): ranks every key inside its short run of equal top
//                             bits, which finishes the stable sort of the full key with one
//                             more read + write instead of one per remaining digit.
// sortPairsDB picks the plan: top digits + fix-up (3 + 1 launches for 10M Morton64 keys),
// more top digits if the fix-up reports a run above its 256-key limit, or the plain LSD
// over all digits.  Every plan produces the same stable order.
#include "abx_common.cuh"

#include <cstdlib>

namespace abx
{

namespace
{

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

constexpr unsigned kFlagAgg = 1u << 30;
constexpr unsigned kFlagIncl = 1u << 31;
constexpr unsigned kValueMask = (1u << 30) - 1;
// see sortPairsImpl: tile-shape table.  Measured on B200, 10M pairs (scripts/tune_sort.py):
// MATCH.ANY ranking 1.14 ms (u64) / 0.55 ms (u32); 8 ballots per key 0.86 / 0.40 ms.
constexpr int kDefaultConfig64 = 7;
constexpr int kDefaultConfig32 = 7;

template <int BITS>
struct Radix
{
  static constexpr int kBins = 1 << BITS;
  static constexpr int kPerThread = (kBins + kSortThreads - 1) / kSortThreads;
  static_assert(kBins % kSortThreads == 0 || kBins < kSortThreads, "bins must tile the block");
};

template <typename KeyT>
__device__ __forceinline__ unsigned digitOf(KeyT k, int shift, unsigned mask)
{
  return (unsigned)(k >> shift) & mask;
}

__device__ __forceinline__ unsigned ldVolatile(unsigned const *p)
{
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stVolatile(unsigned *p, unsigned v)
{
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- 1. histograms of every digit in one read of the keys -------------------
// digit p of a key is (key >> shifts.s[p]) & (2^BITS - 1); windows may overlap (an LSD pass over
// a window that re-covers already sorted bits is harmless) which lets the driver place the
// digits wherever the key's significant bits are
struct SortShifts
{
  int s[8];
};

template <typename KeyT, int BITS, int MAXP>
__global__ void __launch_bounds__(kSortThreads)
    radixHistogramKernel(KeyT const *__restrict__ keys, int64_t n, SortShifts shifts, int npass,
                         unsigned *__restrict__ hist /*[npass][BINS]*/)
{
  constexpr int BINS = 1 << BITS;
  __shared__ unsigned sh[MAXP * BINS];
  for (int i = threadIdx.x; i < npass * BINS; i += kSortThreads)
    sh[i] = 0;
  __syncthreads();
  int64_t const stride = (int64_t)gridDim.x * kSortThreads;
  for (int64_t i = (int64_t)blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride)
  {
    KeyT k = keys[i];
#pragma unroll
    for (int p = 0; p < MAXP; ++p)
      if (p < npass)
        atomicAdd(&sh[p * BINS + digitOf(k, shifts.s[p], BINS - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * BINS; i += kSortThreads)
  {
    unsigned c = sh[i];
    if (c)
      atomicAdd(&hist[i], c);
  }
}

// block-wide exclusive scan of one value per thread (kSortThreads threads)
__device__ __forceinline__ unsigned blockExclusiveScan(unsigned v, unsigned *warp_sums /*[kSortWarps]*/,
                                                       unsigned &total)
{
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += t;
  }
  if (lane == 31)
    warp_sums[warp] = incl;
  __syncthreads();
  unsigned prefix = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w)
  {
    unsigned s = warp_sums[w];
    if (w < warp)
      prefix += s;
    tot += s;
  }
  total = tot;
  __syncthreads(); // warp_sums reusable
  return prefix + incl - v;
}

template <int WARPS>
__device__ __forceinline__ unsigned blockExclusiveScanT(unsigned v, unsigned *warp_sums /*[WARPS]*/, unsigned &total)
{
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += t;
  }
  if (lane == 31)
    warp_sums[warp] = incl;
  __syncthreads();
  unsigned prefix = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w)
  {
    unsigned s = warp_sums[w];
    if (w < warp)
      prefix += s;
    tot += s;
  }
  total = tot;
  __syncthreads();
  return prefix + incl - v;
}

// ---- 2. exclusive scan of each histogram -----------------------------------
template <int BITS>
__global__ void __launch_bounds__(kSortThreads) radixScanHistKernel(unsigned *__restrict__ hist)
{
  constexpr int BINS = 1 << BITS;
  constexpr int PER = BINS / kSortThreads;
  __shared__ unsigned warp_sums[kSortWarps];
  unsigned *h = hist + (size_t)blockIdx.x * BINS;
  unsigned v[PER];
  unsigned local = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i)
  {
    v[i] = h[threadIdx.x * PER + i];
    local += v[i];
  }
  unsigned total;
  unsigned excl = blockExclusiveScan(local, warp_sums, total);
#pragma unroll
  for (int i = 0; i < PER; ++i)
  {
    h[threadIdx.x * PER + i] = excl;
    excl += v[i];
  }
}

// ---- 3. one onesweep digit pass ----------------------------------------------
// Configurable tile shape: THREADS x ITEMS keys per tile, MINB resident blocks per SM
// (register cap).  BITS = 8: thread d < 256 owns digit d in the per-digit steps.
template <typename KeyT, int BITS, int THREADS, int ITEMS>
struct PassSmem
{
  static constexpr int BINS = 1 << BITS;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int TILE = THREADS * ITEMS;
  KeyT keys[TILE];
  unsigned vals[TILE];
  unsigned warp_hist[WARPS][BINS]; // per-warp digit counts, then per-warp offsets inside the digit
  unsigned digit_excl[BINS];       // first position of digit d in the tile's sorted order
  unsigned global_off[BINS];       // global position = global_off[d] + position in tile
  unsigned warp_sums[WARPS];
  unsigned tile;
};

// lanes of the warp holding the same digit: either one MATCH.ANY or BITS ballots
template <int BITS, bool BALLOT>
__device__ __forceinline__ unsigned matchDigit(unsigned d)
{
  if (!BALLOT)
    return __match_any_sync(0xffffffffu, d);
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < BITS; ++b)
  {
    bool const bit = (d >> b) & 1u;
    unsigned const bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// DEBUG != 0 variants give WRONG results on purpose; they exist to time the kernel with one
// phase removed (ABX_SORT_CONFIG=10..13, scripts/tune_sort.py): 1 no ranking chain, 2 no
// look-back, 3 no output stores, 4 no staging in shared memory
template <typename KeyT, int BITS, int THREADS, int ITEMS, int MINB, bool BALLOT, int DEBUG = 0>
__global__ void __launch_bounds__(THREADS, MINB)
    onesweepPassKernel(KeyT const *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                       unsigned const *__restrict__ vals_in /* may be null: iota */, unsigned *__restrict__ vals_out,
                       unsigned n, int shift, unsigned const *__restrict__ bin_base /*[BINS] exclusive*/,
                       unsigned *tile_state /*[tiles][BINS], zeroed*/, unsigned *tile_counter /*zeroed*/)
{
  using Smem = PassSmem<KeyT, BITS, THREADS, ITEMS>;
  constexpr int BINS = Smem::BINS;
  constexpr int WARPS = Smem::WARPS;
  constexpr int TILE = Smem::TILE;
  constexpr unsigned MASK = BINS - 1;
  static_assert(THREADS >= BINS, "one thread per digit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // tiles are taken in launch order so that a tile only ever waits on tiles that
  // have already started (forward progress of the look-back)
  if (tid == 0)
    sm.tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < WARPS * BINS; i += THREADS)
    (&sm.warp_hist[0][0])[i] = 0;
  __syncthreads();
  unsigned const tile = sm.tile;
  unsigned const tile_base = tile * (unsigned)TILE;
  unsigned const valid = min((unsigned)TILE, n - tile_base);

  // warp-striped load: element order inside the tile is warp*(32*ITEMS) + j*32 + lane
  KeyT key[ITEMS];
  unsigned val[ITEMS];
  unsigned const warp_base = warp * (32 * ITEMS);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
  {
    unsigned const local = warp_base + j * 32 + lane;
    bool const ok = local < valid;
    key[j] = ok ? keys_in[tile_base + local] : (KeyT)~(KeyT)0; // padding sorts to the very end of the tile
    val[j] = ok ? (vals_in ? vals_in[tile_base + local] : tile_base + local) : 0u;
  }

  // rank inside (warp, digit).  All ITEMS match-any votes are issued back to back
  // (independent), then the per-warp digit counters are updated in item order.
  unsigned peers[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    peers[j] = matchDigit<BITS, BALLOT>(digitOf(key[j], shift, MASK));
  unsigned short rank[ITEMS];
  unsigned const lanemask_lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
  {
    unsigned const d = digitOf(key[j], shift, MASK);
    if (DEBUG == 1)
    {
      rank[j] = (unsigned short)__popc(peers[j] & lanemask_lt);
      if (lane == 0)
        sm.warp_hist[warp][j] = 32;
      continue;
    }
    int const leader = __ffs(peers[j]) - 1;
    unsigned base = 0;
    if (lane == leader)
    {
      base = sm.warp_hist[warp][d];
      sm.warp_hist[warp][d] = base + __popc(peers[j]);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = (unsigned short)(base + __popc(peers[j] & lanemask_lt));
    __syncwarp();
  }
  __syncthreads();

  // per digit (thread d): offsets of each warp inside the digit and the tile total;
  // publish the tile's aggregate right away, look back later
  unsigned count = 0;
  unsigned *my_state = tile_state + (size_t)tile * BINS;
  if (tid < BINS)
  {
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
    {
      unsigned c = sm.warp_hist[w][tid];
      sm.warp_hist[w][tid] = count;
      count += c;
    }
    stVolatile(&my_state[tid], count | (tile == 0 ? kFlagIncl : kFlagAgg));
  }
  // position of each digit inside the tile's sorted order
  {
    unsigned total;
    unsigned excl = blockExclusiveScanT<WARPS>(tid < BINS ? count : 0u, sm.warp_sums, total);
    if (tid < BINS)
      sm.digit_excl[tid] = excl;
  }
  __syncthreads();

  // stage the tile in shared memory in sorted order (overlaps the predecessors'
  // progress: the look-back below rarely has to spin)
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
  {
    unsigned const d = digitOf(key[j], shift, MASK);
    unsigned pos = sm.digit_excl[d] + sm.warp_hist[warp][d] + rank[j];
    if (DEBUG != 0)
      pos = min(pos, (unsigned)TILE - 1);
    if (DEBUG == 4)
      continue;
    sm.keys[pos] = key[j];
    sm.vals[pos] = val[j];
  }

  // decoupled look-back for digit d
  if (tid < BINS)
  {
    unsigned excl = 0;
    if (tile != 0 && DEBUG != 2)
    {
      // windowed look-back: kWindow predecessor states are fetched with independent
      // loads and consumed in order, so a chain of W unresolved tiles (the first wave
      // of every pass) costs W / kWindow L2 round trips instead of W
      constexpr int kWindow = 8;
      int t = (int)tile - 1;
      bool done = false;
      while (!done)
      {
        unsigned v[kWindow];
#pragma unroll
        for (int i = 0; i < kWindow; ++i)
          v[i] = (t - i >= 0) ? ldVolatile(&tile_state[(size_t)(t - i) * BINS + tid]) : kFlagIncl;
#pragma unroll
        for (int i = 0; i < kWindow; ++i)
        {
          if (done)
            break;
          if (v[i] & kFlagIncl)
          {
            excl += v[i] & kValueMask;
            done = true;
          }
          else if (v[i] & kFlagAgg)
          {
            excl += v[i] & kValueMask;
            --t;
          }
          else
            break; // not published yet: re-fetch from this tile
        }
      }
      stVolatile(&my_state[tid], (excl + count) | kFlagIncl);
    }
    sm.global_off[tid] = bin_base[tid] + excl - sm.digit_excl[tid];
  }
  __syncthreads();

  // coalesced runs out
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
  {
    unsigned const pos = j * THREADS + tid;
    if (pos < valid)
    {
      KeyT const k = sm.keys[pos];
      unsigned g = sm.global_off[digitOf(k, shift, MASK)] + pos;
      if (DEBUG != 0)
        g = tile_base + pos; // keep the (wrong) stores in bounds
      if (DEBUG == 3)
      {
        if (g == 0xffffffffu)
          keys_out[0] = k;
        continue;
      }
      keys_out[g] = k;
      vals_out[g] = sm.vals[pos];
    }
  }
}

template <typename KeyT, int BITS, int THREADS, int ITEMS, int MINB, bool BALLOT, int DEBUG = 0>
abx_status launchPasses(cudaStream_t s, int passes, SortShifts const &shifts, KeyT *const keys[2],
                        unsigned *const vals[2], int &cur, int64_t n, bool iota_vals, unsigned *hist,
                        unsigned *counters, unsigned *states, int tiles)
{
  constexpr int BINS = 1 << BITS;
  auto kernel = onesweepPassKernel<KeyT, BITS, THREADS, ITEMS, MINB, BALLOT, DEBUG>;
  size_t const smem = sizeof(PassSmem<KeyT, BITS, THREADS, ITEMS>);
  static PerDeviceOnce attr; // function attributes are per device
  if (attr.needed())
  {
    ABX_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABX_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr.done();
  }
  for (int p = 0; p < passes; ++p)
  {
    ABX_LAUNCH_TAGGED(sizeof(KeyT) == 8 ? "onesweepPassKernel<u64>" : "onesweepPassKernel<u32>", kernel, tiles,
                      THREADS, smem, s, keys[cur], keys[cur ^ 1],
                      (p == 0 && iota_vals) ? (unsigned const *)nullptr : vals[cur], vals[cur ^ 1], (unsigned)n,
                      shifts.s[p], hist + (size_t)p * BINS, states + (size_t)p * tiles * BINS, counters + p);
    cur ^= 1;
  }
  return ABX_OK;
}

inline int sortConfig(bool wide)
{
  int const config = ABX_TUNE_INT("ABX_SORT_CONFIG", -1);
  return config >= 0 ? config : (wide ? kDefaultConfig64 : kDefaultConfig32);
}

// Stable LSD passes over the given digit windows (lowest first).  keys[cur]/vals[cur] hold the
// input (vals ignored when iota_vals) and, on return, the output; every pass flips cur.
// tile shapes: {threads, items, min blocks/SM}; ABX_SORT_CONFIG picks one (tuning aid)
template <typename KeyT>
abx_status runPasses(cudaStream_t s, KeyT *const keys[2], unsigned *const vals[2], int &cur, int64_t n,
                     bool iota_vals, SortShifts const &shifts, int passes)
{
  constexpr int BITS = 8;
  constexpr int BINS = 1 << BITS;
  constexpr int MAXP = 8;
  if (passes <= 0)
    return ABX_OK;
  int const cfg = sortConfig(sizeof(KeyT) == 8);
  int tile_keys;
  switch (cfg)
  {
  case 1: tile_keys = 256 * 8; break;
  case 2: tile_keys = 512 * 8; break;
  case 3: tile_keys = 384 * 12; break;
  case 4: tile_keys = 512 * 12; break;
  case 5: tile_keys = 256 * 16; break;
  case 6: tile_keys = 256 * 16; break;
  case 7: case 10: case 11: case 12: case 13: tile_keys = 384 * 12; break;
  case 8: tile_keys = 512 * 8; break;
  case 9: tile_keys = 256 * 8; break;
  default: tile_keys = 256 * 16; break;
  }
  int const tiles = divUp(n, tile_keys);
  TempBuffer<unsigned> ctrl; // [passes*BINS hist][passes counters][passes * tiles * BINS states]
  size_t const hist_words = (size_t)passes * BINS;
  size_t const ctrl_words = hist_words + passes + (size_t)passes * tiles * BINS;
  ABX_TRY(ctrl.alloc(ctrl_words, s));
  ABX_CUDA_TRY(cudaMemsetAsync(ctrl.ptr, 0, ctrl_words * sizeof(unsigned), s));
  unsigned *hist = ctrl.ptr;
  unsigned *counters = ctrl.ptr + hist_words;
  unsigned *states = counters + passes;

  int const hist_grid = (int)std::min<int64_t>(divUp(n, kSortThreads * 8), kNumSMs * 8);
  ABX_LAUNCH_TAGGED(sizeof(KeyT) == 8 ? "radixHistogramKernel<u64>" : "radixHistogramKernel<u32>",
                    (radixHistogramKernel<KeyT, BITS, MAXP>), hist_grid, kSortThreads, 0, s, keys[cur], n, shifts,
                    passes, hist);
  ABX_LAUNCH((radixScanHistKernel<BITS>), passes, kSortThreads, 0, s, hist);

#define ABX_PASSES(T, I, M, B, D)                                                                                     \
  return launchPasses<KeyT, BITS, T, I, M, B, D>(s, passes, shifts, keys, vals, cur, n, iota_vals, hist, counters,    \
                                                 states, tiles)
#ifdef ABX_TUNING
  // alternative tile shapes and the phase-removal variants (DEBUG != 0: wrong results on purpose,
  // scripts/tune_sort.py) exist in the tuning build only
  switch (cfg)
  {
  case 1: ABX_PASSES(256, 8, 5, false, 0);
  case 2: ABX_PASSES(512, 8, 2, false, 0);
  case 3: ABX_PASSES(384, 12, 2, false, 0);
  case 4: ABX_PASSES(512, 12, 2, false, 0);
  case 5: ABX_PASSES(256, 16, 3, false, 0);
  case 6: ABX_PASSES(256, 16, 2, true, 0);
  case 7: ABX_PASSES(384, 12, 2, true, 0);
  case 8: ABX_PASSES(512, 8, 2, true, 0);
  case 9: ABX_PASSES(256, 8, 5, true, 0);
  case 10: ABX_PASSES(384, 12, 2, true, 1);
  case 11: ABX_PASSES(384, 12, 2, true, 2);
  case 12: ABX_PASSES(384, 12, 2, true, 3);
  case 13: ABX_PASSES(384, 12, 2, true, 4);
  default: ABX_PASSES(256, 16, 2, false, 0);
  }
#else
  static_assert(kDefaultConfig64 == 7 && kDefaultConfig32 == 7, "release build compiles configuration 7 only");
  ABX_PASSES(384, 12, 2, true, 0);
#endif
#undef ABX_PASSES
}

// digit windows covering bits [lo, hi) of the key, lowest first; the top window is pulled down so
// that it ends exactly at hi (overlapping its predecessor) rather than spilling over dead bits
inline int coverBits(int lo, int hi, SortShifts &shifts)
{
  int const passes = (hi - lo + 7) / 8;
  for (int p = 0; p < passes; ++p)
    shifts.s[p] = std::max(lo, std::min(lo + 8 * p, hi - 8));
  if (hi - lo < 8)
    shifts.s[0] = std::max(0, hi - 8);
  return passes;
}

// ---- 4. segment fix-up: finishes a sort whose top bits are already in place ----
// After LSD passes over the top bits only, keys with the same prefix (key >> prefix_shift) form
// a contiguous run in original order.  For n keys spread over >= n/8 prefixes the runs are a
// handful of keys long, and ranking every key inside its run finishes the sort with one more
// read+write of the data instead of one per remaining digit (5 of the 8 digits of a Morton64
// key at 10M points).  Block b owns the runs that START in its tile of kFixTile positions; a run
// longer than kMaxRun raises *overflow (the driver then redoes the sort with more LSD digits),
// so the window a block needs is its tile plus kMaxRun keys.
constexpr int kFixThreads = 256;
constexpr int kFixItems = 8;
constexpr int kFixTile = kFixThreads * kFixItems;
constexpr int kMaxRun = 256;
static_assert(kMaxRun <= kFixThreads, "one thread per halo position");

template <typename KeyT>
__global__ void __launch_bounds__(kFixThreads)
    segmentFixKernel(KeyT const *__restrict__ keys_in, unsigned const *__restrict__ vals_in,
                     KeyT *__restrict__ keys_out, unsigned *__restrict__ vals_out, unsigned n, int prefix_shift,
                     unsigned *__restrict__ overflow)
{
  // sk[j] holds the key at position t0 - 1 + j
  __shared__ KeyT sk[kFixTile + kMaxRun + 1];
  unsigned const t0 = blockIdx.x * (unsigned)kFixTile;
  unsigned const t1 = min(n, t0 + (unsigned)kFixTile);
  unsigned const wend = min(n, t1 + (unsigned)kMaxRun); // window is [t0 - 1, wend)
  int const count = (int)(wend - t0) + 1;
  for (int j = threadIdx.x; j < count; j += kFixThreads)
  {
    KeyT k;
    if (j == 0 && t0 == 0)
      k = ~keys_in[0]; // position -1: a prefix that differs from key 0's
    else
      k = keys_in[t0 - 1 + j];
    sk[j] = k;
  }
  __syncthreads();
  int const last = count - 1; // sk index of the last loaded key
  // a run longer than kMaxRun that starts in this tile has two keys kMaxRun apart with the same
  // prefix inside the window: found with one comparison per key; the block then gives up at once
  // (the driver discards the output), instead of every key walking its run
  {
    bool long_run = false;
    for (int j = 1 + (int)threadIdx.x; j + kMaxRun <= last; j += kFixThreads)
      long_run |= (sk[j] >> prefix_shift) == (sk[j + kMaxRun] >> prefix_shift);
    if (__syncthreads_or(long_run))
    {
      if (threadIdx.x == 0)
        atomicExch(overflow, 1u);
      return;
    }
  }
  bool over = false;
  // j = sk index of the key this thread places; tile positions, then the halo
#pragma unroll 1
  for (int it = 0; it <= kFixItems; ++it)
  {
    int const j = 1 + it * kFixThreads + (int)threadIdx.x;
    bool const halo = it == kFixItems;
    if (j > last || (!halo && j > (int)(t1 - t0)))
      continue;
    KeyT const mine = sk[j];
    KeyT const pre = mine >> prefix_shift;
    // run start: walk back to the first key with this prefix
    int a = j;
    while (a > 0 && (sk[a - 1] >> prefix_shift) == pre && j - a <= kMaxRun)
      --a;
    if (a == 0)
      continue; // the run began in an earlier tile: that tile's block places it
    if (halo && a > (int)(t1 - t0))
      continue; // starts in the next tile
    if (!halo && j - a > kMaxRun)
    {
      over = true;
      continue;
    }
    int b = j + 1; // run end (exclusive)
    while (b <= last && (sk[b] >> prefix_shift) == pre && b - a <= kMaxRun)
      ++b;
    if (b - a > kMaxRun || (b > last && t0 - 1 + (unsigned)b < n))
    {
      over = true;
      continue;
    }
    int rank = 0;
    for (int i = a; i < b; ++i)
    {
      KeyT const o = sk[i];
      rank += (o < mine || (o == mine && i < j)) ? 1 : 0;
    }
    unsigned const src = t0 - 1 + (unsigned)j;
    unsigned const dst = t0 - 1 + (unsigned)(a + rank);
    keys_out[dst] = mine;
    vals_out[dst] = vals_in[src];
  }
  if (__syncthreads_or(over) && threadIdx.x == 0)
    atomicExch(overflow, 1u);
}

// Starting level of the fix-up path.  kSampleKeys keys taken at a regular stride are inserted
// into a shared-memory hash set by their prefix (key >> shift[i]); out[i] = samples whose prefix
// was already present.  Uniform-ish data gives ~0 at 10M keys, clustered data (most keys in few
// prefixes: every run would overflow the fix-up) gives hundreds; a handful of long runs is
// invisible here and is caught by the fix-up's own overflow flag.
constexpr int kSampleKeys = 2048;
constexpr int kSampleSlots = 4096;

template <typename KeyT>
__global__ void __launch_bounds__(1024)
    prefixSampleKernel(KeyT const *__restrict__ keys, unsigned n, int shift0, int shift1, unsigned *__restrict__ out)
{
  __shared__ unsigned long long table[kSampleSlots];
  __shared__ unsigned dup;
  unsigned const stride = max(1u, n / (unsigned)kSampleKeys);
  for (int level = 0; level < 2; ++level)
  {
    int const shift = level == 0 ? shift0 : shift1;
    for (int i = threadIdx.x; i < kSampleSlots; i += blockDim.x)
      table[i] = ~0ull;
    if (threadIdx.x == 0)
      dup = 0;
    __syncthreads();
    for (unsigned k = threadIdx.x; k < (unsigned)kSampleKeys && (unsigned long long)k * stride < n; k += blockDim.x)
    {
      unsigned long long const pre = (unsigned long long)(keys[(size_t)k * stride] >> shift);
      unsigned slot = (unsigned)((pre * 0x9E3779B97F4A7C15ull) >> 52) & (kSampleSlots - 1);
      while (true)
      {
        unsigned long long const old = atomicCAS(&table[slot], ~0ull, pre);
        if (old == ~0ull)
          break;
        if (old == pre)
        {
          atomicAdd(&dup, 1u);
          break;
        }
        slot = (slot + 1) & (kSampleSlots - 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
      out[level] = dup;
    __syncthreads();
  }
}

template <typename KeyT>
abx_status sortPairsDB(cudaStream_t s, KeyT *const keys[2], unsigned *const vals[2], int &cur, int64_t n,
                       bool iota_vals, int key_bits, int approx_top_bits, bool fixup)
{
  cur = 0;
  if (n <= 0)
    return ABX_OK;
  if (n >= (int64_t)kValueMask)
  {
    setError("sort: n must be < 2^30");
    return ABX_ERR_ARG;
  }
  int const width = (int)sizeof(KeyT) * 8;
  key_bits = std::max(1, std::min(key_bits, width));
  SortShifts shifts;
  if (approx_top_bits > 0)
  {
    // ordering hint only (predicate sorting): the top bits of the key decide
    int const passes = coverBits(std::max(0, key_bits - approx_top_bits), key_bits, shifts);
    return runPasses<KeyT>(s, keys, vals, cur, n, iota_vals, shifts, passes);
  }
  int const full = (key_bits + 7) / 8;
  // top digits such that the runs left to the fix-up average <= 8 keys
  int top = 1;
  while (top < full && ((int64_t)1 << (8 * top)) < n / 8)
    ++top;
  int const fix_mode = ABX_TUNE_INT("ABX_SORT_FIXUP", 1);
  TempBuffer<unsigned> flag;
  if (fixup && fix_mode && top + 1 < full && n >= 2 * kSampleKeys)
  {
    // pick the starting level from a sample (one small launch + one 8-byte read back)
    ABX_TRY(flag.alloc(2, s));
    int const lo0 = key_bits - 8 * top;
    int const lo1 = std::max(0, key_bits - 8 * (top + 2));
    ABX_LAUNCH_TAGGED("prefixSampleKernel", (prefixSampleKernel<KeyT>), 1, 1024, 0, s, keys[cur], (unsigned)n, lo0, lo1,
                      flag.ptr);
    unsigned dup[2] = {0, 0};
    ABX_CUDA_TRY(cudaMemcpyAsync(dup, flag.ptr, sizeof(dup), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    // two samples share a prefix with probability p => a key's run holds ~n*p keys; dup ~ m^2/2 * p.
    // Runs of a few keys on average are where the 256-key limit starts to be hit on clustered clouds
    // (GanTao at 10M: the 40-bit level sampled at ~40 keys per run and still overflowed).
    double const m = (double)std::min<int64_t>(kSampleKeys, n);
    unsigned const limit = (unsigned)std::max(2.0, 4.0 * m * m / 2.0 / (double)n);
    if (dup[0] > limit)
      top += 2;
    if (dup[0] > limit && dup[1] > limit)
      top = full; // plain LSD
  }
  while (fixup && fix_mode && top + 1 < full)
  {
    int const lo = key_bits - 8 * top;
    int const passes = coverBits(lo, key_bits, shifts);
    ABX_TRY(runPasses<KeyT>(s, keys, vals, cur, n, iota_vals, shifts, passes));
    iota_vals = false;
    if (!flag.ptr)
      ABX_TRY(flag.alloc(2, s));
    ABX_CUDA_TRY(cudaMemsetAsync(flag.ptr, 0, sizeof(unsigned), s));
    ABX_LAUNCH_TAGGED("segmentFixKernel", (segmentFixKernel<KeyT>), divUp(n, kFixTile), kFixThreads, 0, s, keys[cur],
                      vals[cur], keys[cur ^ 1], vals[cur ^ 1], (unsigned)n, lo, flag.ptr);
    unsigned over = 0;
    ABX_CUDA_TRY(cudaMemcpyAsync(&over, flag.ptr, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    if (!over)
    {
      cur ^= 1;
      return ABX_OK;
    }
    // some prefix holds more than kMaxRun keys (clustered data): the fix-up output is void,
    // keys[cur] is still sorted by its top bits in stable order; take two more digits
    top += 2;
  }
  int const passes = coverBits(0, key_bits, shifts);
  return runPasses<KeyT>(s, keys, vals, cur, n, iota_vals, shifts, passes);
}

// caller's buffers in and out
template <typename KeyT>
abx_status sortPairsInPlace(cudaStream_t s, KeyT *keys, unsigned *vals, int64_t n, bool iota_vals, int key_bits,
                            bool fixup)
{
  if (n <= 0)
    return ABX_OK;
  TempBuffer<KeyT> keys_alt;
  TempBuffer<unsigned> vals_alt;
  ABX_TRY(keys_alt.alloc(n, s));
  ABX_TRY(vals_alt.alloc(n, s));
  KeyT *k[2] = {keys, keys_alt.ptr};
  unsigned *v[2] = {vals, vals_alt.ptr};
  int cur = 0;
  ABX_TRY(sortPairsDB<KeyT>(s, k, v, cur, n, iota_vals, key_bits, 0, fixup));
  if (cur != 0)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(keys, keys_alt.ptr, sizeof(KeyT) * n, cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(vals, vals_alt.ptr, sizeof(unsigned) * n, cudaMemcpyDeviceToDevice, s));
  }
  return ABX_OK;
}

// ---- exclusive scan (reduce-then-scan, three small launches) ----------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
    scanTileSumsKernel(int32_t const *__restrict__ in, int64_t n, int32_t *__restrict__ tile_sums,
                       unsigned long long *__restrict__ total64 /* may be null; zeroed */)
{
  __shared__ int warp_sums[kScanThreads / 32];
  int64_t const base = (int64_t)blockIdx.x * kScanTile;
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    int64_t const i = base + j * kScanThreads + threadIdx.x;
    if (i < n)
      sum += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0)
    warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w)
      t += warp_sums[w];
    tile_sums[blockIdx.x] = t;
    // 64-bit total of the (non-negative) inputs: a tile of 2048 counts below 2^20 each cannot wrap
    if (total64 && t)
      atomicAdd(total64, (unsigned long long)(unsigned)t);
  }
}

// single block: exclusive scan of the tile sums in place
constexpr int kScanOffsetThreads = 1024;
__global__ void __launch_bounds__(kScanOffsetThreads) scanTileOffsetsKernel(int32_t *tile_sums, int tiles)
{
  __shared__ unsigned warp_sums[kScanOffsetThreads / 32];
  __shared__ unsigned carry_s;
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += kScanOffsetThreads)
  {
    int const i = base + threadIdx.x;
    unsigned v = i < tiles ? (unsigned)tile_sums[i] : 0u;
    unsigned total;
    unsigned excl = blockExclusiveScanT<kScanOffsetThreads / 32>(v, warp_sums, total);
    unsigned const carry = carry_s;
    if (i < tiles)
      tile_sums[i] = (int32_t)(carry + excl);
    __syncthreads();
    if (threadIdx.x == 0)
      carry_s = carry + total;
    __syncthreads();
  }
}

// out[i] = sum(in[0..i)) for i < n_plus_1 (the last input element is not read
// into the total, matching KokkosExt::exclusive_scan over an offsets array whose
// last slot is scratch: out[n] = total of in[0..n)).  A thread owns kScanItems consecutive
// elements: two 16-byte loads and stores when the arrays are 16-byte aligned and the
// thread's elements are all in range.
__global__ void __launch_bounds__(kScanThreads)
    scanApplyKernel(int32_t const *__restrict__ in, int32_t *__restrict__ out, int64_t n_in, int64_t n_out,
                    int32_t const *__restrict__ tile_offsets, bool aligned16)
{
  static_assert(kScanItems == 8, "two int4 per thread");
  __shared__ unsigned warp_sums[kSortWarps];
  int64_t const base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  bool const vec = aligned16 && base + kScanItems <= n_in;
  if (vec)
  {
    int4 const a = *reinterpret_cast<int4 const *>(in + base), b = *reinterpret_cast<int4 const *>(in + base + 4);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
  else
  {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      v[j] = base + j < n_in ? in[base + j] : 0;
  }
  unsigned local = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    local += (unsigned)v[j];
  unsigned total;
  unsigned excl = blockExclusiveScan(local, warp_sums, total) + (unsigned)tile_offsets[blockIdx.x];
  int o[kScanItems];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    o[j] = (int)excl;
    excl += (unsigned)v[j];
  }
  if (vec) // base + 8 <= n_in < n_out
  {
    *reinterpret_cast<int4 *>(out + base) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4 *>(out + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
  }
  else
  {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (base + j < n_out)
        out[base + j] = o[j];
  }
}

} // namespace

abx_status sortPairsU64(cudaStream_t s, uint64_t *keys, uint32_t *vals, int64_t n, bool iota_vals, int key_bits,
                        bool fixup)
{
  return sortPairsInPlace<unsigned long long>(s, (unsigned long long *)keys, vals, n, iota_vals, key_bits, fixup);
}
abx_status sortPairsU32(cudaStream_t s, uint32_t *keys, uint32_t *vals, int64_t n, bool iota_vals, int key_bits,
                        bool fixup)
{
  return sortPairsInPlace<unsigned>(s, keys, vals, n, iota_vals, key_bits, fixup);
}
abx_status sortPairsU64DB(cudaStream_t s, uint64_t *const keys[2], uint32_t *const vals[2], int *cur, int64_t n,
                          bool iota_vals, int key_bits)
{
  return sortPairsDB<unsigned long long>(s, (unsigned long long *const *)keys, vals, *cur, n, iota_vals, key_bits, 0,
                                         true);
}
abx_status sortPairsU32DB(cudaStream_t s, uint32_t *const keys[2], uint32_t *const vals[2], int *cur, int64_t n,
                          bool iota_vals, int key_bits, int approx_top_bits)
{
  return sortPairsDB<unsigned>(s, keys, vals, *cur, n, iota_vals, key_bits, approx_top_bits, true);
}

// out has n_plus_1 entries: out[i] = in[0] + ... + in[i-1]; in[n_plus_1-1] is ignored.
// total64 (optional, device): receives the sum of the inputs accumulated in 64 bits.
abx_status exclusiveScanI32(cudaStream_t s, int32_t const *in, int32_t *out, int64_t n_plus_1,
                            unsigned long long *total64)
{
  if (total64)
    ABX_CUDA_TRY(cudaMemsetAsync(total64, 0, sizeof(unsigned long long), s));
  if (n_plus_1 <= 0)
    return ABX_OK;
  int64_t const n_in = n_plus_1 - 1;
  int const tiles = divUp(n_plus_1, kScanTile);
  TempBuffer<int32_t> tile_sums;
  ABX_TRY(tile_sums.alloc(tiles, s));
  ABX_LAUNCH(scanTileSumsKernel, tiles, kScanThreads, 0, s, in, n_in, tile_sums.ptr, total64);
  ABX_LAUNCH(scanTileOffsetsKernel, 1, kScanOffsetThreads, 0, s, tile_sums.ptr, tiles);
  bool const aligned16 = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
  ABX_LAUNCH(scanApplyKernel, tiles, kScanThreads, 0, s, in, out, n_in, n_plus_1, tile_sums.ptr, aligned16);
  return ABX_OK;
}

} // namespace abx
