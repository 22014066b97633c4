// abx_sort.cu -- hand-written onesweep LSD radix sort of (key, index) pairs and
// the exclusive scan used by the CRS path.
//
// Replaces thrust::sort_by_key in the reference's sortObjects
// (misc/ArborX_SortUtils.hpp:28-43, kokkos_ext/ArborX_KokkosExtSort.hpp:115-148).
// Stable: equal keys keep their original order, so the tree
// built on top is deterministic and bit-comparable with the oracle.
//
// Structure (one launch per digit + two small launches up front):
//   1. radixHistogramKernel   reads the keys once, builds all per-digit histograms
//   2. radixScanHistKernel    exclusive scan of each 2^BITS-bin histogram
//   3. onesweepPassKernel     per digit: tiles rank their keys with warp match-any
//                             multisplit, obtain their global digit offsets with a
//                             decoupled look-back over per-tile digit counts, stage
//                             the tile in shared memory in sorted order and write
//                             it out in coalesced runs.
#include "abx_common.cuh"

namespace abx
{

namespace
{

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kSortThreads * kItems; // 4096 keys per tile

constexpr unsigned kFlagAgg = 1u << 30;
constexpr unsigned kFlagIncl = 1u << 31;
constexpr unsigned kValueMask = (1u << 30) - 1;

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
template <typename KeyT, int BITS, int PASSES>
__global__ void __launch_bounds__(kSortThreads)
    radixHistogramKernel(KeyT const *__restrict__ keys, int64_t n, unsigned *__restrict__ hist /*[PASSES][BINS]*/)
{
  constexpr int BINS = 1 << BITS;
  __shared__ unsigned sh[PASSES * BINS];
  for (int i = threadIdx.x; i < PASSES * BINS; i += kSortThreads)
    sh[i] = 0;
  __syncthreads();
  int64_t const stride = (int64_t)gridDim.x * kSortThreads;
  for (int64_t i = (int64_t)blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride)
  {
    KeyT k = keys[i];
#pragma unroll
    for (int p = 0; p < PASSES; ++p)
      atomicAdd(&sh[p * BINS + digitOf(k, p * BITS, BINS - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PASSES * BINS; i += kSortThreads)
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
template <typename KeyT, int BITS>
struct PassSmem
{
  static constexpr int BINS = 1 << BITS;
  KeyT keys[kTile];
  unsigned vals[kTile];
  unsigned warp_hist[kSortWarps][BINS]; // per-warp digit counts, then per-warp offsets inside the digit
  unsigned digit_excl[BINS];            // first position of digit d in the tile's sorted order
  unsigned global_off[BINS];            // global position = global_off[d] + position in tile
  unsigned warp_sums[kSortWarps];
  unsigned tile;
};

template <typename KeyT, int BITS>
__global__ void __launch_bounds__(kSortThreads)
    onesweepPassKernel(KeyT const *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                       unsigned const *__restrict__ vals_in /* may be null: iota */, unsigned *__restrict__ vals_out,
                       unsigned n, int shift, unsigned const *__restrict__ bin_base /*[BINS] exclusive*/,
                       unsigned *tile_state /*[tiles][BINS], zeroed*/, unsigned *tile_counter /*zeroed*/)
{
  constexpr int BINS = 1 << BITS;
  constexpr int PER = BINS / kSortThreads;
  constexpr unsigned MASK = BINS - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PassSmem<KeyT, BITS> &sm = *reinterpret_cast<PassSmem<KeyT, BITS> *>(smem_raw);

  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // tiles are taken in launch order so that a tile only ever waits on tiles that
  // have already started (forward progress of the look-back)
  if (tid == 0)
    sm.tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < kSortWarps * BINS; i += kSortThreads)
    (&sm.warp_hist[0][0])[i] = 0;
  __syncthreads();
  unsigned const tile = sm.tile;
  unsigned const tile_base = tile * (unsigned)kTile;
  unsigned const valid = min((unsigned)kTile, n - tile_base);

  // warp-striped load: element order inside the tile is warp*512 + j*32 + lane
  KeyT key[kItems];
  unsigned val[kItems];
  unsigned const warp_base = warp * (32 * kItems);
#pragma unroll
  for (int j = 0; j < kItems; ++j)
  {
    unsigned const local = warp_base + j * 32 + lane;
    bool const ok = local < valid;
    key[j] = ok ? keys_in[tile_base + local] : (KeyT)~(KeyT)0; // padding sorts to the very end of the tile
    val[j] = ok ? (vals_in ? vals_in[tile_base + local] : tile_base + local) : 0u;
  }

  // rank inside (warp, digit): match-any multisplit, one step per item
  unsigned short rank[kItems];
  unsigned const lanemask_lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kItems; ++j)
  {
    unsigned const d = digitOf(key[j], shift, MASK);
    unsigned const peers = __match_any_sync(0xffffffffu, d);
    int const leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (lane == leader)
    {
      base = sm.warp_hist[warp][d];
      sm.warp_hist[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[j] = (unsigned short)(base + __popc(peers & lanemask_lt));
    __syncwarp();
  }
  __syncthreads();

  // per digit: offsets of each warp inside the digit, tile total, look-back
  unsigned count[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i)
  {
    int const d = tid * PER + i;
    unsigned sum = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
    {
      unsigned c = sm.warp_hist[w][d];
      sm.warp_hist[w][d] = sum;
      sum += c;
    }
    count[i] = sum;
  }
  unsigned excl_tiles[PER];
  {
    unsigned *my_state = tile_state + (size_t)tile * BINS;
#pragma unroll
    for (int i = 0; i < PER; ++i)
    {
      int const d = tid * PER + i;
      if (tile == 0)
      {
        stVolatile(&my_state[d], count[i] | kFlagIncl);
        excl_tiles[i] = 0;
      }
      else
        stVolatile(&my_state[d], count[i] | kFlagAgg);
    }
    if (tile != 0)
    {
#pragma unroll
      for (int i = 0; i < PER; ++i)
      {
        int const d = tid * PER + i;
        unsigned excl = 0;
        int t = (int)tile - 1;
        while (true)
        {
          unsigned v = ldVolatile(&tile_state[(size_t)t * BINS + d]);
          if (v & kFlagIncl)
          {
            excl += v & kValueMask;
            break;
          }
          if (v & kFlagAgg)
          {
            excl += v & kValueMask;
            --t;
          }
        }
        excl_tiles[i] = excl;
        stVolatile(&my_state[d], (excl + count[i]) | kFlagIncl);
      }
    }
  }
  // position of each digit inside the tile's sorted order
  {
    unsigned local = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i)
      local += count[i];
    unsigned total;
    unsigned excl = blockExclusiveScan(local, sm.warp_sums, total);
#pragma unroll
    for (int i = 0; i < PER; ++i)
    {
      int const d = tid * PER + i;
      sm.digit_excl[d] = excl;
      sm.global_off[d] = bin_base[d] + excl_tiles[i] - excl;
      excl += count[i];
    }
  }
  __syncthreads();

  // stage the tile in shared memory in sorted order
#pragma unroll
  for (int j = 0; j < kItems; ++j)
  {
    unsigned const d = digitOf(key[j], shift, MASK);
    unsigned const pos = sm.digit_excl[d] + sm.warp_hist[warp][d] + rank[j];
    sm.keys[pos] = key[j];
    sm.vals[pos] = val[j];
  }
  __syncthreads();

  // coalesced runs out
#pragma unroll
  for (int j = 0; j < kItems; ++j)
  {
    unsigned const pos = j * kSortThreads + tid;
    if (pos < valid)
    {
      KeyT const k = sm.keys[pos];
      unsigned const g = sm.global_off[digitOf(k, shift, MASK)] + pos;
      keys_out[g] = k;
      vals_out[g] = sm.vals[pos];
    }
  }
}

template <typename KeyT, int BITS, int PASSES>
abx_status sortPairsImpl(cudaStream_t s, KeyT *keys, unsigned *vals, int64_t n, bool iota_vals)
{
  constexpr int BINS = 1 << BITS;
  static_assert(PASSES % 2 == 0, "ping-pong must end in the caller's buffers");
  if (n <= 0)
    return ABX_OK;
  if (n >= (int64_t)kValueMask)
  {
    setError("sort: n must be < 2^30");
    return ABX_ERR_ARG;
  }
  int const tiles = divUp(n, kTile);
  TempBuffer<KeyT> keys_alt;
  TempBuffer<unsigned> vals_alt;
  TempBuffer<unsigned> ctrl; // [PASSES*BINS hist][PASSES counters][PASSES * tiles * BINS states]
  size_t const hist_words = (size_t)PASSES * BINS;
  size_t const ctrl_words = hist_words + PASSES + (size_t)PASSES * tiles * BINS;
  ABX_TRY(keys_alt.alloc(n, s));
  ABX_TRY(vals_alt.alloc(n, s));
  ABX_TRY(ctrl.alloc(ctrl_words, s));
  ABX_CUDA_TRY(cudaMemsetAsync(ctrl.ptr, 0, ctrl_words * sizeof(unsigned), s));
  unsigned *hist = ctrl.ptr;
  unsigned *counters = ctrl.ptr + hist_words;
  unsigned *states = counters + PASSES;

  int const hist_grid = (int)std::min<int64_t>(divUp(n, kSortThreads * 8), kNumSMs * 8);
  ABX_LAUNCH_TAGGED(sizeof(KeyT) == 8 ? "radixHistogramKernel<u64>" : "radixHistogramKernel<u32>",
                    (radixHistogramKernel<KeyT, BITS, PASSES>), hist_grid, kSortThreads, 0, s, keys, n, hist);
  ABX_LAUNCH((radixScanHistKernel<BITS>), PASSES, kSortThreads, 0, s, hist);

  auto kernel = onesweepPassKernel<KeyT, BITS>;
  size_t const smem = sizeof(PassSmem<KeyT, BITS>);
  static bool attr_set = false;
  if (!attr_set)
  {
    ABX_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  KeyT *kin = keys, *kout = keys_alt.ptr;
  unsigned *vin = vals, *vout = vals_alt.ptr;
  for (int p = 0; p < PASSES; ++p)
  {
    ABX_LAUNCH_TAGGED(sizeof(KeyT) == 8 ? "onesweepPassKernel<u64>" : "onesweepPassKernel<u32>", kernel, tiles,
                      kSortThreads, smem, s, kin, kout, (p == 0 && iota_vals) ? (unsigned const *)nullptr : vin,
               vout, (unsigned)n, p * BITS, hist + (size_t)p * BINS, states + (size_t)p * tiles * BINS, counters + p);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  return ABX_OK;
}

// ---- exclusive scan (reduce-then-scan, three small launches) ----------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
    scanTileSumsKernel(int32_t const *__restrict__ in, int64_t n, int32_t *__restrict__ tile_sums)
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
  }
}

// single block: exclusive scan of the tile sums in place
__global__ void __launch_bounds__(kScanThreads) scanTileOffsetsKernel(int32_t *tile_sums, int tiles)
{
  __shared__ unsigned warp_sums[kSortWarps];
  __shared__ unsigned carry_s;
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += kScanThreads)
  {
    int const i = base + threadIdx.x;
    unsigned v = i < tiles ? (unsigned)tile_sums[i] : 0u;
    unsigned total;
    unsigned excl = blockExclusiveScan(v, warp_sums, total);
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
// last slot is scratch: out[n] = total of in[0..n))
__global__ void __launch_bounds__(kScanThreads)
    scanApplyKernel(int32_t const *__restrict__ in, int32_t *__restrict__ out, int64_t n_in, int64_t n_out,
                    int32_t const *__restrict__ tile_offsets)
{
  __shared__ unsigned warp_sums[kSortWarps];
  int64_t const base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  unsigned local = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    int64_t const i = base + j;
    v[j] = i < n_in ? in[i] : 0;
    local += (unsigned)v[j];
  }
  unsigned total;
  unsigned excl = blockExclusiveScan(local, warp_sums, total) + (unsigned)tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    int64_t const i = base + j;
    if (i < n_out)
      out[i] = (int32_t)excl;
    excl += (unsigned)v[j];
  }
}

} // namespace

abx_status sortPairsU64(cudaStream_t s, uint64_t *keys, uint32_t *vals, int64_t n, bool iota_vals)
{
  return sortPairsImpl<unsigned long long, 8, 8>(s, (unsigned long long *)keys, vals, n, iota_vals);
}
abx_status sortPairsU32(cudaStream_t s, uint32_t *keys, uint32_t *vals, int64_t n, bool iota_vals)
{
  return sortPairsImpl<unsigned, 8, 4>(s, keys, vals, n, iota_vals);
}

// out has n_plus_1 entries: out[i] = in[0] + ... + in[i-1]; in[n_plus_1-1] is ignored.
abx_status exclusiveScanI32(cudaStream_t s, int32_t const *in, int32_t *out, int64_t n_plus_1)
{
  if (n_plus_1 <= 0)
    return ABX_OK;
  int64_t const n_in = n_plus_1 - 1;
  int const tiles = divUp(n_plus_1, kScanTile);
  TempBuffer<int32_t> tile_sums;
  ABX_TRY(tile_sums.alloc(tiles, s));
  ABX_LAUNCH(scanTileSumsKernel, tiles, kScanThreads, 0, s, in, n_in, tile_sums.ptr);
  ABX_LAUNCH(scanTileOffsetsKernel, 1, kScanThreads, 0, s, tile_sums.ptr, tiles);
  ABX_LAUNCH(scanApplyKernel, tiles, kScanThreads, 0, s, in, out, n_in, n_plus_1, tile_sums.ptr);
  return ABX_OK;
}

} // namespace abx
