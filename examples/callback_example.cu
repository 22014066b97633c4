// Device callbacks over a tree built by libabx.so (include/ArborX_B200_Callbacks.cuh), checked against the
// CRS results of the same predicates.  Mirrors the reference's examples/callback (counting and early-exit
// callbacks) and the nearest-callback form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --extended-lambda -fmad=false -I include \
//        examples/callback_example.cu -o callback_example -L arborx_b200/lib -labx
#include <ArborX_B200_Callbacks.cuh>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                                       \
  do                                                                                                                   \
  {                                                                                                                    \
    if ((x) != ABX_OK)                                                                                                 \
    {                                                                                                                  \
      std::printf("FAILED %s: %s\n", #x, abx_last_error());                                                            \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

struct CountAndSum
{
  int *count;
  unsigned long long *sum;
  __device__ void operator()(int64_t q, unsigned value) const
  {
    atomicAdd(count + q, 1);
    atomicAdd(sum + q, (unsigned long long)value);
  }
};
struct FirstOnly
{
  int *count;
  __device__ abx::cb::Control operator()(int64_t q, unsigned) const
  {
    atomicAdd(count + q, 1);
    return abx::cb::Control::early_exit;
  }
};
// output form: for every match emit (value, squared distance to the sphere centre) -- a custom output type
struct Hit
{
  unsigned value;
  float d2;
};
struct EmitHit
{
  float const *spheres;
  float const *pts;
  template <class Out>
  __device__ void operator()(int64_t q, unsigned value, Out &out) const
  {
    float const dx = pts[3 * value] - spheres[4 * q], dy = pts[3 * value + 1] - spheres[4 * q + 1],
                dz = pts[3 * value + 2] - spheres[4 * q + 2];
    out(Hit{value, dx * dx + dy * dy + dz * dz});
    if (value % 7 == 0) // a callback may emit any number of results per match
      out(Hit{value, -1.f});
  }
};
struct NearestSum
{
  float *dist_sum;
  int *last;
  __device__ void operator()(int64_t q, unsigned value, float d) const
  {
    dist_sum[q] += d; // one thread per query, ascending order
    last[q] = (int)value;
  }
};

int main()
{
  int const side = 24, n = side * side * side, q = 5000, k = 7;
  std::vector<float> pts(3 * n), spheres(4 * q), qpts(3 * q);
  for (int i = 0; i < n; ++i)
  {
    pts[3 * i] = float(i % side);
    pts[3 * i + 1] = float((i / side) % side);
    pts[3 * i + 2] = float(i / (side * side));
  }
  unsigned state = 12345u;
  auto rnd = [&] { state = state * 1664525u + 1013904223u; return float(state >> 8) / float(1 << 24); };
  for (int i = 0; i < q; ++i)
  {
    for (int d = 0; d < 3; ++d)
      qpts[3 * i + d] = spheres[4 * i + d] = rnd() * float(side - 1);
    spheres[4 * i + 3] = 0.5f + 2.0f * rnd();
  }
  cudaStream_t s;
  cudaStreamCreate(&s);
  float *d_pts, *d_spheres, *d_qpts, *d_dsum;
  int *d_count, *d_first, *d_last;
  unsigned long long *d_sum;
  cudaMalloc(&d_pts, pts.size() * 4);
  cudaMalloc(&d_spheres, spheres.size() * 4);
  cudaMalloc(&d_qpts, qpts.size() * 4);
  cudaMalloc(&d_count, q * 4);
  cudaMalloc(&d_first, q * 4);
  cudaMalloc(&d_last, q * 4);
  cudaMalloc(&d_sum, q * 8);
  cudaMalloc(&d_dsum, q * 4);
  cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_spheres, spheres.data(), spheres.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_qpts, qpts.data(), qpts.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_count, 0, q * 4);
  cudaMemset(d_first, 0, q * 4);
  cudaMemset(d_sum, 0, q * 8);
  cudaMemset(d_dsum, 0, q * 4);

  abx_bvh *bvh = nullptr;
  CHECK(abx_bvh_build(s, ABX_PRIM_POINT3F, d_pts, n, &bvh));
  CHECK(abx::cb::query(bvh, s, abx::cb::intersects_spheres(d_spheres, q), CountAndSum{d_count, d_sum}));
  CHECK(abx::cb::query(bvh, s, abx::cb::intersects_spheres(d_spheres, q), FirstOnly{d_first}));
  CHECK(abx::cb::query(bvh, s, abx::cb::nearest(d_qpts, q, k), NearestSum{d_dsum, d_last}));
  // an extended lambda works as well
  int *d_total;
  cudaMalloc(&d_total, 4);
  cudaMemset(d_total, 0, 4);
  CHECK(abx::cb::query(bvh, s, abx::cb::intersects_spheres(d_spheres, q),
                       [=] __device__(int64_t, unsigned) { atomicAdd(d_total, 1); }));

  int32_t *hoff_dev = nullptr;
  Hit *hits_dev = nullptr;
  int64_t n_hits = 0;
  CHECK(abx::cb::query_crs<Hit>(bvh, s, abx::cb::intersects_spheres(d_spheres, q), EmitHit{d_spheres, d_pts}, &hoff_dev,
                                &hits_dev, &n_hits));

  // reference answers: CRS queries through the C ABI
  abx_policy pol = {0, 1};
  int32_t *off = nullptr, *koff = nullptr;
  uint32_t *idx = nullptr, *kidx = nullptr;
  float *kdist = nullptr;
  int64_t nnz = 0, knnz = 0;
  CHECK(abx_query_spatial_crs(bvh, s, ABX_PRED_SPHERE3F, d_spheres, q, &pol, nullptr, nullptr, &off, &idx, &nnz));
  CHECK(abx_query_nearest_crs(bvh, s, d_qpts, q, k, nullptr, &pol, nullptr, nullptr, &koff, &kidx, &kdist, &knnz));
  cudaStreamSynchronize(s);
  std::vector<int32_t> h_off(q + 1), h_koff(q + 1);
  std::vector<uint32_t> h_idx(nnz), h_kidx(knnz);
  std::vector<float> h_kdist(knnz), h_dsum(q);
  std::vector<int> h_count(q), h_first(q), h_last(q);
  std::vector<unsigned long long> h_sum(q);
  int h_total = 0;
  cudaMemcpy(h_off.data(), off, (q + 1) * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_idx.data(), idx, nnz * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_koff.data(), koff, (q + 1) * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_kidx.data(), kidx, knnz * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_kdist.data(), kdist, knnz * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_count.data(), d_count, q * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_first.data(), d_first, q * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_last.data(), d_last, q * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_sum.data(), d_sum, q * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_dsum.data(), d_dsum, q * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&h_total, d_total, 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < q; ++i)
  {
    unsigned long long want_sum = 0;
    for (int j = h_off[i]; j < h_off[i + 1]; ++j)
      want_sum += h_idx[j];
    int const want = h_off[i + 1] - h_off[i];
    float want_d = 0.f;
    for (int j = h_koff[i]; j < h_koff[i + 1]; ++j)
      want_d += h_kdist[j];
    bad += h_count[i] != want;
    bad += h_sum[i] != want_sum;
    bad += h_first[i] != (want > 0 ? 1 : 0);
    bad += h_dsum[i] != want_d;
    bad += h_last[i] != (int)h_kidx[h_koff[i + 1] - 1];
  }
  bad += h_total != (int)nnz;
  // output-form rows: same values as the CRS rows (as sets), plus the extra record for multiples of 7
  {
    std::vector<int32_t> h_hoff(q + 1);
    std::vector<Hit> h_hits(n_hits);
    cudaMemcpy(h_hoff.data(), hoff_dev, (q + 1) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_hits.data(), hits_dev, n_hits * sizeof(Hit), cudaMemcpyDeviceToHost);
    for (int i = 0; i < q; ++i)
    {
      unsigned long long sum = 0, want_sum = 0;
      int plain = 0, extra = 0, want_extra = 0;
      for (int j = h_hoff[i]; j < h_hoff[i + 1]; ++j)
      {
        if (h_hits[j].d2 < 0.f)
          ++extra;
        else
        {
          ++plain;
          sum += h_hits[j].value;
          float const r = spheres[4 * i + 3];
          bad += !(h_hits[j].d2 <= r * r * 1.0001f);
        }
      }
      for (int j = h_off[i]; j < h_off[i + 1]; ++j)
      {
        want_sum += h_idx[j];
        want_extra += h_idx[j] % 7 == 0;
      }
      bad += plain != h_off[i + 1] - h_off[i];
      bad += sum != want_sum;
      bad += extra != want_extra;
    }
    std::printf("output-form results %lld\n", (long long)n_hits);
  }
  std::printf("queries %d, matches %lld, mismatches %d\n", q, (long long)nnz, bad);
  abx_bvh_destroy(bvh);
  if (bad)
    return 1;
  std::printf("CALLBACKS OK\n");
  return 0;
}
