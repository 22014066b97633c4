// Device callbacks of include/ArborX_B200_Callbacks.cuh on caller-supplied inputs, results written to a file that
// tests/test_callbacks_oracle.py compares with the CPU oracle:
//   attach(predicates, data)        callback(data, value)                       -> matches per attached id
//   query_per_thread                one query per thread of a user kernel        -> matches per query
//   ordered_intersects_rays         callback(query, value, distance), all hits   -> the visiting order per ray
//   ordered_intersects_rays         the same with early exit after the first hit -> first hit per ray
//   callback_check in.bin out.bin
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

constexpr int kCap = 64;

struct Tag
{
  int id;
};
struct CountByTag
{
  int *count;
  __device__ void operator()(Tag const &t, unsigned) const { atomicAdd(count + t.id, 1); }
};
__global__ void perThreadKernel(abx::cb::DeviceTree tree, float const *spheres, int q, int *counts)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q)
    return;
  int c = 0;
  auto pred = abx::cb::make_sphere(spheres[4 * i], spheres[4 * i + 1], spheres[4 * i + 2], spheres[4 * i + 3]);
  abx::cb::query_per_thread(tree, pred, [&](unsigned) { ++c; });
  counts[i] = c;
}
struct RecordAll
{
  int *count;
  unsigned *vals;
  float *dist;
  __device__ void operator()(int64_t q, unsigned value, float d) const
  {
    int const j = count[q]++; // one thread per query
    if (j < kCap)
    {
      vals[q * kCap + j] = value;
      dist[q * kCap + j] = d;
    }
  }
};
struct FirstHit
{
  int *val;
  float *dist;
  __device__ abx::cb::Control operator()(int64_t q, unsigned value, float d) const
  {
    val[q] = (int)value;
    dist[q] = d;
    return abx::cb::Control::early_exit;
  }
};

template <class T>
static T *upload(std::vector<T> const &h)
{
  T *d = nullptr;
  cudaMalloc(&d, sizeof(T) * (h.size() ? h.size() : 1));
  cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
  return d;
}
template <class T>
static std::vector<T> download(T const *d, size_t n)
{
  std::vector<T> h(n);
  cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost);
  return h;
}

int main(int argc, char **argv)
{
  if (argc < 3)
    return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f)
    return 2;
  int hdr[4];
  if (std::fread(hdr, sizeof(int), 4, f) != 4)
    return 2;
  int const n = hdr[0], kind = hdr[1], qs = hdr[2], qr = hdr[3];
  int const stride = kind == ABX_PRIM_POINT3F ? 3 : 6;
  std::vector<float> prims((size_t)n * stride), spheres((size_t)qs * 4), rays((size_t)qr * 6);
  std::vector<int> tags(qs);
  if (std::fread(prims.data(), sizeof(float), prims.size(), f) != prims.size() ||
      std::fread(spheres.data(), sizeof(float), spheres.size(), f) != spheres.size() ||
      std::fread(rays.data(), sizeof(float), rays.size(), f) != rays.size() ||
      std::fread(tags.data(), sizeof(int), tags.size(), f) != tags.size())
    return 2;
  std::fclose(f);
  float *d_prims = upload(prims), *d_spheres = upload(spheres), *d_rays = upload(rays);
  std::vector<Tag> htags(qs);
  for (int i = 0; i < qs; ++i)
    htags[i].id = tags[i];
  Tag *d_tags = upload(htags);
  cudaStream_t s;
  cudaStreamCreate(&s);
  abx_bvh *bvh = nullptr;
  CHECK(abx_bvh_build(s, kind, d_prims, n, &bvh));

  int *d_attach = upload(std::vector<int>(qs, 0));
  CHECK(abx::cb::query(bvh, s, abx::cb::attach(abx::cb::intersects_spheres(d_spheres, qs), d_tags), CountByTag{d_attach}));

  int *d_pt = upload(std::vector<int>(qs, -1));
  abx::cb::DeviceTree tree;
  CHECK(abx::cb::device_tree(bvh, &tree));
  if (qs > 0)
    perThreadKernel<<<(qs + 127) / 128, 128, 0, s>>>(tree, d_spheres, qs, d_pt);

  int *d_ocount = upload(std::vector<int>(qr, 0));
  unsigned *d_ovals = upload(std::vector<unsigned>((size_t)qr * kCap, 0u));
  float *d_odist = upload(std::vector<float>((size_t)qr * kCap, 0.f));
  CHECK(abx::cb::query(bvh, s, abx::cb::ordered_intersects_rays(d_rays, qr), RecordAll{d_ocount, d_ovals, d_odist}));
  int *d_fval = upload(std::vector<int>(qr, -1));
  float *d_fdist = upload(std::vector<float>(qr, -1.f));
  CHECK(abx::cb::query(bvh, s, abx::cb::ordered_intersects_rays(d_rays, qr), FirstHit{d_fval, d_fdist}));
  if (cudaStreamSynchronize(s) != cudaSuccess)
  {
    std::printf("FAILED: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  FILE *o = std::fopen(argv[2], "wb");
  auto put = [&](auto const &v) { std::fwrite(v.data(), sizeof(v[0]), v.size(), o); };
  put(download(d_attach, qs));
  put(download(d_pt, qs));
  put(download(d_ocount, qr));
  put(download(d_ovals, (size_t)qr * kCap));
  put(download(d_odist, (size_t)qr * kCap));
  put(download(d_fval, qr));
  put(download(d_fdist, qr));
  std::fclose(o);
  abx_bvh_destroy(bvh);
  std::printf("CALLBACK CHECK WRITTEN\n");
  return 0;
}
