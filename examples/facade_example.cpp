// Facade smoke program: the reference's examples/simple_intersection and
// examples/dbscan/example_dbscan.cpp rewritten against include/ArborX_B200.hpp.
// Built by tests/test_facade.py (g++ host compile, links libabx.so + cudart); run on the GPU box.
#include <ArborX_B200.hpp>

#include <cstdio>
#include <vector>

int main()
{
  using namespace ArborX;
  Cuda space;
  // test/tstQueryTreeDegenerate.cpp:335-369 (duplicated leaves)
  std::vector<Box<>> boxes(4);
  boxes[0] = Box<>{{{0, 0, 0}}, {{0, 0, 0}}};
  for (int i = 1; i < 4; ++i)
    boxes[i] = Box<>{{{1, 1, 1}}, {{1, 1, 1}}};
  DeviceView<Box<>> d_boxes;
  d_boxes.assign(boxes);
  BoundingVolumeHierarchy bvh(space, d_boxes);
  std::vector<Intersects<Sphere<>>> preds = {intersects(Sphere<>{{{0, 0, 0}}, 1.f}), intersects(Sphere<>{{{1, 1, 1}}, 1.f}),
                                            intersects(Sphere<>{{{.5f, .5f, .5f}}, 1.f})};
  DeviceView<Intersects<Sphere<>>> d_preds;
  d_preds.assign(preds);
  DeviceView<int> indices, offsets;
  query(bvh, space, d_preds, indices, offsets);
  auto off = offsets.to_host();
  std::printf("offsets:");
  for (int o : off)
    std::printf(" %d", o);
  std::printf("\n");
  bool ok = off.size() == 4 && off[0] == 0 && off[1] == 1 && off[2] == 4 && off[3] == 8;

  // examples/dbscan/example_dbscan.cpp:33-90 (z = 0)
  std::vector<Point<>> cloud = {{{4, 3, 0}}, {{0, 0, 0}}, {{0, 1, 0}}, {{1, 1, 0}}, {{1, 0, 0}},
                                {{3, 3, 0}}, {{3, 4, 0}}, {{4, 4, 0}}, {{4, 0, 0}}, {{2, 2, 0}}};
  DeviceView<Point<>> d_cloud;
  d_cloud.assign(cloud);
  DeviceView<int> labels;
  dbscan(space, d_cloud, 1.0, 2, labels);
  auto l = labels.to_host();
  std::printf("labels:");
  for (int v : l)
    std::printf(" %d", v);
  std::printf("\n");
  ok = ok && l[1] == l[2] && l[2] == l[3] && l[3] == l[4] && l[0] == l[5] && l[5] == l[6] && l[6] == l[7] &&
       l[0] != l[1] && l[8] == -1 && l[9] == -1;

  // nearest
  DeviceView<Point<>> q;
  q.assign({{{0.1f, 0.f, 0.f}}, {{3.9f, 3.9f, 0.f}}});
  BoundingVolumeHierarchy tree(space, d_cloud);
  DeviceView<float> dist;
  tree.query(space, q, 2, indices, offsets, &dist);
  auto ki = indices.to_host();
  std::printf("knn: %d %d | %d %d\n", ki[0], ki[1], ki[2], ki[3]);
  ok = ok && ki[0] == 1 && ki[2] == 7;
  // the same queries through BruteForce (spatial/ArborX_BruteForce.hpp) and with nearest(Sphere, k)
  BruteForce brute(space, d_cloud);
  DeviceView<int> bi, bo;
  brute.query(space, q, 2, bi, bo, &dist);
  auto bki = bi.to_host();
  ok = ok && brute.size() == 10 && bki[0] == 1 && bki[2] == 7;
  DeviceView<Sphere<>> qs;
  qs.assign({{{{0.1f, 0.f, 0.f}}, 0.05f}, {{{3.9f, 3.9f, 0.f}}, 0.05f}});
  tree.query(space, qs, 2, indices, offsets, &dist);
  auto si = indices.to_host();
  auto sd = dist.to_host();
  std::printf("knn(sphere): %d %d | %d %d  d0 = %g\n", si[0], si[1], si[2], si[3], sd[0]);
  ok = ok && si[0] == 1 && si[2] == 7 && sd[0] > 0.04f && sd[0] < 0.06f;
  bool threw = false;
  try
  {
    dbscan(space, d_cloud, -1.0, 2, labels);
  }
  catch (SearchException const &)
  {
    threw = true;
  }
  ok = ok && threw;
  // MinimumSpanningTree / hdbscan on the equidistant points of tstMinimumSpanningTree.cpp:90-124 (k = 3:
  // weights 2, 1, 1, 2)
  DeviceView<Point<>> line;
  line.assign({{{0.f, 0.f, 0.f}}, {{1.f, 0.f, 0.f}}, {{2.f, 0.f, 0.f}}, {{3.f, 0.f, 0.f}}, {{4.f, 0.f, 0.f}}});
  Experimental::MinimumSpanningTree mst(space, line, 3);
  auto mw = mst.weights.to_host();
  float total = 0.f;
  for (float w : mw)
    total += w;
  std::printf("mst(k = 3): %zu edges, total weight %g, %d rounds\n", mw.size(), total, mst.iterations);
  ok = ok && mw.size() == 4 && total == 6.f;
  auto dendrogram = Experimental::hdbscan(space, line, 3, Experimental::DendrogramImplementation::UNION_FIND);
  auto dp = dendrogram._parents.to_host();
  auto dh = dendrogram._parent_heights.to_host();
  ok = ok && dp.size() == 9 && dh.size() == 4 && dp[3] == -1 && dh[0] == 1.f && dh[3] == 2.f;
  auto hybrid = Experimental::hdbscan(space, line, 3); // BORUVKA: the dendrogram grows with the rounds, on the device
  auto hp = hybrid._parents.to_host();
  int roots = 0;
  for (int i = 0; i < 4; ++i)
    roots += hp[i] == -1;
  ok = ok && hp.size() == 9 && roots == 1;
  std::printf(ok ? "FACADE OK\n" : "FACADE FAILED\n");
  return ok ? 0 : 1;
}
