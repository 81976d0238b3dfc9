// Host-side proof that the region-binned decomposition is exact: for every test ray the voxel sequence (and the
// enter/exit ranges) produced by enumerateSegments + resumeSegment equals, bit for bit, the sequence of the
// sequential walk (walkLine = ohm/LineWalkCompute.h:345-413).  Compiled by nvcc for the host; needs no GPU.
#include "ohmb200_regions.cuh"

#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace ohmb200;

struct Visit
{
  int r[3], l[3];
  double enter, exit;
};

static Geom makeGeom(double res, int dx, int dy, int dz, double ox, double oy, double oz)
{
  Geom g;
  g.res = res;
  g.dim[0] = dx;
  g.dim[1] = dy;
  g.dim[2] = dz;
  g.origin[0] = ox;
  g.origin[1] = oy;
  g.origin[2] = oz;
  g.vpr = (uint32_t)(dx * dy * dz);
  for (int a = 0; a < 3; ++a)
  {
    g.region_size[a] = g.dim[a] * res;
  }
  return g;
}

static long long g_rays = 0, g_visits = 0, g_segments = 0;

static bool checkRay(const Geom &g, const double start[3], const double end[3], unsigned walk_flags)
{
  Key skey, ekey;
  if (!voxelKey(g, start, skey) || !voxelKey(g, end, ekey))
  {
    return true;
  }
  std::vector<Visit> seq;
  walkLine(g, start, end, skey, ekey, walk_flags, [&](const Key &k, double enter, double exit) {
    Visit v;
    memcpy(v.r, k.r, sizeof(v.r));
    memcpy(v.l, k.l, sizeof(v.l));
    v.enter = enter;
    v.exit = exit;
    seq.push_back(v);
  });
  RayRec rec;
  if (!makeRayRec(rec, g, start, end, walk_flags))
  {
    return true;
  }
  const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
  const double len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const double length = (len2 > 1e-6) ? sqrt(len2) : 0;
  std::vector<Visit> seg;
  std::vector<uint32_t> fast_idx, tile_idx;
  const TileLayout tl = makeTileLayout(g);
  bool tile_ok = true;
  int last_flags = 0;
  enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
    ++g_segments;
    const int total[3] = { rec.total[0], rec.total[1], rec.total[2] };
    const int local0[3] = { rec.local[0], rec.local[1], rec.local[2] };
    resumeSegment<true>(rec.initial, rec.delta, local0, total, rec.flags, st, n, length, g,
                        [&](const int l[3], double enter, double exit, bool last_of_ray) {
                          Visit v;
                          memcpy(v.r, r, sizeof(v.r));
                          memcpy(v.l, l, sizeof(v.l));
                          v.enter = enter;
                          v.exit = exit;
                          seg.push_back(v);
                          last_flags += last_of_ray ? 1 : 0;
                        });
    // the hot-path variant must visit the same voxels (as linear indices inside the region)
    resumeSegmentFast(rec.initial, rec.delta, entry, total, rec.flags, st, n, g,
                      [&](uint32_t idx) { fast_idx.push_back(idx); });
    // and so must the counter-tile walker: its running byte offset must address the counter of the same voxel, and
    // its running increment must select the right half of the word
    resumeSegmentTile(rec.initial, rec.delta, entry, total, rec.flags, st, n, tl, 0u, [&](uint32_t offset, uint32_t one) {
      const uint32_t half = offset >> 1;
      tile_ok = tile_ok && (offset & 1u) == 0 && one == ((half & 1u) ? 0x10000u : 1u);
      tile_ok = tile_ok && half < 2u * tl.words && tileHalf(tl, tileVoxel(tl, half)) == half;
      tile_idx.push_back(tileVoxel(tl, half));
    });
  });
  ++g_rays;
  g_visits += (long long)seq.size();
  bool ok = seq.size() == seg.size();
  for (size_t i = 0; ok && i < seq.size(); ++i)
  {
    ok = memcmp(seq[i].r, seg[i].r, sizeof(seq[i].r)) == 0 && memcmp(seq[i].l, seg[i].l, sizeof(seq[i].l)) == 0;
    // enter/exit must be the same doubles (the excluded-start walk reports the first exit as the first enter)
    ok = ok && memcmp(&seq[i].exit, &seg[i].exit, sizeof(double)) == 0;
    ok = ok && (i == 0 || memcmp(&seq[i].enter, &seg[i].enter, sizeof(double)) == 0);
  }
  ok = ok && (seq.empty() || last_flags == 1);
  ok = ok && fast_idx.size() == seg.size() && tile_ok && tile_idx == fast_idx;
  for (size_t i = 0; ok && i < seg.size(); ++i)
  {
    ok = fast_idx[i] == (uint32_t)(seg[i].l[0] + seg[i].l[1] * g.dim[0] + seg[i].l[2] * g.dim[0] * g.dim[1]);
  }
  if (!ok)
  {
    fprintf(stderr, "MISMATCH flags %u: (%.17g %.17g %.17g) -> (%.17g %.17g %.17g): sequential %zu visits, segments %zu\n",
            walk_flags, start[0], start[1], start[2], end[0], end[1], end[2], seq.size(), seg.size());
    for (size_t i = 0; i < seq.size() && i < seg.size(); ++i)
    {
      if (memcmp(seq[i].r, seg[i].r, sizeof(seq[i].r)) || memcmp(seq[i].l, seg[i].l, sizeof(seq[i].l)) ||
          seq[i].exit != seg[i].exit)
      {
        fprintf(stderr, "  first difference at visit %zu: seq r(%d %d %d) l(%d %d %d) exit %.17g | seg r(%d %d %d) l(%d %d %d) exit %.17g\n",
                i, seq[i].r[0], seq[i].r[1], seq[i].r[2], seq[i].l[0], seq[i].l[1], seq[i].l[2], seq[i].exit, seg[i].r[0],
                seg[i].r[1], seg[i].r[2], seg[i].l[0], seg[i].l[1], seg[i].l[2], seg[i].exit);
        break;
      }
    }
  }
  return ok;
}

int main()
{
  int failures = 0;
  std::mt19937_64 rng(1153297050u);
  const Geom geoms[] = {
    makeGeom(0.1, 32, 32, 32, 0, 0, 0),       makeGeom(0.1, 32, 32, 32, 0.05, 0.05, 0.05),
    makeGeom(0.25, 16, 16, 16, 0, 0, 0),      makeGeom(0.2, 20, 24, 28, -0.3, 0.7, 0.11),
    makeGeom(0.05, 32, 32, 32, 0, 0, 0),      makeGeom(1.0, 5, 7, 3, 0, 0, 0),
  };
  const unsigned flag_sets[] = { 0u, kExcludeStartVoxel, kExcludeEndVoxel, kExcludeStartVoxel | kExcludeEndVoxel };
  for (const Geom &g : geoms)
  {
    const double scale = g.res / 0.1;
    std::uniform_real_distribution<double> far(-40.0 * scale, 40.0 * scale), near(-1.0 * scale, 1.0 * scale);
    for (unsigned flags : flag_sets)
    {
      // random long rays from a common sensor and from random origins
      for (int i = 0; i < 4000; ++i)
      {
        double s[3] = { 0.05 * scale, 0.05 * scale, 0.05 * scale };
        double e[3] = { far(rng), far(rng), far(rng) };
        failures += !checkRay(g, s, e, flags);
        double s2[3] = { far(rng), far(rng), far(rng) };
        failures += !checkRay(g, s2, e, flags);
      }
      // short rays (LineWalkTests.cpp Random)
      for (int i = 0; i < 4000; ++i)
      {
        double s[3] = { near(rng), near(rng), near(rng) };
        double e[3] = { near(rng), near(rng), near(rng) };
        failures += !checkRay(g, s, e, flags);
      }
      // lattice rays: exact voxel-boundary ties (LineWalkTests.cpp Walk), also across region boundaries
      for (int sx = -1; sx <= 1; ++sx)
        for (int sy = -1; sy <= 1; ++sy)
          for (int sz = -1; sz <= 1; ++sz)
            for (int len = 1; len <= 70; len += 3)
              for (int off = 0; off < 3; ++off)
              {
                double s[3] = { off * g.res, off * 2 * g.res, -off * g.res };
                double e[3] = { s[0] + sx * len * g.res, s[1] + sy * len * g.res, s[2] + sz * len * g.res };
                failures += !checkRay(g, s, e, flags);
                double e2[3] = { s[0] + sx * len * g.res, s[1] + sy * len * g.res * 0.5, s[2] + sz * len * g.res * 0.25 };
                failures += !checkRay(g, s, e2, flags);
              }
      // axis-aligned and degenerate rays
      for (int a = 0; a < 3; ++a)
      {
        for (int i = 0; i < 200; ++i)
        {
          double s[3] = { near(rng), near(rng), near(rng) };
          double e[3] = { s[0], s[1], s[2] };
          e[a] = far(rng);
          failures += !checkRay(g, s, e, flags);
        }
      }
      double p[3] = { 0.31 * scale, 0.2 * scale, 0.1 * scale };
      double q[3] = { 0.31 * scale + 1e-9, 0.2 * scale, 0.1 * scale };
      failures += !checkRay(g, p, p, flags);
      failures += !checkRay(g, p, q, flags);
    }
  }
  printf("rays %lld visits %lld segments %lld failures %d\n", g_rays, g_visits, g_segments, failures);
  return failures ? 1 : 0;
}
