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

static long long g_rays = 0, g_visits = 0, g_segments = 0, g_crossings = 0;

struct SegmentTuple
{
  int r[3], st[3], entry[3], n;
  bool operator==(const SegmentTuple &o) const
  {
    return memcmp(r, o.r, sizeof(r)) == 0 && memcmp(st, o.st, sizeof(st)) == 0 && memcmp(entry, o.entry, sizeof(entry)) == 0 &&
           n == o.n;
  }
};

// The segments of a ray rebuilt from its crossings taken ONE AT A TIME (crossingOf): each crossing yields its rank in
// walk order, the walk position after it and the per-axis step counts there, without reference to any other crossing;
// a segment is what lies between two crossings of consecutive rank.  Must equal enumerateSegments' output exactly.
static bool segmentsFromCuts(const RayRec &rec, const Geom &g, const int cut[3], std::vector<SegmentTuple> &out);
static bool segmentsFromCrossings(const RayRec &rec, const Geom &g, std::vector<SegmentTuple> &out)
{
  return segmentsFromCuts(rec, g, g.dim, out);
}

// ... and cut finer: wherever a local coordinate passes a multiple of cut[axis] (a divisor of the region dimension).
static bool segmentsFromCuts(const RayRec &rec, const Geom &g, const int cut[3], std::vector<SegmentTuple> &out)
{
  const int total[3] = { rec.total[0], rec.total[1], rec.total[2] };
  const int T = total[0] + total[1] + total[2];
  const bool exclude_start = (rec.flags & kRecExcludeStart) != 0, exclude_end = (rec.flags & kRecExcludeEnd) != 0;
  auto state_at = [&](const int st[3], SegmentTuple &t) {
    for (int a = 0; a < 3; ++a)
    {
      const int dir = (rec.flags & (1u << a)) ? -1 : 1;
      t.st[a] = st[a];
      const int pos = ((int)rec.local[a] + dir * st[a]) % g.dim[a];
      t.entry[a] = pos < 0 ? pos + g.dim[a] : pos;
      t.r[a] = (int)(int16_t)((int)rec.region[a] + dir * crossingsWithin(rec, g, a, st[a]));
    }
  };
  if (T == 0)
  {
    if (!exclude_end)
    {
      SegmentTuple t;
      const int zero[3] = { 0, 0, 0 };
      state_at(zero, t);
      t.n = 1;
      out.push_back(t);
    }
    return true;
  }
  int count = 0;
  for (int a = 0; a < 3; ++a)
  {
    count += cutsWithin(rec, a, cut[a], total[a]);
  }
  std::vector<Crossing> by_rank((size_t)count + 1);
  std::vector<char> have((size_t)count + 1, 0);
  by_rank[0].rank = 0;
  by_rank[0].position = 0;
  by_rank[0].stepped[0] = by_rank[0].stepped[1] = by_rank[0].stepped[2] = 0;
  have[0] = 1;
  for (int a = 0; a < 3; ++a)
  {
    Crossing c;
    for (int j = 0; cutOf(rec, cut, a, j, c); ++j)
    {
      ++g_crossings;
      if (c.rank < 1 || c.rank > count || have[c.rank])
      {
        return false;  // the ranks must be a permutation of 1 .. count
      }
      have[c.rank] = 1;
      by_rank[c.rank] = c;
    }
  }
  const int q_first = exclude_start ? 1 : 0, q_last = exclude_end ? T - 1 : T;
  for (int i = 0; i <= count; ++i)
  {
    if (!have[i] || (i > 0 && by_rank[i].position <= by_rank[i - 1].position))
    {
      return false;
    }
    const int begin = by_rank[i].position, end = (i < count) ? by_rank[i + 1].position - 1 : T;
    const int lo = begin > q_first ? begin : q_first, hi = end < q_last ? end : q_last;
    if (hi < lo)
    {
      continue;
    }
    SegmentTuple t;
    if (lo == begin)
    {
      state_at(by_rank[i].stepped, t);
    }
    else
    {
      // the excluded start voxel: the segment begins one step into the ray (the first step: the earliest exit time)
      double t0[3];
      for (int a = 0; a < 3; ++a)
      {
        t0[a] = total[a] ? rec.initial[a] : (double)INFINITY;
      }
      int st[3] = { 0, 0, 0 };
      st[selectNextAxis(t0)] = 1;
      state_at(st, t);
    }
    t.n = hi - lo + 1;
    out.push_back(t);
  }
  return true;
}

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
  std::vector<SegmentTuple> enumerated, rebuilt;
  enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
    ++g_segments;
    SegmentTuple tuple;
    memcpy(tuple.r, r, sizeof(tuple.r));
    memcpy(tuple.st, st, sizeof(tuple.st));
    memcpy(tuple.entry, entry, sizeof(tuple.entry));
    tuple.n = n;
    enumerated.push_back(tuple);
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
  const bool crossings_ok = segmentsFromCrossings(rec, g, rebuilt) && rebuilt == enumerated;
  if (!crossings_ok)
  {
    fprintf(stderr, "crossings: %zu segments rebuilt, %zu enumerated\n", rebuilt.size(), enumerated.size());
  }
  ok = ok && crossings_ok;
  // Cut at half regions: every piece stays inside one region, and walking the pieces one after the other visits the
  // voxels of the sequential walk.
  if ((g.dim[0] % 2) == 0 && (g.dim[1] % 2) == 0 && (g.dim[2] % 2) == 0)
  {
    const int half[3] = { g.dim[0] / 2, g.dim[1] / 2, g.dim[2] / 2 };
    std::vector<SegmentTuple> pieces;
    bool halves_ok = segmentsFromCuts(rec, g, half, pieces);
    size_t at = 0;
    const int total[3] = { rec.total[0], rec.total[1], rec.total[2] };
    for (const SegmentTuple &p : pieces)
    {
      halves_ok = halves_ok && p.n <= half[0] + half[1] + half[2];
      resumeSegmentFast(rec.initial, rec.delta, p.entry, total, rec.flags, p.st, p.n, g, [&](uint32_t idx) {
        halves_ok = halves_ok && at < seq.size() && memcmp(seq[at].r, p.r, sizeof(p.r)) == 0 &&
                    idx == (uint32_t)(seq[at].l[0] + seq[at].l[1] * g.dim[0] + seq[at].l[2] * g.dim[0] * g.dim[1]);
        ++at;
      });
    }
    halves_ok = halves_ok && at == seq.size();
    if (!halves_ok)
    {
      fprintf(stderr, "half-region cuts: %zu pieces, %zu of %zu visits matched\n", pieces.size(), at, seq.size());
    }
    ok = ok && halves_ok;
  }
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

// The layout of the shared-memory counter tile and the thread -> ray mapping of the per-ray kernels: pure index maths,
// checked exhaustively.
static int checkLayouts()
{
  int bad = 0;
  const int dims[][3] = { { 32, 32, 32 }, { 16, 16, 16 }, { 16, 24, 8 }, { 8, 5, 3 }, { 12, 10, 6 }, { 5, 7, 3 },
                          { 9, 32, 4 }, { 64, 2, 2 }, { 255, 3, 1 }, { 1, 1, 1 }, { 2, 255, 2 } };
  for (const auto &d : dims)
  {
    for (int variant = 0; variant < 3; ++variant)
    {
      const Geom g = makeGeom(0.1, d[0], d[1], d[2], 0, 0, 0);
      const TileLayout tl = variant == 0 ? makeTileLayout(g) : (variant == 1 ? makeTileLayout(g, 1, 5) : makeTileLayout(g, 0, -1));
      bad += (tl.row & 1) || (tl.slab & 1) || tl.row < d[0] || tl.slab < tl.row * d[1] || (tl.words & 31u);
      bad += (uint32_t)tl.slab * (uint32_t)d[2] > 2u * tl.words;
      std::vector<char> used((size_t)2 * tl.words, 0);
      for (uint32_t v = 0; v < g.vpr; ++v)
      {
        const uint32_t half = tileHalf(tl, v);
        const uint32_t x = v % (uint32_t)d[0], y = (v / (uint32_t)d[0]) % (uint32_t)d[1], z = v / (uint32_t)(d[0] * d[1]);
        bad += half != x + (uint32_t)tl.row * y + (uint32_t)tl.slab * z;  // linear in the coordinates
        bad += half >= used.size() || used[half]++ != 0 || tileVoxel(tl, half) != v;
      }
      // rows -> (y, z) and groups -> rows by multiply-high, as foldTile / groupPosition do on the device
      for (uint32_t r = 0; r < (uint32_t)(d[1] * d[2]); ++r)
      {
        const uint32_t z = (d[1] > 1) ? (uint32_t)(((unsigned long long)r * tl.inv_dy) >> 32) : r;
        bad += z != r / (uint32_t)d[1];
      }
      if (tl.fast)
      {
        bad += (d[0] % 8) != 0 || tl.row_groups != (uint32_t)d[0] / 8u;
        for (uint32_t c = 0; c < g.vpr / 8u; ++c)
        {
          const uint32_t r = (tl.row_groups > 1) ? (uint32_t)(((unsigned long long)c * tl.inv_row_groups) >> 32) : c;
          bad += r != c / tl.row_groups;
          bad += (tileHalf(tl, 8u * c) & 7u) != 0;  // groups of eight voxels are aligned 16-byte groups of the tile
          bad += tileHalf(tl, 8u * c + 7u) != tileHalf(tl, 8u * c) + 7u;
        }
      }
    }
  }
  const uint32_t counts[] = { 0, 1, 31, 2047, 2048, 2049, 4096, 5000, 131072, 131072 + 777 };
  for (uint32_t n : counts)
  {
    std::vector<char> seen(n, 0);
    for (uint32_t i = 0; i < n; ++i)
    {
      const uint32_t ray = rayOfThread(i, n);
      bad += ray >= n || seen[ray]++ != 0;  // a permutation of the rays
    }
  }
  return bad;
}

// The miss ladder of the fold (MissLadder, ohmb200_device.cuh) against the definition — missOnce applied `count` times —
// for every miss rule the mapper has: plain, saturation, each exclusion flag, odd parameters; values on the ladder, off it
// (voxels with hits in their history), unobserved, the clamp, NaN; counts from 0 to the tile counter's maximum.
static int checkMissLadder()
{
  int bad = 0;
  struct Rule
  {
    float miss, min, max, threshold;
    bool sat_min, sat_max;
    unsigned flags;
  };
  const Rule rules[] = {
    { -0.2006707f, -2.0f, 3.511f, 0.0f, false, false, 0u },   { -0.2006707f, -2.0f, 3.511f, 0.0f, true, true, 0u },
    { -0.2006707f, -2.0f, 3.511f, 0.0f, false, false, 1u << 5 }, { -0.2006707f, -2.0f, 3.511f, 0.0f, false, false, 1u << 6 },
    { -0.2006707f, -2.0f, 3.511f, 0.0f, false, false, 1u << 7 }, { -0.4f, -1.9f, 2.0f, 0.0f, false, false, 0u },
    { -0.01f, -2.0f, 3.5f, 0.0f, false, false, 0u },          { 0.3f, -2.0f, 3.5f, 0.0f, false, false, 0u },
    { -2.4079456f, -2.0f, 3.511f, 0.0f, true, false, 0u },    { -1e-9f, -2.0f, 3.5f, 0.5f, false, false, 0u },
  };
  const uint32_t counts[] = { 0, 1, 2, 3, 5, 9, 10, 11, 12, 31, 63, 64, 65, 200, 1000, 32767 };
  std::mt19937 rng(5489u);
  std::uniform_real_distribution<float> any_value(-2.5f, 4.0f);
  for (const Rule &r : rules)
  {
    MapParams p{};
    p.miss_value = r.miss;
    p.hit_value = 2.1972246f;
    p.min_value = r.min;
    p.max_value = r.max;
    p.threshold_value = r.threshold;
    p.sat_min = r.sat_min ? r.min : -FLT_MAX;
    p.sat_max = r.sat_max ? r.max : FLT_MAX;
    MissLadder ladder;
    buildMissLadder(ladder, p, r.flags);
    std::vector<float> values = { INFINITY, r.min, r.max, 0.0f, -0.0f, NAN, r.min + 1e-6f, 2.1972246f };
    for (int k = 0; k < kMissLadder; ++k)
    {
      values.push_back(ladder.value[k]);
      values.push_back(hitOnce(ladder.value[k], p, 0u));              // a voxel that took a hit after k misses
      values.push_back(missOnce(hitOnce(ladder.value[k], p, 0u), p, r.flags));
    }
    for (int k = 0; k < 300; ++k)
    {
      values.push_back(any_value(rng));
    }
    for (float v : values)
    {
      for (uint32_t count : counts)
      {
        float want = v;
        for (uint32_t n = 0; n < count; ++n)
        {
          const float next = missOnce(want, p, r.flags);
          if (memcmp(&next, &want, sizeof(float)) == 0)
          {
            break;
          }
          want = next;
        }
        const float got = missRepeatLadder(ladder, v, count, p, r.flags);
        const float plain = missRepeat(v, count, p, r.flags);
        const bool same = memcmp(&got, &want, sizeof(float)) == 0 || (got == want && got == 0.0f) || (got != got && want != want);
        const bool same_plain = memcmp(&plain, &want, sizeof(float)) == 0 || (plain == want && plain == 0.0f) || (plain != plain && want != want);
        if (!same || !same_plain)
        {
          if (bad < 10)
          {
            fprintf(stderr, "miss ladder: rule miss %g min %g flags %u: v %.9g count %u -> ladder %.9g, missRepeat %.9g, want %.9g\n",
                    r.miss, r.min, r.flags, v, count, got, plain, want);
          }
          ++bad;
        }
      }
    }
  }
  return bad;
}

int main()
{
  int failures = checkLayouts() + checkMissLadder();
  if (failures)
  {
    fprintf(stderr, "layout checks: %d failures\n", failures);
  }
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
  printf("rays %lld visits %lld segments %lld crossings %lld failures %d\n", g_rays, g_visits, g_segments, g_crossings, failures);
  return failures ? 1 : 0;
}
