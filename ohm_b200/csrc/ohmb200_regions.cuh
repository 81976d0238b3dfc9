// ohmb200_regions.cuh — the region-binned walk: exact ray→segment producer and the shared-memory tile consumer.
//
// Why segments can be exact.  walkLineVoxels (ohm/LineWalkCompute.h:345-413) takes its steps in the order of a 3-way
// merge of the per-axis exit-time sequences
//        T_a(m) = (m == 0) ? initial_a : fl(initial_a + fl(delta_a * m))          (LineWalkCompute.h:298-300, 373-376)
// ordered by (time, then HIGHER axis first on ties) (walkSelectNextAxis, :282-289), and stops after total_a steps per
// axis.  T_a is never accumulated, so the walk state after any prefix of the merge is a pure function of the
// per-axis step counts.  A ray can therefore be cut wherever it crosses a region boundary: the producer merges only
// the boundary-crossing steps (a few dozen per ray), finds by search how many steps of the other two axes precede
// each crossing, and emits (region, ray, stepped[3], visits); a consumer lane resumes the walk from `stepped` and
// reproduces exactly the voxels the sequential walk visits inside that region.
#pragma once

#include "ohmb200_device.cuh"

#include <cstring>

namespace ohmb200
{
constexpr unsigned kRecValid = 1u << 3, kRecExcludeStart = 1u << 4, kRecExcludeEnd = 1u << 5;
#ifndef OHMB200_ITEM_SEGMENTS
#define OHMB200_ITEM_SEGMENTS 2048
#endif
constexpr uint32_t kMaxSegmentsPerItem = OHMB200_ITEM_SEGMENTS;  // < 32768: tile counters are 15 bit + flag
constexpr uint32_t kRecordChunk = 64;   // ordered-miss records are reserved per warp in chunks (unused slots of a chunk are
                                        // skipped by linkRecords: 256 per chunk left it four empty slots per record)
// Counter tile addressing.  One u16 counter per voxel (15-bit count + flag bit), two per 32-bit word.  The tile is a
// padded copy of the region, the position of a voxel LINEAR in its coordinates:
//        position(x, y, z) = x + row * y + slab * z          (counters; row and slab are even)
// so a walking lane keeps a running byte offset and adds a constant per step — no index arithmetic per visit.
// Rows are padded to an even length; a slab is padded to 8 counters more than a multiple of 64, so that the lanes of a
// warp that walk one x/y column at different heights (neighbouring beams of one azimuth) do not all share a bank.
// (Measured: the choice of paddings — OHMB200_TILE — does not move walkRegions; the fold's instruction count does.)
// When dim x is a multiple of 8 (`fast`) the eight voxels 8c .. 8c+7 of the region are one aligned 16-byte group of
// the tile: the fold scans the tile with 128-bit loads and skips empty groups.
struct TileLayout
{
  int dx, dy, dz, dxy;  // region dimensions in voxels
  int row, slab;        // tile strides in counters
  uint32_t words;       // 32-bit words of the tile, a multiple of 32
  uint32_t row_words;   // words of a row that hold voxels
  uint32_t row_lanes;   // lanes the word-wise fold spends on a row: the power of two >= row_words, at most 32
  uint32_t row_shift;   // log2(row_lanes)
  uint32_t inv_dy;      // ceil(2^32 / dy): row number -> z with one multiply-high (exact below 2^21 rows)
  uint32_t fast;        // groups of eight voxels are aligned 16-byte groups of the tile
  uint32_t row_groups;  // fast: groups per row (dx / 8), and ceil(2^32 / row_groups)
  uint32_t inv_row_groups;
};
OHMB200_HD inline TileLayout makeTileLayout(const Geom &g, int row_pad_words = 0, int slab_bank = 4)
{
  TileLayout t;
  t.dx = g.dim[0];
  t.dy = g.dim[1];
  t.dz = g.dim[2];
  t.dxy = g.dim[0] * g.dim[1];
  t.row_words = (uint32_t)(g.dim[0] + 1) >> 1;
  const uint32_t row_w = t.row_words + (uint32_t)row_pad_words;
  uint32_t slab_w = row_w * (uint32_t)g.dim[1];
  if (slab_bank >= 0)
  {
    slab_w += ((uint32_t)slab_bank + 32u - (slab_w & 31u)) & 31u;
  }
  t.row = (int)(2u * row_w);
  t.slab = (int)(2u * slab_w);
  t.words = (slab_w * (uint32_t)g.dim[2] + 31u) & ~31u;
  t.row_lanes = 1;
  t.row_shift = 0;
  while (t.row_lanes < t.row_words && t.row_lanes < 32u)
  {
    t.row_lanes <<= 1;
    ++t.row_shift;
  }
  t.inv_dy = (g.dim[1] > 1) ? (uint32_t)((0x100000000ull + (uint32_t)g.dim[1] - 1u) / (uint32_t)g.dim[1]) : 0u;
  t.fast = (g.dim[0] % 8 == 0 && row_w % 4u == 0 && slab_w % 4u == 0) ? 1u : 0u;
  t.row_groups = (uint32_t)g.dim[0] >> 3;
  t.inv_row_groups = (t.row_groups > 1) ? (uint32_t)((0x100000000ull + t.row_groups - 1u) / t.row_groups) : 0u;
  return t;
}
// counter position of voxel v (index inside the region) and back; divisions: not for the per-visit path
OHMB200_HD __forceinline__ uint32_t tileHalf(const TileLayout &t, uint32_t v)
{
  const uint32_t z = v / (uint32_t)t.dxy;
  const uint32_t r = v - z * (uint32_t)t.dxy;
  const uint32_t y = r / (uint32_t)t.dx;
  return (r - y * (uint32_t)t.dx) + y * (uint32_t)t.row + z * (uint32_t)t.slab;
}
OHMB200_HD __forceinline__ uint32_t tileVoxel(const TileLayout &t, uint32_t half)
{
  const uint32_t z = half / (uint32_t)t.slab;
  const uint32_t r = half - z * (uint32_t)t.slab;
  const uint32_t y = r / (uint32_t)t.row;
  return (r - y * (uint32_t)t.row) + y * (uint32_t)t.dx + z * (uint32_t)t.dxy;
}
constexpr uint32_t kStageSegments = 96;  // segments per ray that pass A hands to pass B without a second enumeration
constexpr uint32_t kTileFlag = 0x8000u;

// Walk constants of one ray (64 bytes).
struct RayRec
{
  double initial[3];
  double delta[3];
  int16_t region[3];  // start voxel key
  uint8_t local[3];
  uint8_t flags;      // bit a (0..2): axis a walks in -direction; kRecValid | kRecExcludeStart | kRecExcludeEnd
  uint16_t total[3];  // steps to take per axis
};
static_assert(sizeof(RayRec) == 64, "RayRec must be 64 bytes");

// One ray's visits inside one region (16 bytes).
struct Segment
{
  uint32_t ray;
  uint16_t stepped[3];  // per-axis steps already taken when the segment starts
  uint16_t visits;      // voxels to visit
  uint32_t entry;       // local voxel coordinates of the first voxel: x | y << 8 | z << 16
};
static_assert(sizeof(Segment) == 16, "Segment must be 16 bytes");

struct WorkItem
{
  uint32_t slot;
  uint32_t begin;  // segment range
  uint32_t end;
  uint32_t shared;  // region split over several items: fold with CAS
};

// The ray a thread of the per-ray segment kernels (prepSegments, emitSegments, markTsdfNear) works on.  Lidar data comes
// in firing order — the beams of one azimuth, then the next azimuth — so 32 consecutive rays span the whole elevation
// fan: 4 m rays into the floor next to 40 m rays to the far wall, and a warp of them runs as long as its longest ray
// with a third of its lanes busy (ncu: 13.7 of 32 in prepSegments).  Inside blocks of 2048 rays the lanes of a warp
// take rays 64 apart instead — the same beam of neighbouring azimuths for any sensor of up to 64 beams (two beams for
// 128), rays of nearly equal length.  Everything indexed by THREAD (the staged segments) stays coalesced; the ray
// records are 64-byte reads either way.  Unordered input loses nothing.
OHMB200_HD __forceinline__ uint32_t rayOfThread(uint32_t i, uint32_t n)
{
  const uint32_t base = i & ~2047u;
  if (base + 2048u > n)
  {
    return i;
  }
  const uint32_t j = i & 2047u;
  return base + (j & 31u) * 64u + (j >> 5);
}

OHMB200_HD __forceinline__ double stepTime(double initial, double delta, int m)
{
  return m == 0 ? initial : initial + delta * m;
}

// (tb, axis b) is taken before (ta, axis a) by walkSelectNextAxis.
OHMB200_HD __forceinline__ bool stepPrecedes(double tb, int b, double ta, int a)
{
  return tb < ta || (tb == ta && b > a);
}

// Build the walk constants of a ray; false when the ray is not walked.
OHMB200_HD inline bool makeRayRec(RayRec &rec, const Geom &g, const double start[3], const double end[3],
                                  unsigned walk_flags)
{
  Key skey, ekey;
  rec.flags = 0;
  if (!voxelKey(g, start, skey) || !voxelKey(g, end, ekey))
  {
    return false;
  }
  Walk w;
  walkInit(w, g, start, end, skey, ekey);
  unsigned flags = kRecValid;
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    rec.initial[a] = w.initial[a];
    rec.delta[a] = w.delta[a];
    rec.region[a] = (int16_t)skey.r[a];
    rec.local[a] = (uint8_t)skey.l[a];
    const int total = abs(w.remaining[a]);
    ok = ok && total <= 0xFFFF;
    rec.total[a] = (uint16_t)total;
    // the walk steps along sign(dir); remaining has the same sign whenever it is non-zero
    flags |= (w.dir[a] < 0) ? (1u << a) : 0u;
  }
  flags |= (walk_flags & kExcludeStartVoxel) ? kRecExcludeStart : 0u;
  flags |= (walk_flags & kExcludeEndVoxel) ? kRecExcludeEnd : 0u;
  rec.flags = ok ? (uint8_t)flags : 0;
  return ok;
}

// Enumerate the per-region segments of a ray in walk order.  emit(region[3], stepped[3], entry_local[3], visits).
template <typename Emit>
OHMB200_HD inline void enumerateSegments(const RayRec &rec, const Geom &g, Emit &&emit)
{
  int dir[3], total[3], l[3], r[3], st[3];
  int T = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    dir[a] = (rec.flags & (1u << a)) ? -1 : 1;
    total[a] = rec.total[a];
    l[a] = rec.local[a];
    r[a] = rec.region[a];
    st[a] = 0;
    T += total[a];
  }
  const bool exclude_start = (rec.flags & kRecExcludeStart) != 0;
  const bool exclude_end = (rec.flags & kRecExcludeEnd) != 0;
  const double inv_delta[3] = { 1.0 / rec.delta[0], 1.0 / rec.delta[1], 1.0 / rec.delta[2] };  // estimates only
  if (T == 0)
  {
    // start and end share a voxel: only the end-voxel visit can happen (LineWalkCompute.h:392-410)
    if (!exclude_end)
    {
      emit(r, st, l, 1);
    }
    return;
  }
  const int q_last = exclude_end ? T - 1 : T;
  int q = 0;
  if (exclude_start)
  {
    // take the first step for real
    double t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      t[a] = total[a] ? rec.initial[a] : (double)INFINITY;
    }
    const int a0 = selectNextAxis(t);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      if (a == a0)
      {
        st[a] = 1;
        l[a] += dir[a];
        if (l[a] < 0)
        {
          l[a] = g.dim[a] - 1;
          r[a] = (int)(int16_t)(r[a] - 1);
        }
        else if (l[a] >= g.dim[a])
        {
          l[a] = 0;
          r[a] = (int)(int16_t)(r[a] + 1);
        }
      }
    }
    q = 1;
  }
  while (q <= q_last)
  {
    // First region-boundary crossing among the axes, in walk order.
    int ca = -1;
    double ct = 0;
    int ck = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const int rem = total[a] - st[a];
      const int k = (dir[a] > 0) ? g.dim[a] - l[a] : l[a] + 1;  // steps on this axis until the local key wraps
      if (rem > 0 && k <= rem)
      {
        const double t = stepTime(rec.initial[a], rec.delta[a], st[a] + k - 1);
        if (ca < 0 || stepPrecedes(t, a, ct, ca))
        {
          ca = a;
          ct = t;
          ck = k;
        }
      }
    }
    if (ca < 0)
    {
      emit(r, st, l, q_last - q + 1);  // the ray ends inside this region
      return;
    }
    // Steps of the other axes that precede the crossing step.
    int nst[3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
      if (b == ca)
      {
        nst[b] = st[b] + ck;
        continue;
      }
      const int rem = total[b] - st[b];
      const int k = (dir[b] > 0) ? g.dim[b] - l[b] : l[b] + 1;
      // candidate new step counts n in [st, hi]; step number n is taken at time T_b(n - 1) and "step n precedes the
      // crossing" is monotone in n.  Start from the estimate (ct - initial) / delta and walk to the boundary with
      // exact evaluations (typically two).
      int lo = st[b];
      const int hi = st[b] + min(rem, k - 1);
      if (lo < hi)
      {
        const double guess = (ct - rec.initial[b]) * inv_delta[b];
        int n = (guess >= (double)hi) ? hi : ((guess > (double)lo) ? (int)guess + 1 : lo);  // NaN -> lo
        n = min(max(n, lo), hi);
        if (n > lo && !stepPrecedes(stepTime(rec.initial[b], rec.delta[b], n - 1), b, ct, ca))
        {
          do
          {
            --n;
          } while (n > lo && !stepPrecedes(stepTime(rec.initial[b], rec.delta[b], n - 1), b, ct, ca));
        }
        else
        {
          while (n < hi && stepPrecedes(stepTime(rec.initial[b], rec.delta[b], n), b, ct, ca))
          {
            ++n;
          }
        }
        lo = n;
      }
      nst[b] = lo;
    }
    const int q_exit = nst[0] + nst[1] + nst[2];  // first position inside the next region
    const int n = min(q_exit - 1, q_last) - q + 1;
    if (n > 0)
    {
      emit(r, st, l, n);
    }
    if (q_exit > q_last)
    {
      return;
    }
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
      if (b == ca)
      {
        l[b] = (dir[b] > 0) ? 0 : g.dim[b] - 1;
        r[b] = (int)(int16_t)(r[b] + dir[b]);
      }
      else
      {
        l[b] += dir[b] * (nst[b] - st[b]);
      }
      st[b] = nst[b];
    }
    q = q_exit;
  }
}

// ---- one crossing at a time (groundwork for a producer with one thread per crossing; not yet used by the kernels) ----
// enumerateSegments walks a ray's region crossings in order, each found from the one before.  They do not depend on
// each other: crossing j of axis a IS step k1_a + j dim_a of that axis (k1_a = steps until the local coordinate first
// wraps), taken at time T_a(k1_a + j dim_a - 1); the steps of another axis b that precede it are counted by the same
// estimate-and-check search enumerateSegments uses; and the number of crossings of axis b among those steps is a
// closed form of that count — so the crossing's rank in the merged order of all crossings, the walk position right
// after it and the state of the walk there follow without looking at any other crossing.  A segment is what lies
// between two crossings of consecutive rank: with one thread per crossing writing (position, steps) at its rank and a
// second pass taking differences, the ~1 M segments of a sweep become ~1 M balanced threads instead of 131 k threads
// that loop over 4 to 60 segments each.  tests/cpp/segments_host_test.cu proves the equivalence on the host.
struct Crossing
{
  int rank;         // 1-based position among the ray's crossings in walk order (rank 0 = the start of the ray)
  int position;     // walk position (steps taken) of the first voxel after the crossing
  int stepped[3];   // per-axis steps taken at that position
};

// steps on `axis` until the local coordinate first passes a multiple of `cut` (cut = the region dimension: until it wraps)
OHMB200_HD __forceinline__ int firstCutStep(const RayRec &rec, int axis, int cut)
{
  const int within = (int)rec.local[axis] % cut;
  return (rec.flags & (1u << axis)) ? within + 1 : cut - within;
}

// cuts of `axis` among its first `steps` steps
OHMB200_HD __forceinline__ int cutsWithin(const RayRec &rec, int axis, int cut, int steps)
{
  const int first = firstCutStep(rec, axis, cut);
  return steps >= first ? 1 + (steps - first) / cut : 0;
}

// region crossings of `axis` among its first `steps` steps
OHMB200_HD __forceinline__ int crossingsWithin(const RayRec &rec, const Geom &g, int axis, int steps)
{
  return cutsWithin(rec, axis, g.dim[axis], steps);
}

// Cut j (0-based) of `axis`, the walk being cut wherever a local coordinate passes a multiple of cut[axis]; false when
// the ray has no such cut.  cut = the region dimensions gives the region crossings; a divisor of them cuts the segments
// shorter (every region crossing is still a cut), which bounds the serial length of a lane's walk.
OHMB200_HD inline bool cutOf(const RayRec &rec, const int cut[3], int axis, int j, Crossing &out)
{
  const int step = firstCutStep(rec, axis, cut[axis]) + j * cut[axis];  // its number among the steps of the axis
  if (j < 0 || step > (int)rec.total[axis])
  {
    return false;
  }
  const double ct = stepTime(rec.initial[axis], rec.delta[axis], step - 1);
  out.rank = 1 + j;
  out.position = 0;
#pragma unroll
  for (int b = 0; b < 3; ++b)
  {
    int n = step;
    if (b != axis)
    {
      // the largest n in [0, total_b] whose step n (taken at T_b(n - 1)) precedes the cut; "precedes" is monotone
      const int hi = (int)rec.total[b];
      n = 0;
      if (hi > 0)
      {
        const double guess = (ct - rec.initial[b]) / rec.delta[b];  // an estimate only
        n = (guess >= (double)hi) ? hi : ((guess > 0.0) ? (int)guess + 1 : 0);  // NaN -> 0
        n = min(max(n, 0), hi);
        while (n > 0 && !stepPrecedes(stepTime(rec.initial[b], rec.delta[b], n - 1), b, ct, axis))
        {
          --n;
        }
        while (n < hi && stepPrecedes(stepTime(rec.initial[b], rec.delta[b], n), b, ct, axis))
        {
          ++n;
        }
      }
      out.rank += cutsWithin(rec, b, cut[b], n);
    }
    out.stepped[b] = n;
    out.position += n;
  }
  return true;
}

// Crossing j (0-based) of `axis`; false when the ray has no such crossing.
OHMB200_HD inline bool crossingOf(const RayRec &rec, const Geom &g, int axis, int j, Crossing &out)
{
  return cutOf(rec, g.dim, axis, j, out);
}

// Resume a segment's walk from its per-axis step counts and call visit(l, enter, exit, last_of_ray) for each of
// its `visits` voxels (l = local voxel coordinates inside the segment's region).  kTimes: also track the
// enter/exit ranges (traversal layer); `length` is the walk's length (exit range of the end voxel).
template <bool kTimes, typename Visit>
OHMB200_HD inline void resumeSegment(const double init[3], const double delta[3], const int local0[3], const int total[3],
                                     uint32_t flags, const int st_in[3], int visits, double length, const Geom &g,
                                     Visit &&visit)
{
  int dir[3], l[3], rem[3], st[3];
  double tn[3];
  double last_time = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    st[a] = st_in[a];
    dir[a] = (flags & (1u << a)) ? -1 : 1;
    rem[a] = total[a] - st[a];
    const int pos = (local0[a] + dir[a] * st[a]) % g.dim[a];
    l[a] = pos < 0 ? pos + g.dim[a] : pos;
    tn[a] = rem[a] > 0 ? stepTime(init[a], delta[a], st[a]) : (double)INFINITY;
    if (kTimes && st[a] > 0)
    {
      last_time = fmax(last_time, stepTime(init[a], delta[a], st[a] - 1));
    }
  }
  const int q_last = total[0] + total[1] + total[2] - ((flags & kRecExcludeEnd) ? 1 : 0);
  int axis = selectNextAxis(tn);
  for (int v = 0; v < visits; ++v)
  {
    double t_exit = 0;
    bool last_of_ray = false;
    if (kTimes)
    {
      const bool at_end = (rem[0] | rem[1] | rem[2]) == 0;
      t_exit = at_end ? length : ((axis == 0) ? tn[0] : ((axis == 1) ? tn[1] : tn[2]));
      last_of_ray = (st[0] + st[1] + st[2]) == q_last;
    }
    visit(l, last_time, t_exit, last_of_ray);
    last_time = t_exit;
    if (v + 1 < visits)
    {
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        if (a == axis)
        {
          ++st[a];
          --rem[a];
          l[a] += dir[a];
          tn[a] = rem[a] > 0 ? init[a] + delta[a] * st[a] : (double)INFINITY;
        }
      }
      axis = selectNextAxis(tn);
    }
  }
}

// resumeSegment without ranges: a running linear voxel index instead of coordinates, fp64 step counters (no int->double
// conversion in the loop) and a branch per stepped axis.  visit(idx) receives the voxel index inside the region.
// Same arithmetic, same order of comparisons.  The kernels use resumeSegmentTile (below), which addresses the counter
// tile directly; this form stays as its reference in tests/cpp/segments_host_test.cu.
template <typename Visit>
OHMB200_HD __forceinline__ void resumeSegmentFast(const double init[3], const double delta[3], const int entry[3],
                                                  const int total[3], uint32_t flags, const int st_in[3], int visits,
                                                  const Geom &g, Visit &&visit)
{
  // entry = local voxel coordinates at which the segment starts (the producer stores them with the segment)
  int s0 = st_in[0], s1 = st_in[1], s2 = st_in[2];
  const int d0 = (flags & 1u) ? -1 : 1, d1 = (flags & 2u) ? -1 : 1, d2 = (flags & 4u) ? -1 : 1;
  const int stride1 = g.dim[0], stride2 = g.dim[0] * g.dim[1];
  int idx = entry[0] + entry[1] * stride1 + entry[2] * stride2;
  const int step0 = d0, step1 = d1 * stride1, step2 = d2 * stride2;
  double m0 = (double)s0, m1 = (double)s1, m2 = (double)s2;
  double t0 = (s0 < total[0]) ? (s0 == 0 ? init[0] : init[0] + delta[0] * m0) : (double)INFINITY;
  double t1 = (s1 < total[1]) ? (s1 == 0 ? init[1] : init[1] + delta[1] * m1) : (double)INFINITY;
  double t2 = (s2 < total[2]) ? (s2 == 0 ? init[2] : init[2] + delta[2] * m2) : (double)INFINITY;
  for (int v = 0;;)
  {
    visit((uint32_t)idx);
    if (++v >= visits)
    {
      break;
    }
    // walkSelectNextAxis: strict '<', ties go to the higher axis
    const bool x_first = t0 < t1;
    const bool low_first = x_first ? (t0 < t2) : (t1 < t2);
    if (!low_first)
    {
      ++s2;
      m2 += 1.0;
      idx += step2;
      t2 = (s2 < total[2]) ? init[2] + delta[2] * m2 : (double)INFINITY;
    }
    else if (x_first)
    {
      ++s0;
      m0 += 1.0;
      idx += step0;
      t0 = (s0 < total[0]) ? init[0] + delta[0] * m0 : (double)INFINITY;
    }
    else
    {
      ++s1;
      m1 += 1.0;
      idx += step1;
      t1 = (s1 < total[1]) ? init[1] + delta[1] * m1 : (double)INFINITY;
    }
  }
}

// High word of a double.  For integer values below 2^20 it identifies the value (the low word is zero).
OHMB200_HD __forceinline__ int hiWord(double x)
{
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  unsigned long long u;
  memcpy(&u, &x, sizeof(u));
  return (int)(u >> 32);
#endif
}

// The walk of the counting kernels: resumeSegmentFast against the counter tile.  visit(offset, one) receives the BYTE
// offset of the voxel's u16 counter in the tile (plus tile_base: the kernels pass the tile's shared-memory address, so
// that `offset & ~3` is the address of the word) and the increment of that counter inside its 32-bit word (1 or 1 << 16);
// both are carried along — a constant is added per step, the increment swaps halves on x steps (tile rows and slabs
// have even strides) — so a visit costs one AND, the shared-memory atomic and the flag test.  The per-axis step count
// lives only in the double m (the multiplier of LineWalkCompute.h:373-376); "axis finished" compares its high word
// with that of the axis total (both are integers below 2^16).  Same arithmetic, same order of comparisons as
// walkSelectNextAxis: strict '<', ties go to the higher axis.
// (Measured alternatives that lost: computing the next exit time one step ahead and testing the flag of a visit one
// step later shorten the dependent chain but add instructions — the walk phase is bound by issue slots, +8 %.)
template <typename Visit>
OHMB200_HD __forceinline__ void resumeSegmentTile(const double init[3], const double delta[3], const int entry[3],
                                                  const int total[3], uint32_t flags, const int st_in[3], int visits,
                                                  const TileLayout &tl, uint32_t tile_base, Visit &&visit)
{
  int offset = (int)tile_base + 2 * (entry[0] + entry[1] * tl.row + entry[2] * tl.slab);
  uint32_t one = (entry[0] & 1) ? 0x10000u : 1u;
  const int step0 = (flags & 1u) ? -2 : 2;
  const int step1 = (flags & 2u) ? -2 * tl.row : 2 * tl.row;
  const int step2 = (flags & 4u) ? -2 * tl.slab : 2 * tl.slab;
  double m0 = (double)st_in[0], m1 = (double)st_in[1], m2 = (double)st_in[2];
  const int end0 = hiWord((double)total[0]), end1 = hiWord((double)total[1]), end2 = hiWord((double)total[2]);
  double t0 = (st_in[0] < total[0]) ? (st_in[0] == 0 ? init[0] : init[0] + delta[0] * m0) : (double)INFINITY;
  double t1 = (st_in[1] < total[1]) ? (st_in[1] == 0 ? init[1] : init[1] + delta[1] * m1) : (double)INFINITY;
  double t2 = (st_in[2] < total[2]) ? (st_in[2] == 0 ? init[2] : init[2] + delta[2] * m2) : (double)INFINITY;
  for (int v = 0;;)
  {
    visit((uint32_t)offset, one);
    if (++v >= visits)
    {
      break;
    }
    const bool x_before_y = t0 < t1, x_before_z = t0 < t2, y_before_z = t1 < t2;
    if (x_before_y && x_before_z)
    {
      m0 += 1.0;
      offset += step0;
      one = (one << 16) | (one >> 16);
      const double t = init[0] + delta[0] * m0;
      t0 = (hiWord(m0) != end0) ? t : (double)INFINITY;
    }
    else if (!x_before_y && y_before_z)
    {
      m1 += 1.0;
      offset += step1;
      const double t = init[1] + delta[1] * m1;
      t1 = (hiWord(m1) != end1) ? t : (double)INFINITY;
    }
    else
    {
      m2 += 1.0;
      offset += step2;
      const double t = init[2] + delta[2] * m2;
      t2 = (hiWord(m2) != end2) ? t : (double)INFINITY;
    }
  }
}

}  // namespace ohmb200
