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

namespace ohmb200
{
constexpr unsigned kRecValid = 1u << 3, kRecExcludeStart = 1u << 4, kRecExcludeEnd = 1u << 5;
constexpr uint32_t kMaxSegmentsPerItem = 2048;  // < 32768: tile counters are 15 bit + flag
constexpr uint32_t kRecordChunk = 256;
// Counter tile addressing.  Two u16 counters per 32-bit word; the word index is XOR-swizzled in bits 2..4 with a hash of
// the higher bits (the y and z coordinates of the voxel), so lanes walking the same x/y column at different heights —
// a lidar's elevation fan — spread over the shared-memory banks instead of queueing on one.  Groups of four words stay
// together (128-bit accesses of the fold remain valid).  A tile holds tileWords() words (a multiple of 32).
OHMB200_HD __forceinline__ uint32_t tileWords(uint32_t vpr)
{
  return (((vpr + 1u) >> 1) + 31u) & ~31u;
}
OHMB200_HD __forceinline__ uint32_t tileWord(uint32_t voxel)
{
  const uint32_t w = voxel >> 1;
  const uint32_t u = w >> 5;
  return w ^ (((u ^ (u >> 3) ^ (u >> 6)) & 7u) << 2);
}
OHMB200_HD __forceinline__ uint32_t tileGroup(uint32_t group)  // group = word index / 4
{
  const uint32_t u = group >> 3;
  return group ^ ((u ^ (u >> 3) ^ (u >> 6)) & 7u);
}
constexpr uint32_t kStageSegments = 96;  // segments per ray that pass A hands to pass B without a second enumeration         // ordered-miss records are reserved per warp in chunks
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

// The hot-path variant of resumeSegment: no ranges, a running linear voxel index instead of coordinates, fp64 step
// counters (no int->double conversion in the loop) and a branch per stepped axis instead of predicating all three.
// visit(idx) receives the voxel index inside the region.  Same arithmetic, same order of comparisons.
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

}  // namespace ohmb200
