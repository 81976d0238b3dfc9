// ohmb200_device.cuh — device-side map model, key maths and the exact fp64 voxel walk.
//
// Everything here is compiled with --fmad=false: the parity target is ohm's CPU mapper, whose x86-64 build
// never contracts a*b+c into an FMA, and the voxel sequence of a ray is decided by comparisons of such
// expressions.  All fp64 operators used (+ - * / sqrt floor) are IEEE-754 correctly rounded on sm_100a, so a
// ray produces bit-identical exit times — hence identical voxel keys — on the device and in the CPU mapper.
//
// Semantics follow (reference file:line, csiro-robotics/ohm @ 4e2e769):
//   point -> key        ohm/MapCoord.h:37-93, ohm/MapRegion.cpp:32-69
//   key -> centre       ohm/OccupancyMap.h:757-778
//   key stepping/diff   ohm/OccupancyMap.h:827-846,887-901
//   voxel walk          ohm/LineWalkCompute.h:162-413 (driven as ohm/LineWalk.h:112-129 does on the CPU)
#pragma once

#include <cfloat>

#include <cuda_runtime.h>
#include <stdint.h>

// Key maths and the walk are also compiled for the host by tests/cpp/segments_host_test.cu (no GPU needed there).
#define OHMB200_HD __host__ __device__

namespace ohmb200
{
constexpr uint64_t kEmptyKey = ~0ull;
// A slot whose region was evicted to the host store (GpuLayerCache's eviction, ohmgpu/GpuLayerCache.cpp:429-633):
// probes walk over it, an insert may take it.
constexpr uint64_t kTombKey = ~0ull - 1ull;
OHMB200_HD __forceinline__ bool isRegionKey(unsigned long long k)
{
  return k < kTombKey;
}
constexpr uint32_t kHitFlag = 0x80000000u;   // pending word: voxel has sample updates in this batch
constexpr uint32_t kInvalidVoxel = 0xFFFFFFFFu;

// Ray filter result flags (ohm/RayFilter.h:22-29)
constexpr unsigned kRffInvalid = 1u, kRffClippedStart = 2u, kRffClippedEnd = 4u;
// Walk flags (ohm/LineWalk.h:51-57)
constexpr unsigned kExcludeStartVoxel = 1u, kExcludeEndVoxel = 2u;

struct Geom
{
  double res;
  double region_size[3];
  double origin[3];
  int dim[3];
  uint32_t vpr;  // voxels per region
};

struct MapParams
{
  float hit_value, miss_value, min_value, max_value, threshold_value;
  float sat_min, sat_max;  // lowest()/max() when saturation is disabled (RayMapperOccupancy.cpp:94-95)
  int filter_kind;
  double filter_range;
  double clip_box[6];
  float sensor_noise, adaptation_rate, reinit_threshold, initial_intensity_cov;
  uint32_t reinit_count, sample_threshold;
  int ndt_tm;
  float tsdf_max_weight, tsdf_trunc, tsdf_dropoff, tsdf_sparsity;
};

// Region table + layer slabs.  A region's slab slot IS its hash-table index: slabs are initialised to the layer
// clear values once, so inserting a region is a single 64-bit CAS and needs no per-region setup.
struct DeviceMap
{
  unsigned long long *keys;  // [capacity] packed region keys, kEmptyKey when free
  uint32_t capacity;
  uint32_t *region_stamp;    // [capacity] last batch stamp that walked the region
  uint32_t *pending;         // [capacity * vpr] per-batch miss counters / kHitFlag|run
  float *occupancy;          // layer slabs, nullptr when the layer is absent
  uint2 *mean;
  float *traversal;
  uint32_t *touch_time;
  uint32_t *incident;
  float *covariance;         // 6 floats per voxel
  float2 *intensity;
  uint2 *hit_miss;
  float2 *tsdf;
  uint2 *secondary;          // {f32 m2, u16 range_mean | u16 count << 16}
  uint32_t *voxel_bits;      // [capacity][(vpr + 31) / 32] persistent bit per voxel (NDT / TSDF maps), else nullptr
  unsigned long long *region_count;  // device counter of occupied slots
  int *table_full;                   // set when an insert found no free slot
  uint32_t *new_slots;               // slots of the regions inserted since new_count was reset (paging), or nullptr
  uint32_t *new_count;
  int part_rank;                     // this GPU's index among `part_world` region owners (multi-GPU sharding)
  int part_world;                    // 1 = the map owns every region
};

struct Key
{
  int r[3];  // region
  int l[3];  // local voxel
};

OHMB200_HD __forceinline__ unsigned long long packRegion(int x, int y, int z)
{
  return (unsigned long long)(uint16_t)x | ((unsigned long long)(uint16_t)y << 16) |
         ((unsigned long long)(uint16_t)z << 32);
}

__host__ __device__ __forceinline__ uint32_t hashRegion(unsigned long long k)
{
  k *= 0x9E3779B97F4A7C15ull;
  return (uint32_t)(k >> 32) ^ (uint32_t)k;
}

// Region -> owning GPU: (rx + 2 ry + 4 rz) mod world — with 8 owners the parity octants, so the regions around the
// sensor (which carry most of a sweep: the two hottest hold 9 % each, measured on config 2) always land on different
// GPUs.  Every rank sees every ray, so locality between neighbouring regions buys nothing here, balance does: the
// heaviest owner carries 1.09 / 1.26 / 1.37 x the mean at 2 / 4 / 8 owners (2x2x2-block ownership: 1.15 / 1.67 / 2.56).
__host__ __device__ __forceinline__ int regionOwner(int rx, int ry, int rz, int world)
{
  const int v = rx + 2 * ry + 4 * rz;
  const int r = v % world;
  return r < 0 ? r + world : r;
}

__device__ __forceinline__ bool ownsRegion(const DeviceMap &m, const int r[3])
{
  return m.part_world <= 1 || regionOwner(r[0], r[1], r[2], m.part_world) == m.part_rank;
}

// Find or insert a region; returns its slot, or -1 when the table is full.  Open addressing, linear probing; the slot
// of an evicted region (kTombKey) is taken by the first insert whose probe sequence passes it and ends on an empty
// slot — every inserter of one key picks its slot by the same rule from the same table, so two of them meet on the
// same compare-and-swap; the loser of any CAS simply probes again.
__device__ inline int regionSlot(const DeviceMap &m, unsigned long long key)
{
  for (int attempt = 0; attempt < 64; ++attempt)
  {
    uint32_t h = hashRegion(key) % m.capacity;
    int tomb = -1;
    bool retry = false;
    for (uint32_t probe = 0; probe < m.capacity; ++probe)
    {
      const unsigned long long k = m.keys[h];
      if (k == key)
      {
        return (int)h;
      }
      if (k == kTombKey && tomb < 0)
      {
        tomb = (int)h;
      }
      if (k == kEmptyKey)
      {
        const uint32_t target = tomb >= 0 ? (uint32_t)tomb : h;
        const unsigned long long expected = tomb >= 0 ? kTombKey : kEmptyKey;
        const unsigned long long old = atomicCAS(&m.keys[target], expected, key);
        if (old == expected)
        {
          atomicAdd(m.region_count, 1ull);
          if (m.new_slots)
          {
            const uint32_t at = atomicAdd(m.new_count, 1u);
            if (at < m.capacity)
            {
              m.new_slots[at] = target;
            }
          }
          return (int)target;
        }
        if (old == key)
        {
          return (int)target;
        }
        retry = true;  // somebody else took the slot for another region
        break;
      }
      h = (h + 1 == m.capacity) ? 0 : h + 1;
    }
    if (!retry)
    {
      if (tomb >= 0)
      {
        // no empty slot on the way, but a free one: take it (the probe went all the way round, the key is absent)
        const unsigned long long old = atomicCAS(&m.keys[tomb], kTombKey, key);
        if (old == kTombKey)
        {
          atomicAdd(m.region_count, 1ull);
          if (m.new_slots)
          {
            const uint32_t at = atomicAdd(m.new_count, 1u);
            if (at < m.capacity)
            {
              m.new_slots[at] = (uint32_t)tomb;
            }
          }
          return tomb;
        }
        if (old == key)
        {
          return tomb;
        }
        continue;
      }
      break;
    }
  }
  *m.table_full = 1;
  return -1;
}

// Lookup only.
__device__ inline int regionFind(const DeviceMap &m, unsigned long long key)
{
  uint32_t h = hashRegion(key) % m.capacity;
  for (uint32_t probe = 0; probe < m.capacity; ++probe)
  {
    const unsigned long long k = m.keys[h];
    if (k == key)
    {
      return (int)h;
    }
    if (k == kEmptyKey)
    {
      return -1;
    }
    h = (h + 1 == m.capacity) ? 0 : h + 1;
  }
  return -1;
}

// ohm/MapCoord.h:85-93
OHMB200_HD __forceinline__ int pointToRegionCoord(double coord, double resolution)
{
  return (int)floor(coord / resolution + 0.5);
}

// ohm/MapCoord.h:41-80
OHMB200_HD __forceinline__ int pointToRegionVoxel(double coord, double voxel_resolution, double region_resolution)
{
  const double epsilon = (double)1e-6f;
  if (-epsilon <= coord && coord < 0)
  {
    coord = 0;
  }
  else if (coord >= region_resolution && coord - epsilon < region_resolution)
  {
    coord -= epsilon;
  }
  return (int)floor(coord / voxel_resolution);
}

// ohm/OccupancyMap.cpp:859-886 -> ohm/MapRegion.cpp:32-69.  false => null key.
OHMB200_HD inline bool voxelKey(const Geom &g, const double p[3], Key &key)
{
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    const int rc = (int)(int16_t)pointToRegionCoord(p[a] - g.origin[a], g.region_size[a]);
    const double centre = rc * g.region_size[a];
    const double region_min = centre - 0.5 * g.region_size[a];
    const double local = p[a] - g.origin[a] - region_min;
    const int q = pointToRegionVoxel(local, g.res, g.region_size[a]);
    key.r[a] = rc;
    key.l[a] = q;
    ok = ok && (0 <= q && q < g.dim[a]);
  }
  return ok;
}

// ohm/OccupancyMap.h:757-778
OHMB200_HD __forceinline__ double voxelCentreAxis(const Geom &g, int region, int local, int a)
{
  double c = (double)(float)region;
  c *= g.region_size[a];
  c -= 0.5 * g.region_size[a];
  c += g.origin[a];
  c += (double)local * g.res;
  c += 0.5 * g.res;
  return c;
}

OHMB200_HD __forceinline__ uint32_t voxelIndex(const Geom &g, const Key &k)
{
  // ohm/MapChunk.h:47-50
  return (uint32_t)k.l[0] + (uint32_t)k.l[1] * g.dim[0] + (uint32_t)k.l[2] * g.dim[0] * g.dim[1];
}

// ohm/RayFilter.cpp:15-55.  Returns false for a rejected ray; may move `end` and set kRffClippedEnd.
OHMB200_HD inline bool applyRayFilter(const MapParams &p, double start[3], double end[3], unsigned &filter_flags)
{
  if (p.filter_kind == 0)
  {
    return true;
  }
  bool good = true;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    good = good && !isnan(start[a]) && !isinf(start[a]) && !isnan(end[a]) && !isinf(end[a]);
  }
  if (p.filter_kind == 3)
  {
    // clipBounded (RayFilter.cpp:57-76) -> Aabb::clipLine / rayIntersect / contains (Aabb.h:309-450).
    // No NaN/inf screening here, exactly like the reference filter.
    const double *lo = p.clip_box, *hi = p.clip_box + 3;
    const double origin[3] = { start[0], start[1], start[2] };
    double dir[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
    const double d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2];
    unsigned clip = 0;
    if (!(d2 < 1e-9))
    {
      const double length = sqrt(d2);
      double inv[3];
      bool sign[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        dir[a] /= length;
        inv[a] = 1.0 / dir[a];
        sign[a] = dir[a] < 0.0;
      }
      double t0 = ((sign[0] ? hi[0] : lo[0]) - origin[0]) * inv[0];
      double t1 = ((sign[0] ? lo[0] : hi[0]) - origin[0]) * inv[0];
      double tmin = ((sign[1] ? hi[1] : lo[1]) - origin[1]) * inv[1];
      double tmax = ((sign[1] ? lo[1] : hi[1]) - origin[1]) * inv[1];
      bool miss = (t0 > tmax) || (tmin > t1);
      t0 = (tmin > t0 || isnan(t0)) ? tmin : t0;
      t1 = (tmax < t1 || isnan(t1)) ? tmax : t1;
      tmin = ((sign[2] ? hi[2] : lo[2]) - origin[2]) * inv[2];
      tmax = ((sign[2] ? lo[2] : hi[2]) - origin[2]) * inv[2];
      miss = miss || (t0 > tmax) || (tmin > t1);
      t0 = (tmin > t0 || isnan(t0)) ? tmin : t0;
      t1 = (tmax < t1 || isnan(t1)) ? tmax : t1;
      if (!miss)
      {
        if (t0 > 0 && t0 < length)
        {
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            start[a] = origin[a] + dir[a] * t0;
          }
          clip |= 1u;
        }
        if (t1 > 0 && t1 < length)
        {
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            end[a] = origin[a] + dir[a] * t1;
          }
          clip |= 2u;
        }
      }
    }
    if (clip)
    {
      bool s_in = true, e_in = true;
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        s_in = s_in && !(hi[a] < start[a]) && !(lo[a] > start[a]);
        e_in = e_in && !(hi[a] < end[a]) && !(lo[a] > end[a]);
      }
      if (!s_in && !e_in)
      {
        return false;
      }
    }
    filter_flags |= (clip & 1u) ? kRffClippedStart : 0u;
    filter_flags |= (clip & 2u) ? kRffClippedEnd : 0u;
    return true;
  }
  double ray[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
  const double len2 = (ray[0] * ray[0] + ray[1] * ray[1]) + ray[2] * ray[2];
  const double range = p.filter_range;
  if (p.filter_kind == 1)
  {
    good = good && (range <= 0 || len2 <= range * range);
    if (!good)
    {
      filter_flags |= kRffInvalid;
    }
    return good;
  }
  // clipRayFilter
  if (good && range > 0 && len2 > range * range)
  {
    const double len = sqrt(len2);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      ray[a] /= len;
      end[a] = start[a] + ray[a] * range;
    }
    filter_flags |= kRffClippedEnd;
  }
  if (!good)
  {
    filter_flags |= kRffInvalid;
  }
  return good;
}

// State of one ray's voxel walk.  time_next[a] is always recomputed as initial + delta * |stepped| (never
// accumulated, LineWalkCompute.h:298-300), so the walk can be resumed anywhere from `stepped` alone.
struct Walk
{
  double initial[3];
  double delta[3];
  double time_next[3];
  double length;
  int remaining[3];
  int stepped[3];
  int dir[3];  // +1 / -1
  Key cur;
  Key end;
  int axis;
  unsigned limit;
};

OHMB200_HD __forceinline__ int selectNextAxis(const double t[3])
{
  int axis = 0;
  axis = (t[axis] < t[1]) ? axis : 1;
  axis = (t[axis] < t[2]) ? axis : 2;
  return axis;
}

// LineWalkCompute.h:188-280 + the set-up half of walkLineVoxels (:351-379).
OHMB200_HD inline void walkInit(Walk &w, const Geom &g, const double start[3], const double end[3], const Key &skey,
                                const Key &ekey)
{
  double dir[3], inv[3];
  int sign[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    dir[a] = end[a] - start[a];
  }
  double length = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
  length = (length > 1e-6) ? sqrt(length) : 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    sign[a] = dir[a] < 0;
    dir[a] /= length;
    inv[a] = (length > 0) ? 1 / dir[a] : 0;
  }
  w.length = length;
  w.limit = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    const double centre = voxelCentreAxis(g, skey.r[a], skey.l[a], a);
    double vmin = centre - 0.5 * g.res;
    double vmax = centre + 0.5 * g.res;
    const double initial = ((sign[a] ? vmin : vmax) - start[a]) * inv[a];
    const int sd = -2 * sign[a] + 1;
    const double shift = sd * g.res;
    vmin += shift;
    vmax += shift;
    double delta = ((sign[a] ? vmin : vmax) - start[a]) * inv[a];
    if (delta != (double)INFINITY)
    {
      delta -= initial;
    }
    w.initial[a] = initial;
    w.delta[a] = delta;
    w.dir[a] = sd;
    w.stepped[a] = 0;
    w.remaining[a] = ekey.l[a] - skey.l[a] + (int)(int16_t)(ekey.r[a] - skey.r[a]) * g.dim[a];
    w.limit |= (w.remaining[a] == 0) ? (1u << a) : 0u;
    w.time_next[a] = w.remaining[a] ? initial : (double)INFINITY;
  }
  w.cur = skey;
  w.end = ekey;
  w.axis = selectNextAxis(w.time_next);
}

// LineWalkCompute.h:291-307 (+ key stepping OccupancyMap.h:827-846)
OHMB200_HD __forceinline__ void walkStep(Walk &w, const Geom &g)
{
  // Select by predication rather than dynamic indexing to keep the state in registers.
  const int a = w.axis;
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    if (i == a)
    {
      const int d = w.dir[i];
      int local = w.cur.l[i] + d;
      int region = w.cur.r[i];
      if (local < 0)
      {
        --region;
        local = g.dim[i] - 1;
      }
      else if (local >= g.dim[i])
      {
        ++region;
        local = 0;
      }
      w.cur.l[i] = local;
      w.cur.r[i] = (int)(int16_t)region;
      w.remaining[i] -= d;
      w.stepped[i] += d;
      w.time_next[i] = w.remaining[i] ? w.initial[i] + w.delta[i] * abs(w.stepped[i]) : (double)INFINITY;
      w.limit |= (w.remaining[i] == 0) ? (1u << i) : 0u;
    }
  }
  w.axis = selectNextAxis(w.time_next);
}

OHMB200_HD __forceinline__ bool walkAtEnd(const Walk &w)
{
  return w.cur.r[0] == w.end.r[0] && w.cur.r[1] == w.end.r[1] && w.cur.r[2] == w.end.r[2] &&
         w.cur.l[0] == w.end.l[0] && w.cur.l[1] == w.end.l[1] && w.cur.l[2] == w.end.l[2];
}

OHMB200_HD __forceinline__ double walkNextTime(const Walk &w)
{
  return (w.axis == 0) ? w.time_next[0] : ((w.axis == 1) ? w.time_next[1] : w.time_next[2]);
}

// Drives `visit(key, enter, exit)` exactly as walkLineVoxels (LineWalkCompute.h:345-413) would.
template <typename Visit>
OHMB200_HD inline unsigned walkLine(const Geom &g, const double start[3], const double end[3], const Key &skey,
                                    const Key &ekey, unsigned flags, Visit &&visit)
{
  Walk w;
  walkInit(w, g, start, end, skey, ekey);
  double last_time = 0;
  unsigned count = 0;
  // A well-formed walk takes exactly |dx|+|dy|+|dz| steps.  The bound only matters for non-finite input that slipped
  // past a disabled filter (the reference documents a hang there, ohmgpu/GpuMap.h:209-210); it never alters a
  // valid walk.
  const unsigned max_steps = (unsigned)(abs(w.remaining[0]) + abs(w.remaining[1]) + abs(w.remaining[2]));
  if (flags & kExcludeStartVoxel)
  {
    last_time = walkNextTime(w);
    ++count;
    walkStep(w, g);
  }
  while (w.limit < 7u && !walkAtEnd(w) && count <= max_steps)
  {
    const double t = walkNextTime(w);
    visit(w.cur, last_time, t);
    last_time = t;
    ++count;
    walkStep(w, g);
  }
  if ((flags & kExcludeEndVoxel) == 0u)
  {
    visit(ekey, last_time, w.length);
    ++count;
  }
  return count;
}

// ---------------------------------------------------------------------------------------------------------
// Per-voxel arithmetic
// ---------------------------------------------------------------------------------------------------------

// One miss applied to `v` with the CPU mapper's flag logic (RayMapperOccupancy.cpp:150-166 +
// VoxelOccupancyCompute.h:110-120).  A pure function of v, so k identical misses commute with each other.
OHMB200_HD __forceinline__ float missOnce(float v, const MapParams &p, unsigned ray_flags)
{
  const float uninit = INFINITY;
  const bool unobserved = v == uninit;
  const bool is_free = !unobserved && v < p.threshold_value;
  const bool occupied = !unobserved && v >= p.threshold_value;
  float adj = p.miss_value;
  adj = (unobserved && (ray_flags & (1u << 5))) ? uninit : adj;
  adj = (is_free && (ray_flags & (1u << 6))) ? 0.0f : adj;
  adj = (occupied && (ray_flags & (1u << 7))) ? 0.0f : adj;
  const float base = (!unobserved) ? v : 0.0f;
  adj = (unobserved || (p.sat_min < v && v < p.sat_max)) ? adj : 0.0f;
  return (base != uninit) ? fmaxf(p.min_value, base + adj) : base;
}

// `count` consecutive misses.  Stops early at a fixed point (min clamp, saturation, exclusion).
// The usual case — no exclusion flags, saturation off (missIsPlain) — is v -> max(min, v + miss) after the first miss
// has turned "unobserved" into 0 + miss: a four-instruction step instead of missOnce's flag logic, and when the count
// is so large that the walk down certainly ends on the clamp (with room for the rounding of every add) the answer is
// min itself.  (A fold that ran missOnce ten times per voxel of a fresh map was 60 % of walkRegions.)
OHMB200_HD __forceinline__ bool missIsPlain(const MapParams &p, unsigned ray_flags)
{
  return (ray_flags & 0xE0u) == 0 && p.sat_min == -FLT_MAX && p.sat_max == FLT_MAX;
}

OHMB200_HD __forceinline__ float missRepeatPlain(float v, uint32_t count, float miss_value, float min_value)
{
  if (count == 0)
  {
    return v;
  }
  if (v == INFINITY)
  {
    v = fmaxf(min_value, 0.0f + miss_value);
    --count;
  }
  if (miss_value < 0 && count > 2)
  {
    const float steps = (float)(count - 1u);
    const float slack = 1e-6f * steps * (fabsf(v) + fabsf(min_value) + fabsf(miss_value));
    if (v + steps * miss_value <= min_value - slack)
    {
      return min_value;
    }
  }
  while (count)
  {
    const float n = fmaxf(min_value, v + miss_value);
    if (n == v)
    {
      break;
    }
    v = n;
    --count;
  }
  return v;
}

OHMB200_HD __forceinline__ float missRepeat(float v, uint32_t count, const MapParams &p, unsigned ray_flags)
{
  if (missIsPlain(p, ray_flags))
  {
    return missRepeatPlain(v, count, p.miss_value, p.min_value);
  }
  while (count)
  {
    const float n = missOnce(v, p, ray_flags);
    if (n == v)
    {
      break;
    }
    v = n;
    --count;
  }
  return v;
}

// The miss ladder: T[k] = log-odds of a voxel that was unobserved and then took k misses (T[0] = +inf), up to the
// first fixed point of the miss rule (the min clamp, an exclusion flag) or the end of the table.  missOnce is a pure
// function of the value, so this one sequence answers "count misses on v" for every voxel whose history is misses
// only — the free space along the rays, nearly every voxel a ray walks through — whatever the flags:
// v == T[j]  =>  the result is T[min(j + count, fix)].  j is guessed as v / miss and CHECKED bit for bit, so a wrong
// guess (a voxel that has seen hits) only falls back to the loop.  Built once per CTA in shared memory.
constexpr int kMissLadder = 64;
struct MissLadder
{
  float value[kMissLadder];
  float inv_miss;  // 1 / miss value: the guess of j
  uint32_t fix;    // value[fix] is a fixed point of the miss rule; kMissLadder if the table ends before one
};

OHMB200_HD __forceinline__ void buildMissLadder(MissLadder &ladder, const MapParams &p, unsigned ray_flags)
{
  float v = INFINITY;
  ladder.fix = kMissLadder;
  ladder.inv_miss = 1.0f / p.miss_value;
  for (int k = 0; k < kMissLadder; ++k)
  {
    ladder.value[k] = v;
    const float n = missOnce(v, p, ray_flags);
    if (n == v && ladder.fix == kMissLadder)
    {
      ladder.fix = (uint32_t)k;
    }
    v = n;
  }
}

// Branch-free lookup: the value after `count` misses when v is on the ladder (ok = true), so that the lookups of the
// eight voxels of a group overlap; off the ladder (a voxel with hits in its history, a table without a fixed point)
// ok = false and the caller runs missRepeat.
OHMB200_HD __forceinline__ float missLadderLookup(const MissLadder &ladder, float v, uint32_t count, bool &ok)
{
  const uint32_t fix = ladder.fix;
  const bool has_fix = fix < (uint32_t)kMissLadder;
  const uint32_t last = has_fix ? fix : (uint32_t)kMissLadder - 1u;
  float guess = (v == INFINITY) ? 0.0f : rintf(v * ladder.inv_miss);
  guess = fminf(fmaxf(guess, 0.0f), (float)(kMissLadder - 1));  // NaN -> 0, and value[0] = inf != NaN
  const uint32_t j = (uint32_t)guess;
  const bool on_ladder = ladder.value[j] == v;
  const bool at_fix = v == ladder.value[last] && has_fix;  // e.g. free space on the min clamp, wherever the guess fell
  const uint32_t k = j + min(count, (uint32_t)kMissLadder);
  ok = count == 0 || at_fix || (on_ladder && (has_fix || k < (uint32_t)kMissLadder));
  const float after = ladder.value[min(k, last)];
  return (count == 0 || at_fix) ? v : after;
}

OHMB200_HD __forceinline__ float missRepeatLadder(const MissLadder &ladder, float v, uint32_t count, const MapParams &p,
                                                  unsigned ray_flags)
{
  bool ok;
  const float after = missLadderLookup(ladder, v, count, ok);
  return ok ? after : missRepeat(v, count, p, ray_flags);
}

// RayMapperOccupancy.cpp:262-279 + VoxelOccupancyCompute.h:44-54
OHMB200_HD __forceinline__ float hitOnce(float v, const MapParams &p, unsigned ray_flags)
{
  const float uninit = INFINITY;
  const bool unobserved = v == uninit;
  const bool is_free = !unobserved && v < p.threshold_value;
  const bool occupied = !unobserved && v >= p.threshold_value;
  float adj = p.hit_value;
  adj = (unobserved && (ray_flags & (1u << 5))) ? uninit : adj;
  adj = (is_free && (ray_flags & (1u << 6))) ? 0.0f : adj;
  adj = (occupied && (ray_flags & (1u << 7))) ? 0.0f : adj;
  const float base = (!unobserved) ? v : 0.0f;
  adj = (unobserved || (p.sat_min < v && v < p.sat_max)) ? adj : 0.0f;
  return (base != uninit) ? fminf(base + adj, p.max_value) : base;
}

// ohm/VoxelMeanCompute.h:69-152 instantiated as the CPU mappers do: <glm::dvec3, double>.
__device__ __forceinline__ void subVoxelToLocal(uint32_t coord, double resolution, double out[3])
{
  const double mean_resolution = resolution / (double)1023;
  const double offset = (double)0.5f * resolution;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    out[a] = (int)((coord >> (10 * a)) & 1023u) * mean_resolution - offset;
  }
}

__device__ __forceinline__ uint32_t subVoxelCoord(const double local[3], double resolution)
{
  const double mean_resolution = resolution / (double)1023;
  const double offset = (double)0.5f * resolution;
  uint32_t pattern = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    int pos = pointToRegionCoord(local[a] + offset, mean_resolution);
    pos = (pos >= 0 ? (pos < 1024 ? pos : 1023) : 0);
    pattern |= ((uint32_t)pos) << (10 * a);
  }
  return pattern | (1u << 31);
}

__device__ __forceinline__ uint32_t subVoxelUpdate(uint32_t coord, uint32_t count, const double local[3],
                                                   double resolution)
{
  double mean[3];
  subVoxelToLocal(coord, resolution, mean);
  const double w = (double)1 / (double)(count + 1u);
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    mean[a] += (local[a] - mean[a]) * w;
  }
  return subVoxelCoord(mean, resolution);
}

// ohm/VoxelIncidentCompute.h:35-112 (float arithmetic; sqrtf and '/' are IEEE under -prec-sqrt/-prec-div)
__device__ inline uint32_t updateIncidentNormal(uint32_t packed, float ix, float iy, float iz, uint32_t count)
{
  float nx = (2.0f * ((float)((packed >> 0) & 0x3FFFu) / 16383.0f)) - 1.0f;
  float ny = (2.0f * ((float)((packed >> 15) & 0x3FFFu) / 16383.0f)) - 1.0f;
  nx = fmaxf(-1.0f, fminf(nx, 1.0f));
  ny = fmaxf(-1.0f, fminf(ny, 1.0f));
  float nz = fmaxf(-1.0f, fminf(1.0f - (nx * nx + ny * ny), 1.0f));
  const bool set = (packed & (1u << 30)) != 0;
  nx = set ? nx : 0.0f;
  ny = set ? ny : 0.0f;
  nz = set ? sqrtf(nz) : 0.0f;
  nz *= (packed & (1u << 31)) ? -1.0f : 1.0f;

  count = ((nx != 0 || ny != 0 || nz != 0) && count) ? count : 0;
  const float w = 1.0f / (float)(count + 1u);
  float len2 = ix * ix + iy * iy + iz * iz;
  // the reference forms this reciprocal in double (its `sqrt(float)` is the double overload) and narrows it
  float s = (len2 > 1e-6f) ? (float)(1.0 / sqrt((double)len2)) : 0.0f;
  ix *= s;
  iy *= s;
  iz *= s;
  nx += (ix - nx) * w;
  ny += (iy - ny) * w;
  nz += (iz - nz) * w;
  len2 = nx * nx + ny * ny + nz * nz;
  s = (len2 > 1e-6f) ? (float)(1.0 / sqrt((double)len2)) : 0.0f;
  nx *= s;
  ny *= s;
  nz *= s;

  const float ex = 0.5f * (fmaxf(-1.0f, fminf(nx, 1.0f)) + 1.0f);
  const float ey = 0.5f * (fmaxf(-1.0f, fminf(ny, 1.0f)) + 1.0f);
  uint32_t n = 0;
  n |= ((uint32_t)(ex * 16383.0f) & 0x3FFFu) << 0;
  n |= ((uint32_t)(ey * 16383.0f) & 0x3FFFu) << 15;
  n &= ~((1u << 30) | (1u << 31));
  n |= (nz < 0) ? (1u << 31) : 0u;
  n |= (ex != 0.0f || ey != 0.0f || nz != 0.0f) ? (1u << 30) : 0u;
  return n;
}

// ohm/VoxelTouchTimeCompute.h:18-27
__device__ __forceinline__ uint32_t encodeTouchTime(double timebase, double timestamp)
{
  // x86-64 converts double -> unsigned through a 64-bit truncation; mirror that rather than saturating.
  return (uint32_t)(long long)((timestamp - timebase) / 0.001);
}

}  // namespace ohmb200
