// ohmb200.cu — kernels, device-resident map runtime and the C ABI of libohmb200.so.
//
// Pipeline of one integrateRays batch (all on the map's compute stream, no host synchronisation):
//
//   prepSamples   1 thread/ray   filter, sample-voxel key (fp64), region find-or-insert -> (voxel id, ray) pair
//   radix sort                   pairs by voxel id (stable: ray order is preserved inside a voxel)
//   markRuns      1 thread/pair  run heads -> pending[voxel] = kHitFlag | run, compact run list
//   walkRays      1 thread/ray   exact fp64 voxel walk; per visit: unflagged voxel -> RED.ADD pending (miss count),
//                                flagged voxel -> (run, ray) record chained on the run (replayed in ray order)
//   applySamples  1 thread/run   replays hits (+mean, incident normal, touch time) interleaved, in ray order, with
//                                the recorded misses of that voxel — bit-exact with the sequential CPU mapper
//   resolveMisses 1 CTA/region   pending miss counts -> occupancy (k identical clamped adds commute), clear pending
//
// Reference semantics: ohm/RayMapperOccupancy.cpp:68-339 (the CPU mapper is the parity target);
// replaces ohmgpu/GpuMap.cpp:540-1191 + ohmgpu/gpu/RegionUpdate.cl:158-494 + ohmgpu/GpuLayerCache.cpp.
#include "ohmb200.h"
#include "ohmb200_device.cuh"
#include "ohmb200_regions.cuh"
#include "ohmb200_ndt.cuh"

#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace ohmb200;

// ---------------------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------------------
#define PAGING_TRACE(...)                          \
  do                                               \
  {                                                \
    static const bool on = getenv("OHMB200_TRACE_PAGING") != nullptr; \
    if (on)                                        \
    {                                              \
      fprintf(stderr, "[paging] " __VA_ARGS__);    \
      fputc('\n', stderr);                         \
      fflush(stderr);                              \
    }                                              \
  } while (0)

static thread_local char g_last_error[512] = "";

static int setError(int code, const char *fmt, ...)
{
  va_list args;
  va_start(args, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, args);
  va_end(args);
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do                                                                                                     \
  {                                                                                                      \
    cudaError_t err__ = (expr);                                                                          \
    if (err__ != cudaSuccess)                                                                            \
    {                                                                                                    \
      return setError(OHMB200_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                      __LINE__);                                                                         \
    }                                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// Device-side batch state
// ---------------------------------------------------------------------------------------------------------
struct Counters
{
  unsigned long long rays_accepted;
  unsigned long long voxel_visits;
  unsigned long long sample_updates;
  unsigned long long ordered_records;
  unsigned long long region_count;
  // per batch (reset before each batch)
  uint32_t record_count;
  uint32_t run_count;
  uint32_t touched_count;
  uint32_t record_overflow;
  uint32_t segment_total;  // region-binned path
  uint32_t item_count;
  uint32_t work_next;
  uint32_t segment_overflow;
  uint32_t gauss_count;  // NDT: reserved Gaussian-visit record slots
  uint32_t heavy_count;  // NDT: runs set aside for the warp-per-run replay
  uint32_t new_count;    // regions inserted by this batch (paging: the ones with a chunk in the host store are restored)
  // sticky
  int table_full;
  int overflow_seen;  // bit 0: an ordered-record list overflowed (results incomplete); bit 1: the segment list did (batch dropped)
  unsigned long long sample_voxels;  // S': distinct sample voxels, summed over batches (the NDT byte model's unit)
  unsigned long long owned_visits;   // exchange: voxel visits of the segments routed to THIS rank (voxel_visits counts the sender's)
  uint32_t batch_stamp;  // stamp of the last batch planRegions saw (read on the device: a replayed graph bakes no stamp)
};
constexpr int kPerBatchCounterWords = 11;  // record_count .. new_count

struct Batch
{
  const double *rays;        // [2n] x 3 doubles: origin, sample
  const float *intensities;  // [n] or null
  const double *timestamps;  // [n] or null
  uint32_t n;
  uint32_t heavy_run;  // NDT: a run with this many hits + recorded misses is replayed by a warp (OHMB200_HEAVY_RUN)
  unsigned ray_flags;
  uint32_t stamp;
  double time_base;
  // sample pairs
  uint32_t *keys_in, *keys_out;  // voxel ids
  uint32_t *vals_in, *vals_out;  // ray indices
  uint32_t *run_list;            // [n] sorted index of each run head
  int32_t *run_head;             // [n] head of the record chain of the run starting at sorted index i (-1 = none)
  uint32_t *interval_count;      // [2n + 1] misses that precede sorted hit i inside its run; [n + i]: see tail_overflow
  uint32_t *tail_overflow;       // = interval_count + n: misses after the last hit of the run starting at sorted index i
  uint32_t *interval_offset;     // [2n + 1] NDT: exclusive scan of interval_count
  uint32_t *sorted_rays;         // [record_capacity] NDT: rays of the ordered-miss records, grouped by interval
  // ordered miss records
  uint32_t *record_ray;
  int32_t *record_next;
  uint32_t record_capacity;
  uint32_t *touched_list;  // [capacity] region slots walked this batch
  double *last_exit;       // [n] exit range of the last walked voxel (traversal layer only)
  // region-binned path
  RayRec *recs;            // [n] walk constants
  double *ray_length;      // [n] (traversal layer only)
  uint32_t *record_vid;    // [record_capacity] voxel id of an ordered-miss record (linked to its run afterwards)
  uint32_t *seg_count;     // [capacity] segments per region slot
  uint32_t *seg_offset;    // [capacity] exclusive scan of seg_count
  uint32_t *seg_cursor;    // [capacity] fill cursors
  uint32_t *sample_begin;  // [capacity] the region's range of the sorted sample pairs (markRuns); begin == end: none
  uint32_t *sample_end;
  Segment *segments;       // [seg_capacity] binned by region slot
  uint32_t seg_capacity;
  uint4 *stage;            // [kStageSegments][stage_stride] segments found by pass A, plane k = k-th segment of each ray
  uint32_t *stage_count;   // [n] segments pass A found for the ray (> kStageSegments: pass B enumerates again)
  uint32_t stage_stride;
  uint32_t *cross_pos;     // (unused: reserved for a per-crossing producer, see crossingOf in ohmb200_regions.cuh)
  uint32_t stage_by_ray;   // the staged planes are indexed by ray, not by thread (always 0: prepSegments stages by thread)
  WorkItem *items;         // (region, segment range) work list
  uint32_t item_capacity;
  unsigned long long *gauss_keys;          // NDT: (voxel id << 32 | ray) of visits to voxels with an established Gaussian
  uint32_t gauss_capacity;
  unsigned long long *record_keys;         // TSDF: (voxel id << 32 | ray) of every visit that must be replayed in order
  unsigned long long *record_keys_sorted;
  Counters *counters;
};

// ---------------------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warpAggregatedInc(uint32_t *counter)
{
  // One atomic per converged group instead of one per lane.
  const unsigned mask = __activemask();
  const int leader = __ffs(mask) - 1;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader)
  {
    base = atomicAdd(counter, (uint32_t)__popc(mask));
  }
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

// b.last_exit[] after carryLastExit: the exit range of the last voxel the ray walked, or that of the nearest earlier ray
// of the batch that walked one — the reference's last_exit_range is a variable of the whole integrateRays call and a
// ray that visits no voxel (start and sample in one voxel) finds the previous ray's value in it
// (ohm/RayMapperOccupancy.cpp:79,190,309).  NaN = no ray before it walked anything: the initial 0.
__device__ __forceinline__ double staleExit(double v)
{
  return isnan(v) ? 0.0 : v;
}

struct CarryValid
{
  __device__ __forceinline__ double operator()(double before, double here) const
  {
    return isnan(here) ? before : here;
  }
};

__device__ __forceinline__ void loadRay(const Batch &b, uint32_t i, double start[3], double end[3])
{
  const double *r = b.rays + (size_t)i * 6;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    start[a] = r[a];
    end[a] = r[3 + a];
  }
}

// Filter + sample voxel of every ray.
__global__ void prepSamples(DeviceMap dm, Geom g, MapParams mp, Batch b, int mode)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool accepted = false;
  if (i < b.n)
  {
    double start[3], end[3];
    loadRay(b, i, start, end);
    unsigned filter_flags = 0;
    uint32_t vid = kInvalidVoxel;
    if (applyRayFilter(mp, start, end, filter_flags))
    {
      accepted = true;
      // RayMapperOccupancy.cpp:223,234 / RayMapperNdt.cpp:270,284
      const bool include_sample_in_ray = (filter_flags & kRffClippedEnd) || (b.ray_flags & OHMB200_RF_END_POINT_AS_FREE);
      bool hit = !include_sample_in_ray;
      if (mode == OHMB200_MODE_OCCUPANCY)
      {
        hit = hit && !(b.ray_flags & OHMB200_RF_EXCLUDE_SAMPLE);
      }
      Key skey, ekey;
      // walkSegmentKeys drops rays with a null key, but the sample update still resolves its own key.
      (void)skey;
      if (hit && voxelKey(g, end, ekey) && ownsRegion(dm, ekey.r))
      {
        const int slot = regionSlot(dm, packRegion(ekey.r[0], ekey.r[1], ekey.r[2]));
        if (slot >= 0)
        {
          vid = (uint32_t)slot * g.vpr + voxelIndex(g, ekey);
        }
      }
    }
    b.keys_in[i] = vid;
    b.vals_in[i] = i;
  }
  const unsigned n_acc = __reduce_add_sync(0xffffffffu, accepted ? 1u : 0u);
  if ((threadIdx.x & 31) == 0 && n_acc)
  {
    atomicAdd(&b.counters->rays_accepted, (unsigned long long)n_acc);
  }
}

// Run heads of the sorted (voxel, ray) pairs.
__global__ void markRuns(DeviceMap dm, Batch b, uint32_t vpr)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n)
  {
    return;
  }
  const uint32_t vid = b.keys_out[i];
  if (vid == kInvalidVoxel)
  {
    return;
  }
  if (i == 0 || b.keys_out[i - 1] != vid)
  {
    if (dm.pending)
    {
      dm.pending[vid] = kHitFlag | i;
    }
    const uint32_t r = warpAggregatedInc(&b.counters->run_count);
    b.run_list[r] = i;
  }
  if (b.sample_begin)
  {
    // region boundaries of the sorted pairs (keys are slot * vpr + voxel, so a region is one contiguous range)
    const uint32_t slot = vid / vpr;
    const uint32_t prev = (i > 0) ? b.keys_out[i - 1] : kInvalidVoxel;
    if (i == 0 || prev / vpr != slot)
    {
      b.sample_begin[slot] = i;
    }
    const uint32_t next = (i + 1 < b.n) ? b.keys_out[i + 1] : kInvalidVoxel;
    if (next == kInvalidVoxel || next / vpr != slot)
    {
      b.sample_end[slot] = i + 1;
    }
  }
}

// Exact voxel walk of every ray: the miss side of RayMapperOccupancy::integrateRays.
__global__ void __launch_bounds__(128) walkRays(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned visits = 0;
  if (i < b.n)
  {
    double start[3], end[3];
    loadRay(b, i, start, end);
    unsigned filter_flags = 0;
    Key skey, ekey;
    double last_exit = 0;
    bool walked_any = false;
    if (applyRayFilter(mp, start, end, filter_flags) && !(b.ray_flags & OHMB200_RF_EXCLUDE_RAY) &&
        voxelKey(g, start, skey) && voxelKey(g, end, ekey))
    {
      const bool include_sample_in_ray = (filter_flags & kRffClippedEnd) || (b.ray_flags & OHMB200_RF_END_POINT_AS_FREE);
      unsigned walk_flags = (!include_sample_in_ray) ? kExcludeEndVoxel : 0u;
      walk_flags |= (b.ray_flags & OHMB200_RF_EXCLUDE_ORIGIN) ? kExcludeStartVoxel : 0u;

      unsigned long long last_region = kEmptyKey;
      int slot = -1;
      walkLine(g, start, end, skey, ekey, walk_flags, [&](const Key &k, double enter, double exit) {
        const unsigned long long rk = packRegion(k.r[0], k.r[1], k.r[2]);
        if (rk != last_region)
        {
          last_region = rk;
          // Regions owned by another GPU are walked (the voxel sequence is a property of the whole ray) but not
          // updated here: their owner applies the same visits.
          slot = ownsRegion(dm, k.r) ? regionSlot(dm, rk) : -2;
          if (slot >= 0 && __ldcg(&dm.region_stamp[slot]) != b.stamp)
          {
            if (atomicExch(&dm.region_stamp[slot], b.stamp) != b.stamp)
            {
              b.touched_list[warpAggregatedInc(&b.counters->touched_count)] = (uint32_t)slot;
            }
          }
        }
        last_exit = exit;
        walked_any = true;
        if (slot < 0)
        {
          visits += (slot == -1) ? 1u : 0u;
          return;
        }
        ++visits;
        const uint32_t vid = (uint32_t)slot * g.vpr + voxelIndex(g, k);
        const uint32_t p = __ldcg(&dm.pending[vid]);
        if (p & kHitFlag)
        {
          // This voxel also receives samples in this batch: keep the miss ordered against them.
          const uint32_t run = p & ~kHitFlag;
          const uint32_t rec = warpAggregatedInc(&b.counters->record_count);
          if (rec < b.record_capacity)
          {
            b.record_ray[rec] = i;
            b.record_next[rec] = atomicExch(&b.run_head[run], (int32_t)rec);
          }
          else
          {
            atomicAdd(&b.tail_overflow[run], 1u);
            b.counters->record_overflow = 1;
          }
        }
        else
        {
          atomicAdd(&dm.pending[vid], 1u);
        }
        if (dm.traversal)
        {
          atomicAdd(&dm.traversal[vid], (float)(exit - enter));
        }
      });
    }
    if (b.last_exit)
    {
      b.last_exit[i] = walked_any ? last_exit : nan("");  // NaN: "the previous ray's" (carryLastExit)
    }
  }
  __syncwarp();
  const unsigned total = __reduce_add_sync(0xffffffffu, visits);
  if ((threadIdx.x & 31) == 0 && total)
  {
    atomicAdd(&b.counters->voxel_visits, (unsigned long long)total);
  }
}

__device__ __forceinline__ void unpackRegion(unsigned long long k, int r[3])
{
  r[0] = (int)(int16_t)(k & 0xffffu);
  r[1] = (int)(int16_t)((k >> 16) & 0xffffu);
  r[2] = (int)(int16_t)((k >> 32) & 0xffffu);
}

// Sample-voxel updates, one thread per voxel, replayed in ray order: RayMapperOccupancy.cpp:234-335.
__global__ void __launch_bounds__(128) applySamples(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned samples = 0, ordered = 0;
  if (t == 0 && !b.counters->segment_overflow)
  {
    b.counters->sample_voxels += b.counters->run_count;  // S' of the byte model; the run count is final here
  }
  if (t < b.counters->run_count && !b.counters->segment_overflow)  // (a batch whose segment list overflowed is dropped whole)
  {
    const uint32_t head = b.run_list[t];
    const uint32_t vid = b.keys_out[head];
    uint32_t k = 1;
    while (head + k < b.n && b.keys_out[head + k] == vid)
    {
      ++k;
    }
    // Sort this voxel's recorded misses into the intervals between its hits.
    uint32_t tail = b.tail_overflow[head];
    for (int32_t rec = b.run_head[head]; rec >= 0; rec = b.record_next[rec])
    {
      const uint32_t ray = b.record_ray[rec];
      uint32_t lo = 0, hi = k;  // first hit whose ray index is greater than `ray`
      while (lo < hi)
      {
        const uint32_t mid = (lo + hi) >> 1;
        if (b.vals_out[head + mid] < ray)
        {
          lo = mid + 1;
        }
        else
        {
          hi = mid;
        }
      }
      if (lo < k)
      {
        ++b.interval_count[head + lo];
      }
      else
      {
        ++tail;
      }
      ++ordered;
    }

    const uint32_t slot = vid / g.vpr;
    const uint32_t local = vid - slot * g.vpr;
    Key key;
    unpackRegion(dm.keys[slot], key.r);
    key.l[0] = (int)(local % (uint32_t)g.dim[0]);
    key.l[1] = (int)((local / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]);
    key.l[2] = (int)(local / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1]));
    double centre[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      centre[a] = voxelCentreAxis(g, key.r[a], key.l[a], a);
    }

    float value = dm.occupancy[vid];
    uint2 mean = dm.mean ? dm.mean[vid] : make_uint2(0, 0);
    uint32_t incident = dm.incident ? dm.incident[vid] : 0;
    uint32_t touch = 0;
    bool touch_set = false;
    float traversal_add = 0.0f;
    for (uint32_t j = 0; j < k; ++j)
    {
      value = missRepeat(value, b.interval_count[head + j], mp, b.ray_flags);
      const uint32_t ray = b.vals_out[head + j];
      double start[3], end[3];
      loadRay(b, ray, start, end);
      value = hitOnce(value, mp, b.ray_flags);
      uint32_t sample_count = 0;
      if (dm.mean)
      {
        const double local_pt[3] = { end[0] - centre[0], end[1] - centre[1], end[2] - centre[2] };
        mean.x = subVoxelUpdate(mean.x, mean.y, local_pt, g.res);
        sample_count = mean.y;
        ++mean.y;
      }
      if (dm.traversal)
      {
        const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
        const double len = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        traversal_add += (float)(len - staleExit(b.last_exit[ray]));
      }
      if (dm.touch_time && b.timestamps)
      {
        touch = encodeTouchTime(b.time_base, b.timestamps[ray]);
        touch_set = true;
      }
      if (dm.incident)
      {
        incident = updateIncidentNormal(incident, (float)(start[0] - end[0]), (float)(start[1] - end[1]),
                                        (float)(start[2] - end[2]), sample_count);
      }
      ++samples;
    }
    value = missRepeat(value, tail, mp, b.ray_flags);

    dm.occupancy[vid] = value;
    if (dm.mean)
    {
      dm.mean[vid] = mean;
    }
    if (dm.incident)
    {
      dm.incident[vid] = incident;
    }
    if (touch_set)
    {
      dm.touch_time[vid] = touch;
    }
    if (dm.traversal)
    {
      atomicAdd(&dm.traversal[vid], traversal_add);
    }
    if (dm.pending)
    {
      dm.pending[vid] = 0;
    }
  }
  __syncwarp();
  const unsigned s = __reduce_add_sync(0xffffffffu, samples);
  const unsigned o = __reduce_add_sync(0xffffffffu, ordered);
  if ((threadIdx.x & 31) == 0)
  {
    if (s)
    {
      atomicAdd(&b.counters->sample_updates, (unsigned long long)s);
    }
    if (o)
    {
      atomicAdd(&b.counters->ordered_records, (unsigned long long)o);
    }
  }
}

// kRfStopOnFirstOccupied (ohm/RayMapperOccupancy.cpp:183,234): a ray stops adjusting voxels — and loses its sample — once
// it has passed a voxel that was occupied WHEN THE RAY GOT THERE.  That makes every ray depend on the voxels as the rays
// before it left them, so the batch is integrated the way the CPU mapper does it: one ray after the other, each voxel
// read, adjusted and written in walk order, by ONE thread.  Exact and slow by construction; the flag belongs to
// ohm::ClearingPattern (ohm/ClearingPattern.h:45), whose batches are a few hundred rays.  Occupancy mode only
// (RayMapperNdt never raises stop_adjustments, RayMapperNdt.cpp:230; RayMapperTsdf ignores ray flags).
__global__ void __launch_bounds__(32) integrateOrdered(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  if (blockIdx.x != 0 || threadIdx.x != 0)
  {
    return;
  }
  unsigned long long accepted = 0, visits = 0, samples = 0;
  const float uninit = INFINITY;
  double last_exit = 0;  // per call, NOT per ray: a ray that visits no voxel leaves the previous ray's (RayMapperOccupancy.cpp:79,190)
  for (uint32_t i = 0; i < b.n; ++i)
  {
    double start[3], end[3];
    loadRay(b, i, start, end);
    unsigned filter_flags = 0;
    if (!applyRayFilter(mp, start, end, filter_flags))
    {
      continue;
    }
    ++accepted;
    const bool include_sample_in_ray = (filter_flags & kRffClippedEnd) || (b.ray_flags & OHMB200_RF_END_POINT_AS_FREE);
    unsigned walk_flags = (!include_sample_in_ray) ? kExcludeEndVoxel : 0u;
    walk_flags |= (b.ray_flags & OHMB200_RF_EXCLUDE_ORIGIN) ? kExcludeStartVoxel : 0u;
    bool stop = false;
    Key skey, ekey;
    if (!(b.ray_flags & OHMB200_RF_EXCLUDE_RAY) && voxelKey(g, start, skey) && voxelKey(g, end, ekey))
    {
      unsigned long long last_region = kEmptyKey;
      int slot = -1;
      walkLine(g, start, end, skey, ekey, walk_flags, [&](const Key &k, double enter, double exit) {
        const unsigned long long rk = packRegion(k.r[0], k.r[1], k.r[2]);
        if (rk != last_region)
        {
          last_region = rk;
          slot = regionSlot(dm, rk);
        }
        last_exit = exit;
        ++visits;
        if (slot < 0)
        {
          return;
        }
        const size_t vid = (size_t)slot * g.vpr + voxelIndex(g, k);
        const float v = dm.occupancy[vid];
        const bool occupied = v != uninit && v >= mp.threshold_value;
        // occupancyAdjustMiss with null_update = stop (VoxelOccupancyCompute.h:110-120)
        dm.occupancy[vid] = stop ? ((v != uninit) ? fmaxf(mp.min_value, v + 0.0f) : v) : missOnce(v, mp, b.ray_flags);
        if (dm.traversal)
        {
          dm.traversal[vid] += (float)(exit - enter);
        }
        stop = stop || ((b.ray_flags & OHMB200_RF_STOP_ON_FIRST_OCCUPIED) && occupied);
      });
    }
    if (stop || include_sample_in_ray || (b.ray_flags & OHMB200_RF_EXCLUDE_SAMPLE) || !voxelKey(g, end, ekey))
    {
      continue;
    }
    const int slot = regionSlot(dm, packRegion(ekey.r[0], ekey.r[1], ekey.r[2]));
    if (slot < 0)
    {
      continue;
    }
    const size_t vid = (size_t)slot * g.vpr + voxelIndex(g, ekey);
    dm.occupancy[vid] = hitOnce(dm.occupancy[vid], mp, b.ray_flags);
    uint32_t sample_count = 0;
    if (dm.mean)
    {
      uint2 mean = dm.mean[vid];
      const double local_pt[3] = { end[0] - voxelCentreAxis(g, ekey.r[0], ekey.l[0], 0),
                                   end[1] - voxelCentreAxis(g, ekey.r[1], ekey.l[1], 1),
                                   end[2] - voxelCentreAxis(g, ekey.r[2], ekey.l[2], 2) };
      mean.x = subVoxelUpdate(mean.x, mean.y, local_pt, g.res);
      sample_count = mean.y;
      ++mean.y;
      dm.mean[vid] = mean;
    }
    if (dm.traversal)
    {
      const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
      dm.traversal[vid] += (float)(sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) - last_exit);
    }
    if (dm.touch_time && b.timestamps)
    {
      dm.touch_time[vid] = encodeTouchTime(b.time_base, b.timestamps[i]);
    }
    if (dm.incident)
    {
      dm.incident[vid] = updateIncidentNormal(dm.incident[vid], (float)(start[0] - end[0]), (float)(start[1] - end[1]),
                                              (float)(start[2] - end[2]), sample_count);
    }
    ++samples;
  }
  atomicAdd(&b.counters->rays_accepted, accepted);
  atomicAdd(&b.counters->voxel_visits, visits);
  atomicAdd(&b.counters->sample_updates, samples);
}

// Fold the per-voxel miss counts of every region walked this batch into the occupancy layer.
__global__ void __launch_bounds__(256) resolveMisses(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t n_regions = b.counters->touched_count;
  for (uint32_t r = blockIdx.x; r < n_regions; r += gridDim.x)
  {
    const size_t base = (size_t)b.touched_list[r] * g.vpr;
    for (uint32_t v = threadIdx.x; v < g.vpr; v += blockDim.x)
    {
      const uint32_t c = dm.pending[base + v];
      if (c != 0 && !(c & kHitFlag))
      {
        dm.occupancy[base + v] = missRepeat(dm.occupancy[base + v], c, mp, b.ray_flags);
        dm.pending[base + v] = 0;
      }
    }
  }
}

#include "ohmb200_region_kernels.cuh"
#include "ohmb200_tsdf_kernels.cuh"
#include "ohmb200_exchange.cuh"

__global__ void fillFloat(float *dst, size_t n, float value)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    dst[i] = value;
  }
}

// Gather region chunks of one layer into a contiguous staging buffer (16-byte vectors).
__global__ void gatherRegions(const uint4 *slab, const uint32_t *slots, uint4 *dst, size_t vec_per_region)
{
  const uint4 *src = slab + (size_t)slots[blockIdx.x] * vec_per_region;
  uint4 *out = dst + (size_t)blockIdx.x * vec_per_region;
  for (size_t i = threadIdx.x; i < vec_per_region; i += blockDim.x)
  {
    out[i] = src[i];
  }
}

// RayMapperSecondarySample::integrateRays (ohm/RayMapperSecondarySample.cpp:37-74), step 1: the voxel of every ray's END
// point (region find-or-insert) as a sort key, the range |end - start| (glm::length) beside it.
__global__ void __launch_bounds__(128) prepSecondary(DeviceMap dm, Geom g, Batch b, double *ranges)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n)
  {
    return;
  }
  double start[3], end[3];
  loadRay(b, i, start, end);
  const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
  ranges[i] = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  uint32_t vid = kInvalidVoxel;
  Key key;
  if (voxelKey(g, end, key) && ownsRegion(dm, key.r))
  {
    const int slot = regionSlot(dm, packRegion(key.r[0], key.r[1], key.r[2]));
    if (slot >= 0)
    {
      vid = (uint32_t)slot * g.vpr + voxelIndex(g, key);
    }
  }
  b.keys_in[i] = vid;
  b.vals_in[i] = i;
}

// Step 2, one thread per voxel run of the sorted pairs: addSecondarySample (VoxelSecondarySample.h:87-99) in ray order.
__global__ void __launch_bounds__(128) applySecondary(DeviceMap dm, Batch b, const double *ranges)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned samples = 0;
  if (t < b.counters->run_count)
  {
    const uint32_t head = b.run_list[t];
    const uint32_t vid = b.keys_out[head];
    const uint2 raw = dm.secondary[vid];
    float m2 = __uint_as_float(raw.x);
    uint32_t range_mean_q = raw.y & 0xffffu, count = raw.y >> 16;
    const double quantisation = 1000.0;
    const double max_range = (65535 - 1u) / quantisation;
    for (uint32_t j = head; j < b.n && b.keys_out[j] == vid; ++j)
    {
      double range = ranges[b.vals_out[j]];
      range = (range < max_range) ? range : max_range;
      double range_mean = range_mean_q / quantisation;
      count = (count + 1u) & 0xffffu;  // uint16_t ++
      const double delta = range - range_mean;
      range_mean += delta / count;
      range_mean_q = (uint32_t)(uint16_t)(range_mean * quantisation);
      const double delta2 = range - range_mean;
      m2 += (float)(delta * delta2);
      ++samples;
    }
    dm.secondary[vid] = make_uint2(__float_as_uint(m2), range_mean_q | (count << 16));
  }
  samples = __reduce_add_sync(0xffffffffu, samples);
  if ((threadIdx.x & 31u) == 0 && samples)
  {
    atomicAdd(&b.counters->sample_updates, (unsigned long long)samples);
  }
}

// ohm::RaysQuery::onExecute (ohm/RaysQuery.cpp:109-199; the OpenCL form is ohmgpu/gpu/RaysQuery.cl) on the resident map,
// one thread per ray: walk until the first occupied voxel.  Same fp64 walk as the mappers, so ranges, volumes, the
// terminal state and the terminal key are those of the CPU query on the same map.  A region that is not resident reads
// as unobserved.  (The reference never resets terminal state/key between rays: a ray whose keys are out of range
// repeats the previous ray's answer there; here it reports kNull.)
__global__ void __launch_bounds__(128) raysQuery(DeviceMap dm, Geom g, MapParams mp, const double *rays, uint32_t n,
                                                 double volume_coefficient, double *ranges, double *volumes,
                                                 int *states, int32_t *keys)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
  {
    return;
  }
  double start[3], end[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    start[a] = rays[(size_t)i * 6 + a];
    end[a] = rays[(size_t)i * 6 + 3 + a];
  }
  double unobserved_volume = 0.0;
  float range = 0.0f;
  int state = -2;  // OccupancyType::kNull
  Key terminal;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    terminal.r[a] = terminal.l[a] = 0;
  }
  unsigned filter_flags = 0;
  Key skey, ekey;
  if (applyRayFilter(mp, start, end, filter_flags) && voxelKey(g, start, skey) && voxelKey(g, end, ekey))
  {
    Walk w;
    walkInit(w, g, start, end, skey, ekey);
    unsigned long long last_region = kEmptyKey;
    int slot = -1;
    double last_time = 0;
    bool go_on = true;
    // the visit lambda of RaysQuery.cpp:129-158
    auto visit = [&](const Key &key, double enter_range, double exit_range) {
      const unsigned long long region = packRegion(key.r[0], key.r[1], key.r[2]);
      if (region != last_region)
      {
        slot = regionFind(dm, region);
        last_region = region;
      }
      const float value = (slot >= 0) ? dm.occupancy[(size_t)slot * g.vpr + voxelIndex(g, key)] : INFINITY;
      const bool is_unobserved = value == INFINITY;
      const bool is_occupied = !is_unobserved && value > mp.threshold_value;
      unobserved_volume +=
        is_unobserved ?
          (volume_coefficient * (exit_range * exit_range * exit_range - enter_range * enter_range * enter_range)) :
          0.0;
      range = (!is_occupied) ? (float)exit_range : range;
      state = is_unobserved ? -1 : (is_occupied ? 1 : 0);
      terminal = key;
      go_on = !is_occupied;
    };
    // walkLineVoxels with an aborting visitor (LineWalkCompute.h:392-410)
    const unsigned max_steps = (unsigned)(abs(w.remaining[0]) + abs(w.remaining[1]) + abs(w.remaining[2]));
    unsigned count = 0;
    while (go_on && w.limit < 7u && !walkAtEnd(w) && count <= max_steps)
    {
      const double t = walkNextTime(w);
      visit(w.cur, last_time, t);
      last_time = t;
      ++count;
      walkStep(w, g);
    }
    if (go_on)
    {
      visit(ekey, last_time, w.length);
    }
  }
  ranges[i] = (double)range;
  volumes[i] = unobserved_volume;
  states[i] = state;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    keys[(size_t)i * 6 + a] = terminal.r[a];
    keys[(size_t)i * 6 + 3 + a] = terminal.l[a];
  }
}

// ohm::LineKeysQuery::onExecute (ohm/LineKeysQuery.cpp:103-123; OpenCL form ohmgpu/gpu/LineKeys.cl): the voxel keys
// along each line, calculateSegmentKeys(..., include_end_point = true) = walkSegmentKeys with no flags.  Pure geometry
// (resolution, region dimensions, origin); no ray filter, no map data.  kFill = false counts the keys of every line,
// kFill = true writes them at the offsets an exclusive scan of the counts gave.
template <bool kFill>
__global__ void __launch_bounds__(128) lineKeys(Geom g, const double *rays, uint32_t n, uint32_t *counts,
                                                const uint32_t *offsets, int32_t *keys)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
  {
    return;
  }
  double start[3], end[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    start[a] = rays[(size_t)i * 6 + a];
    end[a] = rays[(size_t)i * 6 + 3 + a];
  }
  Key skey, ekey;
  uint32_t count = 0;
  if (voxelKey(g, start, skey) && voxelKey(g, end, ekey))  // a null key: no voxels (LineWalk.h:119-122)
  {
    int32_t *out = kFill ? keys + (size_t)offsets[i] * 6 : nullptr;
    count = walkLine(g, start, end, skey, ekey, 0u, [&](const Key &key, double, double) {
      if (kFill)
      {
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          out[a] = key.r[a];
          out[3 + a] = key.l[a];
        }
        out += 6;
      }
    });
  }
  if (!kFill)
  {
    counts[i] = count;
  }
}

// ohmb200_clear: wipe only the regions that exist (the slabs of free slots are clean already), then free the table.
struct ClearTable
{
  uint32_t *base[OHMB200_LAYER_COUNT + 3];   // layer slabs, then voxel_bits, near bits, pending
  uint32_t words[OHMB200_LAYER_COUNT + 3];   // 32-bit words per region
  uint32_t fill[OHMB200_LAYER_COUNT + 3];
  int count;
};

// Resets every layer chunk of one slot to its clear value (one CTA).
__device__ __forceinline__ void clearSlot(const ClearTable &table, uint32_t slot)
{
  for (int l = 0; l < table.count; ++l)
  {
    uint32_t *chunk = table.base[l] + (size_t)slot * table.words[l];
    const uint32_t fill = table.fill[l];
    if ((table.words[l] & 3u) == 0)
    {
      uint4 *chunk4 = reinterpret_cast<uint4 *>(chunk);
      for (uint32_t w = threadIdx.x; w < (table.words[l] >> 2); w += blockDim.x)
      {
        chunk4[w] = make_uint4(fill, fill, fill, fill);
      }
    }
    else
    {
      for (uint32_t w = threadIdx.x; w < table.words[l]; w += blockDim.x)
      {
        chunk[w] = fill;
      }
    }
  }
}

__global__ void __launch_bounds__(256) clearLiveRegions(DeviceMap dm, ClearTable table)
{
  for (uint32_t slot = blockIdx.x; slot < dm.capacity; slot += gridDim.x)
  {
    const unsigned long long key = dm.keys[slot];
    if (key == kEmptyKey)
    {
      continue;
    }
    if (key != kTombKey)  // the chunks of an evicted slot were cleared when it was evicted
    {
      clearSlot(table, slot);
    }
    __syncthreads();  // every thread has read keys[slot] before it is freed
    if (threadIdx.x == 0)
    {
      dm.keys[slot] = kEmptyKey;
      dm.region_stamp[slot] = 0;
    }
  }
}

// Eviction (GpuLayerCache.cpp:536-600): the chunks of the listed slots have been copied to the host store; clear them
// for their next tenant.  The keys themselves are rewritten by the host (tombstones, see makeRoom).
__global__ void __launch_bounds__(256) clearEvictedSlots(DeviceMap dm, ClearTable table, const uint32_t *slots, uint32_t count)
{
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x)
  {
    clearSlot(table, slots[i]);
    if (threadIdx.x == 0)
    {
      dm.region_stamp[slots[i]] = 0;
    }
  }
}

// Region key -> slab slot for a list of keys (0xFFFFFFFF and *missing = 1 when a key is not resident).
__global__ void lookupSlots(DeviceMap dm, const unsigned long long *keys, uint32_t count, uint32_t *slots, int *missing)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
  {
    return;
  }
  const int slot = regionFind(dm, keys[i]);
  slots[i] = (uint32_t)slot;
  if (slot < 0)
  {
    *missing = 1;
  }
}

// gatherRegions for a slot list that may hold 0xFFFFFFFF (left untouched).
__global__ void gatherRegionsChecked(const uint4 *slab, const uint32_t *slots, uint4 *dst, size_t vec_per_region)
{
  const uint32_t slot = slots[blockIdx.x];
  if (slot == 0xFFFFFFFFu)
  {
    return;
  }
  const uint4 *src = slab + (size_t)slot * vec_per_region;
  uint4 *out = dst + (size_t)blockIdx.x * vec_per_region;
  for (size_t i = threadIdx.x; i < vec_per_region; i += blockDim.x)
  {
    out[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host-side map
// ---------------------------------------------------------------------------------------------------------
static const size_t kLayerBytes[OHMB200_LAYER_COUNT] = { 4, 8, 4, 4, 4, 24, 8, 8, 8, 8 };

enum KernelId
{
  kKPrep = 0,
  kKSort,
  kKMark,
  kKWalk,
  kKSamples,
  kKResolve,
  kKGather,
  kKFill,
  kKPrepRays,
  kKPrepSegments,
  kKPlan,
  kKEmit,
  kKWalkRegions,
  kKLink,
  kKScatter,
  kKTsdfMark,
  kKTsdfClear,
  kKTsdfReplay,
  kKNdtGauss,
  kKNdtClamp,
  kKOrdered,
  kKExPrepRays,
  kKExRoute,
  kKExSegments,
  kKExWait,
  kKExBin,
  kKExBinSamples,
  kKExEmit,
  kKernelCount
};
static const char *kKernelNames[kKernelCount] = { "prepSamples",  "radixSort",     "markRuns",      "walkRays",
                                                  "applySamples", "resolveMisses", "gatherRegions", "fillFloat",
                                                  "prepRays",     "prepSegments",  "planRegions",   "emitSegments",  "walkRegions",
                                                  "linkRecords",  "scatterRecords", "markTsdfNear",  "clearTouchedBits", "replayTsdf",
                                                  "ndtGaussianMisses", "ndtClampGaussians", "integrateOrdered",
                                                  "exPrepRays", "exRouteSamples", "exPrepSegments", "exWait", "exBinSegments", "exBinSamples", "exEmit" };
static_assert(kKernelCount <= OHMB200_KERNEL_SLOTS, "raise OHMB200_KERNEL_SLOTS");

struct ohmb200_map
{
  int device = 0;
  int mode = 0;
  int sm_count = 148;
  int algo = 1;           // 1 = region-binned walk (shared-memory tiles), 0 = one thread per ray (global counters)
  size_t tile_bytes = 0;  // dynamic shared memory of walkRegions
  TileLayout tile;        // layout of the shared-memory counter tile
  int walk_ctas_per_sm = 1;
  bool slab_walk = false;   // occupancy maps without traversal: walkRegionsSlab (one 1024-thread CTA per SM, slab staged by TMA)
  size_t slab_walk_bytes = 0;
  uint32_t heavy_run = 16;
  uint32_t seg_factor = 96;     // segments per ray the batch's segment list is sized for (doubled after an overflow)
  uint32_t record_factor = 24;  // ordered-miss records per ray, likewise
  uint32_t *tsdf_near = nullptr;  // TSDF: per-batch bit per voxel, "visited near a sample in this batch"
  size_t voxel_bit_bytes = 0;     // size of dm.voxel_bits (and of tsdf_near)
  ohmb200_params params{};
  Geom geom{};
  MapParams mp{};
  DeviceMap dm{};
  void *layer_slab[OHMB200_LAYER_COUNT] = {};
  size_t region_layer_bytes[OHMB200_LAYER_COUNT] = {};
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // sample path of a batch, concurrent with the segment path
  cudaEvent_t fork_event = nullptr, join_event = nullptr;
  cudaStream_t stream = nullptr;
  Counters *d_counters = nullptr;
  Counters *h_counters = nullptr;  // pinned
  double first_ray_time = -1.0;
  uint32_t stamp = 0;
  // CUDA graphs of whole batches (occupancy / NDT, region-binned path): a batch is ~30 stream operations; replaying
  // it as one graph launch removes the gaps between them and most of the host cost.  Keyed by everything a kernel
  // argument is built from; dropped whenever parameters, scratch buffers or the partition change.
  struct BatchGraph
  {
    const void *rays, *intensities, *timestamps;
    size_t n;
    unsigned ray_flags;
    double time_base;
    cudaGraphExec_t exec;
    uint64_t launches;
  };
  std::vector<BatchGraph> graphs;
  std::vector<BatchGraph> seen;  // shapes met once (exec unused): a shape is recorded the second time it comes
  bool capturing = false;
  bool use_graphs = true;
  uint64_t rays_in = 0;
  uint64_t batches = 0;
  uint64_t launches = 0;
  // batch scratch
  Batch batch{};
  size_t scratch_rays = 0;
  void *cub_temp = nullptr;
  size_t cub_temp_bytes = 0;
  void *carry_temp = nullptr;  // scan storage of carryLastExit (traversal layer)
  size_t carry_temp_bytes = 0;
  int sort_bits = 32;
  // host -> device input staging (double buffered)
  double *d_rays[2] = {};
  float *d_intensities[2] = {};
  double *d_timestamps[2] = {};
  size_t in_capacity[2] = {};
  cudaEvent_t in_ready[2] = {};
  cudaEvent_t in_free[2] = {};
  int next_input = 0;
  // gather staging
  void *d_gather = nullptr;
  size_t gather_bytes = 0;
  uint32_t *d_gather_slots = nullptr;
  size_t gather_slots_cap = 0;
  // asynchronous download (ohmb200_read_regions_async): two staging sets, D2H on its own stream
  cudaStream_t download_stream = nullptr;
  void *d_stage[2] = { nullptr, nullptr };
  size_t stage_bytes[2] = { 0, 0 };
  unsigned long long *d_stage_keys[2] = { nullptr, nullptr };
  unsigned long long *h_stage_keys[2] = { nullptr, nullptr };  // pinned
  uint32_t *d_stage_slots[2] = { nullptr, nullptr };
  size_t stage_keys_cap[2] = { 0, 0 };
  cudaEvent_t stage_gathered[2] = { nullptr, nullptr };
  cudaEvent_t stage_done[2] = { nullptr, nullptr };
  bool stage_pending[2] = { false, false };
  int stage_next = 0;
  void *d_query = nullptr;  // staging of ohmb200_rays_query / ohmb200_line_keys_query
  size_t query_bytes = 0;
  void *d_query_keys = nullptr;
  size_t query_keys_bytes = 0;
  double *d_secondary_ranges = nullptr;  // ohmb200_integrate_secondary scratch
  size_t secondary_bytes = 0;
  void *d_secondary_rays = nullptr;
  size_t secondary_rays_bytes = 0;
  int *d_lookup_missing = nullptr;  // device flag: an asynchronous download named a region that is not resident
  // Paging (the GpuLayerCache of this build, ohmgpu/GpuLayerCache.cpp:429-633): when the region table runs out of
  // slots the least recently walked regions are copied to this host store and their slots freed; a region that is
  // touched again is restored before the batch updates it.  Value = every enabled layer's chunk, in layer order.
  std::unordered_map<unsigned long long, std::vector<char>> store;
  size_t store_layer_offset[OHMB200_LAYER_COUNT] = {};
  size_t store_chunk_bytes = 0;
  uint64_t regions_bound = 0;  // upper bound of the resident regions, without asking the device
  // ... kept tight by a copy of the device's region counter queued behind every batch and read once it has landed
  unsigned long long *h_region_snap = nullptr;  // pinned
  cudaEvent_t snap_event = nullptr;
  bool snap_pending = false;
  uint64_t snap_batch = 0;     // batches queued when the pending snapshot was taken
  uint64_t known_regions = 0;  // the last snapshot that landed, and the batch count it belongs to
  uint64_t known_batch = 0;
  uint32_t region_reserve = 0; // free slots a batch may need (ohmb200_set_region_reserve)
  uint64_t evicted = 0, paged_in = 0;
  // Multi-GPU routed exchange (ohmb200_exchange.cuh)
  struct Exchange
  {
    bool open = false, connected = false, pending = false;
    bool local_peers = false;  // some peer map lives in this process: its sends are queued by the same host thread
    bool forked = false;       // the pending step's sample branch runs on the side stream (join_event recorded)
    int rank = 0, world = 1;
    uint32_t per = 0, seg_cap = 0, step = 0;
    char *arena = nullptr;
    size_t arena_bytes = 0, parity_bytes = 0;
    char *peer_base[kMaxWorld] = {};
    bool peer_mapped[kMaxWorld] = {};  // opened with cudaIpcOpenMemHandle (to be closed)
    uint32_t *out_counts = nullptr;    // [2 * kMaxWorld] segments / samples sent to each owner this step
    unsigned long long *smp_key = nullptr;  // sender-side parking of the own rays' samples (see ExStep)
    uint32_t *smp_voxel = nullptr, *smp_owner = nullptr;
    double *smp_last_exit = nullptr;
    int *abort = nullptr;
    uint32_t *d_step = nullptr;        // == step, counted on the device
    uint32_t *d_barrier = nullptr;     // ohmb200_exchange_barrier calls so far, counted on the device
    ExMailbox *mailbox_now = nullptr;  // this rank's mailbox of the step being integrated
    cudaStream_t stream = nullptr;     // the per-ray broadcast (copy engines) runs here, beside the cut
    cudaEvent_t prepped = nullptr, bcast_done = nullptr;
    // CUDA graphs of whole steps (send + integrate), one per (buffers, size, flags, parity): see launchBatch's graphs
    struct StepGraph
    {
      const void *rays, *intensities, *timestamps;
      size_t n;
      unsigned ray_flags;
      int parity;
      double time_base;
      cudaGraphExec_t exec;
      uint64_t launches;
    };
    std::vector<StepGraph> graphs, seen;
    bool capturing = false, replayed = false;
    StepGraph current{};
    uint64_t launches_before = 0;
    int host_buf = -1;                 // input staging buffer of the pending step (ohmb200_exchange_send), -1: device rays
    size_t n_own = 0;
    unsigned ray_flags = 0;
    bool has_timestamps = false, has_intensities = false;
  } ex;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> event_pool;
  struct Span
  {
    int kernel;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;
  size_t events_used = 0;
  double kernel_ms[kKernelCount] = {};
  uint64_t kernel_launches[kKernelCount] = {};
};

namespace
{
struct KernelScope
{
  ohmb200_map *m;
  int id;
  cudaEvent_t a = nullptr, b = nullptr;
  KernelScope(ohmb200_map *map, int kernel)
    : m(map)
    , id(kernel)
  {
    ++m->launches;
    if (m->profiling)
    {
      a = take();
      b = take();
      cudaEventRecord(a, m->stream);
    }
  }
  ~KernelScope()
  {
    if (m->profiling)
    {
      cudaEventRecord(b, m->stream);
      m->spans.push_back({ id, a, b });
    }
  }
  cudaEvent_t take()
  {
    if (m->events_used == m->event_pool.size())
    {
      cudaEvent_t e;
      cudaEventCreate(&e);
      m->event_pool.push_back(e);
    }
    return m->event_pool[m->events_used++];
  }
};

void drainSpans(ohmb200_map *m)
{
  if (m->spans.empty())
  {
    return;
  }
  cudaStreamSynchronize(m->stream);
  for (const auto &s : m->spans)
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess)
    {
      m->kernel_ms[s.kernel] += ms;
      ++m->kernel_launches[s.kernel];
    }
  }
  m->spans.clear();
  m->events_used = 0;
}

void dropBatchGraphs(ohmb200_map *m)
{
  for (auto &g : m->graphs)
  {
    cudaGraphExecDestroy(g.exec);
  }
  m->graphs.clear();
  m->seen.clear();
  // the recorded exchange steps bake the same things in (parameters, scratch pointers, the partition)
  if (!m->ex.capturing)
  {
    for (auto &g : m->ex.graphs)
    {
      cudaGraphExecDestroy(g.exec);
    }
    m->ex.graphs.clear();
    m->ex.seen.clear();
  }
}

void refreshParams(ohmb200_map *m)
{
  dropBatchGraphs(m);
  const ohmb200_params &p = m->params;
  Geom &g = m->geom;
  g.res = p.resolution;
  g.vpr = 1;
  for (int a = 0; a < 3; ++a)
  {
    g.dim[a] = p.region_dim[a];
    g.region_size[a] = p.region_dim[a] * p.resolution;  // OccupancyMap.cpp:204-206
    g.origin[a] = p.origin[a];
    g.vpr *= (uint32_t)p.region_dim[a];
  }
  MapParams &mp = m->mp;
  mp.hit_value = p.hit_value;
  mp.miss_value = p.miss_value;
  mp.min_value = p.min_value;
  mp.max_value = p.max_value;
  mp.threshold_value = p.threshold_value;
  mp.sat_min = p.saturate_min ? p.min_value : -3.402823466e+38f;
  mp.sat_max = p.saturate_max ? p.max_value : 3.402823466e+38f;
  mp.filter_kind = p.filter_kind;
  mp.filter_range = p.filter_range;
  memcpy(mp.clip_box, p.clip_box, sizeof(mp.clip_box));
  mp.sensor_noise = p.sensor_noise;
  mp.adaptation_rate = p.adaptation_rate;
  mp.reinit_threshold = p.reinit_threshold;
  mp.initial_intensity_cov = p.initial_intensity_cov;
  mp.reinit_count = p.reinit_count;
  mp.sample_threshold = p.sample_threshold;
  mp.ndt_tm = p.ndt_tm;
  mp.tsdf_max_weight = p.tsdf_max_weight;
  mp.tsdf_trunc = p.tsdf_trunc;
  mp.tsdf_dropoff = p.tsdf_dropoff;
  mp.tsdf_sparsity = p.tsdf_sparsity;
}

int initialiseSlabs(ohmb200_map *m)
{
  const size_t voxels = (size_t)m->dm.capacity * m->geom.vpr;
  CUDA_TRY(cudaMemsetAsync(m->dm.keys, 0xFF, sizeof(unsigned long long) * m->dm.capacity, m->stream));
  CUDA_TRY(cudaMemsetAsync(m->dm.region_stamp, 0, sizeof(uint32_t) * m->dm.capacity, m->stream));
  if (m->dm.pending)
  {
    CUDA_TRY(cudaMemsetAsync(m->dm.pending, 0, sizeof(uint32_t) * voxels, m->stream));
  }
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    if (!m->layer_slab[l])
    {
      continue;
    }
    if (l == OHMB200_LAYER_OCCUPANCY)
    {
      KernelScope scope(m, kKFill);
      fillFloat<<<m->sm_count * 8, 256, 0, m->stream>>>((float *)m->layer_slab[l], voxels, INFINITY);
    }
    else
    {
      CUDA_TRY(cudaMemsetAsync(m->layer_slab[l], 0, kLayerBytes[l] * voxels, m->stream));
    }
  }
  if (m->dm.voxel_bits)
  {
    CUDA_TRY(cudaMemsetAsync(m->dm.voxel_bits, 0, m->voxel_bit_bytes, m->stream));
  }
  if (m->tsdf_near)
  {
    CUDA_TRY(cudaMemsetAsync(m->tsdf_near, 0, m->voxel_bit_bytes, m->stream));
  }
  CUDA_TRY(cudaMemsetAsync(m->d_counters, 0, sizeof(Counters), m->stream));
  CUDA_TRY(cudaGetLastError());
  return OHMB200_OK;
}

template <typename T>
int deviceAlloc(T *&ptr, size_t count)
{
  void *p = nullptr;
  CUDA_TRY(cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1)));
  ptr = (T *)p;
  return OHMB200_OK;
}

int pageInNewRegions(ohmb200_map *m);
int ensureRoom(ohmb200_map *m);

int ensureScratch(ohmb200_map *m, size_t n)
{
  if (n <= m->scratch_rays)
  {
    return OHMB200_OK;
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  dropBatchGraphs(m);
  Batch &b = m->batch;
  cudaFree(b.keys_in);
  cudaFree(b.keys_out);
  cudaFree(b.vals_in);
  cudaFree(b.vals_out);
  cudaFree(b.run_list);
  cudaFree(b.run_head);
  cudaFree(b.interval_count);
  cudaFree(b.interval_offset);
  cudaFree(b.sorted_rays);
  b.interval_offset = nullptr;
  b.sorted_rays = nullptr;
  cudaFree(b.record_ray);
  cudaFree(b.record_next);
  cudaFree(b.last_exit);
  cudaFree(m->cub_temp);
  b.last_exit = nullptr;
  // Grow with headroom, in 16 Ki-ray steps: a stream of sweeps whose size creeps up (clipped rays vary from sweep to
  // sweep) must not pay ~50 cudaFree/cudaMalloc calls — tens of milliseconds — every time it sets a new maximum.
  const size_t cap = std::max<size_t>(((n + n / 8 + 16383) / 16384) * 16384, 4096);
  int rc = 0;
  rc |= deviceAlloc(b.keys_in, cap);
  rc |= deviceAlloc(b.keys_out, cap);
  rc |= deviceAlloc(b.vals_in, cap);
  rc |= deviceAlloc(b.vals_out, cap);
  rc |= deviceAlloc(b.run_list, cap);
  rc |= deviceAlloc(b.run_head, cap);
  rc |= deviceAlloc(b.interval_count, 2 * cap + 1);
  b.tail_overflow = b.interval_count + cap;  // re-pointed at interval_count + n by every batch
  // Ordered-record lists are sized by the rays whose visits this map applies: on an exchange map that is about one
  // rank's share (twice it, for imbalance), not the whole step's rays — the lists are cleared every batch, and at 8 GPUs a
  // list sized by all rays was a 113 MB memset per step.
  const size_t record_rays = m->ex.open ? std::min<size_t>(cap, 2 * (size_t)m->ex.per + 16384) : cap;
  b.record_capacity = (uint32_t)std::min<size_t>(std::max<size_t>(record_rays * m->record_factor, 1u << 20), 1u << 28);
  rc |= deviceAlloc(b.record_ray, b.record_capacity);
  rc |= deviceAlloc(b.record_next, b.record_capacity);
  if (m->dm.traversal)
  {
    rc |= deviceAlloc(b.last_exit, cap);
    cudaFree(m->carry_temp);
    m->carry_temp = nullptr;
    m->carry_temp_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, m->carry_temp_bytes, b.last_exit, b.last_exit, CarryValid(), (int)cap, m->stream);
    rc |= cudaMalloc(&m->carry_temp, std::max<size_t>(m->carry_temp_bytes, 16)) == cudaSuccess ? 0 : 1;
  }
  if (m->algo == 1)
  {
    cudaFree(b.recs);
    cudaFree(b.ray_length);
    cudaFree(b.record_vid);
    cudaFree(b.segments);
    cudaFree(b.items);
    cudaFree(b.stage);
    cudaFree(b.stage_count);
    cudaFree(b.cross_pos);
    b.cross_pos = nullptr;
    b.ray_length = nullptr;
    b.recs = nullptr;
    b.stage = nullptr;
    b.stage_count = nullptr;
    if (!m->ex.open)
    {
      rc |= deviceAlloc(b.recs, cap);
      if (m->dm.traversal)
      {
        rc |= deviceAlloc(b.ray_length, cap);
      }
    }
    rc |= deviceAlloc(b.record_vid, b.record_capacity);
    b.seg_capacity = (uint32_t)std::min<size_t>(cap * m->seg_factor, 0xFFFFFFF0u);
    if (m->ex.open)
    {
      // an exchange map bins what its inboxes hold; the per-ray arrays (walk constants, lengths) live in the arena and
      // the segments arrive cut: no staging planes
      b.seg_capacity = (uint32_t)std::min<size_t>((size_t)m->ex.world * m->ex.seg_cap, 0xFFFFFFF0u);
    }
    rc |= deviceAlloc(b.segments, b.seg_capacity);
    b.stage_stride = (uint32_t)cap;
    if (!m->ex.open)
    {
      rc |= deviceAlloc(b.stage, (size_t)kStageSegments * cap);
      rc |= deviceAlloc(b.stage_count, cap);
    }
    b.item_capacity = m->dm.capacity + b.seg_capacity / 512 + 16;
    rc |= deviceAlloc(b.items, b.item_capacity);
    if (m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM)
    {
      cudaFree(b.gauss_keys);
      rc |= deviceAlloc(b.interval_offset, 2 * cap + 1);
      rc |= deviceAlloc(b.sorted_rays, b.record_capacity);
      b.gauss_capacity = (uint32_t)std::min<size_t>(std::max<size_t>(record_rays * (m->record_factor * 2u / 3u), 1u << 20), 1u << 28);
      rc |= deviceAlloc(b.gauss_keys, b.gauss_capacity);
    }
    if (m->mode == OHMB200_MODE_TSDF)
    {
      cudaFree(b.record_keys);
      cudaFree(b.record_keys_sorted);
      rc |= deviceAlloc(b.record_keys, b.record_capacity);
      rc |= deviceAlloc(b.record_keys_sorted, b.record_capacity);
    }
  }
  if (rc)
  {
    return OHMB200_E_CUDA;
  }
  m->cub_temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, m->cub_temp_bytes, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)cap, 0,
                                  m->sort_bits, m->stream);
  if (b.interval_offset)
  {
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, b.interval_count, b.interval_offset, (int)(2 * cap + 1), m->stream);
    m->cub_temp_bytes = std::max(m->cub_temp_bytes, scan_bytes);
  }
  if (m->mode == OHMB200_MODE_TSDF && m->algo == 1)
  {
    size_t keys_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, keys_bytes, b.record_keys, b.record_keys_sorted, (int)b.record_capacity, 0, 64,
                                   m->stream);
    m->cub_temp_bytes = std::max(m->cub_temp_bytes, keys_bytes);
  }
  CUDA_TRY(cudaMalloc(&m->cub_temp, std::max<size_t>(m->cub_temp_bytes, 16)));
  m->scratch_rays = cap;
  return OHMB200_OK;
}

// Rays that walked no voxel take the exit range of the nearest earlier ray that did (see staleExit).
int carryLastExit(ohmb200_map *m, size_t n, cudaStream_t s, double *last_exit = nullptr)
{
  size_t temp = m->carry_temp_bytes;
  last_exit = last_exit ? last_exit : m->batch.last_exit;
  CUDA_TRY(cub::DeviceScan::InclusiveScan(m->carry_temp, temp, last_exit, last_exit, CarryValid(), (int)n, s));
  ++m->launches;
  return OHMB200_OK;
}

// The second half of an occupancy / NDT batch on the region-binned path: the segments are binned by region, the sample
// pairs sorted (both streams joined).  Walk -> (NDT: Gaussian misses) -> link records -> sample replay.
int launchWalkAndReplay(ohmb200_map *m, const Batch &b, cudaStream_t s, size_t n, bool has_samples)
{
  const unsigned threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  {
    KernelScope scope(m, kKWalkRegions);
    if ((m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM))
    {
      CUDA_TRY(cudaMemsetAsync(b.gauss_keys, 0xFF, sizeof(unsigned long long) * b.gauss_capacity, s));
      const auto kernel = m->dm.traversal ? walkRegionsNdt<true> : walkRegionsNdt<false>;
      kernel<<<m->sm_count * m->walk_ctas_per_sm, kWalkThreads, m->tile_bytes, s>>>(m->dm, m->geom, m->mp, b, m->tile);
    }
    else if (m->slab_walk && !m->dm.traversal)
    {
      walkRegionsSlab<<<m->sm_count, kSlabThreads, m->slab_walk_bytes, s>>>(m->dm, m->geom, m->mp, b, m->tile, has_samples ? 1 : 0);
    }
    else
    {
      const auto kernel = m->dm.traversal ? walkRegions<true> : walkRegions<false>;
      kernel<<<m->sm_count * m->walk_ctas_per_sm, kWalkThreads, m->tile_bytes, s>>>(m->dm, m->geom, m->mp, b, m->tile,
                                                                                 has_samples ? 1 : 0);
    }
  }
  if (m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM)
  {
    if (m->ex.open)
    {
      // the rays of the other ranks (read by the Gaussian-miss and replay kernels) follow their walk constants: stage 3
      KernelScope scope(m, kKExWait);
      exWait<<<1, 32, 0, s>>>(m->ex.mailbox_now, m->ex.world, m->ex.d_step, 3, m->ex.abort);
    }
    {
      KernelScope scope(m, kKNdtGauss);
      ndtGaussianMisses<<<m->sm_count * 8, 128, 0, s>>>(m->dm, m->geom, m->mp, b);
    }
    {
      KernelScope scope(m, kKNdtClamp);
      ndtClampGaussians<<<m->sm_count * 4, 256, 0, s>>>(m->dm, m->mp, b);
    }
  }
  if (has_samples)
  {
    const bool ndt_mode = m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM;
    {
      KernelScope scope(m, kKLink);
      linkRecords<<<m->sm_count * 4, 256, 0, s>>>(b, ndt_mode ? 1 : 0, m->geom.vpr);
    }
    if (ndt_mode)
    {
      // group the records by interval: exclusive scan of the interval counts, then one scatter
      KernelScope scope(m, kKScatter);
      size_t temp = m->cub_temp_bytes;
      cub::DeviceScan::ExclusiveSum(m->cub_temp, temp, b.interval_count, b.interval_offset, (int)(2 * n + 1), s);
      scatterRecords<<<m->sm_count * 4, 256, 0, s>>>(b);
    }
    {
      KernelScope scope(m, kKSamples);
      if ((m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM))
      {
        applySamplesNdt<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b);
        applySamplesNdtHeavy<<<m->sm_count * 8, 128, 0, s>>>(m->dm, m->geom, m->mp, b);
      }
      else
      {
        applySamples<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b);
      }
    }
  }
  return OHMB200_OK;
}

int launchBatch(ohmb200_map *m, const double *d_rays, size_t n, const float *d_intensities, const double *d_timestamps,
                unsigned ray_flags)
{
  if (n == 0)
  {
    return OHMB200_OK;
  }
  if (m->ex.open)
  {
    return setError(OHMB200_E_INVALID, "the map has an open exchange: integrate through ohmb200_exchange_send / "
                                       "ohmb200_exchange_integrate (or ohmb200_exchange_close first)");
  }
  int rc = ensureScratch(m, n);
  if (rc)
  {
    return rc;
  }
  Batch &b = m->batch;
  b.rays = d_rays;
  b.intensities = d_intensities;
  b.timestamps = (m->dm.touch_time) ? d_timestamps : nullptr;
  b.n = (uint32_t)n;
  b.heavy_run = m->heavy_run;
  b.ray_flags = ray_flags;
  b.stamp = ++m->stamp;
  b.time_base = m->first_ray_time;
  b.counters = m->d_counters;
  cudaStream_t s = m->stream;
  const unsigned threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  const bool has_samples = m->mode != OHMB200_MODE_TSDF;

  if ((ray_flags & OHMB200_RF_STOP_ON_FIRST_OCCUPIED) && m->mode == OHMB200_MODE_OCCUPANCY)
  {
    // order-dependent across rays: the exact, sequential path (see integrateOrdered)
    if (m->dm.part_world > 1)
    {
      return setError(OHMB200_E_INVALID, "kRfStopOnFirstOccupied needs the whole map on one GPU: a ray's stop depends on "
                                         "voxels of every region it crosses");
    }
    if (!m->store.empty())
    {
      return setError(OHMB200_E_INVALID, "kRfStopOnFirstOccupied is not available while part of the map is paged out");
    }
    CUDA_TRY(cudaMemsetAsync(&m->d_counters->record_count, 0, sizeof(uint32_t) * kPerBatchCounterWords, s));
    {
      KernelScope scope(m, kKOrdered);
      integrateOrdered<<<1, 32, 0, s>>>(m->dm, m->geom, m->mp, b);
    }
    CUDA_TRY(cudaGetLastError());
    m->rays_in += n;
    ++m->batches;
    return OHMB200_OK;
  }

  if (m->use_graphs && !m->capturing && m->algo == 1 && has_samples && !m->profiling && m->store.empty())
  {
    for (auto &g : m->graphs)
    {
      if (g.rays == d_rays && g.intensities == d_intensities && g.timestamps == b.timestamps && g.n == n &&
          g.ray_flags == ray_flags && g.time_base == b.time_base)
      {
        CUDA_TRY(cudaGraphLaunch(g.exec, s));
        m->launches += g.launches;
        m->rays_in += n;
        ++m->batches;
        return OHMB200_OK;
      }
    }
    // A shape is recorded the second time it is met (a stream of sweeps that all differ in size or buffer never
    // pays for a capture).  The same code below then runs into a capture instead of the stream.
    bool met_before = false;
    for (auto &g : m->seen)
    {
      met_before = met_before || (g.rays == d_rays && g.intensities == d_intensities && g.timestamps == b.timestamps &&
                                  g.n == n && g.ray_flags == ray_flags && g.time_base == b.time_base);
    }
    if (!met_before)
    {
      if (m->seen.size() >= 16)
      {
        m->seen.erase(m->seen.begin());
      }
      m->seen.push_back({ d_rays, d_intensities, b.timestamps, n, ray_flags, b.time_base, nullptr, 0 });
    }
    if (m->graphs.size() >= 8)
    {
      dropBatchGraphs(m);
    }
    const uint64_t launches_before = m->launches, rays_before = m->rays_in, batches_before = m->batches;
    cudaGraph_t graph = nullptr;
    if (met_before && cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
    {
      m->capturing = true;
      const int rc_capture = launchBatch(m, d_rays, n, d_intensities, d_timestamps, ray_flags);
      m->capturing = false;
      const cudaError_t end = cudaStreamEndCapture(s, &graph);
      cudaGraphExec_t exec = nullptr;
      if (rc_capture == OHMB200_OK && end == cudaSuccess && graph &&
          cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
      {
        cudaGraphDestroy(graph);
        m->graphs.push_back({ d_rays, d_intensities, b.timestamps, n, ray_flags, b.time_base, exec,
                              m->launches - launches_before });
        CUDA_TRY(cudaGraphLaunch(exec, s));
        return OHMB200_OK;
      }
      // could not be recorded: forget it and run the batch directly
      if (graph)
      {
        cudaGraphDestroy(graph);
      }
      cudaGetLastError();
      m->launches = launches_before;
      m->rays_in = rays_before;
      m->batches = batches_before;
      m->use_graphs = false;
    }
  }

  // Reset the per-batch counters (record_count .. segment_overflow are contiguous).
  CUDA_TRY(cudaMemsetAsync(&m->d_counters->record_count, 0, sizeof(uint32_t) * kPerBatchCounterWords, s));
  if (m->algo == 1)
  {
    CUDA_TRY(cudaMemsetAsync(b.seg_count, 0, sizeof(uint32_t) * m->dm.capacity, s));
    CUDA_TRY(cudaMemsetAsync(b.seg_cursor, 0, sizeof(uint32_t) * m->dm.capacity, s));
    CUDA_TRY(cudaMemsetAsync(b.sample_begin, 0, sizeof(uint32_t) * 2 * m->dm.capacity, s));  // begin and end
    if (has_samples)
    {
      // record slots are reserved per warp in chunks; unwritten slots must read as "no record"
      CUDA_TRY(cudaMemsetAsync(b.record_vid, 0xFF, sizeof(uint32_t) * b.record_capacity, s));
    }
    CUDA_TRY(cudaMemsetAsync(b.run_head, 0xFF, sizeof(int32_t) * n, s));
    b.tail_overflow = b.interval_count + n;
    CUDA_TRY(cudaMemsetAsync(b.interval_count, 0, sizeof(uint32_t) * (2 * n + 1), s));
    {
      KernelScope scope(m, kKPrepRays);
      prepRays<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b, m->mode);
    }
    if (b.last_exit)
    {
      rc = carryLastExit(m, n, s);
      if (rc)
      {
        return rc;
      }
    }
    // The sample path (sort -> run heads) and the segment path (plan -> scatter) only meet at walkRegions: run them
    // on two streams.  (With per-kernel profiling on they are serialised so that the event pairs stay meaningful.)
    const bool fork = has_samples && !m->profiling;
    cudaStream_t sample_stream = fork ? m->side_stream : s;
    if (fork)
    {
      CUDA_TRY(cudaEventRecord(m->fork_event, s));
      CUDA_TRY(cudaStreamWaitEvent(sample_stream, m->fork_event, 0));
    }
    if (has_samples)
    {
      {
        KernelScope scope(m, kKSort);
        size_t temp = m->cub_temp_bytes;
        cub::DeviceRadixSort::SortPairs(m->cub_temp, temp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n, 0,
                                        m->sort_bits, sample_stream);
      }
      {
        KernelScope scope(m, kKMark);
        markRuns<<<blocks, threads, 0, sample_stream>>>(m->dm, b, m->geom.vpr);
      }
    }
    {
      KernelScope scope(m, kKPrepSegments);
      b.stage_by_ray = 0;
      prepSegments<<<blocks, threads, 0, s>>>(m->dm, m->geom, b);
    }
    if (!m->store.empty() && !m->capturing)
    {
      // every region of the batch exists now: restore the ones that were evicted before anything updates them
      rc = pageInNewRegions(m);
      if (rc)
      {
        return rc;
      }
    }
    {
      KernelScope scope(m, kKPlan);
      planRegions<<<1, 1024, 0, s>>>(m->dm, b, (uint32_t)(m->sm_count * ((m->slab_walk && !m->dm.traversal) ? 1 : m->walk_ctas_per_sm)));
    }
    {
      KernelScope scope(m, kKEmit);
      emitSegments<<<blocks, threads, 0, s>>>(m->dm, m->geom, b);
    }
    if (fork)
    {
      CUDA_TRY(cudaEventRecord(m->join_event, sample_stream));
      CUDA_TRY(cudaStreamWaitEvent(s, m->join_event, 0));
    }
    if (m->mode == OHMB200_MODE_TSDF)
    {
      const unsigned grid = m->sm_count * m->walk_ctas_per_sm;
      CUDA_TRY(cudaMemsetAsync(b.record_keys, 0xFF, sizeof(unsigned long long) * b.record_capacity, s));
      {
        KernelScope scope(m, kKTsdfMark);
        markTsdfNear<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b, m->tsdf_near);
      }
      {
        KernelScope scope(m, kKWalkRegions);
        walkRegionsTsdf<<<grid, kWalkThreads, m->tile_bytes, s>>>(m->dm, m->geom, m->mp, b, m->tile, m->tsdf_near);
      }
      {
        KernelScope scope(m, kKTsdfClear);
        clearTouchedBits<<<m->sm_count * 4, 256, 0, s>>>(b, m->tsdf_near, (m->geom.vpr + 31u) / 32u);
      }
      // The ordered records are sorted by (voxel, ray) with a host-known count: the one host wait of the TSDF path.
      CUDA_TRY(cudaMemcpyAsync(m->h_counters, m->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      const uint32_t records = std::min(m->h_counters->record_count, b.record_capacity);
      if (records)
      {
        {
          KernelScope scope(m, kKSort);
          size_t temp = m->cub_temp_bytes;
          cub::DeviceRadixSort::SortKeys(m->cub_temp, temp, b.record_keys, b.record_keys_sorted, (int)records, 0, 64, s);
        }
        {
          KernelScope scope(m, kKTsdfReplay);
          replayTsdf<<<(records + 127) / 128, 128, 0, s>>>(m->dm, m->geom, m->mp, b, records);
        }
      }
      CUDA_TRY(cudaGetLastError());
      m->rays_in += n;
      ++m->batches;
      return OHMB200_OK;
    }
    rc = launchWalkAndReplay(m, b, s, n, has_samples);
    if (rc)
    {
      return rc;
    }
    CUDA_TRY(cudaGetLastError());
    m->rays_in += n;
    ++m->batches;
    return OHMB200_OK;
  }
  if (has_samples)
  {
    CUDA_TRY(cudaMemsetAsync(b.run_head, 0xFF, sizeof(int32_t) * n, s));
    b.tail_overflow = b.interval_count + n;
    CUDA_TRY(cudaMemsetAsync(b.interval_count, 0, sizeof(uint32_t) * (2 * n + 1), s));
    {
      KernelScope scope(m, kKPrep);
      prepSamples<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b, m->mode);
    }
    {
      KernelScope scope(m, kKSort);
      size_t temp = m->cub_temp_bytes;
      cub::DeviceRadixSort::SortPairs(m->cub_temp, temp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n, 0,
                                      m->sort_bits, s);
    }
    {
      KernelScope scope(m, kKMark);
      markRuns<<<blocks, threads, 0, s>>>(m->dm, b, m->geom.vpr);
    }
  }
  {
    KernelScope scope(m, kKWalk);
    walkRays<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b);
  }
  if (b.last_exit)
  {
    rc = carryLastExit(m, n, s);
    if (rc)
    {
      return rc;
    }
  }
  if (has_samples)
  {
    KernelScope scope(m, kKSamples);
    applySamples<<<blocks, threads, 0, s>>>(m->dm, m->geom, m->mp, b);
  }
  {
    KernelScope scope(m, kKResolve);
    resolveMisses<<<m->sm_count * 4, 256, 0, s>>>(m->dm, m->geom, m->mp, b);
  }
  CUDA_TRY(cudaGetLastError());
  m->rays_in += n;
  ++m->batches;
  return OHMB200_OK;
}

int pullCounters(ohmb200_map *m)
{
  CUDA_TRY(cudaMemcpyAsync(m->h_counters, m->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return OHMB200_OK;
}

int findSlots(ohmb200_map *m, const int16_t *keys_xyz, size_t count, std::vector<uint32_t> &slots)
{
  std::vector<unsigned long long> table(m->dm.capacity);
  CUDA_TRY(cudaMemcpyAsync(table.data(), m->dm.keys, sizeof(unsigned long long) * table.size(), cudaMemcpyDeviceToHost,
                           m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  slots.resize(count);
  for (size_t i = 0; i < count; ++i)
  {
    const unsigned long long key = (unsigned long long)(uint16_t)keys_xyz[3 * i] |
                                   ((unsigned long long)(uint16_t)keys_xyz[3 * i + 1] << 16) |
                                   ((unsigned long long)(uint16_t)keys_xyz[3 * i + 2] << 32);
    uint32_t h = hashRegion(key) % m->dm.capacity;
    bool found = false;
    for (uint32_t probe = 0; probe < m->dm.capacity; ++probe)
    {
      if (table[h] == key)
      {
        found = true;
        break;
      }
      if (table[h] == kEmptyKey)
      {
        break;
      }
      h = (h + 1 == m->dm.capacity) ? 0 : h + 1;  // (walks over tombstones)
    }
    if (!found)
    {
      return setError(OHMB200_E_NOT_FOUND, "region (%d,%d,%d) is not resident", keys_xyz[3 * i], keys_xyz[3 * i + 1],
                      keys_xyz[3 * i + 2]);
    }
    slots[i] = h;
  }
  return OHMB200_OK;
}

// ---- paging -------------------------------------------------------------------------------------------------
ClearTable makeClearTable(ohmb200_map *m)
{
  ClearTable table{};
  auto add = [&](void *base, size_t bytes_per_region, uint32_t fill) {
    if (base)
    {
      table.base[table.count] = (uint32_t *)base;
      table.words[table.count] = (uint32_t)(bytes_per_region / 4);
      table.fill[table.count] = fill;
      ++table.count;
    }
  };
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    add(m->layer_slab[l], m->region_layer_bytes[l], l == OHMB200_LAYER_OCCUPANCY ? 0x7f800000u : 0u);  // +inf
  }
  const size_t bit_bytes = sizeof(uint32_t) * ((m->geom.vpr + 31u) / 32u);
  add(m->dm.voxel_bits, bit_bytes, 0u);
  add(m->tsdf_near, bit_bytes, 0u);
  add(m->dm.pending, sizeof(uint32_t) * m->geom.vpr, 0u);
  return table;
}

int ensureGather(ohmb200_map *m, size_t bytes, size_t slots)
{
  if (bytes > m->gather_bytes)
  {
    cudaFree(m->d_gather);
    m->gather_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_gather, bytes));
    m->gather_bytes = bytes;
  }
  if (slots > m->gather_slots_cap)
  {
    cudaFree(m->d_gather_slots);
    m->gather_slots_cap = 0;
    CUDA_TRY(cudaMalloc(&m->d_gather_slots, sizeof(uint32_t) * slots));
    m->gather_slots_cap = slots;
  }
  return OHMB200_OK;
}

// Copies the chunks of `layer` of the listed slots to dst (host), chunk after chunk.
int downloadSlots(ohmb200_map *m, int layer, const uint32_t *slots, size_t count, char *dst, size_t dst_stride)
{
  const size_t chunk = m->region_layer_bytes[layer];
  const size_t piece = std::max<size_t>(1, std::min<size_t>(count, (64u << 20) / chunk));
  std::vector<char> staging(piece * chunk);
  for (size_t first = 0; first < count; first += piece)
  {
    const size_t n = std::min(piece, count - first);
    int rc = ensureGather(m, chunk * n, n);
    if (rc)
    {
      return rc;
    }
    for (size_t i = 0; i < n; ++i)
    {
      CUDA_TRY(cudaMemcpyAsync((char *)m->d_gather + i * chunk, (const char *)m->layer_slab[layer] + (size_t)slots[first + i] * chunk,
                               chunk, cudaMemcpyDeviceToDevice, m->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(staging.data(), m->d_gather, chunk * n, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    for (size_t i = 0; i < n; ++i)
    {
      memcpy(dst + (first + i) * dst_stride, staging.data() + i * chunk, chunk);
    }
  }
  return OHMB200_OK;
}

// Frees at least `need` slots of the region table: the least recently walked regions go to the host store
// (GpuLayerCache evicts its oldest entry the same way, GpuLayerCache.cpp:536-600).  Waits for the queued work.
int makeRoom(ohmb200_map *m, size_t need)
{
  CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  const uint32_t capacity = m->dm.capacity;
  std::vector<unsigned long long> keys(capacity);
  std::vector<uint32_t> stamps(capacity);
  CUDA_TRY(cudaMemcpy(keys.data(), m->dm.keys, sizeof(unsigned long long) * capacity, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(stamps.data(), m->dm.region_stamp, sizeof(uint32_t) * capacity, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> live;
  for (uint32_t h = 0; h < capacity; ++h)
  {
    if (isRegionKey(keys[h]))
    {
      live.push_back(h);
    }
  }
  const size_t free_slots = capacity - live.size();
  PAGING_TRACE("makeRoom: need %zu, live %zu of %u", need, live.size(), capacity);
  if (free_slots >= need)
  {
    m->regions_bound = live.size();
    return OHMB200_OK;
  }
  // oldest first; a little more than asked so that the next batches do not come straight back
  std::sort(live.begin(), live.end(), [&](uint32_t a, uint32_t b) { return stamps[a] != stamps[b] ? stamps[a] < stamps[b] : a < b; });
  const size_t count = std::min(live.size(), need - free_slots + capacity / 8u);
  std::vector<uint32_t> victims(live.begin(), live.begin() + count);
  PAGING_TRACE("makeRoom: evicting %zu regions, chunk %zu bytes", count, m->store_chunk_bytes);
  // (the chunks enter the store only when every layer has arrived: a failure half-way must not leave stale copies of
  // regions that are still resident)
  std::vector<std::vector<char>> evicted(count);
  std::vector<std::vector<char> *> chunks(count);
  for (size_t i = 0; i < count; ++i)
  {
    evicted[i].resize(m->store_chunk_bytes);
    chunks[i] = &evicted[i];
  }
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    if (!m->layer_slab[l])
    {
      continue;
    }
    // gather the layer of all victims, then hand each region its piece
    std::vector<char> layer_data(count * m->region_layer_bytes[l]);
    int rc = downloadSlots(m, l, victims.data(), count, layer_data.data(), m->region_layer_bytes[l]);
    if (rc)
    {
      return rc;
    }
    for (size_t i = 0; i < count; ++i)
    {
      memcpy(chunks[i]->data() + m->store_layer_offset[l], layer_data.data() + i * m->region_layer_bytes[l], m->region_layer_bytes[l]);
    }
  }
  int rc = ensureGather(m, 16, count);
  if (rc)
  {
    return rc;
  }
  for (size_t i = 0; i < count; ++i)
  {
    m->store[keys[victims[i]]] = std::move(evicted[i]);
  }
  CUDA_TRY(cudaMemcpyAsync(m->d_gather_slots, victims.data(), sizeof(uint32_t) * count, cudaMemcpyHostToDevice, m->stream));
  const ClearTable clear_table = makeClearTable(m);
  const unsigned clear_grid = (unsigned)std::min<size_t>(count, (size_t)m->sm_count * 16u);
  clearEvictedSlots<<<clear_grid, 256, 0, m->stream>>>(m->dm, clear_table, m->d_gather_slots, (uint32_t)count);
  CUDA_TRY(cudaGetLastError());
  // The new key table: victims become tombstones; a tombstone followed by an empty slot ends no probe sequence and
  // becomes empty itself (backwards, twice for the wrap-around).
  for (uint32_t v : victims)
  {
    keys[v] = kTombKey;
  }
  for (int pass = 0; pass < 2; ++pass)
  {
    for (uint32_t h = capacity; h-- > 0;)
    {
      if (keys[h] == kTombKey && keys[(h + 1 == capacity) ? 0 : h + 1] == kEmptyKey)
      {
        keys[h] = kEmptyKey;
      }
    }
  }
  CUDA_TRY(cudaMemcpyAsync(m->dm.keys, keys.data(), sizeof(unsigned long long) * capacity, cudaMemcpyHostToDevice, m->stream));
  const unsigned long long resident = live.size() - count;
  CUDA_TRY(cudaMemcpyAsync(&m->d_counters->region_count, &resident, sizeof(resident), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  m->evicted += count;
  m->regions_bound = resident;
  PAGING_TRACE("makeRoom: done, %llu resident, %zu stored", resident, m->store.size());
  return OHMB200_OK;
}

// Restores the chunks of `key` (in the host store) into `slot` and forgets the stored copy.
int pageIn(ohmb200_map *m, unsigned long long key, uint32_t slot)
{
  auto it = m->store.find(key);
  if (it == m->store.end())
  {
    return OHMB200_OK;
  }
  PAGING_TRACE("pageIn: key %llx -> slot %u", key, slot);
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    if (m->layer_slab[l])
    {
      CUDA_TRY(cudaMemcpyAsync((char *)m->layer_slab[l] + (size_t)slot * m->region_layer_bytes[l],
                               it->second.data() + m->store_layer_offset[l], m->region_layer_bytes[l], cudaMemcpyHostToDevice,
                               m->stream));
    }
  }
  if (m->dm.voxel_bits)
  {
    recomputeVoxelBits<<<1, 256, 0, m->stream>>>(m->dm, m->geom, m->mp, m->mode == OHMB200_MODE_TSDF, slot);
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));  // the stored copy is pageable memory: gone only after the copies
  m->store.erase(it);
  ++m->paged_in;
  return OHMB200_OK;
}

// After the region discovery of a batch (prepRays, prepSegments): the regions it created that have a stored copy.
int pageInNewRegions(ohmb200_map *m)
{
  uint32_t count = 0;
  CUDA_TRY(cudaMemcpyAsync(&count, &m->d_counters->new_count, sizeof(count), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  count = std::min(count, m->dm.capacity);
  PAGING_TRACE("pageInNewRegions: %u new regions, %zu stored", count, m->store.size());
  if (count == 0)
  {
    return OHMB200_OK;
  }
  std::vector<uint32_t> slots(count);
  std::vector<unsigned long long> keys(m->dm.capacity);
  CUDA_TRY(cudaMemcpyAsync(slots.data(), m->dm.new_slots, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(keys.data(), m->dm.keys, sizeof(unsigned long long) * keys.size(), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  for (uint32_t slot : slots)
  {
    if (slot < m->dm.capacity && isRegionKey(keys[slot]))
    {
      int rc = pageIn(m, keys[slot], slot);
      if (rc)
      {
        return rc;
      }
    }
  }
  return OHMB200_OK;
}

// Before a batch is queued: make sure the table has room for the regions it may create.  The resident count is only
// asked of the device (a wait) when the pessimistic bound — every batch so far used its whole reserve — says the
// table could be full.
int ensureRoom(ohmb200_map *m)
{
  if (m->algo != 1)
  {
    return OHMB200_OK;  // the per-ray fallback path updates regions while it discovers them: no paging there
  }
  const uint64_t reserve = m->region_reserve;
  if (m->snap_pending && cudaEventQuery(m->snap_event) == cudaSuccess)
  {
    m->known_regions = *m->h_region_snap;
    m->known_batch = m->snap_batch;
    m->snap_pending = false;
  }
  // every batch since the last known count may have used its whole reserve, and so may this one
  m->regions_bound = m->known_regions + reserve * (m->batches - m->known_batch + 1);
  PAGING_TRACE("ensureRoom: %llu regions after batch %llu, now batch %llu, reserve %llu, capacity %u",
               (unsigned long long)m->known_regions, (unsigned long long)m->known_batch, (unsigned long long)m->batches,
               (unsigned long long)reserve, m->dm.capacity);
  if (m->regions_bound <= m->dm.capacity)
  {
    return OHMB200_OK;
  }
  int rc = pullCounters(m);  // waits for the queued batches
  if (rc)
  {
    return rc;
  }
  m->known_regions = m->h_counters->region_count;
  m->known_batch = m->batches;
  m->snap_pending = false;
  if (m->known_regions + reserve > m->dm.capacity)
  {
    rc = makeRoom(m, reserve);
    m->known_regions = m->regions_bound;  // makeRoom leaves the resident count there
  }
  return rc;
}

// Behind a batch: a copy of the region counter for ensureRoom (no wait, read when it has landed).
void snapshotRegionCount(ohmb200_map *m)
{
  if (m->algo != 1 || m->snap_pending || !m->h_region_snap)
  {
    return;
  }
  if (cudaMemcpyAsync(m->h_region_snap, &m->d_counters->region_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                      m->stream) == cudaSuccess &&
      cudaEventRecord(m->snap_event, m->stream) == cudaSuccess)
  {
    m->snap_pending = true;
    m->snap_batch = m->batches;
  }
}

#include "ohmb200_exchange_host.inl"

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char *ohmb200_last_error(void)
{
  return g_last_error;
}

const char *ohmb200_version(void)
{
  return "ohmb200 0.1.0 sm_100a";
}

int ohmb200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  int usable = 0;
  for (int d = 0; d < n; ++d)
  {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
    {
      ++usable;
    }
  }
  return usable;
}

void ohmb200_default_params(ohmb200_params *p, double resolution)
{
  // ohm/OccupancyMap.cpp:195-213, ohm/private/NdtMapDetail.h:24-40, ohm/NdtMap.cpp:194-213, ohm/VoxelTsdf.h:27-37
  memset(p, 0, sizeof(*p));
  p->resolution = resolution;
  p->region_dim[0] = p->region_dim[1] = p->region_dim[2] = 32;
  p->hit_value = logf(0.9f / (1.0f - 0.9f));
  p->miss_value = logf(0.45f / (1.0f - 0.45f));
  p->min_value = -2.0f;
  p->max_value = 3.511f;
  p->threshold_value = logf(0.5f / (1.0f - 0.5f));
  p->layers = 1u << OHMB200_LAYER_OCCUPANCY;
  p->filter_kind = OHMB200_FILTER_GOOD_RAY;
  p->filter_range = 1e10;
  p->sensor_noise = 0.05f;
  p->adaptation_rate = 0.2f;
  p->reinit_threshold = logf(0.2f / (1.0f - 0.2f));
  p->reinit_count = 100;
  p->sample_threshold = 3;
  p->initial_intensity_cov = 1.0f;
  p->tsdf_max_weight = 1e4f;
  p->tsdf_trunc = 0.1f;
  p->tsdf_dropoff = 0.0f;
  p->tsdf_sparsity = 1.0f;
}

ohmb200_map *ohmb200_create(const ohmb200_params *params, int mode, size_t device_bytes, int device)
{
  if (!params || params->resolution <= 0 || mode < 0 || mode > OHMB200_MODE_TSDF)
  {
    setError(OHMB200_E_INVALID, "ohmb200_create: bad parameters");
    return nullptr;
  }
  for (int a = 0; a < 3; ++a)
  {
    if (params->region_dim[a] < 1 || params->region_dim[a] > 255)
    {
      setError(OHMB200_E_INVALID, "ohmb200_create: region_dim must be in [1,255]");
      return nullptr;
    }
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
  {
    cudaGetLastError();
    setError(OHMB200_E_NO_DEVICE, "ohmb200_create: no CUDA device %d (the CUDA path is mandatory; there is no CPU fallback)",
             device);
    return nullptr;
  }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10)
  {
    setError(OHMB200_E_NO_DEVICE, "ohmb200_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
             prop.major, prop.minor);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess)
  {
    setError(OHMB200_E_CUDA, "cudaSetDevice(%d) failed", device);
    return nullptr;
  }

  ohmb200_map *m = new ohmb200_map();
  m->device = device;
  m->mode = mode;
  m->sm_count = prop.multiProcessorCount;
  m->params = *params;
  // Layers the mapper requires (ohmgpu/GpuNdtMap.cpp:80-94, ohmapp/OhmAppGpu.cpp:192-201).
  if (mode == OHMB200_MODE_TSDF)
  {
    m->params.layers |= 1u << OHMB200_LAYER_TSDF;
  }
  else
  {
    m->params.layers |= 1u << OHMB200_LAYER_OCCUPANCY;
  }
  if (mode == OHMB200_MODE_NDT || mode == OHMB200_MODE_NDT_TM)
  {
    m->params.layers |= (1u << OHMB200_LAYER_MEAN) | (1u << OHMB200_LAYER_COVARIANCE);
  }
  if (mode == OHMB200_MODE_NDT_TM)
  {
    m->params.layers |= (1u << OHMB200_LAYER_INTENSITY) | (1u << OHMB200_LAYER_HIT_MISS);
    m->params.ndt_tm = 1;
  }
  refreshParams(m);

  // Walk algorithm: the region-binned path needs the u16 counter tile of a region to fit in shared memory.
  m->tile = makeTileLayout(m->geom);
  if (const char *env = getenv("OHMB200_TILE"))  // "row padding in words,slab bank": layout experiments
  {
    int row_pad = 1, slab_bank = 5;
    if (sscanf(env, "%d,%d", &row_pad, &slab_bank) == 2 && row_pad >= 0 && row_pad < 64 && slab_bank < 32)
    {
      m->tile = makeTileLayout(m->geom, row_pad, slab_bank);
    }
  }
  m->tile_bytes = sizeof(uint32_t) * m->tile.words;
  m->algo = 1;
  if (const char *env = getenv("OHMB200_ALGO"))
  {
    m->algo = atoi(env) ? 1 : 0;
  }
  if (const char *env = getenv("OHMB200_GRAPHS"))
  {
    m->use_graphs = atoi(env) != 0;
  }
  if (const char *env = getenv("OHMB200_HEAVY_RUN"))
  {
    m->heavy_run = (uint32_t)std::max(1, atoi(env));
  }
  const bool ndt = mode == OHMB200_MODE_NDT || mode == OHMB200_MODE_NDT_TM;
  if (ndt)
  {
    m->tile_bytes += sizeof(uint32_t) * (m->tile.words >> 4);  // + one "established Gaussian" bit per counter
    m->algo = 1;  // NDT runs on the region-binned path only
  }
  const bool tsdf = mode == OHMB200_MODE_TSDF;
  if (tsdf)
  {
    m->algo = 1;
  }
  if (m->tile_bytes > 200u * 1024u)
  {
    m->algo = 0;
  }
  if ((ndt || tsdf) && m->algo == 0)
  {
    setError(OHMB200_E_INVALID, "ohmb200_create: mode %d needs a region whose counter tile fits in shared memory", mode);
    delete m;
    return nullptr;
  }
  if (m->algo == 1)
  {
    if (cudaFuncSetAttribute(walkRegions<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tile_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(walkRegions<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tile_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(walkRegionsNdt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tile_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(walkRegionsNdt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tile_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(walkRegionsTsdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tile_bytes) != cudaSuccess)
    {
      cudaGetLastError();
      m->algo = 0;
    }
    if (const char *env = getenv("OHMB200_CARVEOUT"))  // percent of the SM's shared memory + L1 to prefer as shared (experiments)
    {
      const int pct = atoi(env);
      cudaFuncSetAttribute(walkRegions<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(walkRegions<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(walkRegionsNdt<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(walkRegionsNdt<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(walkRegionsTsdf, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    {
      // the slab-staged walk: tile + the region's whole occupancy slab in one CTA's shared memory
      const size_t need = m->tile_bytes + sizeof(float) * m->geom.vpr;
      // OHMB200_WALK=slab selects it (A/B runs).  Measured on config 2: 278 us against the tile kernel's 183 us — with one CTA
      // per SM no other CTA fills the per-item latencies (segment pops, queue build, barriers), which costs far more than
      // the shared-memory fold saves; the two-CTA tile kernel stays the default.
      const char *env = getenv("OHMB200_WALK");
      const bool want = env && strcmp(env, "slab") == 0;
      if (want && mode == OHMB200_MODE_OCCUPANCY && m->tile.fast && (m->geom.vpr % 4u) == 0 && need + 8u * 1024u <= 227u * 1024u &&
          cudaFuncSetAttribute(walkRegionsSlab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) == cudaSuccess)
      {
        m->slab_walk = true;
        m->slab_walk_bytes = need;
      }
      cudaGetLastError();
    }
    m->walk_ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(kWalkCtasPerSm, (226u * 1024u) / (m->tile_bytes + 40u * 1024u)));  // + ~38 KB static: staged segments, queue, ladder
  }
  size_t bytes_per_region = (m->algo == 0) ? sizeof(uint32_t) * m->geom.vpr : 3 * sizeof(uint32_t);  // pending / counters
  if (tsdf)
  {
    bytes_per_region += sizeof(uint32_t) * ((m->geom.vpr + 31u) / 32u);
  }
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    if (m->params.layers & (1u << l))
    {
      m->region_layer_bytes[l] = kLayerBytes[l] * m->geom.vpr;
      bytes_per_region += m->region_layer_bytes[l];
    }
  }
  if (device_bytes == 0)
  {
    device_bytes = (size_t)8 << 30;
  }
  size_t free_bytes = 0, total_bytes = 0;
  cudaMemGetInfo(&free_bytes, &total_bytes);
  device_bytes = std::min(device_bytes, (size_t)(free_bytes * 0.9));
  size_t capacity = std::max<size_t>(device_bytes / bytes_per_region, 8);
  capacity = std::min<size_t>(capacity, (size_t)0xFFFFFFF0u / m->geom.vpr);  // voxel ids are 32-bit
  m->dm.capacity = (uint32_t)capacity;
  m->sort_bits = 32;

  bool ok = true;
  ok = ok && cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&m->fork_event, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&m->join_event, cudaEventDisableTiming) == cudaSuccess;
  m->stream = m->own_stream;
  for (int i = 0; i < 2 && ok; ++i)
  {
    ok = ok && cudaEventCreateWithFlags(&m->in_ready[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&m->in_free[i], cudaEventDisableTiming) == cudaSuccess;
  }
  const size_t voxels = capacity * m->geom.vpr;
  ok = ok && cudaMalloc(&m->dm.keys, sizeof(unsigned long long) * capacity) == cudaSuccess;
  ok = ok && cudaMalloc(&m->dm.region_stamp, sizeof(uint32_t) * capacity) == cudaSuccess;
  ok = ok && cudaMalloc(&m->dm.new_slots, sizeof(uint32_t) * capacity) == cudaSuccess;
  if (m->algo == 0)
  {
    ok = ok && cudaMalloc(&m->dm.pending, sizeof(uint32_t) * voxels) == cudaSuccess;
  }
  else
  {
    ok = ok && cudaMalloc(&m->batch.seg_count, sizeof(uint32_t) * capacity) == cudaSuccess;
    ok = ok && cudaMalloc(&m->batch.seg_offset, sizeof(uint32_t) * capacity) == cudaSuccess;
    ok = ok && cudaMalloc(&m->batch.seg_cursor, sizeof(uint32_t) * capacity) == cudaSuccess;
    ok = ok && cudaMalloc(&m->batch.sample_begin, sizeof(uint32_t) * 2 * capacity) == cudaSuccess;
    m->batch.sample_end = ok ? m->batch.sample_begin + capacity : nullptr;
  }
  ok = ok && cudaMalloc(&m->batch.touched_list, sizeof(uint32_t) * capacity) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_counters, sizeof(Counters)) == cudaSuccess;
  ok = ok && cudaMallocHost(&m->h_counters, sizeof(Counters)) == cudaSuccess;
  for (int l = 0; l < OHMB200_LAYER_COUNT && ok; ++l)
  {
    if (m->params.layers & (1u << l))
    {
      ok = ok && cudaMalloc(&m->layer_slab[l], kLayerBytes[l] * voxels) == cudaSuccess;
    }
  }
  if (tsdf || ndt)
  {
    m->voxel_bit_bytes = sizeof(uint32_t) * ((m->geom.vpr + 31u) / 32u) * capacity;
    ok = ok && cudaMalloc(&m->dm.voxel_bits, m->voxel_bit_bytes) == cudaSuccess;
    if (tsdf)
    {
      ok = ok && cudaMalloc(&m->tsdf_near, m->voxel_bit_bytes) == cudaSuccess;
    }
  }
  if (!ok)
  {
    setError(OHMB200_E_CUDA, "ohmb200_create: device allocation failed (%zu region slots x %zu bytes): %s", capacity,
             bytes_per_region, cudaGetErrorString(cudaGetLastError()));
    ohmb200_destroy(m);
    return nullptr;
  }
  m->dm.occupancy = (float *)m->layer_slab[OHMB200_LAYER_OCCUPANCY];
  m->dm.mean = (uint2 *)m->layer_slab[OHMB200_LAYER_MEAN];
  m->dm.traversal = (float *)m->layer_slab[OHMB200_LAYER_TRAVERSAL];
  m->dm.touch_time = (uint32_t *)m->layer_slab[OHMB200_LAYER_TOUCH_TIME];
  m->dm.incident = (uint32_t *)m->layer_slab[OHMB200_LAYER_INCIDENT];
  m->dm.covariance = (float *)m->layer_slab[OHMB200_LAYER_COVARIANCE];
  m->dm.intensity = (float2 *)m->layer_slab[OHMB200_LAYER_INTENSITY];
  m->dm.hit_miss = (uint2 *)m->layer_slab[OHMB200_LAYER_HIT_MISS];
  m->dm.tsdf = (float2 *)m->layer_slab[OHMB200_LAYER_TSDF];
  m->dm.secondary = (uint2 *)m->layer_slab[OHMB200_LAYER_SECONDARY];
  m->dm.region_count = &m->d_counters->region_count;
  m->dm.table_full = &m->d_counters->table_full;
  m->dm.new_count = &m->d_counters->new_count;
  for (int l = 0; l < OHMB200_LAYER_COUNT; ++l)
  {
    m->store_layer_offset[l] = m->store_chunk_bytes;
    m->store_chunk_bytes += m->layer_slab[l] ? m->region_layer_bytes[l] : 0;
  }
  cudaMallocHost(&m->h_region_snap, sizeof(unsigned long long));
  cudaEventCreateWithFlags(&m->snap_event, cudaEventDisableTiming);
  // free slots a batch may need: half of a small table, 4096 of a large one (a 64x2048 sweep at 0.1 m creates ~1000)
  m->region_reserve = std::min<uint32_t>(4096u, std::max<uint32_t>(1u, m->dm.capacity / 2u));
  m->dm.part_rank = 0;
  m->dm.part_world = 1;
  if (initialiseSlabs(m) != OHMB200_OK || cudaStreamSynchronize(m->stream) != cudaSuccess)
  {
    ohmb200_destroy(m);
    return nullptr;
  }
  return m;
}

void ohmb200_destroy(ohmb200_map *m)
{
  if (!m)
  {
    return;
  }
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  exchangeClose(m);
  dropBatchGraphs(m);
  Batch &b = m->batch;
  void *to_free[] = { m->dm.keys,       m->dm.region_stamp, m->dm.new_slots,    m->dm.pending,      b.touched_list,   m->d_counters,
                      b.keys_in,        b.keys_out,         b.vals_in,          b.vals_out,       b.run_list,
                      b.run_head,       b.interval_count,   b.interval_offset,  b.sorted_rays,   b.record_ray,     b.record_next,
                      b.last_exit,      m->carry_temp,      m->cub_temp,        m->d_rays[0],       m->d_rays[1],     m->d_intensities[0],
                      m->d_intensities[1], m->d_timestamps[0], m->d_timestamps[1], m->d_gather,    m->d_gather_slots,
                      b.recs,           b.ray_length,       b.record_vid,       b.seg_count,      b.seg_offset,
                      b.seg_cursor,     b.sample_begin,     b.segments,         b.items,            b.record_keys,    b.record_keys_sorted, b.gauss_keys,
                      b.stage,          b.stage_count,      m->tsdf_near,       m->dm.voxel_bits,   b.cross_pos };
  for (void *p : to_free)
  {
    if (p)
    {
      cudaFree(p);
    }
  }
  for (void *p : m->layer_slab)
  {
    if (p)
    {
      cudaFree(p);
    }
  }
  if (m->h_counters)
  {
    cudaFreeHost(m->h_counters);
    cudaFreeHost(m->h_region_snap);
    if (m->snap_event)
    {
      cudaEventDestroy(m->snap_event);
    }
  }
  for (cudaEvent_t e : m->event_pool)
  {
    cudaEventDestroy(e);
  }
  for (int i = 0; i < 2; ++i)
  {
    if (m->in_ready[i])
    {
      cudaEventDestroy(m->in_ready[i]);
    }
    if (m->in_free[i])
    {
      cudaEventDestroy(m->in_free[i]);
    }
  }
  if (m->own_stream)
  {
    cudaStreamDestroy(m->own_stream);
  }
  for (int i = 0; i < 2; ++i)
  {
    cudaFree(m->d_stage[i]);
    cudaFree(m->d_stage_keys[i]);
    cudaFree(m->d_stage_slots[i]);
    cudaFreeHost(m->h_stage_keys[i]);
    if (m->stage_gathered[i])
    {
      cudaEventDestroy(m->stage_gathered[i]);
      cudaEventDestroy(m->stage_done[i]);
    }
  }
  cudaFree(m->d_lookup_missing);
  cudaFree(m->d_query);
  cudaFree(m->d_query_keys);
  cudaFree(m->d_secondary_ranges);
  cudaFree(m->d_secondary_rays);
  if (m->download_stream)
  {
    cudaStreamDestroy(m->download_stream);
  }
  if (m->copy_stream)
  {
    cudaStreamDestroy(m->copy_stream);
  }
  if (m->side_stream)
  {
    cudaStreamDestroy(m->side_stream);
  }
  if (m->fork_event)
  {
    cudaEventDestroy(m->fork_event);
  }
  if (m->join_event)
  {
    cudaEventDestroy(m->join_event);
  }
  cudaGetLastError();
  delete m;
}

int ohmb200_set_params(ohmb200_map *m, const ohmb200_params *p)
{
  if (!m || !p)
  {
    return setError(OHMB200_E_INVALID, "null argument");
  }
  if (p->resolution != m->params.resolution || memcmp(p->region_dim, m->params.region_dim, sizeof(p->region_dim)) != 0)
  {
    return setError(OHMB200_E_INVALID, "resolution and region dimensions are fixed at creation");
  }
  const uint32_t layers = m->params.layers;
  const int ndt_tm = m->params.ndt_tm;
  const bool bits_stale = m->dm.voxel_bits && (p->sample_threshold != m->params.sample_threshold ||
                                               p->tsdf_trunc != m->params.tsdf_trunc);
  m->params = *p;
  m->params.layers = layers;
  m->params.ndt_tm = ndt_tm;
  refreshParams(m);
  if (bits_stale)
  {
    // the persistent per-voxel bits are a function of the stored layers AND these parameters
    cudaSetDevice(m->device);
    recomputeVoxelBits<<<m->dm.capacity, 256, 0, m->stream>>>(m->dm, m->geom, m->mp, m->mode == OHMB200_MODE_TSDF, 0u);
    CUDA_TRY(cudaGetLastError());
  }
  return OHMB200_OK;
}

int ohmb200_get_params(const ohmb200_map *m, ohmb200_params *p)
{
  if (!m || !p)
  {
    return setError(OHMB200_E_INVALID, "null argument");
  }
  *p = m->params;
  return OHMB200_OK;
}

size_t ohmb200_integrate_device(ohmb200_map *m, const double *d_rays, size_t element_count, const float *d_intensities,
                                const double *d_timestamps, unsigned ray_flags)
{
  if (!m || !d_rays || element_count < 2)
  {
    setError(OHMB200_E_INVALID, "ohmb200_integrate_device: bad arguments");
    return 0;
  }
  cudaSetDevice(m->device);
  if (d_timestamps && m->first_ray_time < 0)
  {
    // OccupancyMap::updateFirstRayTime(*timestamps) (RayMapperOccupancy.cpp:99-103)
    double t0 = 0;
    if (cudaMemcpyAsync(&t0, d_timestamps, sizeof(double), cudaMemcpyDeviceToHost, m->stream) != cudaSuccess ||
        cudaStreamSynchronize(m->stream) != cudaSuccess)
    {
      setError(OHMB200_E_CUDA, "failed to read the first timestamp");
      return 0;
    }
    m->first_ray_time = t0;
  }
  if (ensureRoom(m) != OHMB200_OK || launchBatch(m, d_rays, element_count / 2, d_intensities, d_timestamps, ray_flags) != OHMB200_OK)
  {
    return 0;
  }
  snapshotRegionCount(m);
  return element_count;
}

size_t ohmb200_integrate(ohmb200_map *m, const double *rays, size_t element_count, const float *intensities,
                         const double *timestamps, unsigned ray_flags)
{
  if (!m || !rays || element_count < 2)
  {
    setError(OHMB200_E_INVALID, "ohmb200_integrate: bad arguments");
    return 0;
  }
  cudaSetDevice(m->device);
  const size_t n = element_count / 2;
  const int buf = m->next_input;
  m->next_input ^= 1;
  // Drop inputs the map has no layer for (GpuMap.cpp:566-575).
  if (!(m->params.layers & (1u << OHMB200_LAYER_INTENSITY)))
  {
    intensities = nullptr;
  }
  if (!(m->params.layers & (1u << OHMB200_LAYER_TOUCH_TIME)))
  {
    timestamps = nullptr;
  }
  if (timestamps && m->first_ray_time < 0)
  {
    m->first_ray_time = timestamps[0];
  }
  if (n > m->in_capacity[buf])
  {
    cudaEventSynchronize(m->in_free[buf]);
    cudaFree(m->d_rays[buf]);
    cudaFree(m->d_intensities[buf]);
    cudaFree(m->d_timestamps[buf]);
    const size_t cap = std::max<size_t>(((n + n / 8 + 16383) / 16384) * 16384, 4096);
    if (cudaMalloc(&m->d_rays[buf], sizeof(double) * 6 * cap) != cudaSuccess ||
        cudaMalloc(&m->d_intensities[buf], sizeof(float) * cap) != cudaSuccess ||
        cudaMalloc(&m->d_timestamps[buf], sizeof(double) * cap) != cudaSuccess)
    {
      setError(OHMB200_E_CUDA, "input staging allocation failed");
      m->in_capacity[buf] = 0;
      return 0;
    }
    m->in_capacity[buf] = cap;
  }
  // Upload on the copy stream once the kernels that last read this buffer are done; the other buffer's batch may
  // still be running on the compute stream (upload/compute overlap).
  cudaStreamWaitEvent(m->copy_stream, m->in_free[buf], 0);
  bool ok = cudaMemcpyAsync(m->d_rays[buf], rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, m->copy_stream) ==
            cudaSuccess;
  if (intensities)
  {
    ok = ok && cudaMemcpyAsync(m->d_intensities[buf], intensities, sizeof(float) * n, cudaMemcpyHostToDevice,
                               m->copy_stream) == cudaSuccess;
  }
  if (timestamps)
  {
    ok = ok && cudaMemcpyAsync(m->d_timestamps[buf], timestamps, sizeof(double) * n, cudaMemcpyHostToDevice,
                               m->copy_stream) == cudaSuccess;
  }
  ok = ok && cudaEventRecord(m->in_ready[buf], m->copy_stream) == cudaSuccess;
  ok = ok && cudaStreamWaitEvent(m->stream, m->in_ready[buf], 0) == cudaSuccess;
  if (!ok)
  {
    setError(OHMB200_E_CUDA, "ray upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  int rc = ensureRoom(m);
  rc = rc ? rc : launchBatch(m, m->d_rays[buf], n, intensities ? m->d_intensities[buf] : nullptr,
                             timestamps ? m->d_timestamps[buf] : nullptr, ray_flags);
  cudaEventRecord(m->in_free[buf], m->stream);
  snapshotRegionCount(m);
  // The caller's arrays are only guaranteed to be read during the call (GpuMap.cpp:843-862).
  cudaEventSynchronize(m->in_ready[buf]);
  return rc == OHMB200_OK ? element_count : 0;
}

int ohmb200_sync(ohmb200_map *m)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  cudaSetDevice(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
  int rc = pullCounters(m);
  if (rc)
  {
    return rc;
  }
  if (m->h_counters->table_full)
  {
    return setError(OHMB200_E_CACHE_FULL, "region table full (%u slots): raise device_bytes", m->dm.capacity);
  }
  if (m->ex.open)
  {
    int aborted = 0;
    CUDA_TRY(cudaMemcpyAsync(&aborted, m->ex.abort, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (aborted)
    {
      CUDA_TRY(cudaMemsetAsync(m->ex.abort, 0, sizeof(int), m->stream));
      return setError(OHMB200_E_CUDA, "exchange: a peer's records did not arrive within the wait bound; that step was dropped on this rank");
    }
  }
  if (const int seen = m->h_counters->overflow_seen)
  {
    // reported once; the lists are grown for the batches that follow
    CUDA_TRY(cudaMemsetAsync(&m->d_counters->overflow_seen, 0, sizeof(int), m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (seen & 2)
    {
      m->seg_factor = std::min<uint32_t>(m->seg_factor * 2u, 4096u);
    }
    if (seen & 1)
    {
      m->record_factor = std::min<uint32_t>(m->record_factor * 2u, 1024u);
    }
    m->scratch_rays = 0;  // the next batch reallocates its scratch with the new factors
    if (seen & 1)
    {
      return setError(OHMB200_E_OVERFLOW, "a per-batch list of ordered records overflowed: the results of that batch are "
                                          "incomplete (the list has been enlarged; use smaller batches)");
    }
    return setError(OHMB200_E_OVERFLOW, "a batch cut into more region segments than its list holds was dropped whole: "
                                        "integrate it again (the list has been enlarged to %u segments per ray)", m->seg_factor);
  }
  return OHMB200_OK;
}

size_t ohmb200_region_count(ohmb200_map *m)
{
  if (!m || pullCounters(m) != OHMB200_OK)
  {
    return 0;
  }
  return (size_t)m->h_counters->region_count + m->store.size();
}

size_t ohmb200_enumerate_regions(ohmb200_map *m, int16_t *keys_xyz, size_t capacity)
{
  if (!m)
  {
    return 0;
  }
  cudaSetDevice(m->device);
  std::vector<unsigned long long> table(m->dm.capacity);
  if (cudaMemcpyAsync(table.data(), m->dm.keys, sizeof(unsigned long long) * table.size(), cudaMemcpyDeviceToHost,
                      m->stream) != cudaSuccess ||
      cudaStreamSynchronize(m->stream) != cudaSuccess)
  {
    setError(OHMB200_E_CUDA, "region table download failed");
    return 0;
  }
  struct R
  {
    int16_t x, y, z;
  };
  std::vector<R> regions;
  for (unsigned long long k : table)
  {
    if (isRegionKey(k))
    {
      regions.push_back({ (int16_t)(k & 0xffff), (int16_t)((k >> 16) & 0xffff), (int16_t)((k >> 32) & 0xffff) });
    }
  }
  for (const auto &kv : m->store)
  {
    const unsigned long long k = kv.first;
    regions.push_back({ (int16_t)(k & 0xffff), (int16_t)((k >> 16) & 0xffff), (int16_t)((k >> 32) & 0xffff) });
  }
  std::sort(regions.begin(), regions.end(), [](const R &a, const R &b) {
    if (a.z != b.z)
    {
      return a.z < b.z;
    }
    if (a.y != b.y)
    {
      return a.y < b.y;
    }
    return a.x < b.x;
  });
  for (size_t i = 0; i < regions.size() && i < capacity && keys_xyz; ++i)
  {
    keys_xyz[3 * i] = regions[i].x;
    keys_xyz[3 * i + 1] = regions[i].y;
    keys_xyz[3 * i + 2] = regions[i].z;
  }
  return regions.size();
}

size_t ohmb200_region_layer_bytes(const ohmb200_map *m, int layer)
{
  if (!m || layer < 0 || layer >= OHMB200_LAYER_COUNT)
  {
    return 0;
  }
  return m->region_layer_bytes[layer];
}

int ohmb200_read_regions(ohmb200_map *m, int layer, const int16_t *keys_xyz, size_t count, void *dst, size_t bytes)
{
  if (!m || layer < 0 || layer >= OHMB200_LAYER_COUNT || !m->layer_slab[layer] || (!keys_xyz && count) || !dst)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_read_regions: bad arguments or layer %d absent", layer);
  }
  const size_t chunk = m->region_layer_bytes[layer];
  if (bytes < chunk * count)
  {
    return setError(OHMB200_E_INVALID, "destination too small: %zu < %zu", bytes, chunk * count);
  }
  if (count == 0)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  if (!m->store.empty())
  {
    // some regions live in the host store (paging): those are copied from there; the resident ones are looked up with
    // ONE download of the key table and gathered together
    std::vector<size_t> resident;
    std::vector<int16_t> resident_keys;
    for (size_t i = 0; i < count; ++i)
    {
      const int16_t *k = keys_xyz + 3 * i;
      const auto it = m->store.find(packRegion(k[0], k[1], k[2]));
      if (it != m->store.end())
      {
        memcpy((char *)dst + i * chunk, it->second.data() + m->store_layer_offset[layer], chunk);
        continue;
      }
      resident.push_back(i);
      resident_keys.insert(resident_keys.end(), k, k + 3);
    }
    if (resident.empty())
    {
      return OHMB200_OK;
    }
    std::vector<uint32_t> resident_slots;
    int rc1 = findSlots(m, resident_keys.data(), resident.size(), resident_slots);
    if (rc1)
    {
      return rc1;
    }
    std::vector<char> gathered(resident.size() * chunk);
    rc1 = downloadSlots(m, layer, resident_slots.data(), resident.size(), gathered.data(), chunk);
    if (rc1)
    {
      return rc1;
    }
    for (size_t j = 0; j < resident.size(); ++j)
    {
      memcpy((char *)dst + resident[j] * chunk, gathered.data() + j * chunk, chunk);
    }
    return OHMB200_OK;
  }
  std::vector<uint32_t> slots;
  int rc = findSlots(m, keys_xyz, count, slots);
  if (rc)
  {
    return rc;
  }
  if (chunk * count > m->gather_bytes)
  {
    cudaFree(m->d_gather);
    m->gather_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_gather, chunk * count));
    m->gather_bytes = chunk * count;
  }
  if (count > m->gather_slots_cap)
  {
    cudaFree(m->d_gather_slots);
    m->gather_slots_cap = 0;
    CUDA_TRY(cudaMalloc(&m->d_gather_slots, sizeof(uint32_t) * count));
    m->gather_slots_cap = count;
  }
  CUDA_TRY(cudaMemcpyAsync(m->d_gather_slots, slots.data(), sizeof(uint32_t) * count, cudaMemcpyHostToDevice, m->stream));
  if (chunk % 16 == 0)
  {
    KernelScope scope(m, kKGather);
    gatherRegions<<<(unsigned)count, 256, 0, m->stream>>>((const uint4 *)m->layer_slab[layer], m->d_gather_slots,
                                                          (uint4 *)m->d_gather, chunk / 16);
  }
  else
  {
    for (size_t i = 0; i < count; ++i)
    {
      CUDA_TRY(cudaMemcpyAsync((char *)m->d_gather + i * chunk, (const char *)m->layer_slab[layer] + (size_t)slots[i] * chunk,
                               chunk, cudaMemcpyDeviceToDevice, m->stream));
    }
  }
  CUDA_TRY(cudaMemcpyAsync(dst, m->d_gather, chunk * count, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return OHMB200_OK;
}

int ohmb200_read_region(ohmb200_map *m, const int16_t key_xyz[3], int layer, void *dst, size_t bytes)
{
  return ohmb200_read_regions(m, layer, key_xyz, 1, dst, bytes);
}

size_t ohmb200_integrate_secondary_device(ohmb200_map *m, const double *d_rays, size_t element_count)
{
  if (!m || !d_rays || element_count < 2)
  {
    setError(OHMB200_E_INVALID, "ohmb200_integrate_secondary_device: bad arguments");
    return 0;
  }
  if (!m->dm.secondary)
  {
    setError(OHMB200_E_INVALID, "the map has no secondary-sample layer (OHMB200_LAYER_SECONDARY)");
    return 0;
  }
  cudaSetDevice(m->device);
  const size_t n = element_count / 2;
  if (ensureScratch(m, n) != OHMB200_OK || ensureRoom(m) != OHMB200_OK)
  {
    return 0;
  }
  if (sizeof(double) * n > m->secondary_bytes)
  {
    cudaFree(m->d_secondary_ranges);
    m->secondary_bytes = 0;
    if (cudaMalloc(&m->d_secondary_ranges, sizeof(double) * m->scratch_rays) != cudaSuccess)
    {
      setError(OHMB200_E_CUDA, "scratch allocation failed");
      return 0;
    }
    m->secondary_bytes = sizeof(double) * m->scratch_rays;
  }
  Batch &b = m->batch;
  b.rays = d_rays;
  b.intensities = nullptr;
  b.timestamps = nullptr;
  b.n = (uint32_t)n;
  b.counters = m->d_counters;
  cudaStream_t s = m->stream;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  bool ok = cudaMemsetAsync(&m->d_counters->record_count, 0, sizeof(uint32_t) * kPerBatchCounterWords, s) == cudaSuccess;
  {
    KernelScope scope(m, kKPrepRays);
    prepSecondary<<<blocks, 128, 0, s>>>(m->dm, m->geom, b, m->d_secondary_ranges);
  }
  if (!m->store.empty() && pageInNewRegions(m) != OHMB200_OK)  // the sample regions exist now; none is updated yet
  {
    return 0;
  }
  {
    KernelScope scope(m, kKSort);
    size_t temp = m->cub_temp_bytes;
    ok = ok && cub::DeviceRadixSort::SortPairs(m->cub_temp, temp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n, 0,
                                               m->sort_bits, s) == cudaSuccess;
  }
  {
    // run heads only (no per-region sample ranges: those belong to the main mapper's batch)
    Batch runs = b;
    runs.sample_begin = nullptr;
    KernelScope scope(m, kKMark);
    markRuns<<<blocks, 128, 0, s>>>(m->dm, runs, m->geom.vpr);
  }
  {
    KernelScope scope(m, kKSamples);
    applySecondary<<<blocks, 128, 0, s>>>(m->dm, b, m->d_secondary_ranges);
  }
  if (!ok || cudaGetLastError() != cudaSuccess)
  {
    setError(OHMB200_E_CUDA, "secondary-sample batch failed to launch");
    return 0;
  }
  m->rays_in += n;
  ++m->batches;
  return element_count;
}

size_t ohmb200_integrate_secondary(ohmb200_map *m, const double *rays, size_t element_count)
{
  if (!m || !rays || element_count < 2)
  {
    setError(OHMB200_E_INVALID, "ohmb200_integrate_secondary: bad arguments");
    return 0;
  }
  cudaSetDevice(m->device);
  const size_t n = element_count / 2;
  if (cudaStreamSynchronize(m->stream) != cudaSuccess)  // the staging buffer may still feed the previous call
  {
    setError(OHMB200_E_CUDA, "device error before the secondary-sample batch");
    return 0;
  }
  if (sizeof(double) * 6 * n > m->secondary_rays_bytes)
  {
    cudaFree(m->d_secondary_rays);
    m->secondary_rays_bytes = 0;
    if (cudaMalloc(&m->d_secondary_rays, sizeof(double) * 6 * n) != cudaSuccess)
    {
      setError(OHMB200_E_CUDA, "input staging allocation failed");
      return 0;
    }
    m->secondary_rays_bytes = sizeof(double) * 6 * n;
  }
  if (cudaMemcpyAsync(m->d_secondary_rays, rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, m->stream) != cudaSuccess)
  {
    setError(OHMB200_E_CUDA, "ray upload failed");
    return 0;
  }
  return ohmb200_integrate_secondary_device(m, (const double *)m->d_secondary_rays, 2 * n) ? element_count : 0;
}

int ohmb200_rays_query_device(ohmb200_map *m, const double *d_rays, size_t element_count, double volume_coefficient,
                              double *d_ranges, double *d_unobserved_volumes, int *d_terminal_states,
                              int32_t *d_terminal_keys)
{
  if (!m || !d_rays || !d_ranges || !d_unobserved_volumes || !d_terminal_states || !d_terminal_keys)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_rays_query_device: null argument");
  }
  if (!m->dm.occupancy)
  {
    return setError(OHMB200_E_INVALID, "the rays query needs the occupancy layer");
  }
  const size_t n = element_count / 2;
  if (n == 0)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  // in stream order: the query sees every batch queued before it
  raysQuery<<<(unsigned)((n + 127) / 128), 128, 0, m->stream>>>(m->dm, m->geom, m->mp, d_rays, (uint32_t)n,
                                                                volume_coefficient, d_ranges, d_unobserved_volumes,
                                                                d_terminal_states, d_terminal_keys);
  CUDA_TRY(cudaGetLastError());
  return OHMB200_OK;
}

int ohmb200_rays_query(ohmb200_map *m, const double *rays, size_t element_count, double volume_coefficient,
                       double *ranges, double *unobserved_volumes, int *terminal_states, int32_t *terminal_keys)
{
  if (!m || !rays || !ranges || !unobserved_volumes || !terminal_states || !terminal_keys)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_rays_query: null argument");
  }
  const size_t n = element_count / 2;
  if (n == 0)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  // one device block: rays | ranges | volumes | keys | states
  const size_t bytes = n * (6 * sizeof(double) + 2 * sizeof(double) + 6 * sizeof(int32_t) + sizeof(int));
  if (bytes > m->query_bytes)
  {
    cudaFree(m->d_query);
    m->query_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_query, bytes));
    m->query_bytes = bytes;
  }
  double *d_rays = (double *)m->d_query;
  double *d_ranges = d_rays + 6 * n;
  double *d_volumes = d_ranges + n;
  int32_t *d_keys = (int32_t *)(d_volumes + n);
  int *d_states = (int *)(d_keys + 6 * n);
  cudaStream_t s = m->stream;
  CUDA_TRY(cudaMemcpyAsync(d_rays, rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s));
  const int rc = ohmb200_rays_query_device(m, d_rays, 2 * n, volume_coefficient, d_ranges, d_volumes, d_states, d_keys);
  if (rc)
  {
    return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(ranges, d_ranges, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(unobserved_volumes, d_volumes, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(terminal_keys, d_keys, sizeof(int32_t) * 6 * n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(terminal_states, d_states, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return OHMB200_OK;
}

int ohmb200_line_keys_query(ohmb200_map *m, const double *rays, size_t element_count, uint64_t *result_indices,
                            uint64_t *result_counts, int32_t *keys, size_t key_capacity, size_t *total_keys)
{
  if (!m || !rays || !result_indices || !result_counts || !total_keys || (!keys && key_capacity))
  {
    return setError(OHMB200_E_INVALID, "ohmb200_line_keys_query: null argument");
  }
  const size_t n = element_count / 2;
  *total_keys = 0;
  if (n == 0)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  cudaStream_t s = m->stream;
  // staging: rays | counts[n + 1] | offsets[n + 1] | scan temp; the keys live in a second block
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n + 1), s);
  const size_t head_bytes = sizeof(double) * 6 * n + sizeof(uint32_t) * 2 * (n + 1) + scan_bytes + 256;
  if (head_bytes > m->query_bytes)
  {
    cudaFree(m->d_query);
    m->query_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_query, head_bytes));
    m->query_bytes = head_bytes;
  }
  double *d_rays = (double *)m->d_query;
  uint32_t *d_counts = (uint32_t *)(d_rays + 6 * n);
  uint32_t *d_offsets = d_counts + (n + 1);
  void *d_scan = (void *)(((uintptr_t)(d_offsets + (n + 1)) + 255) & ~(uintptr_t)255);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  CUDA_TRY(cudaMemcpyAsync(d_rays, rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(d_counts + n, 0, sizeof(uint32_t), s));
  lineKeys<false><<<blocks, 128, 0, s>>>(m->geom, d_rays, (uint32_t)n, d_counts, nullptr, nullptr);
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_counts, d_offsets, (int)(n + 1), s));
  std::vector<uint32_t> h_counts(n + 1), h_offsets(n + 1);
  CUDA_TRY(cudaMemcpyAsync(h_counts.data(), d_counts, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_offsets.data(), d_offsets, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i)
  {
    result_indices[i] = h_offsets[i];
    result_counts[i] = h_counts[i];
  }
  const size_t total = h_offsets[n];
  *total_keys = total;
  if (key_capacity < total || total == 0)
  {
    return OHMB200_OK;  // sizes only: the caller allocates 6 * total int32 and calls again
  }
  if (sizeof(int32_t) * 6 * total > m->query_keys_bytes)
  {
    cudaFree(m->d_query_keys);
    m->query_keys_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_query_keys, sizeof(int32_t) * 6 * total));
    m->query_keys_bytes = sizeof(int32_t) * 6 * total;
  }
  lineKeys<true><<<blocks, 128, 0, s>>>(m->geom, d_rays, (uint32_t)n, nullptr, d_offsets, (int32_t *)m->d_query_keys);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(keys, m->d_query_keys, sizeof(int32_t) * 6 * total, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return OHMB200_OK;
}

int ohmb200_read_regions_async(ohmb200_map *m, int layer, const int16_t *keys_xyz, size_t count, void *dst, size_t bytes)
{
  if (!m || layer < 0 || layer >= OHMB200_LAYER_COUNT || !m->layer_slab[layer] || (!keys_xyz && count) || !dst)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_read_regions_async: bad arguments or layer %d absent", layer);
  }
  const size_t chunk = m->region_layer_bytes[layer];
  if (bytes < chunk * count)
  {
    return setError(OHMB200_E_INVALID, "destination too small: %zu < %zu", bytes, chunk * count);
  }
  if (chunk % 16 != 0)
  {
    return setError(OHMB200_E_INVALID, "asynchronous download needs 16-byte multiple chunks");
  }
  if (!m->store.empty())
  {
    return ohmb200_read_regions(m, layer, keys_xyz, count, dst, bytes);  // part of the map is in the host store
  }
  if (count == 0)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  if (!m->download_stream)
  {
    CUDA_TRY(cudaStreamCreateWithFlags(&m->download_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i)
    {
      CUDA_TRY(cudaEventCreateWithFlags(&m->stage_gathered[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&m->stage_done[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMalloc(&m->d_lookup_missing, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(m->d_lookup_missing, 0, sizeof(int), m->stream));
  }
  const int buf = m->stage_next;
  m->stage_next ^= 1;
  if (m->stage_pending[buf])
  {
    CUDA_TRY(cudaEventSynchronize(m->stage_done[buf]));  // the staging set is reused every second call
    m->stage_pending[buf] = false;
  }
  if (chunk * count > m->stage_bytes[buf])
  {
    cudaFree(m->d_stage[buf]);
    m->stage_bytes[buf] = 0;
    CUDA_TRY(cudaMalloc(&m->d_stage[buf], chunk * count));
    m->stage_bytes[buf] = chunk * count;
  }
  if (count > m->stage_keys_cap[buf])
  {
    cudaFree(m->d_stage_keys[buf]);
    cudaFree(m->d_stage_slots[buf]);
    cudaFreeHost(m->h_stage_keys[buf]);
    m->stage_keys_cap[buf] = 0;
    CUDA_TRY(cudaMalloc(&m->d_stage_keys[buf], sizeof(unsigned long long) * count));
    CUDA_TRY(cudaMalloc(&m->d_stage_slots[buf], sizeof(uint32_t) * count));
    CUDA_TRY(cudaMallocHost(&m->h_stage_keys[buf], sizeof(unsigned long long) * count));
    m->stage_keys_cap[buf] = count;
  }
  for (size_t i = 0; i < count; ++i)
  {
    m->h_stage_keys[buf][i] = (unsigned long long)(uint16_t)keys_xyz[3 * i] |
                              ((unsigned long long)(uint16_t)keys_xyz[3 * i + 1] << 16) |
                              ((unsigned long long)(uint16_t)keys_xyz[3 * i + 2] << 32);
  }
  cudaStream_t s = m->stream;
  CUDA_TRY(cudaMemcpyAsync(m->d_stage_keys[buf], m->h_stage_keys[buf], sizeof(unsigned long long) * count,
                           cudaMemcpyHostToDevice, s));
  lookupSlots<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(m->dm, m->d_stage_keys[buf], (uint32_t)count,
                                                               m->d_stage_slots[buf], m->d_lookup_missing);
  {
    KernelScope scope(m, kKGather);
    gatherRegionsChecked<<<(unsigned)count, 256, 0, s>>>((const uint4 *)m->layer_slab[layer], m->d_stage_slots[buf],
                                                         (uint4 *)m->d_stage[buf], chunk / 16);
  }
  CUDA_TRY(cudaGetLastError());
  // The snapshot is taken in stream order (later batches do not disturb it); the D2H runs beside them.
  CUDA_TRY(cudaEventRecord(m->stage_gathered[buf], s));
  CUDA_TRY(cudaStreamWaitEvent(m->download_stream, m->stage_gathered[buf], 0));
  CUDA_TRY(cudaMemcpyAsync(dst, m->d_stage[buf], chunk * count, cudaMemcpyDeviceToHost, m->download_stream));
  CUDA_TRY(cudaEventRecord(m->stage_done[buf], m->download_stream));
  m->stage_pending[buf] = true;
  return OHMB200_OK;
}

int ohmb200_download_wait(ohmb200_map *m)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  if (!m->download_stream)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->download_stream));
  m->stage_pending[0] = m->stage_pending[1] = false;
  int missing = 0;
  CUDA_TRY(cudaMemcpyAsync(&missing, m->d_lookup_missing, sizeof(int), cudaMemcpyDeviceToHost, m->download_stream));
  CUDA_TRY(cudaMemsetAsync(m->d_lookup_missing, 0, sizeof(int), m->download_stream));
  CUDA_TRY(cudaStreamSynchronize(m->download_stream));
  if (missing)
  {
    return setError(OHMB200_E_NOT_FOUND, "an asynchronous download named a region that is not resident");
  }
  return OHMB200_OK;
}

__global__ void insertRegion(DeviceMap dm, unsigned long long key, int *slot_out)
{
  *slot_out = regionSlot(dm, key);
}

int ohmb200_write_region(ohmb200_map *m, const int16_t key_xyz[3], int layer, const void *src, size_t bytes)
{
  if (!m || !key_xyz || layer < 0 || layer >= OHMB200_LAYER_COUNT || !m->layer_slab[layer] || !src)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_write_region: bad arguments or layer absent");
  }
  const size_t chunk = m->region_layer_bytes[layer];
  if (bytes != chunk)
  {
    return setError(OHMB200_E_INVALID, "chunk size mismatch: %zu != %zu", bytes, chunk);
  }
  cudaSetDevice(m->device);
  const unsigned long long key = (unsigned long long)(uint16_t)key_xyz[0] | ((unsigned long long)(uint16_t)key_xyz[1] << 16) |
                                 ((unsigned long long)(uint16_t)key_xyz[2] << 32);
  if (pullCounters(m) == OHMB200_OK && m->h_counters->region_count >= m->dm.capacity && m->algo == 1)
  {
    int rc_room = makeRoom(m, 1);
    if (rc_room)
    {
      return rc_room;
    }
  }
  int *d_slot = (int *)&m->d_counters->record_count;  // scratch word, reset before every batch
  DeviceMap dm_insert = m->dm;
  dm_insert.new_slots = nullptr;  // the per-batch list of new regions is not this call's business (and is not reset here)
  insertRegion<<<1, 1, 0, m->stream>>>(dm_insert, key, d_slot);
  int slot = -1;
  CUDA_TRY(cudaMemcpyAsync(&slot, d_slot, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (slot < 0)
  {
    return setError(OHMB200_E_CACHE_FULL, "region table full");
  }
  ++m->known_regions;  // (at most) one more resident region than ensureRoom knows of
  int rc_page = pageIn(m, key, (uint32_t)slot);  // its other layers, if the region was evicted
  if (rc_page)
  {
    return rc_page;
  }
  CUDA_TRY(cudaMemcpyAsync((char *)m->layer_slab[layer] + (size_t)slot * chunk, src, chunk, cudaMemcpyHostToDevice,
                           m->stream));
  if (m->dm.voxel_bits && (layer == OHMB200_LAYER_TSDF || layer == OHMB200_LAYER_MEAN))
  {
    recomputeVoxelBits<<<1, 256, 0, m->stream>>>(m->dm, m->geom, m->mp, m->mode == OHMB200_MODE_TSDF, (uint32_t)slot);
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return OHMB200_OK;
}

int ohmb200_clear(ohmb200_map *m)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  cudaSetDevice(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
  const ClearTable table = makeClearTable(m);
  m->store.clear();
  m->regions_bound = 0;
  m->known_regions = 0;
  m->known_batch = m->batches;
  m->snap_pending = false;
  clearLiveRegions<<<std::min<unsigned>(m->dm.capacity, (unsigned)m->sm_count * 16u), 256, 0, m->stream>>>(m->dm, table);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(m->d_counters, 0, sizeof(Counters), m->stream));
  // (queued, like everything else: what follows on the map's stream sees the empty map; no host wait)
  return OHMB200_OK;
}

double ohmb200_first_ray_time(const ohmb200_map *m)
{
  return m ? m->first_ray_time : -1.0;
}

int ohmb200_set_first_ray_time(ohmb200_map *m, double t)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  m->first_ray_time = t;
  return OHMB200_OK;
}

int ohmb200_get_stats(ohmb200_map *m, ohmb200_stats *stats)
{
  if (!m || !stats)
  {
    return setError(OHMB200_E_INVALID, "null argument");
  }
  cudaSetDevice(m->device);
  int rc = pullCounters(m);
  if (rc)
  {
    return rc;
  }
  stats->rays_in = m->rays_in;
  stats->rays_accepted = m->h_counters->rays_accepted;
  stats->voxel_visits = m->h_counters->voxel_visits;
  stats->sample_updates = m->h_counters->sample_updates;
  stats->ordered_records = m->h_counters->ordered_records;
  stats->regions = m->h_counters->region_count + m->store.size();
  stats->region_capacity = m->dm.capacity;
  stats->batches = m->batches;
  stats->kernel_launches = m->launches;
  stats->sample_voxels = m->h_counters->sample_voxels;
  stats->owned_visits = m->ex.open ? m->h_counters->owned_visits : m->h_counters->voxel_visits;
  return OHMB200_OK;
}

int ohmb200_remove_region(ohmb200_map *m, const int16_t key_xyz[3])
{
  if (!m || !key_xyz)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_remove_region: bad arguments");
  }
  cudaSetDevice(m->device);
  const unsigned long long key = packRegion(key_xyz[0], key_xyz[1], key_xyz[2]);
  if (m->store.erase(key))
  {
    return OHMB200_OK;  // it lived in the host store
  }
  CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
  std::vector<uint32_t> slot;
  int rc = findSlots(m, key_xyz, 1, slot);  // waits for the queued batches
  if (rc)
  {
    return rc;
  }
  rc = ensureGather(m, 16, 1);
  if (rc)
  {
    return rc;
  }
  // the slot is cleared for its next tenant and becomes a tombstone (probe sequences that pass it stay intact)
  CUDA_TRY(cudaMemcpyAsync(m->d_gather_slots, slot.data(), sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
  clearEvictedSlots<<<1, 256, 0, m->stream>>>(m->dm, makeClearTable(m), m->d_gather_slots, 1u);
  CUDA_TRY(cudaGetLastError());
  const unsigned long long tomb = kTombKey;
  CUDA_TRY(cudaMemcpyAsync(m->dm.keys + slot[0], &tomb, sizeof(tomb), cudaMemcpyHostToDevice, m->stream));
  rc = pullCounters(m);
  if (rc)
  {
    return rc;
  }
  const unsigned long long resident = m->h_counters->region_count ? m->h_counters->region_count - 1 : 0;
  CUDA_TRY(cudaMemcpyAsync(&m->d_counters->region_count, &resident, sizeof(resident), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return OHMB200_OK;
}

int ohmb200_set_region_reserve(ohmb200_map *m, uint32_t free_slots)
{
  if (!m || free_slots == 0 || free_slots > m->dm.capacity)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_set_region_reserve: 1 .. region capacity");
  }
  m->region_reserve = free_slots;
  return OHMB200_OK;
}

int ohmb200_paging_stats(ohmb200_map *m, uint64_t *resident, uint64_t *stored, uint64_t *evicted, uint64_t *paged_in)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  cudaSetDevice(m->device);
  int rc = pullCounters(m);
  if (rc)
  {
    return rc;
  }
  if (resident)
  {
    *resident = m->h_counters->region_count;
  }
  if (stored)
  {
    *stored = m->store.size();
  }
  if (evicted)
  {
    *evicted = m->evicted;
  }
  if (paged_in)
  {
    *paged_in = m->paged_in;
  }
  return OHMB200_OK;
}

int ohmb200_set_stream(ohmb200_map *m, void *cuda_stream)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  cudaSetDevice(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
  return OHMB200_OK;
}

int ohmb200_set_partition(ohmb200_map *m, int rank, int world)
{
  if (!m || world < 1 || rank < 0 || rank >= world)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_set_partition: need 0 <= rank < world");
  }
  cudaSetDevice(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  m->dm.part_rank = rank;
  m->dm.part_world = world;
  dropBatchGraphs(m);
  return OHMB200_OK;
}

int ohmb200_region_owner(const int16_t key_xyz[3], int world)
{
  return (key_xyz && world >= 1) ? regionOwner(key_xyz[0], key_xyz[1], key_xyz[2], world) : -1;
}

int ohmb200_set_profiling(ohmb200_map *m, int enabled)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  drainSpans(m);
  m->profiling = enabled != 0;
  return OHMB200_OK;
}

int ohmb200_kernel_times(ohmb200_map *m, ohmb200_kernel_time *out, int capacity, int reset)
{
  if (!m || !out)
  {
    return setError(OHMB200_E_INVALID, "null argument");
  }
  cudaSetDevice(m->device);
  drainSpans(m);
  int n = 0;
  for (int k = 0; k < kKernelCount && n < capacity; ++k)
  {
    if (m->kernel_launches[k] == 0)
    {
      continue;
    }
    memset(&out[n], 0, sizeof(out[n]));
    strncpy(out[n].name, kKernelNames[k], sizeof(out[n].name) - 1);
    out[n].ms = m->kernel_ms[k];
    out[n].launches = m->kernel_launches[k];
    ++n;
  }
  if (reset)
  {
    memset(m->kernel_ms, 0, sizeof(m->kernel_ms));
    memset(m->kernel_launches, 0, sizeof(m->kernel_launches));
  }
  return n;
}

int ohmb200_exchange_open(ohmb200_map *m, int rank, int world, size_t max_rays_per_rank, ohmb200_exchange_handle *handle)
{
  return exchangeOpen(m, rank, world, max_rays_per_rank, handle);
}

int ohmb200_exchange_connect(ohmb200_map *m, const ohmb200_exchange_handle *handles, int count)
{
  return exchangeConnect(m, handles, count);
}

size_t ohmb200_exchange_send_device(ohmb200_map *m, const double *d_rays, size_t element_count, const float *d_intensities,
                                    const double *d_timestamps, unsigned ray_flags)
{
  return exchangeSend(m, d_rays, element_count, d_intensities, d_timestamps, ray_flags) == OHMB200_OK ? element_count : 0;
}

size_t ohmb200_exchange_send(ohmb200_map *m, const double *rays, size_t element_count, const float *intensities,
                             const double *timestamps, unsigned ray_flags)
{
  return exchangeSendHost(m, rays, element_count, intensities, timestamps, ray_flags) == OHMB200_OK ? element_count : 0;
}

int ohmb200_exchange_integrate(ohmb200_map *m)
{
  return exchangeIntegrate(m);
}

int ohmb200_exchange_close(ohmb200_map *m)
{
  return exchangeClose(m);
}

int ohmb200_exchange_barrier(ohmb200_map *m)
{
  return exchangeBarrier(m);
}

int ohmb200_exchange_last_counts(ohmb200_map *m, uint32_t *segments, uint32_t *samples, int capacity)
{
  if (!m || !m->ex.open || capacity < m->ex.world)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_last_counts: no open exchange, or capacity < world");
  }
  cudaSetDevice(m->device);
  uint32_t counts[2 * kMaxWorld];
  CUDA_TRY(cudaMemcpyAsync(counts, m->ex.out_counts, sizeof(counts), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  for (int r = 0; r < m->ex.world; ++r)
  {
    if (segments)
    {
      segments[r] = counts[r];
    }
    if (samples)
    {
      samples[r] = counts[kMaxWorld + r];
    }
  }
  return OHMB200_OK;
}

}  // extern "C"
