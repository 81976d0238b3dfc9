// ohmb200_tsdf_kernels.cuh — GpuTsdfMap: RayMapperTsdf::integrateRays (ohm/RayMapperTsdf.cpp:87-182) on the
// region-binned pipeline.  Included by ohmb200.cu after ohmb200_region_kernels.cuh.
//
// calculateTsdf (ohm/VoxelTsdfCompute.h:87-136) is order dependent: each update is a clamped weighted mean, so the
// reference's own GPU kernel (one 64-bit CAS per visit, arbitrary order) is only approximately the CPU result.  Here:
//   * A visit is "far" when sdf >= trunc * (1 + 8u(max_weight + 2)), u = 2^-24, with dropoff disabled.  For a voxel
//     that is unobserved (w = 0, d = 0) or already saturated at d == +trunc, a far visit yields exactly
//     (d, w) <- (trunc, min(w + 1, max_weight)) whatever its sdf: the margin covers the four roundings of
//     (sdf*1 + d*w) / (w + 1) for every w <= max_weight.  Such visits commute: count them, apply the count.
//   * Every other voxel — it has a near visit in this batch, or its stored distance is not the saturated value — is
//     flagged by a first pass; ALL its visits of the batch are recorded, sorted by (voxel, ray) and replayed in ray order.
// Free space (the bulk of the visits) takes the counting path; surfaces take the ordered path.  Result: bit-exact.
#pragma once

namespace ohmb200
{
// ohm/VoxelTsdfCompute.h:57-69 computeDistance<glm::dvec3>
__device__ __forceinline__ float tsdfDistance(const double sensor[3], const double sample[3], const double centre[3],
                                              float distance_g)
{
  const double s2v[3] = { centre[0] - sensor[0], centre[1] - sensor[1], centre[2] - sensor[2] };
  const double s2s[3] = { sample[0] - sensor[0], sample[1] - sensor[1], sample[2] - sensor[2] };
  const float distance_g_v = (float)((s2v[0] * s2s[0] + s2v[1] * s2s[1]) + s2v[2] * s2s[2]) / distance_g;
  return distance_g - distance_g_v;
}

// ohm/VoxelTsdfCompute.h:87-136 calculateTsdf
__device__ __forceinline__ void tsdfUpdate(float sdf, const MapParams &p, float &weight, float &distance)
{
  const float initial_weight = weight;
  float updated_weight = 1.0f;
  updated_weight *= (p.tsdf_dropoff > 0) ? ((p.tsdf_trunc + sdf) / (p.tsdf_trunc - p.tsdf_dropoff)) : 1.0f;
  updated_weight = fmaxf(updated_weight, 0.0f);
  updated_weight *= (p.tsdf_sparsity > 0 && fabsf(sdf) < p.tsdf_trunc) ? p.tsdf_sparsity : 1.0f;
  const float new_weight = initial_weight + updated_weight;
  const bool near_zero = fabsf(new_weight) < 0.00001f;
  const float new_sdf = (!near_zero) ? (sdf * updated_weight + distance * initial_weight) / new_weight : 0.0f;
  distance = (!near_zero) ? ((new_sdf > 0.0f) ? fminf(p.tsdf_trunc, new_sdf) : fmaxf(-p.tsdf_trunc, new_sdf)) : distance;
  weight = (!near_zero) ? fminf(new_weight, p.tsdf_max_weight) : initial_weight;
}

__device__ __forceinline__ float tsdfFarThreshold(const MapParams &p)
{
  return p.tsdf_trunc * (1.0f + 8.0f * 5.9604645e-8f * (p.tsdf_max_weight + 2.0f));
}

// A voxel `k` steps before the end of a walk has sdf >= res * sqrt((k/sqrt3 - sqrt3/2)^2 - 3/4): its key is k
// axis-steps from the end key (L2 >= L1/sqrt3), the end voxel's centre is within a half diagonal of the (filtered) end
// point, the ray passes through the voxel (perpendicular offset <= half diagonal), and the unfiltered sample lies on
// the ray at or beyond the filtered end.  Voxels further than this many steps from the end are far for certain, so the
// mark pass never has to look at them (+2 steps of slack for the float evaluation of sdf).
__device__ __forceinline__ int tsdfNearSteps(const MapParams &p, const Geom &g)
{
  const double ratio = (double)tsdfFarThreshold(p) / g.res;
  return (int)ceil(1.7320508075688772 * (sqrt(ratio * ratio + 0.75) + 0.8660254037844386)) + 2;
}

// Stored state for which far visits commute.
__device__ __forceinline__ bool tsdfOrderFree(float2 v, const MapParams &p)
{
  return (v.x == 0.0f && v.y == 0.0f) || v.y == p.tsdf_trunc;
}
}  // namespace ohmb200

struct TsdfRayGeometry
{
  double sensor[3], sample[3];
  float distance_g;
};

__device__ __forceinline__ void tsdfLoadRay(const Batch &b, uint32_t ray, TsdfRayGeometry &geo)
{
  loadRay(b, ray, geo.sensor, geo.sample);  // the UNFILTERED sensor and sample (RayMapperTsdf.cpp:164-165)
  const double d[3] = { geo.sample[0] - geo.sensor[0], geo.sample[1] - geo.sensor[1], geo.sample[2] - geo.sensor[2] };
  geo.distance_g = (float)sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
}

// Pass 1: flag every voxel whose visits must be replayed in order.  flags: one bit per voxel, [capacity][vpr/32].
// Pass 2 (kCount): count far visits of unflagged voxels in the tile, record every visit of flagged voxels, fold.
template <bool kCount>
__global__ void __launch_bounds__(kWalkThreads, 2) walkRegionsTsdf(const __grid_constant__ DeviceMap dm,
                                                                   const __grid_constant__ Geom g,
                                                                   const __grid_constant__ MapParams mp,
                                                                   const __grid_constant__ Batch b, uint32_t *flags)
{
  extern __shared__ uint32_t tile[];
  __shared__ WorkItem item;
  __shared__ unsigned long long record_chunk[kWalkThreads / 32];
  const uint32_t words = (g.vpr + 1u) >> 1;
  const uint32_t flag_words = (g.vpr + 31u) >> 5;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const float far_threshold = tsdfFarThreshold(mp);
  const bool all_ordered = mp.tsdf_dropoff > 0;  // weights depend on sdf: nothing commutes
  const int near_steps = tsdfNearSteps(mp, g);
  if ((tid & 31u) == 0)
  {
    record_chunk[warp] = (unsigned long long)kRecordChunk;
  }
  uint32_t *work_counter = kCount ? &b.counters->work_next : &b.counters->run_count;  // run_count is free in TSDF mode
  for (;;)
  {
    __syncthreads();
    if (tid == 0)
    {
      const uint32_t w = atomicAdd(work_counter, 1u);
      if (w < min(b.counters->item_count, b.item_capacity))
      {
        item = b.items[w];
      }
      else
      {
        item.slot = 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    if (item.slot == 0xFFFFFFFFu)
    {
      return;
    }
    const uint32_t slot = item.slot;
    const uint32_t vbase = slot * g.vpr;
    uint32_t *region_flags = flags + (size_t)slot * flag_words;
    int region[3];
    unpackRegion(dm.keys[slot], region);
    if (kCount)
    {
      for (uint32_t w = tid; w < words; w += blockDim.x)
      {
        tile[w] = 0;
      }
      __syncthreads();
      for (uint32_t v = tid; v < g.vpr; v += blockDim.x)
      {
        if ((region_flags[v >> 5] >> (v & 31u)) & 1u)
        {
          atomicOr(&tile[v >> 1], kTileFlag << ((v & 1u) * 16u));
        }
      }
      __syncthreads();
    }
    else
    {
      // stored states that are not order-free (only the first item of a region needs to do this; all do, idempotently)
      for (uint32_t v = tid; v < g.vpr; v += blockDim.x)
      {
        if (all_ordered || !tsdfOrderFree(dm.tsdf[vbase + v], mp))
        {
          atomicOr(&region_flags[v >> 5], 1u << (v & 31u));
        }
      }
    }

    for (uint32_t s = item.begin + tid; s < item.end; s += blockDim.x)
    {
      const uint4 raw = reinterpret_cast<const uint4 *>(b.segments)[s];
      const uint32_t ray = raw.x;
      const int st[3] = { (int)(raw.y & 0xffffu), (int)(raw.y >> 16), (int)(raw.z & 0xffffu) };
      const int visits = (int)(raw.z >> 16);
      const RayRec *rp = b.recs + ray;
      const uint4 tail = reinterpret_cast<const uint4 *>(rp)[3];
      const uint32_t rflags = (tail.z >> 8) & 0xffu;
      const int total[3] = { (int)(tail.z >> 16), (int)(tail.w & 0xffffu), (int)(tail.w >> 16) };
      const int local0[3] = { (int)((tail.y >> 16) & 0xffu), (int)(tail.y >> 24), (int)(tail.z & 0xffu) };
      const double init[3] = { rp->initial[0], rp->initial[1], rp->initial[2] };
      const double delta[3] = { rp->delta[0], rp->delta[1], rp->delta[2] };
      const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];
      if (kCount)
      {
        const int entry[3] = { (int)(raw.w & 0xffu), (int)((raw.w >> 8) & 0xffu), (int)((raw.w >> 16) & 0xffu) };
        resumeSegmentFast(init, delta, entry, total, rflags, st, visits, g, [&](uint32_t idx) {
          const uint32_t shift = (idx & 1u) * 16u;
          const uint32_t old = atomicAdd(&tile[idx >> 1], 1u << shift);
          if ((old >> shift) & kTileFlag)
          {
            const unsigned group = __activemask();
            const uint32_t n = (uint32_t)__popc(group);
            const uint32_t rank = (uint32_t)__popc(group & ((1u << (tid & 31u)) - 1u));
            uint32_t at = 0;
            if (rank == 0)
            {
              const unsigned long long state = atomicAdd(&record_chunk[warp], (unsigned long long)n);
              const uint32_t used = (uint32_t)state;
              if (used + n <= kRecordChunk)
              {
                at = (uint32_t)(state >> 32) + used;
              }
              else
              {
                at = atomicAdd(&b.counters->record_count, kRecordChunk);
                atomicExch(&record_chunk[warp], ((unsigned long long)at << 32) | n);
              }
            }
            at = __shfl_sync(group, at, __ffs(group) - 1) + rank;
            if (at < b.record_capacity)
            {
              b.record_keys[at] = ((unsigned long long)(vbase + idx) << 32) | ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              b.counters->overflow_seen = 1;
            }
          }
        });
      }
      else if (!all_ordered)
      {
        // Only the last `near_steps` voxels of a walk can be near the sample: skip segments that end earlier.
        const int steps_total = total[0] + total[1] + total[2];
        const int q_entry = st[0] + st[1] + st[2];
        if (steps_total - (q_entry + visits - 1) <= near_steps)
        {
          const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];
          TsdfRayGeometry geo;
          tsdfLoadRay(b, ray, geo);
          int q = q_entry;
          resumeSegment<false>(init, delta, local0, total, rflags, st, visits, 0.0, g,
                               [&](const int l[3], double, double, bool) {
                                 if (steps_total - q <= near_steps)
                                 {
                                   const double centre[3] = { voxelCentreAxis(g, region[0], l[0], 0),
                                                              voxelCentreAxis(g, region[1], l[1], 1),
                                                              voxelCentreAxis(g, region[2], l[2], 2) };
                                   const float sdf = tsdfDistance(geo.sensor, geo.sample, centre, geo.distance_g);
                                   if (!(sdf >= far_threshold))
                                   {
                                     const uint32_t idx = (uint32_t)(l[0] + l[1] * dx + l[2] * dxy);
                                     atomicOr(&region_flags[idx >> 5], 1u << (idx & 31u));
                                   }
                                 }
                                 ++q;
                               });
        }
      }
    }
    if (!kCount)
    {
      continue;
    }
    __syncthreads();
    // Fold: k commuting far visits -> (trunc, min(w + 1, max) k times).
    float2 *slab = dm.tsdf + (size_t)vbase;
    for (uint32_t v = tid; v < g.vpr; v += blockDim.x)
    {
      const uint32_t half = (tile[v >> 1] >> ((v & 1u) * 16u)) & 0xffffu;
      if (half == 0 || (half & kTileFlag))
      {
        continue;
      }
      unsigned long long *addr = reinterpret_cast<unsigned long long *>(slab + v);
      unsigned long long seen = *reinterpret_cast<volatile unsigned long long *>(addr);
      for (;;)
      {
        float w = __uint_as_float((uint32_t)seen);
        for (uint32_t k = 0; k < half; ++k)
        {
          const float next = fminf(w + 1.0f, mp.tsdf_max_weight);
          if (next == w)
          {
            break;
          }
          w = next;
        }
        const unsigned long long want =
          (unsigned long long)__float_as_uint(w) | ((unsigned long long)__float_as_uint(mp.tsdf_trunc) << 32);
        if (!item.shared)
        {
          *addr = want;
          break;
        }
        const unsigned long long prev = atomicCAS(addr, seen, want);
        if (prev == seen)
        {
          break;
        }
        seen = prev;
      }
    }
  }
}

// Ordered replay: records sorted by (voxel, ray); one thread per voxel run.
__global__ void __launch_bounds__(128) replayTsdf(DeviceMap dm, Geom g, MapParams mp, Batch b, uint32_t count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
  {
    return;
  }
  const unsigned long long key = b.record_keys_sorted[i];
  const uint32_t vid = (uint32_t)(key >> 32);
  if (vid == kInvalidVoxel || (i > 0 && (uint32_t)(b.record_keys_sorted[i - 1] >> 32) == vid))
  {
    return;
  }
  const uint32_t slot = vid / g.vpr;
  const uint32_t local = vid - slot * g.vpr;
  int r[3];
  unpackRegion(dm.keys[slot], r);
  const int l[3] = { (int)(local % (uint32_t)g.dim[0]), (int)((local / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]),
                     (int)(local / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1])) };
  const double centre[3] = { voxelCentreAxis(g, r[0], l[0], 0), voxelCentreAxis(g, r[1], l[1], 1),
                             voxelCentreAxis(g, r[2], l[2], 2) };
  float2 state = dm.tsdf[vid];
  for (uint32_t j = i; j < count; ++j)
  {
    const unsigned long long kj = b.record_keys_sorted[j];
    if ((uint32_t)(kj >> 32) != vid)
    {
      break;
    }
    TsdfRayGeometry geo;
    tsdfLoadRay(b, (uint32_t)kj, geo);
    tsdfUpdate(tsdfDistance(geo.sensor, geo.sample, centre, geo.distance_g), mp, state.x, state.y);
  }
  dm.tsdf[vid] = state;
}
