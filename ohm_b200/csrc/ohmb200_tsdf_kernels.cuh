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

// Persistent per-voxel bit (dm.voxel_bits): "the stored state is not order-free".  Rewritten wherever the TSDF layer is:
// replayTsdf, ohmb200_write_region, a change of the truncation distance.  (The fold only touches order-free voxels and
// leaves them order-free.)
// Pass 1, one thread per ray: flag (per-batch `near` bits) the voxels this ray visits at sdf < far threshold.  Only the
// last tsdfNearSteps() voxels of a walk can qualify, so a lane looks at its last one or two staged segments.
__global__ void __launch_bounds__(128) markTsdfNear(const __grid_constant__ DeviceMap dm, const __grid_constant__ Geom g,
                                                    const __grid_constant__ MapParams mp, const __grid_constant__ Batch b,
                                                    uint32_t *near)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n || mp.tsdf_dropoff > 0)  // with a dropoff nothing commutes: the count pass flags every voxel itself
  {
    return;
  }
  const uint32_t staged = b.stage_count[i];
  if (staged == 0)
  {
    return;
  }
  const uint32_t ray = rayOfThread(i, b.n);  // staged segments are indexed by thread (prepSegments)
  RayRec rec;
  loadRec(rec, b.recs + ray);
  const float far_threshold = tsdfFarThreshold(mp);
  const int near_steps = tsdfNearSteps(mp, g);
  const uint32_t flag_words = (g.vpr + 31u) >> 5;
  const int total[3] = { rec.total[0], rec.total[1], rec.total[2] };
  const int local0[3] = { rec.local[0], rec.local[1], rec.local[2] };
  const int steps_total = total[0] + total[1] + total[2];
  const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];
  TsdfRayGeometry geo;
  tsdfLoadRay(b, ray, geo);
  // returns false when the segment (and so every earlier one) ends too far from the end of the walk
  auto mark_segment = [&](uint32_t slot, const int st[3], int visits) -> bool {
    const int q_entry = st[0] + st[1] + st[2];
    if (steps_total - (q_entry + visits - 1) > near_steps)
    {
      return false;
    }
    int region[3];
    unpackRegion(dm.keys[slot], region);
    uint32_t *region_flags = near + (size_t)slot * flag_words;
    int q = q_entry;
    resumeSegment<false>(rec.initial, rec.delta, local0, total, rec.flags, st, visits, 0.0, g,
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
    return true;
  };
  if (staged <= kStageSegments)
  {
    for (uint32_t k = staged; k-- > 0;)
    {
      const uint4 raw = b.stage[(size_t)k * b.stage_stride + i];
      const int st[3] = { (int)(raw.y & 0xffffu), (int)(raw.y >> 16), (int)(raw.z & 0xffffu) };
      if (!mark_segment(raw.x, st, (int)(raw.z >> 16)))
      {
        break;
      }
    }
  }
  else
  {
    enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
      (void)entry;
      if (!ownsRegion(dm, r))
      {
        return;
      }
      const int slot = regionFind(dm, packRegion(r[0], r[1], r[2]));
      if (slot >= 0)
      {
        mark_segment((uint32_t)slot, st, n);
      }
    });
  }
}

// The fold of walkRegionsTsdf: k commuting far visits of a voxel -> (trunc, min(w + 1, max) k times).
__device__ __forceinline__ void foldTsdfTile(const uint32_t *tile, const TileLayout &layout, float2 *slab,
                                             const MapParams &mp, uint32_t *ticket, uint32_t shared)
{
  const TileLayout tl = foldLayout(layout);
  auto far_visits = [&](float w, uint32_t k) {
    if (w == floorf(w) && w + (float)k <= 16777216.0f)
    {
      return fminf(w + (float)k, mp.tsdf_max_weight);  // integer weights: every +1 is exact
    }
    for (uint32_t n = 0; n < k; ++n)
    {
      const float next = fminf(w + 1.0f, mp.tsdf_max_weight);
      if (next == w)
      {
        break;
      }
      w = next;
    }
    return w;
  };
  const auto voxel_after = [&](unsigned long long old, uint32_t count) {
    const float w = far_visits(__uint_as_float((uint32_t)old), count);
    return (unsigned long long)__float_as_uint(w) | ((unsigned long long)__float_as_uint(mp.tsdf_trunc) << 32);
  };
  if (!tl.fast)
  {
    foldTileVoxels(tile, tl, [&](uint32_t v, uint32_t, uint32_t half) {
      if (half == 0 || (half & kTileFlag))
      {
        return;
      }
      unsigned long long *unit = reinterpret_cast<unsigned long long *>(slab + v);
      if (!shared)
      {
        *unit = voxel_after(*unit, half);
      }
      else
      {
        foldUnitShared(unit, __ldcg(unit), [&](unsigned long long old) { return voxel_after(old, half); });
      }
    });
    return;
  }
  const auto counts = [&](uint32_t, const uint4 &t, uint32_t cnt[8]) {
    const uint32_t w4[4] = { t.x, t.y, t.z, t.w };
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const uint32_t half = (w4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
      cnt[k] = (half & kTileFlag) ? 0u : half;
      any |= cnt[k];
    }
    return any != 0;
  };
  const auto far_voxel = [&](float2 &v, uint32_t count) {
    if (count)
    {
      v.x = far_visits(v.x, count);
      v.y = mp.tsdf_trunc;
    }
  };
  if (!shared)
  {
    // sole writer of the region in this batch: 64-byte read-modify-write of eight voxels at a time
    foldGroupsSole<1>(tile, tl, slab, ticket, counts, far_voxel);
  }
  else
  {
    foldGroupsShared(tile, tl, slab, ticket, counts, far_voxel);
  }
}

// Pass 2: count far visits of unflagged voxels in the tile, record every visit of flagged voxels (stored state not
// order-free, or near a sample in this batch), fold the counts.
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) walkRegionsTsdf(const __grid_constant__ DeviceMap dm,
                                                                   const __grid_constant__ Geom g,
                                                                   const __grid_constant__ MapParams mp,
                                                                   const __grid_constant__ Batch b,
                                                                   const __grid_constant__ TileLayout tl, const uint32_t *near)
{
  extern __shared__ uint32_t tile[];
  __shared__ WorkItem items2[2];  // this work item and the next one (fetched during the walk)
  uint32_t parity = 1;
  __shared__ unsigned long long record_chunk[kWalkThreads / 32];
  __shared__ SegmentQueue queue;
  __shared__ uint32_t fold_ticket;  // next block of the fold (foldGroupsSole / foldGroupsShared)
  const uint32_t words = tl.words;
  const uint32_t tile_base = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t flag_words = (g.vpr + 31u) >> 5;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const bool all_ordered = mp.tsdf_dropoff > 0;  // weights depend on sdf: nothing commutes
  if ((tid & 31u) == 0)
  {
    record_chunk[warp] = (unsigned long long)kRecordChunk;
  }
  if (tid == 0)
  {
    loadWorkItem(b, atomicAdd(&b.counters->work_next, 1u), &items2[0]);
    queueInit(queue);
    queueStage(queue, b, items2[0]);
  }
  for (;;)
  {
    parity ^= 1u;
    __syncthreads();  // the item is in place; the previous fold is done with the tile
    const WorkItem &item = items2[parity];
    if (item.slot == 0xFFFFFFFFu)
    {
      return;
    }
    const uint32_t slot = item.slot;
    const uint32_t vbase = slot * g.vpr;
    {
      uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
      const uint32_t fill = all_ordered ? (kTileFlag | (kTileFlag << 16)) : 0u;
      for (uint32_t w = tid; w < (words >> 2); w += blockDim.x)
      {
        tile4[w] = make_uint4(fill, fill, fill, fill);
      }
    }
    __syncthreads();
    if (!all_ordered)
    {
      const uint32_t *ordered_bits = dm.voxel_bits + (size_t)slot * flag_words;
      const uint32_t *near_bits = near + (size_t)slot * flag_words;
      for (uint32_t w = tid; w < flag_words; w += blockDim.x)
      {
        uint32_t bits = ordered_bits[w] | near_bits[w];
        while (bits)
        {
          const uint32_t half = tileHalf(tl, (w << 5) + (uint32_t)__ffs(bits) - 1u);
          bits &= bits - 1u;
          atomicOr(&tile[half >> 1], kTileFlag << ((half & 1u) * 16u));
        }
      }
    }
    queueBuild(queue, b, item, parity);  // (the copy barrier's phase flips with the item slot)
    uint32_t next_work = 0;
    if (tid == 0)
    {
      next_work = atomicAdd(&b.counters->work_next, 1u);  // the reply arrives while this item is walked
    }

    for (;;)
    {
      uint4 raw;
      const int got = queuePop(queue, b, item, raw);
      if (got == 0)
      {
        break;
      }
      if (got == 1)
      {
        SegmentWalk sw;
        loadSegmentWalk(b, raw, sw);
        const uint32_t ray = sw.ray;
        resumeSegmentTile(sw.init, sw.delta, sw.entry, sw.total, sw.flags, sw.st, sw.visits, tl, tile_base, [&](uint32_t offset, uint32_t one) {
          if (tileAdd(offset, one) & (one << 15))
          {
            const uint32_t at = reserveRecord(&record_chunk[warp], &b.counters->record_count);
            if (at < b.record_capacity)
            {
              b.record_keys[at] = ((unsigned long long)(vbase + tileVoxel(tl, (offset - tile_base) >> 1)) << 32) | ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              atomicOr(&b.counters->overflow_seen, 1);
            }
          }
        });
      }
      __syncwarp();  // every lane of the warp is back together before the next pop
    }
    if (tid == 0)
    {
      loadWorkItem(b, next_work, &items2[parity ^ 1u]);
      fold_ticket = 0;
    }
    __syncthreads();
    if (tid == 0)
    {
      queueStage(queue, b, items2[parity ^ 1u]);  // the next item's segments arrive while this one is folded
    }

    foldTsdfTile(tile, tl, dm.tsdf + (size_t)vbase, mp, &fold_ticket, item.shared);
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
  setVoxelBit(dm, g, vid, !tsdfOrderFree(state, mp));
}

// Clears the per-batch near bits of the regions this batch walked.
__global__ void clearTouchedBits(Batch b, uint32_t *bits, uint32_t words_per_region)
{
  const uint32_t touched = b.counters->touched_count;
  for (uint32_t t = blockIdx.x; t < touched; t += gridDim.x)
  {
    uint32_t *region = bits + (size_t)b.touched_list[t] * words_per_region;
    for (uint32_t w = threadIdx.x; w < words_per_region; w += blockDim.x)
    {
      region[w] = 0;
    }
  }
}

// dm.voxel_bits of the regions in slots [first, first + gridDim.x), from the stored layers.
//   NDT:  bit = voxel mean count >= sample threshold ("established Gaussian")
//   TSDF: bit = stored (weight, distance) is not order-free
__global__ void recomputeVoxelBits(DeviceMap dm, Geom g, MapParams mp, int tsdf_mode, uint32_t first)
{
  const uint32_t slot = first + blockIdx.x;
  if (slot >= dm.capacity || !isRegionKey(dm.keys[slot]))
  {
    return;
  }
  const uint32_t flag_words = (g.vpr + 31u) >> 5;
  const uint32_t vbase = slot * g.vpr;
  for (uint32_t base = (threadIdx.x >> 5) << 5; base < g.vpr; base += blockDim.x)
  {
    const uint32_t v = base + (threadIdx.x & 31u);
    bool bit = false;
    if (v < g.vpr)
    {
      bit = tsdf_mode ? !tsdfOrderFree(dm.tsdf[vbase + v], mp) : dm.mean[vbase + v].y >= mp.sample_threshold;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31u) == 0)
    {
      dm.voxel_bits[(size_t)slot * flag_words + (base >> 5)] = word;
    }
  }
}
