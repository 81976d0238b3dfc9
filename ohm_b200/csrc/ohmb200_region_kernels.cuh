// ohmb200_region_kernels.cuh — kernels of the region-binned pipeline (included by ohmb200.cu after Batch/Counters).
//
//   prepRays      1 thread/ray   filter, keys, walk constants (RayRec), sample pair; pass A of the exact segment
//                                enumeration: region find-or-insert + per-region segment histogram
//   planRegions   1 CTA          exclusive scan of the histogram -> segment offsets; (region, <=8192 segments) work items
//   emitSegments  1 thread/ray   pass B: re-enumerate, scatter 16-byte segments into their region's range
//   walkRegions   persistent CTAs, one work item at a time: zero a u16 counter tile in shared memory, flag this
//                                region's sample voxels, resume every segment's walk against the tile (1 shared-memory
//                                atomic per visit), then fold the tile into the occupancy slab with 128-bit RMW
//   linkRecords   1 thread/record  attach each ordered-miss record to the run of its voxel (binary search)
//   applySamples  (shared with the per-ray path)
#pragma once

namespace ohmb200
{
// Adds 1 to counters[slot] for every calling lane; lanes of the converged group that share a slot issue one atomic.
// Returns this lane's position in its slot (previous value + rank).
__device__ __forceinline__ uint32_t slotAggregatedInc(uint32_t *counters, uint32_t slot)
{
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, slot);
  const int leader = __ffs(peers) - 1;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader)
  {
    base = atomicAdd(&counters[slot], (uint32_t)__popc(peers));
  }
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ uint32_t lowerBound(const uint32_t *keys, uint32_t n, uint32_t value)
{
  uint32_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < value)
    {
      lo = mid + 1;
    }
    else
    {
      hi = mid;
    }
  }
  return lo;
}
}  // namespace ohmb200

__global__ void __launch_bounds__(128) prepRays(DeviceMap dm, Geom g, MapParams mp, Batch b, int mode)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool accepted = false;
  unsigned visits = 0;
  if (i < b.n)
  {
    double start[3], end[3];
    loadRay(b, i, start, end);
    unsigned filter_flags = 0;
    uint32_t vid = kInvalidVoxel;
    RayRec rec;
    rec.flags = 0;
    if (applyRayFilter(mp, start, end, filter_flags))
    {
      accepted = true;
      const bool include_sample_in_ray = (filter_flags & kRffClippedEnd) || (b.ray_flags & OHMB200_RF_END_POINT_AS_FREE);
      bool hit = !include_sample_in_ray && mode != OHMB200_MODE_TSDF;
      if (mode == OHMB200_MODE_OCCUPANCY)
      {
        hit = hit && !(b.ray_flags & OHMB200_RF_EXCLUDE_SAMPLE);
      }
      Key ekey;
      if (hit && voxelKey(g, end, ekey) && ownsRegion(dm, ekey.r))
      {
        const int slot = regionSlot(dm, packRegion(ekey.r[0], ekey.r[1], ekey.r[2]));
        if (slot >= 0)
        {
          vid = (uint32_t)slot * g.vpr + voxelIndex(g, ekey);
        }
      }
      unsigned walk_flags = 0;
      if (mode != OHMB200_MODE_TSDF)
      {
        walk_flags = (!include_sample_in_ray) ? kExcludeEndVoxel : 0u;
        walk_flags |= (b.ray_flags & OHMB200_RF_EXCLUDE_ORIGIN) ? kExcludeStartVoxel : 0u;
      }
      if (mode == OHMB200_MODE_TSDF || !(b.ray_flags & OHMB200_RF_EXCLUDE_RAY))
      {
        if (!makeRayRec(rec, g, start, end, walk_flags) && (rec.flags & kRecValid) == 0)
        {
          rec.flags = 0;
        }
      }
      if (b.ray_length)
      {
        const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
        const double len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        b.ray_length[i] = (len2 > 1e-6) ? sqrt(len2) : 0;  // the walk's own length (LineWalkCompute.h:194-196)
      }
    }
    b.keys_in[i] = vid;
    b.vals_in[i] = i;
    reinterpret_cast<uint4 *>(b.recs + i)[0] = reinterpret_cast<const uint4 *>(&rec)[0];
    reinterpret_cast<uint4 *>(b.recs + i)[1] = reinterpret_cast<const uint4 *>(&rec)[1];
    reinterpret_cast<uint4 *>(b.recs + i)[2] = reinterpret_cast<const uint4 *>(&rec)[2];
    reinterpret_cast<uint4 *>(b.recs + i)[3] = reinterpret_cast<const uint4 *>(&rec)[3];
    if (b.last_exit)
    {
      b.last_exit[i] = 0;
    }
    if (rec.flags & kRecValid)
    {
      // Pass A: count this ray's segments per region (creating regions as they are first entered).
      enumerateSegments(rec, g, [&](const int r[3], const int st[3], int n) {
        (void)st;
        if (!ownsRegion(dm, r))
        {
          return;
        }
        const int slot = regionSlot(dm, packRegion(r[0], r[1], r[2]));
        visits += (unsigned)n;
        if (slot >= 0)
        {
          slotAggregatedInc(b.seg_count, (uint32_t)slot);
        }
      });
    }
  }
  __syncwarp();
  const unsigned n_acc = __reduce_add_sync(0xffffffffu, accepted ? 1u : 0u);
  const unsigned n_vis = __reduce_add_sync(0xffffffffu, visits);
  if ((threadIdx.x & 31) == 0)
  {
    if (n_acc)
    {
      atomicAdd(&b.counters->rays_accepted, (unsigned long long)n_acc);
    }
    if (n_vis)
    {
      atomicAdd(&b.counters->voxel_visits, (unsigned long long)n_vis);
    }
  }
}

// Single CTA: segment offsets per region slot and the work-item list.
__global__ void __launch_bounds__(1024) planRegions(DeviceMap dm, Batch b)
{
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage scan_storage;
  const uint32_t per = (dm.capacity + 1023u) / 1024u;
  const uint32_t first = threadIdx.x * per;
  const uint32_t last = min(first + per, dm.capacity);
  uint32_t sum = 0;
  for (uint32_t s = first; s < last; ++s)
  {
    sum += b.seg_count[s];
  }
  uint32_t base = 0, total = 0;
  Scan(scan_storage).ExclusiveSum(sum, base, total);
  for (uint32_t s = first; s < last; ++s)
  {
    b.seg_offset[s] = base;
    base += b.seg_count[s];
  }
  if (threadIdx.x == 0)
  {
    b.counters->segment_total = total;
    if (total > b.seg_capacity)
    {
      b.counters->segment_overflow = 1;
      b.counters->overflow_seen = 1;
    }
  }
  __syncthreads();
  // Work items: large regions first (they bound the tail), split into <= kMaxSegmentsPerItem pieces.
  for (int pass = 0; pass < 2; ++pass)
  {
    for (uint32_t s = first; s < last; ++s)
    {
      const uint32_t count = b.seg_count[s];
      const bool large = count >= kMaxSegmentsPerItem / 4;
      if (count == 0 || large != (pass == 0))
      {
        continue;
      }
      const uint32_t pieces = (count + kMaxSegmentsPerItem - 1) / kMaxSegmentsPerItem;
      const uint32_t at = atomicAdd(&b.counters->item_count, pieces);
      const uint32_t begin = b.seg_offset[s];
      for (uint32_t p = 0; p < pieces && at + p < b.item_capacity; ++p)
      {
        WorkItem w;
        w.slot = s;
        w.begin = begin + p * kMaxSegmentsPerItem;
        w.end = min(begin + count, w.begin + kMaxSegmentsPerItem);
        w.shared = pieces > 1;
        b.items[at + p] = w;
      }
    }
    __syncthreads();
  }
}

// Pass B: scatter the segments into their region's range.
__global__ void __launch_bounds__(128) emitSegments(DeviceMap dm, Geom g, Batch b)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n)
  {
    return;
  }
  RayRec rec;
  reinterpret_cast<uint4 *>(&rec)[0] = reinterpret_cast<const uint4 *>(b.recs + i)[0];
  reinterpret_cast<uint4 *>(&rec)[1] = reinterpret_cast<const uint4 *>(b.recs + i)[1];
  reinterpret_cast<uint4 *>(&rec)[2] = reinterpret_cast<const uint4 *>(b.recs + i)[2];
  reinterpret_cast<uint4 *>(&rec)[3] = reinterpret_cast<const uint4 *>(b.recs + i)[3];
  if (!(rec.flags & kRecValid))
  {
    return;
  }
  enumerateSegments(rec, g, [&](const int r[3], const int st[3], int n) {
    if (!ownsRegion(dm, r))
    {
      return;
    }
    const int slot = regionFind(dm, packRegion(r[0], r[1], r[2]));
    if (slot < 0)
    {
      return;
    }
    const uint32_t at = b.seg_offset[slot] + slotAggregatedInc(b.seg_cursor, (uint32_t)slot);
    if (at < b.seg_capacity)
    {
      uint4 raw;
      raw.x = i;
      raw.y = (uint32_t)st[0] | ((uint32_t)st[1] << 16);
      raw.z = (uint32_t)st[2] | ((uint32_t)n << 16);
      raw.w = 0;
      reinterpret_cast<uint4 *>(b.segments)[at] = raw;
    }
  });
}

// Persistent CTAs: one (region, segment range) work item at a time against a shared-memory counter tile.
__global__ void __launch_bounds__(256) walkRegions(DeviceMap dm, Geom g, MapParams mp, Batch b, int has_samples)
{
  extern __shared__ uint32_t tile[];
  __shared__ WorkItem item;
  __shared__ uint32_t sample_range[2];
  const uint32_t words = (g.vpr + 1u) >> 1;
  const uint32_t tid = threadIdx.x;
  const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];

  for (;;)
  {
    __syncthreads();
    if (tid == 0)
    {
      const uint32_t w = atomicAdd(&b.counters->work_next, 1u);
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
    for (uint32_t w = tid; w < words; w += blockDim.x)
    {
      tile[w] = 0;
    }
    if (has_samples && tid < 2)
    {
      sample_range[tid] = lowerBound(b.keys_out, b.n, vbase + tid * g.vpr);
    }
    __syncthreads();
    if (has_samples)
    {
      // Voxels that also receive samples in this batch: their misses must stay ordered against the hits.
      for (uint32_t s = sample_range[0] + tid; s < sample_range[1]; s += blockDim.x)
      {
        const uint32_t v = b.keys_out[s] - vbase;
        atomicOr(&tile[v >> 1], kTileFlag << ((v & 1u) * 16u));
      }
      __syncthreads();
    }

    for (uint32_t s = item.begin + tid; s < item.end; s += blockDim.x)
    {
      const uint4 raw = reinterpret_cast<const uint4 *>(b.segments)[s];
      const uint32_t ray = raw.x;
      const int st[3] = { (int)(raw.y & 0xffffu), (int)(raw.y >> 16), (int)(raw.z & 0xffffu) };
      const int visits = (int)(raw.z >> 16);
      const RayRec *rp = b.recs + ray;
      // tail of the record: region[3] i16 | local[3] u8 | flags u8 | total[3] u16
      const uint4 tail = reinterpret_cast<const uint4 *>(rp)[3];
      const uint32_t flags = (tail.z >> 8) & 0xffu;
      const int total[3] = { (int)(tail.z >> 16), (int)(tail.w & 0xffffu), (int)(tail.w >> 16) };
      const int local0[3] = { (int)((tail.y >> 16) & 0xffu), (int)(tail.y >> 24), (int)(tail.z & 0xffu) };
      const double init[3] = { rp->initial[0], rp->initial[1], rp->initial[2] };
      const double delta[3] = { rp->delta[0], rp->delta[1], rp->delta[2] };
      auto visit = [&](const int l[3], double t_enter, double t_exit, bool last_of_ray) {
        const uint32_t idx = (uint32_t)(l[0] + l[1] * dx + l[2] * dxy);
        const uint32_t shift = (idx & 1u) * 16u;
        const uint32_t old = atomicAdd(&tile[idx >> 1], 1u << shift);
        if ((old >> shift) & kTileFlag)
        {
          const uint32_t r = warpAggregatedInc(&b.counters->record_count);
          if (r < b.record_capacity)
          {
            b.record_vid[r] = vbase + idx;
            b.record_ray[r] = ray;
          }
          else
          {
            b.counters->record_overflow = 1;
            b.counters->overflow_seen = 1;
          }
        }
        if (dm.traversal)
        {
          atomicAdd(&dm.traversal[vbase + idx], (float)(t_exit - t_enter));
          if (last_of_ray)
          {
            b.last_exit[ray] = t_exit;
          }
        }
      };
      if (dm.traversal)
      {
        resumeSegment<true>(init, delta, local0, total, flags, st, visits, b.ray_length[ray], g, visit);
      }
      else
      {
        resumeSegment<false>(init, delta, local0, total, flags, st, visits, 0.0, g, visit);
      }
    }
    __syncthreads();

    // Fold the miss counts into the occupancy slab.  k identical misses commute, so the count is all that matters.
    float *occ = dm.occupancy + (size_t)vbase;
    if (!item.shared && (g.vpr & 7u) == 0)
    {
      const uint4 *tile4 = reinterpret_cast<const uint4 *>(tile);
      float4 *occ4 = reinterpret_cast<float4 *>(occ);
      for (uint32_t c = tid; c < (g.vpr >> 3); c += blockDim.x)
      {
        const uint4 t = tile4[c];
        const uint32_t w4[4] = { t.x, t.y, t.z, t.w };
        uint32_t cnt[8];
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
          const uint32_t lo = w4[k] & 0xffffu, hi = w4[k] >> 16;
          cnt[2 * k] = (lo & kTileFlag) ? 0u : lo;
          cnt[2 * k + 1] = (hi & kTileFlag) ? 0u : hi;
          any |= cnt[2 * k] | cnt[2 * k + 1];
        }
        if (any)
        {
          float4 a = occ4[2 * c], d = occ4[2 * c + 1];
          a.x = missRepeat(a.x, cnt[0], mp, b.ray_flags);
          a.y = missRepeat(a.y, cnt[1], mp, b.ray_flags);
          a.z = missRepeat(a.z, cnt[2], mp, b.ray_flags);
          a.w = missRepeat(a.w, cnt[3], mp, b.ray_flags);
          d.x = missRepeat(d.x, cnt[4], mp, b.ray_flags);
          d.y = missRepeat(d.y, cnt[5], mp, b.ray_flags);
          d.z = missRepeat(d.z, cnt[6], mp, b.ray_flags);
          d.w = missRepeat(d.w, cnt[7], mp, b.ray_flags);
          occ4[2 * c] = a;
          occ4[2 * c + 1] = d;
        }
      }
    }
    else
    {
      for (uint32_t v = tid; v < g.vpr; v += blockDim.x)
      {
        const uint32_t half = (tile[v >> 1] >> ((v & 1u) * 16u)) & 0xffffu;
        if (half == 0 || (half & kTileFlag))
        {
          continue;
        }
        if (!item.shared)
        {
          occ[v] = missRepeat(occ[v], half, mp, b.ray_flags);
        }
        else
        {
          // Region split over several work items: serialise the read-modify-write per voxel.
          int *addr = reinterpret_cast<int *>(occ + v);
          int seen = *reinterpret_cast<volatile int *>(addr);
          for (;;)
          {
            const float next = missRepeat(__int_as_float(seen), half, mp, b.ray_flags);
            const int prev = atomicCAS(addr, seen, __float_as_int(next));
            if (prev == seen)
            {
              break;
            }
            seen = prev;
          }
        }
      }
    }
  }
}

// Attach every ordered-miss record to the run (sorted sample pairs) of its voxel.
__global__ void linkRecords(Batch b)
{
  const uint32_t count = min(b.counters->record_count, b.record_capacity);
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x)
  {
    const uint32_t head = lowerBound(b.keys_out, b.n, b.record_vid[r]);
    b.record_next[r] = atomicExch(&b.run_head[head], (int32_t)r);
  }
}
