// ohmb200_region_kernels.cuh — kernels of the region-binned pipeline (included by ohmb200.cu after Batch/Counters).
//
//   prepRays      1 thread/ray   filter, keys, walk constants (RayRec), sample pair; pass A of the exact segment
//                                enumeration: region find-or-insert + per-region segment histogram + touched list
//   planRegions   1 CTA          scan of the touched regions' histogram -> segment offsets; (region, <=2048 segments)
//                                work items, largest regions first
//   emitSegments  1 thread/ray   pass B: re-enumerate, scatter 16-byte segments into their region's range
//   walkRegions   persistent CTAs, one work item at a time: zero a u16 counter tile in shared memory, flag this
//                                region's sample voxels, resume every segment's walk against the tile (1 shared-memory
//                                atomic per visit), then fold the tile into the occupancy slab with 128-bit RMW
//   linkRecords   1 thread/record  attach each ordered-miss record to the run of its voxel (binary search)
//   applySamples  (shared with the per-ray path)
#pragma once

namespace ohmb200
{
// Lanes of the converged group that share a slot issue one atomic for all of them.
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

__device__ __forceinline__ void loadRec(RayRec &rec, const RayRec *src)
{
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    reinterpret_cast<uint4 *>(&rec)[k] = reinterpret_cast<const uint4 *>(src)[k];
  }
}
}  // namespace ohmb200

// Filter, sample voxel and walk constants of every ray.  The sample pairs it writes are all the sort needs, so the
// sample path (radix sort -> run heads) starts right after this kernel, beside prepSegments.
__global__ void __launch_bounds__(128) prepRays(DeviceMap dm, Geom g, MapParams mp, Batch b, int mode)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool accepted = false;
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
        makeRayRec(rec, g, start, end, walk_flags);  // leaves rec.flags == 0 when the ray is not walked
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
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      reinterpret_cast<uint4 *>(b.recs + i)[k] = reinterpret_cast<const uint4 *>(&rec)[k];
    }
    if (b.last_exit)
    {
      // Exit range of the last voxel the walk visits, in closed form (the sample's owner need not have walked it):
      // the end voxel's exit is the walk's length; with the end voxel excluded the last visit ends on the walk's final
      // step, the latest of the per-axis last-step times.  NaN = the ray visits nothing (see staleExit).
      double last = nan("");
      if (rec.flags & kRecValid)
      {
        const int steps = (int)rec.total[0] + (int)rec.total[1] + (int)rec.total[2];
        if (!(rec.flags & kRecExcludeEnd))
        {
          const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
          const double len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
          last = (len2 > 1e-6) ? sqrt(len2) : 0;
        }
        else if (steps - ((rec.flags & kRecExcludeStart) ? 1 : 0) > 0)
        {
          last = -INFINITY;
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            if (rec.total[a])
            {
              last = fmax(last, stepTime(rec.initial[a], rec.delta[a], (int)rec.total[a] - 1));
            }
          }
        }
      }
      b.last_exit[i] = last;
    }
  }
  __syncwarp();
  const unsigned n_acc = __reduce_add_sync(0xffffffffu, accepted ? 1u : 0u);
  if ((threadIdx.x & 31) == 0 && n_acc)
  {
    atomicAdd(&b.counters->rays_accepted, (unsigned long long)n_acc);
  }
}

// Pass A of enumerateSegments, one thread per ray: region find-or-insert, per-region segment histogram, touched list,
// and the segments themselves staged for pass B.
__global__ void __launch_bounds__(128) prepSegments(DeviceMap dm, Geom g, Batch b)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned visits = 0;
  if (i < b.n)
  {
    RayRec rec;
    loadRec(rec, b.recs + rayOfThread(i, b.n));  // (the staged segments are indexed by thread, emitSegments maps back)
    uint32_t staged = 0;
    if (rec.flags & kRecValid)
    {
      // Pass A: count this ray's segments per region (creating regions as they are first entered) and stage them.
      // The two global round trips a segment needs — the hash probe of its region and the reply of the histogram
      // atomic — are kept off the dependent chain: a segment is FINISHED when the next one has been found (the probe
      // was issued a whole boundary search earlier), and the atomic's reply is looked at one segment later still.
      bool pend_valid = false;
      unsigned long long pend_key = 0, pend_probed = 0;
      uint32_t pend_home = 0, pend_y = 0, pend_z = 0, pend_w = 0;
      bool reply_valid = false;
      uint32_t reply_slot = 0, reply_old = 0;
      auto check_reply = [&]() {
        // the adder that found the region's count at zero is the first to enter it in this batch
        if (reply_valid && reply_old == 0u)
        {
          b.touched_list[atomicAdd(&b.counters->touched_count, 1u)] = reply_slot;
        }
        reply_valid = false;
      };
      auto finish = [&]() {
        if (!pend_valid)
        {
          return;
        }
        pend_valid = false;
        const int slot = (pend_probed == pend_key) ? (int)pend_home : regionSlot(dm, pend_key);
        if (slot < 0)
        {
          return;
        }
        check_reply();
        // one reduction per group of lanes that entered the same region
        const unsigned peers = __match_any_sync(__activemask(), slot);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
        {
          reply_old = atomicAdd(&b.seg_count[slot], (uint32_t)__popc(peers));
          reply_slot = (uint32_t)slot;
          reply_valid = true;
        }
        if (staged < kStageSegments)
        {
          b.stage[(size_t)staged * b.stage_stride + i] = make_uint4((uint32_t)slot, pend_y, pend_z, pend_w);
        }
        ++staged;
      };
      enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
        if (!ownsRegion(dm, r))
        {
          return;
        }
        finish();
        visits += (unsigned)n;
        pend_key = packRegion(r[0], r[1], r[2]);
        pend_home = hashRegion(pend_key) % dm.capacity;
        pend_probed = __ldcg(&dm.keys[pend_home]);
        pend_y = (uint32_t)st[0] | ((uint32_t)st[1] << 16);
        pend_z = (uint32_t)st[2] | ((uint32_t)n << 16);
        pend_w = (uint32_t)entry[0] | ((uint32_t)entry[1] << 8) | ((uint32_t)entry[2] << 16);
        pend_valid = true;
      });
      finish();
      check_reply();
    }
    b.stage_count[i] = staged;
  }
  __syncwarp();
  const unsigned n_vis = __reduce_add_sync(0xffffffffu, visits);
  if ((threadIdx.x & 31) == 0 && n_vis)
  {
    atomicAdd(&b.counters->voxel_visits, (unsigned long long)n_vis);
  }
}

// Single CTA over the touched regions only: segment offsets and the work-item list (largest regions first).
// The touched list (slot, count, offset) of the first kPlanCache regions is kept in shared memory, so the five
// size-class passes read no global memory and draw their item indices from a shared-memory counter (a sweep touches
// ~1000 regions; with every pass re-reading three dependent global arrays and a global atomic per region this kernel was
// 18 us of pure latency on the critical path).
constexpr uint32_t kPlanCache = 2048;
__global__ void __launch_bounds__(1024) planRegions(DeviceMap dm, Batch b, uint32_t walk_ctas)
{
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage scan_storage;
  __shared__ uint32_t carry, item_next;
  __shared__ uint32_t s_slot[kPlanCache], s_count[kPlanCache], s_offset[kPlanCache];
  const uint32_t touched = b.counters->touched_count;
  // the batch's age for paging, counted on the device (a replayed batch graph carries no host-side stamp)
  const uint32_t stamp = b.counters->batch_stamp + 1u;
  if (threadIdx.x == 0)
  {
    carry = 0;
    item_next = 0;
  }
  __syncthreads();
  for (uint32_t base = 0; base < touched; base += 1024)
  {
    const uint32_t t = base + threadIdx.x;
    const uint32_t slot = (t < touched) ? b.touched_list[t] : 0;
    const uint32_t count = (t < touched) ? b.seg_count[slot] : 0;
    uint32_t offset = 0, total = 0;
    Scan(scan_storage).ExclusiveSum(count, offset, total);
    if (t < touched)
    {
      b.seg_offset[slot] = carry + offset;
      dm.region_stamp[slot] = stamp;  // walked by this batch: the age paging evicts by
      if (t < kPlanCache)
      {
        s_slot[t] = slot;
        s_count[t] = count;
        s_offset[t] = carry + offset;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      carry += total;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    b.counters->segment_total = carry;
    b.counters->batch_stamp = stamp;
    if (carry > b.seg_capacity)
    {
      b.counters->segment_overflow = 1;
      atomicOr(&b.counters->overflow_seen, 2);
    }
  }
  if (carry > b.seg_capacity)
  {
    // More segments than the batch's list holds: no work items are made, so no walk kernel reads past the list, and
    // the sample replay kernels skip the batch too (they test segment_overflow) — the batch is dropped WHOLE and
    // ohmb200_sync reports OHMB200_E_OVERFLOW (the list is grown for the batches that follow).
    return;
  }
  // Work items in decreasing size classes so that the long items start first and the tail is made of small ones.
  // Item size: kMaxSegmentsPerItem, unless that leaves fewer than two items per persistent CTA (a small batch, or one
  // GPU's share of a sharded map): then the regions are cut finer so that every SM still gets work.
  uint32_t item_max = kMaxSegmentsPerItem;
  if (touched + carry / kMaxSegmentsPerItem < 2u * walk_ctas)
  {
    item_max = min(kMaxSegmentsPerItem, max(512u, ((carry / max(2u * walk_ctas, 1u)) + 63u) & ~63u));
  }
  for (int pass = 0; pass < 5; ++pass)
  {
    const uint32_t hi = (pass == 0) ? 0xFFFFFFFFu : (item_max >> (pass == 1 ? 0 : (pass == 2 ? 1 : (pass == 3 ? 3 : 5))));
    const uint32_t lo = (pass == 4) ? 1u : (item_max >> (pass == 0 ? 0 : (pass == 1 ? 1 : (pass == 2 ? 3 : 5))));
    for (uint32_t t = threadIdx.x; t < touched; t += blockDim.x)
    {
      const bool cached = t < kPlanCache;
      const uint32_t slot = cached ? s_slot[t] : b.touched_list[t];
      const uint32_t count = cached ? s_count[t] : b.seg_count[slot];
      if (count < lo || count >= hi)
      {
        continue;
      }
      const uint32_t pieces = (count + item_max - 1) / item_max;
      const uint32_t at = atomicAdd(&item_next, pieces);
      const uint32_t begin = cached ? s_offset[t] : b.seg_offset[slot];
      for (uint32_t p = 0; p < pieces && at + p < b.item_capacity; ++p)
      {
        WorkItem w;
        w.slot = slot;
        w.begin = begin + p * item_max;
        w.end = min(begin + count, w.begin + item_max);
        w.shared = pieces > 1;
        b.items[at + p] = w;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    b.counters->item_count = item_next;
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
  const uint32_t staged = b.stage_count[i];
  const uint32_t ray = b.stage_by_ray ? i : rayOfThread(i, b.n);
  if (staged <= kStageSegments)
  {
    // Pass B, common case: scatter the segments pass A staged (plane k holds the k-th segment of every ray).
    for (uint32_t k = 0; k < staged; ++k)
    {
      uint4 raw = b.stage[(size_t)k * b.stage_stride + i];
      const uint32_t slot = raw.x;
      if (slot == 0xFFFFFFFFu)
      {
        continue;  // (an empty plane entry; prepSegments stages none)
      }
      const uint32_t at = b.seg_offset[slot] + slotAggregatedInc(b.seg_cursor, slot);
      if (at < b.seg_capacity)
      {
        raw.x = ray;
        reinterpret_cast<uint4 *>(b.segments)[at] = raw;
      }
    }
    return;
  }
  RayRec rec;
  loadRec(rec, b.recs + ray);
  enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
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
      raw.x = ray;
      raw.y = (uint32_t)st[0] | ((uint32_t)st[1] << 16);
      raw.z = (uint32_t)st[2] | ((uint32_t)n << 16);
      raw.w = (uint32_t)entry[0] | ((uint32_t)entry[1] << 8) | ((uint32_t)entry[2] << 16);
      reinterpret_cast<uint4 *>(b.segments)[at] = raw;
    }
  });
}

#ifndef OHMB200_WALK_THREADS
#define OHMB200_WALK_THREADS 512
#define OHMB200_WALK_CTAS 2
#endif
constexpr int kWalkThreads = OHMB200_WALK_THREADS;  // x kWalkCtasPerSm: the registers of an SM at 64 per thread
constexpr int kWalkCtasPerSm = OHMB200_WALK_CTAS;
constexpr uint32_t kLengthBins = 128;  // visits per segment <= 3 * 255; 127+ share a bin

// dm.voxel_bits: one persistent bit per voxel, [capacity][(vpr + 31) / 32] words (NDT: the voxel has an established
// Gaussian; TSDF: the stored state is not order-free).
__device__ __forceinline__ void setVoxelBit(const DeviceMap &dm, const Geom &g, uint32_t vid, bool value)
{
  const uint32_t slot = vid / g.vpr;
  const uint32_t local = vid - slot * g.vpr;
  uint32_t *word = dm.voxel_bits + (size_t)slot * ((g.vpr + 31u) >> 5) + (local >> 5);
  const uint32_t bit = 1u << (local & 31u);
  if (value)
  {
    if (!(*word & bit))
    {
      atomicOr(word, bit);
    }
  }
  else if (*word & bit)
  {
    atomicAnd(word, ~bit);
  }
}

// Slot for one ordered record, for every lane that calls this together.  Slots come from a per-warp chunk
// ((base << 32) | used, in shared memory): one global atomic per kRecordChunk records.
__device__ __forceinline__ uint32_t reserveRecord(unsigned long long *warp_chunk, uint32_t *global_count)
{
  const unsigned group = __activemask();
  const uint32_t n = (uint32_t)__popc(group);
  const uint32_t rank = (uint32_t)__popc(group & ((1u << (threadIdx.x & 31u)) - 1u));
  uint32_t at = 0;
  if (rank == 0)
  {
    const unsigned long long state = atomicAdd(warp_chunk, (unsigned long long)n);
    const uint32_t used = (uint32_t)state;
    if (used + n <= kRecordChunk)
    {
      at = (uint32_t)(state >> 32) + used;
    }
    else
    {
      at = atomicAdd(global_count, kRecordChunk);
      atomicExch(warp_chunk, ((unsigned long long)at << 32) | n);
    }
  }
  return __shfl_sync(group, at, __ffs(group) - 1) + rank;
}

// ---- bulk asynchronous copies (TMA unit, cp.async.bulk + mbarrier) -------------------------------------------------
// The segments of a work item are one contiguous range of 16-byte records (<= 32 KB): ONE cp.async.bulk brings them
// into shared memory while the CTA is still folding the item before — issued by one thread, completion counted in
// bytes on an mbarrier the whole CTA waits on at the top of the next item.  queueBuild's two passes over the segments
// and every queuePop then read shared memory instead of going to L2 / HBM behind a dependent index.
__device__ __forceinline__ uint32_t smemAddress(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbarInit(unsigned long long *mbar, uint32_t arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddress(mbar)), "r"(arrivals) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the async proxy must see the initialised barrier
}

// One thread: expect `bytes` on the barrier and start the global -> shared bulk copy that delivers them.
__device__ __forceinline__ void bulkLoad(void *smem_dst, const void *global_src, uint32_t bytes, unsigned long long *mbar)
{
  // everything the CTA read from / wrote to smem_dst through the generic proxy is ordered before the copy's writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddress(mbar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smemAddress(smem_dst)),
               "l"(global_src), "r"(bytes), "r"(smemAddress(mbar))
               : "memory");
}

// Every thread that reads the copied bytes: wait for the barrier's phase `parity` to complete.
__device__ __forceinline__ void mbarWait(unsigned long long *mbar, uint32_t parity)
{
  const uint32_t addr = smemAddress(mbar);
  uint32_t done = 0;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
  } while (!done);
}

// The segments of a work item, ordered by decreasing visit count and handed to warps 32 at a time: the lanes of a warp
// walk segments of (nearly) equal length, long segments start first, and a warp that finishes early takes more.
struct SegmentQueue
{
  uint4 staged[kMaxSegmentsPerItem];  // the item's segments (bulk copy, see above)
  unsigned long long mbar;            // its completion barrier
  uint16_t order[kMaxSegmentsPerItem];
  uint32_t length_bins[kLengthBins];
  uint32_t next_chunk;
};

// One thread, once per kernel.
__device__ __forceinline__ void queueInit(SegmentQueue &q)
{
  mbarInit(&q.mbar, 1u);
}

// One thread: start bringing the segments of `item` into q.staged.  The previous item's pops are behind a barrier.
__device__ __forceinline__ void queueStage(SegmentQueue &q, const Batch &b, const WorkItem &item)
{
  if (item.slot != 0xFFFFFFFFu)
  {
    bulkLoad(q.staged, b.segments + item.begin, (item.end - item.begin) * (uint32_t)sizeof(Segment), &q.mbar);
  }
}

// Every thread of the CTA calls this (it synchronises); `stage_parity` = number of items this CTA has staged before
// this one, mod 2 (the phase of the copy's barrier).
__device__ __forceinline__ void queueBuild(SegmentQueue &q, const Batch &b, const WorkItem &item, uint32_t stage_parity)
{
  const uint32_t tid = threadIdx.x;
  if (tid < kLengthBins)
  {
    q.length_bins[tid] = 0;
  }
  if (tid == 0)
  {
    q.next_chunk = 0;
  }
  mbarWait(&q.mbar, stage_parity);  // the item's segments are in shared memory
  __syncthreads();
  const uint32_t n_segments = item.end - item.begin;
  const uint32_t *segment_words = reinterpret_cast<const uint32_t *>(q.staged);
  for (uint32_t k = tid; k < n_segments; k += blockDim.x)
  {
    const uint32_t visits = segment_words[4 * k + 2] >> 16;
    atomicAdd(&q.length_bins[kLengthBins - 1u - min(visits, kLengthBins - 1u)], 1u);
  }
  __syncthreads();
  if (tid < 32)
  {
    // exclusive scan of the 128 bins by one warp (4 bins per lane)
    uint32_t c[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      c[k] = q.length_bins[tid * 4 + k];
      sum += c[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      incl += (tid >= (uint32_t)d) ? up : 0u;
    }
    uint32_t base = incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      q.length_bins[tid * 4 + k] = base;
      base += c[k];
    }
  }
  __syncthreads();
  for (uint32_t k = tid; k < n_segments; k += blockDim.x)
  {
    const uint32_t visits = segment_words[4 * k + 2] >> 16;
    q.order[atomicAdd(&q.length_bins[kLengthBins - 1u - min(visits, kLengthBins - 1u)], 1u)] = (uint16_t)k;
  }
  __syncthreads();
}

// Warp-collective.  0: the queue is empty; 1: `raw` holds this lane's segment; 2: this lane has none this round.
__device__ __forceinline__ int queuePop(SegmentQueue &q, const Batch &b, const WorkItem &item, uint4 &raw)
{
  uint32_t k = 0;
  if ((threadIdx.x & 31u) == 0)
  {
    k = atomicAdd(&q.next_chunk, 32u);
  }
  k = __shfl_sync(0xffffffffu, k, 0);
  const uint32_t n_segments = item.end - item.begin;
  if (k >= n_segments)
  {
    return 0;
  }
  k += threadIdx.x & 31u;
  if (k >= n_segments)
  {
    return 2;
  }
  raw = q.staged[q.order[k]];
  return 1;
}

// What a lane needs to resume a segment: the segment itself and the walk constants of its ray.
struct SegmentWalk
{
  uint32_t ray, flags;
  int st[3], visits, entry[3], total[3], local0[3];
  double init[3], delta[3];
};

__device__ __forceinline__ void loadSegmentWalk(const Batch &b, const uint4 &raw, SegmentWalk &w)
{
  w.ray = raw.x;
  w.st[0] = (int)(raw.y & 0xffffu);
  w.st[1] = (int)(raw.y >> 16);
  w.st[2] = (int)(raw.z & 0xffffu);
  w.visits = (int)(raw.z >> 16);
  w.entry[0] = (int)(raw.w & 0xffu);
  w.entry[1] = (int)((raw.w >> 8) & 0xffu);
  w.entry[2] = (int)((raw.w >> 16) & 0xffu);
  const RayRec *rp = b.recs + w.ray;
  // tail of the record: region[3] i16 | local[3] u8 | flags u8 | total[3] u16
  const uint4 tail = reinterpret_cast<const uint4 *>(rp)[3];
  w.flags = (tail.z >> 8) & 0xffu;
  w.total[0] = (int)(tail.z >> 16);
  w.total[1] = (int)(tail.w & 0xffffu);
  w.total[2] = (int)(tail.w >> 16);
  w.local0[0] = (int)((tail.y >> 16) & 0xffu);
  w.local0[1] = (int)(tail.y >> 24);
  w.local0[2] = (int)(tail.z & 0xffu);
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    w.init[a] = rp->initial[a];
    w.delta[a] = rp->delta[a];
  }
}

// atomicAdd on the 32-bit word of the counter tile that holds the u16 counter at shared-memory byte address `at`.
__device__ __forceinline__ uint32_t tileAdd(uint32_t at, uint32_t one)
{
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(at & ~3u), "r"(one) : "memory");
  return old;
}

// The fold's view of the counter tile: fn(v, position, lo, hi) for every word that holds voxels — v = index in the
// region of the word's first voxel (its second is v + 1), position = counter position of that voxel, lo / hi = the two
// raw counters (hi = 0 where the row ends on an odd voxel).  Words that are zero are skipped.  A warp takes 32 /
// row_lanes rows at a time, lanes along x: the slab is then touched in runs of whole rows (128 B of log-odds per row
// of a 32^3 region).
template <typename Fn>
__device__ __forceinline__ void foldTile(const uint32_t *tile, const TileLayout &tl, Fn &&fn)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t rows_per_warp = 32u / tl.row_lanes;
  const uint32_t x0 = lane & (tl.row_lanes - 1u);
  const uint32_t rows = (uint32_t)(tl.dy * tl.dz);
  const uint32_t stride = (blockDim.x >> 5) * rows_per_warp;
  for (uint32_t r = (threadIdx.x >> 5) * rows_per_warp + lane / tl.row_lanes; r < rows; r += stride)
  {
    const uint32_t z = (tl.dy > 1) ? __umulhi(r, tl.inv_dy) : r;
    const uint32_t y = r - z * (uint32_t)tl.dy;
    const uint32_t first_word = ((uint32_t)tl.row >> 1) * y + ((uint32_t)tl.slab >> 1) * z;
    const uint32_t first_voxel = (uint32_t)tl.dx * y + (uint32_t)tl.dxy * z;
    for (uint32_t xw = x0; xw < tl.row_words; xw += tl.row_lanes)
    {
      const uint32_t w = tile[first_word + xw];
      if (w)
      {
        fn(first_voxel + 2u * xw, 2u * (first_word + xw), w & 0xffffu, (2u * xw + 1u < (uint32_t)tl.dx) ? (w >> 16) : 0u);
      }
    }
  }
}

// Tile position (in counters, a multiple of 8) of group c of a fast layout.
__device__ __forceinline__ uint32_t groupPosition(const TileLayout &tl, uint32_t c)
{
  const uint32_t r = (tl.row_groups > 1) ? __umulhi(c, tl.inv_row_groups) : c;
  const uint32_t z = (tl.dy > 1) ? __umulhi(r, tl.inv_dy) : r;
  return 8u * (c - r * tl.row_groups) + (uint32_t)tl.row * (r - z * (uint32_t)tl.dy) + (uint32_t)tl.slab * z;
}

// Fast layouts: group(c, position, counters) for every aligned group of eight voxels (8c .. 8c+7 of the region, at
// counter position `position`, a multiple of 8) whose four tile words are not all zero.
template <typename Group>
__device__ __forceinline__ void foldGroups(const uint32_t *tile, const TileLayout &tl, Group &&group)
{
  const uint4 *tile4 = reinterpret_cast<const uint4 *>(tile);
  const uint32_t groups = ((uint32_t)tl.dxy * (uint32_t)tl.dz) >> 3;
  for (uint32_t c = threadIdx.x; c < groups; c += blockDim.x)
  {
    const uint32_t position = groupPosition(tl, c);
    const uint4 t = tile4[position >> 3];
    if (t.x | t.y | t.z | t.w)
    {
      group(c, position, t);
    }
  }
}

// The fold of a work item that is the sole writer of its region, fast layouts: kGroups groups of eight voxels per
// thread in flight — tile reads, then the slab reads of all of them, then the updates — so that the L2 round trips
// and the (branch-free) per-voxel updates of different voxels overlap.  Voxel = the slab's voxel type (4 or 8 bytes);
// counts(position, counters, cnt[8]) fills the counts to apply and returns whether any is non-zero;
// apply(voxel &, count) updates one voxel.
template <int kGroups, typename Voxel, typename Counts, typename Apply>
__device__ __forceinline__ void foldGroupsSole(const uint32_t *tile, const TileLayout &tl, Voxel *slab, uint32_t *ticket,
                                               Counts &&counts, Apply &&apply)
{
  constexpr int kChunks = (int)sizeof(Voxel) * 8 / 16;  // 16-byte chunks per group
  const uint4 *tile4 = reinterpret_cast<const uint4 *>(tile);
  const uint32_t groups = ((uint32_t)tl.dxy * (uint32_t)tl.dz) >> 3;
  // Warps take blocks of 32 x kGroups groups from a ticket counter in shared memory (zeroed before the barrier that
  // precedes the fold): the voxels a batch touches sit in a few bands of y or z of the region, and a fixed assignment
  // leaves a quarter of the warps with all the work (measured: the others waited 10 k cycles per item at the barrier).
  const uint32_t lane = threadIdx.x & 31u;
  for (;;)
  {
    uint32_t base = 0;
    if (lane == 0)
    {
      base = atomicAdd(ticket, 1u) * (32u * kGroups);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= groups)
    {
      break;
    }
    const uint32_t c0 = base + lane;
    uint32_t cnt[kGroups][8];
    bool any[kGroups];
    union
    {
      uint4 chunk[kChunks];
      Voxel voxel[8];
    } value[kGroups];
#pragma unroll
    for (int u = 0; u < kGroups; ++u)
    {
      const uint32_t c = c0 + (uint32_t)u * 32u;
      any[u] = false;
      if (c < groups)
      {
        const uint32_t position = groupPosition(tl, c);
        const uint4 t = tile4[position >> 3];
        any[u] = (t.x | t.y | t.z | t.w) != 0 && counts(position, t, cnt[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kGroups; ++u)
    {
      if (any[u])
      {
        const uint4 *src = reinterpret_cast<const uint4 *>(slab + 8u * (c0 + (uint32_t)u * 32u));
#pragma unroll
        for (int k = 0; k < kChunks; ++k)
        {
          value[u].chunk[k] = src[k];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGroups; ++u)
    {
      if (any[u])
      {
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
          apply(value[u].voxel[k], cnt[u][k]);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(slab + 8u * (c0 + (uint32_t)u * 32u));
#pragma unroll
        for (int k = 0; k < kChunks; ++k)
        {
          dst[k] = value[u].chunk[k];
        }
      }
    }
  }
}

// The fold of a work item that shares its region with other items (a hot region is cut into several), fast layouts:
// the same scan, the slab updated by 64-bit compare-and-swap — a unit is two 4-byte voxels or one 8-byte voxel — with
// all the swaps of a group in flight together; a swap that lost against another item's fold (rare) is repeated on
// the value it returned.
template <typename Voxel, typename Counts, typename Apply>
__device__ __forceinline__ void foldGroupsShared(const uint32_t *tile, const TileLayout &tl, Voxel *slab, uint32_t *ticket,
                                                 Counts &&counts, Apply &&apply)
{
  constexpr int kUnits = (int)sizeof(Voxel);  // 64-bit units per group of eight voxels
  constexpr int kPerUnit = 8 / kUnits;        // voxels per unit
  const uint4 *tile4 = reinterpret_cast<const uint4 *>(tile);
  const uint32_t groups = ((uint32_t)tl.dxy * (uint32_t)tl.dz) >> 3;
  const uint32_t lane = threadIdx.x & 31u;
  for (;;)
  {
    uint32_t base = 0;
    if (lane == 0)
    {
      base = atomicAdd(ticket, 1u) * 32u;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= groups)
    {
      break;
    }
    const uint32_t c = base + lane;
    uint32_t cnt[8];
    bool any = false;
    if (c < groups)
    {
      const uint32_t position = groupPosition(tl, c);
      const uint4 t = tile4[position >> 3];
      any = (t.x | t.y | t.z | t.w) != 0 && counts(position, t, cnt);
    }
    if (!any)
    {
      continue;
    }
    unsigned long long *units = reinterpret_cast<unsigned long long *>(slab + 8u * c);
    union
    {
      unsigned long long unit[kUnits];
      Voxel voxel[8];
    } old, now;
#pragma unroll
    for (int k = 0; k < kUnits; ++k)
    {
      const uint32_t wanted = (kPerUnit == 2) ? (cnt[2 * k] | cnt[2 * k + 1]) : cnt[k];
      old.unit[k] = wanted ? __ldcg(units + k) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kUnits; ++k)
    {
      now.unit[k] = old.unit[k];
    }
#pragma unroll
    for (int v = 0; v < 8; ++v)
    {
      apply(now.voxel[v], cnt[v]);
    }
    uint32_t lost = 0;
#pragma unroll
    for (int k = 0; k < kUnits; ++k)
    {
      const uint32_t wanted = (kPerUnit == 2) ? (cnt[2 * k] | cnt[2 * k + 1]) : cnt[k];
      if (wanted && now.unit[k] != old.unit[k])
      {
        const unsigned long long got = atomicCAS(units + k, old.unit[k], now.unit[k]);
        lost |= (got != old.unit[k]) ? (1u << k) : 0u;
        old.unit[k] = got;
      }
    }
    if (lost)
    {
#pragma unroll
      for (int k = 0; k < kUnits; ++k)
      {
        while (lost & (1u << k))
        {
          now.unit[k] = old.unit[k];
#pragma unroll
          for (int j = 0; j < kPerUnit; ++j)
          {
            apply(now.voxel[kPerUnit * k + j], cnt[kPerUnit * k + j]);
          }
          if (now.unit[k] == old.unit[k])
          {
            break;
          }
          const unsigned long long got = atomicCAS(units + k, old.unit[k], now.unit[k]);
          if (got == old.unit[k])
          {
            break;
          }
          old.unit[k] = got;
        }
      }
    }
  }
}

// The update of one 64-bit unit of a slab by a work item that shares its region with other items (a hot region is
// cut into several): compare-and-swap, repeated on the value it returned when another item's fold got in between.
template <typename Update>
__device__ __forceinline__ void foldUnitShared(unsigned long long *unit, unsigned long long seen, Update &&update)
{
  for (;;)
  {
    const unsigned long long want = update(seen);
    if (want == seen)
    {
      return;
    }
    const unsigned long long got = atomicCAS(unit, seen, want);
    if (got == seen)
    {
      return;
    }
    seen = got;
  }
}

// ---- the folds: counter tile -> slab ------------------------------------------------------------------------
// The folds take the tile layout through foldLayout(): values the compiler cannot see through, so that none of the
// folds' address arithmetic is hoisted out of the persistent loop (it would stay live through the walk, whose inner
// loop then spills).
__device__ __forceinline__ uint32_t opaque(uint32_t x)
{
  asm volatile("" : "+r"(x));
  return x;
}
__device__ __forceinline__ TileLayout foldLayout(const TileLayout &tl)
{
  TileLayout f = tl;
  f.dx = (int)opaque((uint32_t)tl.dx);
  f.dy = (int)opaque((uint32_t)tl.dy);
  f.dz = (int)opaque((uint32_t)tl.dz);
  f.dxy = (int)opaque((uint32_t)tl.dxy);
  f.row = (int)opaque((uint32_t)tl.row);
  f.slab = (int)opaque((uint32_t)tl.slab);
  f.row_words = opaque(tl.row_words);
  f.row_lanes = opaque(tl.row_lanes);
  f.row_shift = opaque(tl.row_shift);
  f.inv_dy = opaque(tl.inv_dy);
  return f;
}

// The slowest, general form: one voxel at a time (odd region rows: a tile word's two voxels are no aligned pair).
template <typename Voxel>
__device__ __forceinline__ void foldTileVoxels(const uint32_t *tile, const TileLayout &tl, Voxel &&voxel)
{
  foldTile(tile, tl, [&](uint32_t v, uint32_t position, uint32_t lo, uint32_t hi) {
    voxel(v, position, lo);
    voxel(v + 1u, position + 1u, hi);
  });
}

// One voxel's log-odds, `count` misses, by compare-and-swap (or a plain store for the sole writer of the region).
__device__ __forceinline__ void foldLogOdds(float *occ, uint32_t v, uint32_t count, const MapParams &mp, unsigned ray_flags,
                                            const MissLadder &ladder, bool sole)
{
  int *addr = reinterpret_cast<int *>(occ + v);
  int seen = *reinterpret_cast<volatile int *>(addr);
  for (;;)
  {
    const float next = missRepeatLadder(ladder, __int_as_float(seen), count, mp, ray_flags);
    if (__float_as_int(next) == seen)
    {
      break;
    }
    if (sole)
    {
      occ[v] = next;
      break;
    }
    const int prev = atomicCAS(addr, seen, __float_as_int(next));
    if (prev == seen)
    {
      break;
    }
    seen = prev;
  }
}

// Occupancy maps: k identical misses commute, so the count is all that matters (missRepeat).  `kind` (NDT maps; else
// null) marks the voxels with an established Gaussian: those are handled by ndtGaussianMisses / ndtClampGaussians;
// `hit_miss` (NDT-TM; else null) counts every plain miss.  Flagged voxels are replayed with their samples.
__device__ __forceinline__ void foldLogOddsTile(const uint32_t *tile, const uint32_t *kind, const TileLayout &layout,
                                                float *occ, uint2 *hit_miss, const MapParams &mp, unsigned ray_flags,
                                                const MissLadder &ladder, uint32_t *ticket, uint32_t shared)
{
  const TileLayout tl = foldLayout(layout);
  const auto misses = [&](float v, uint32_t count) { return missRepeatLadder(ladder, v, count, mp, ray_flags); };
  if (!tl.fast)
  {
    foldTileVoxels(tile, tl, [&](uint32_t v, uint32_t position, uint32_t half) {
      const bool gauss = kind && ((kind[position >> 5] >> (position & 31u)) & 1u);
      if (half != 0 && !(half & kTileFlag) && !gauss)
      {
        if (hit_miss)
        {
          atomicAdd(&hit_miss[v].y, half);  // every plain NDT miss counts as a miss
        }
        foldLogOdds(occ, v, half, mp, ray_flags, ladder, !shared);
      }
    });
    return;
  }
  // counts of the plain voxels of a group: flagged voxels are replayed with their samples, Gaussian voxels (NDT) are
  // handled by ndtGaussianMisses / ndtClampGaussians; every plain NDT miss counts as a miss (hit_miss, NDT-TM)
  const auto counts = [&](uint32_t position, const uint4 &t, uint32_t cnt[8]) {
    const uint32_t w4[4] = { t.x, t.y, t.z, t.w };
    // position is a multiple of 8: the eight Gaussian bits are one byte of `kind`
    const uint32_t gauss8 = kind ? reinterpret_cast<const uint8_t *>(kind)[position >> 3] : 0u;
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const uint32_t half = (w4[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
      cnt[k] = ((half & kTileFlag) || ((gauss8 >> k) & 1u)) ? 0u : half;
      any |= cnt[k];
    }
    return any != 0;
  };
  const auto log_odds_after = [&](float &v, uint32_t count) {
    bool ok;
    const float after = missLadderLookup(ladder, v, count, ok);
    v = ok ? after : missRepeat(v, count, mp, ray_flags);
  };
  if (hit_miss)
  {
    foldGroups(tile, tl, [&](uint32_t c, uint32_t position, const uint4 &t) {
      uint32_t cnt[8];
      counts(position, t, cnt);
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        if (cnt[k])
        {
          atomicAdd(&hit_miss[c * 8u + (uint32_t)k].y, cnt[k]);
        }
      }
    });
  }
  if (!shared)
  {
    // sole writer of the region's log-odds until this kernel ends
    foldGroupsSole<2>(tile, tl, occ, ticket, counts, log_odds_after);
    return;
  }
  foldGroupsShared(tile, tl, occ, ticket, counts, log_odds_after);
}

// Work item w of the batch, or the "no more work" marker.
__device__ __forceinline__ void loadWorkItem(const Batch &b, uint32_t w, WorkItem *dst)
{
  if (w < min(b.counters->item_count, b.item_capacity))
  {
    *dst = b.items[w];
  }
  else
  {
    dst->slot = 0xFFFFFFFFu;
  }
}

// Persistent CTAs: one (region, segment range) work item at a time against a shared-memory counter tile.
template <bool kTraversal>  // kTraversal: also accumulate the traversal layer (enter / exit ranges of every visit)
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) walkRegions(const __grid_constant__ DeviceMap dm, const __grid_constant__ Geom g,
                                                               const __grid_constant__ MapParams mp, const __grid_constant__ Batch b,
                                                               const __grid_constant__ TileLayout tl, int has_samples)
{
  extern __shared__ uint32_t tile[];
  __shared__ WorkItem items2[2];  // this work item and the next one (fetched during the walk)
  uint32_t parity = 1;
  // per-warp reservation of ordered-miss record slots: (base << 32) | used
  __shared__ unsigned long long record_chunk[kWalkThreads / 32];
  __shared__ SegmentQueue queue;
  __shared__ MissLadder ladder;
  __shared__ uint32_t fold_ticket;  // next block of the fold (foldGroupsSole)
  const uint32_t words = tl.words;
  const uint32_t tile_base = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  if ((tid & 31u) == 0)
  {
    record_chunk[warp] = (unsigned long long)kRecordChunk;  // "full": the first record reserves a chunk
  }
  if (tid == 32)
  {
    buildMissLadder(ladder, mp, b.ray_flags);
  }

  if (tid == 0)
  {
    loadWorkItem(b, atomicAdd(&b.counters->work_next, 1u), &items2[0]);
    queueInit(queue);
    queueStage(queue, b, items2[0]);
  }
#ifdef OHMB200_PHASE_CLOCKS
  long long ph[7] = { 0, 0, 0, 0, 0, 0, 0 }, tc[7];

  unsigned ph_items = 0, ph_shared = 0;
  tc[6] = clock64();
#define PHASE(i) tc[i] = clock64()
#else
#define PHASE(i)
#endif
  for (;;)
  {
    parity ^= 1u;
    __syncthreads();  // the item is in place; the previous fold is done with the tile
    const WorkItem &item = items2[parity];
    if (item.slot == 0xFFFFFFFFu)
    {
#ifdef OHMB200_PHASE_CLOCKS
      if (tid == 0)
      {
        printf("PH %u items %u shared %u top %lld zero %lld build %lld walk %lld wait %lld fold %lld foldshared %lld\n", blockIdx.x,
               ph_items, ph_shared, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6]);
      }
#endif
      return;
    }
    PHASE(0);
    const uint32_t slot = item.slot;
    const uint32_t vbase = slot * g.vpr;
    if ((words & 3u) == 0)
    {
      uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
      for (uint32_t w = tid; w < (words >> 2); w += blockDim.x)
      {
        tile4[w] = make_uint4(0, 0, 0, 0);
      }
    }
    else
    {
      for (uint32_t w = tid; w < words; w += blockDim.x)
      {
        tile[w] = 0;
      }
    }
    __syncthreads();
    PHASE(1);
    if (has_samples)
    {
      // Voxels that also receive samples in this batch: their misses must stay ordered against the hits.
      // (the region's range of the sorted sample pairs comes from markRuns)
      const uint32_t sample_end = b.sample_end[slot];
      for (uint32_t s = b.sample_begin[slot] + tid; s < sample_end; s += blockDim.x)
      {
        const uint32_t half = tileHalf(tl, b.keys_out[s] - vbase);
        atomicOr(&tile[half >> 1], kTileFlag << ((half & 1u) * 16u));
      }
    }
    queueBuild(queue, b, item, parity);  // (the copy barrier's phase flips with the item slot)
    PHASE(2);
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
        auto count_visit = [&](uint32_t offset, uint32_t one) {
          if (tileAdd(offset, one) & (one << 15))
          {
            const uint32_t at = reserveRecord(&record_chunk[warp], &b.counters->record_count);
            if (at < b.record_capacity)
            {
              b.record_vid[at] = vbase + tileVoxel(tl, (offset - tile_base) >> 1);
              b.record_ray[at] = ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              atomicOr(&b.counters->overflow_seen, 1);
            }
          }
        };
        if (kTraversal)
        {
          const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];
          resumeSegment<true>(sw.init, sw.delta, sw.local0, sw.total, sw.flags, sw.st, sw.visits, b.ray_length[ray], g,
                              [&](const int l[3], double t_enter, double t_exit, bool last_of_ray) {
                                const uint32_t idx = (uint32_t)(l[0] + l[1] * dx + l[2] * dxy);
                                count_visit(tile_base + 2u * (uint32_t)(l[0] + l[1] * tl.row + l[2] * tl.slab),
                                            (l[0] & 1) ? 0x10000u : 1u);
                                atomicAdd(&dm.traversal[vbase + idx], (float)(t_exit - t_enter));
                              });
        }
        else
        {
          resumeSegmentTile(sw.init, sw.delta, sw.entry, sw.total, sw.flags, sw.st, sw.visits, tl, tile_base, count_visit);
        }
      }
      __syncwarp();  // every lane of the warp is back together before the next pop
    }
    PHASE(3);
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
    PHASE(4);

    // Fold the miss counts into the occupancy slab.  k identical misses commute, so the count is all that matters.
    foldLogOddsTile(tile, nullptr, tl, dm.occupancy + (size_t)vbase, nullptr, mp, b.ray_flags, ladder, &fold_ticket, item.shared);
#ifdef OHMB200_PHASE_CLOCKS
    PHASE(5);
    ph[0] += tc[0] - tc[6];
    ph[1] += tc[1] - tc[0];
    ph[2] += tc[2] - tc[1];
    ph[3] += tc[3] - tc[2];
    ph[4] += tc[4] - tc[3];
    ph[item.shared ? 6 : 5] += tc[5] - tc[4];
    tc[6] = tc[5];
    ++ph_items;
    ph_shared += item.shared ? 1u : 0u;
#endif
  }
}

// ---- the slab-staged walk (round 2): occupancy maps without a traversal layer ---------------------------------------
// One 1024-thread CTA per SM instead of two of 512: the shared memory that frees holds the region's WHOLE occupancy slab
// (128 KB for a 32^3 region) next to the counter tile.  The slab is brought in by ONE cp.async.bulk (TMA unit) issued at
// the top of the work item — it lands while the segments are walked, 10 us later — the fold then reads and writes shared
// memory only (no L2 round trip per group of voxels: the long_scoreboard stalls of the tile kernel's fold), and ONE bulk
// store writes the slab back while the next item is zeroing its tile.  Work items that share their region with other
// items (hot regions cut in several) cannot stage it — their folds merge through compare-and-swap on global memory — and
// keep the tile kernel's shared fold.  Same walk, same counters, same ladder: results are bit-identical (the parity
// suite passes with it).  MEASURED AND NOT ADOPTED: 278 us against the tile kernel's 183 us on config 2 (profiles/
// walk_r2_ab.md) — see ohmb200_create.  Kept behind OHMB200_WALK=slab as the record of the experiment.
constexpr int kSlabThreads = 1024;

struct PlainQueue  // SegmentQueue without the staged segments (the shared memory goes to the slab)
{
  uint16_t order[kMaxSegmentsPerItem];
  uint32_t length_bins[kLengthBins];
  uint32_t next_chunk;
};

__device__ __forceinline__ void plainQueueBuild(PlainQueue &q, const Batch &b, const WorkItem &item)
{
  const uint32_t tid = threadIdx.x;
  if (tid < kLengthBins)
  {
    q.length_bins[tid] = 0;
  }
  if (tid == 0)
  {
    q.next_chunk = 0;
  }
  __syncthreads();
  const uint32_t n_segments = item.end - item.begin;
  const uint32_t *segment_words = reinterpret_cast<const uint32_t *>(b.segments + item.begin);
  for (uint32_t k = tid; k < n_segments; k += blockDim.x)
  {
    const uint32_t visits = segment_words[4 * k + 2] >> 16;
    atomicAdd(&q.length_bins[kLengthBins - 1u - min(visits, kLengthBins - 1u)], 1u);
  }
  __syncthreads();
  if (tid < 32)
  {
    uint32_t c[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      c[k] = q.length_bins[tid * 4 + k];
      sum += c[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      incl += (tid >= (uint32_t)d) ? up : 0u;
    }
    uint32_t base = incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      q.length_bins[tid * 4 + k] = base;
      base += c[k];
    }
  }
  __syncthreads();
  for (uint32_t k = tid; k < n_segments; k += blockDim.x)
  {
    const uint32_t visits = segment_words[4 * k + 2] >> 16;
    q.order[atomicAdd(&q.length_bins[kLengthBins - 1u - min(visits, kLengthBins - 1u)], 1u)] = (uint16_t)k;
  }
  __syncthreads();
}

__device__ __forceinline__ int plainQueuePop(PlainQueue &q, const Batch &b, const WorkItem &item, uint4 &raw)
{
  uint32_t k = 0;
  if ((threadIdx.x & 31u) == 0)
  {
    k = atomicAdd(&q.next_chunk, 32u);
  }
  k = __shfl_sync(0xffffffffu, k, 0);
  const uint32_t n_segments = item.end - item.begin;
  if (k >= n_segments)
  {
    return 0;
  }
  k += threadIdx.x & 31u;
  if (k >= n_segments)
  {
    return 2;
  }
  raw = reinterpret_cast<const uint4 *>(b.segments)[item.begin + q.order[k]];
  return 1;
}

// One thread: shared -> global bulk store of `bytes` (a bulk async-group of its own), and the wait for the stores
// issued so far to have READ their shared-memory source (the buffer may be overwritten) / to have completed.
__device__ __forceinline__ void bulkStore(void *global_dst, const void *smem_src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(global_dst), "r"(smemAddress(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulkStoreWaitRead()
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulkStoreWaitAll()
{
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(kSlabThreads, 1) walkRegionsSlab(const __grid_constant__ DeviceMap dm, const __grid_constant__ Geom g,
                                                             const __grid_constant__ MapParams mp, const __grid_constant__ Batch b,
                                                             const __grid_constant__ TileLayout tl, int has_samples)
{
  extern __shared__ __align__(128) uint32_t tile[];  // the counter tile, then the region's occupancy slab
  float *slab = reinterpret_cast<float *>(tile + tl.words);
  __shared__ WorkItem items2[2];
  __shared__ unsigned long long record_chunk[kSlabThreads / 32];
  __shared__ PlainQueue queue;
  __shared__ MissLadder ladder;
  __shared__ uint32_t fold_ticket;
  __shared__ unsigned long long slab_mbar;
  uint32_t parity = 1;
  uint32_t slab_phase = 0;  // phase of slab_mbar the next staged slab completes
  const uint32_t words = tl.words;
  const uint32_t tile_base = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t slab_bytes = g.vpr * (uint32_t)sizeof(float);
  if ((tid & 31u) == 0)
  {
    record_chunk[warp] = (unsigned long long)kRecordChunk;
  }
  if (tid == 32)
  {
    buildMissLadder(ladder, mp, b.ray_flags);
  }
  if (tid == 0)
  {
    mbarInit(&slab_mbar, 1u);
    loadWorkItem(b, atomicAdd(&b.counters->work_next, 1u), &items2[0]);
  }
  for (;;)
  {
    parity ^= 1u;
    __syncthreads();  // the item is in place; the previous fold is done with the tile and the slab
    const WorkItem &item = items2[parity];
    if (item.slot == 0xFFFFFFFFu)
    {
      if (tid == 0)
      {
        bulkStoreWaitAll();  // the last slab has reached global memory
      }
      return;
    }
    const uint32_t slot = item.slot;
    const uint32_t vbase = slot * g.vpr;
    const bool sole = item.shared == 0;
    if (tid == 0 && sole)
    {
      bulkStoreWaitRead();  // the previous item's write-back has read the slab buffer
      bulkLoad(slab, dm.occupancy + (size_t)vbase, slab_bytes, &slab_mbar);  // lands while the segments are walked
    }
    {
      uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
      for (uint32_t w = tid; w < (words >> 2); w += blockDim.x)
      {
        tile4[w] = make_uint4(0, 0, 0, 0);
      }
    }
    __syncthreads();
    if (has_samples)
    {
      const uint32_t sample_end = b.sample_end[slot];
      for (uint32_t s = b.sample_begin[slot] + tid; s < sample_end; s += blockDim.x)
      {
        const uint32_t half = tileHalf(tl, b.keys_out[s] - vbase);
        atomicOr(&tile[half >> 1], kTileFlag << ((half & 1u) * 16u));
      }
    }
    plainQueueBuild(queue, b, item);
    uint32_t next_work = 0;
    if (tid == 0)
    {
      next_work = atomicAdd(&b.counters->work_next, 1u);
    }
    for (;;)
    {
      uint4 raw;
      const int got = plainQueuePop(queue, b, item, raw);
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
              b.record_vid[at] = vbase + tileVoxel(tl, (offset - tile_base) >> 1);
              b.record_ray[at] = ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              atomicOr(&b.counters->overflow_seen, 1);
            }
          }
        });
      }
      __syncwarp();
    }
    if (tid == 0)
    {
      loadWorkItem(b, next_work, &items2[parity ^ 1u]);
      fold_ticket = 0;
    }
    __syncthreads();
    if (sole)
    {
      mbarWait(&slab_mbar, slab_phase);  // (long since: the copy was issued a whole walk ago)
      slab_phase ^= 1u;
      foldLogOddsTile(tile, nullptr, tl, slab, nullptr, mp, b.ray_flags, ladder, &fold_ticket, 0u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the fold's writes, before the bulk store reads them
      __syncthreads();
      if (tid == 0)
      {
        bulkStore(dm.occupancy + (size_t)vbase, slab, slab_bytes);
      }
    }
    else
    {
      foldLogOddsTile(tile, nullptr, tl, dm.occupancy + (size_t)vbase, nullptr, mp, b.ray_flags, ladder, &fold_ticket, 1u);
    }
  }
}

// Attach every ordered-miss record to the interval between two hits of its voxel: the run of the voxel in the sorted
// sample pairs (binary search by voxel), then the first hit of the run with a larger ray index (binary search by ray).
// Only the NUMBER of misses per interval is kept here: interval_count[head + j] for the misses before hit j of the
// run, tail_overflow[head] (= interval_count[n + head], the two arrays are contiguous) for the misses after the last hit.
// Occupancy needs no more (misses of an interval commute).  NDT (keep_keys != 0) also leaves each record's interval
// index in record_vid[], for scatterRecords to group the records by interval (a counting sort: these counts, one
// exclusive scan, one scatter).  Slots reserved by a warp but never written keep the kInvalidVoxel fill and are skipped.
__global__ void linkRecords(Batch b, int keep_keys, uint32_t vpr)
{
  const uint32_t count = min(b.counters->record_count, b.record_capacity);
  unsigned linked = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x)
  {
    const uint32_t vid = b.record_vid[r];
    if (vid == kInvalidVoxel)
    {
      continue;
    }
    // binary searches confined to the region's range of the sorted pairs (markRuns): ~7 steps instead of 17
    const uint32_t slot = vid / vpr;
    const uint32_t first = b.sample_begin[slot];
    const uint32_t count_in_region = b.sample_end[slot] - first;
    const uint32_t head = first + lowerBound(b.keys_out + first, count_in_region, vid);
    const uint32_t k = first + lowerBound(b.keys_out + first, count_in_region, vid + 1u) - head;
    const uint32_t ray = b.record_ray[r];
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
    const uint32_t key = (lo < k) ? head + lo : b.n + head;
    atomicAdd(&b.interval_count[key], 1u);
    if (keep_keys)
    {
      b.record_vid[r] = key;
    }
    ++linked;
  }
  linked = __reduce_add_sync(0xffffffffu, linked);
  if ((threadIdx.x & 31u) == 0 && linked)
  {
    atomicAdd(&b.counters->ordered_records, (unsigned long long)linked);
  }
}

// NDT: second half of the counting sort.  interval_offset[] is the exclusive scan of the 2n + 1 interval counts;
// sorted_rays[interval_offset[key] ...] receives the rays of the records of interval `key` (any order inside).
__global__ void scatterRecords(Batch b)
{
  const uint32_t count = min(b.counters->record_count, b.record_capacity);
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x)
  {
    const uint32_t key = b.record_vid[r];
    if (key == kInvalidVoxel)
    {
      continue;
    }
    const uint32_t left = atomicSub(&b.interval_count[key], 1u);  // counts down to 0: no separate cursor array
    b.sorted_rays[b.interval_offset[key] + left - 1u] = b.record_ray[r];
  }
}

// ---------------------------------------------------------------------------------------------------------
// NDT (GpuNdtMap, NdtMode::kOccupancy): RayMapperNdt.cpp:84-407
// ---------------------------------------------------------------------------------------------------------

// Visits to voxels with an established Gaussian (mean count >= sample threshold, no sample in this batch) are not
// evaluated inside the walk — a lane doing ~600 fp64 instructions would stall its whole warp — but recorded as
// (voxel << 32 | ray) and evaluated one thread per record: ndtGaussianMisses adds the NDT term to the occupancy slab,
// ndtClampGaussians then applies occupancyAdjustDown's clamp.  Such a voxel's mean/covariance are constant for the
// whole batch, the terms are <= 0, and sum-then-clamp equals the sequential clamp-each-time.
__global__ void __launch_bounds__(128) ndtGaussianMisses(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t count = min(b.counters->gauss_count, b.gauss_capacity);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    const unsigned long long key = b.gauss_keys[i];
    const uint32_t vid = (uint32_t)(key >> 32);
    if (vid == kInvalidVoxel)
    {
      continue;
    }
    const uint32_t ray = (uint32_t)key;
    double sensor[3], sample[3];
    loadRay(b, ray, sensor, sample);
    unsigned filter_flags = 0;
    applyRayFilter(mp, sensor, sample, filter_flags);  // the mapper hands calculateMissNdt the filtered points
    const uint32_t slot = vid / g.vpr;
    const uint32_t idx = vid - slot * g.vpr;
    int r[3];
    unpackRegion(dm.keys[slot], r);
    const int l[3] = { (int)(idx % (uint32_t)g.dim[0]), (int)((idx / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]),
                       (int)(idx / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1])) };
    const uint2 vm = dm.mean[vid];
    double mean[3];
    subVoxelToLocal(vm.x, g.res, mean);
    float cov[6];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      mean[a] += voxelCentreAxis(g, r[a], l[a], a);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k)
    {
      cov[k] = dm.covariance[(size_t)vid * 6 + k];
    }
    bool valid, is_miss;
    const float adj = ndtMissAdjustment(cov, sensor, sample, mean, mp.adaptation_rate, mp.sensor_noise, valid, is_miss);
    if (valid && adj != 0.0f)
    {
      atomicAdd(&dm.occupancy[vid], adj);
    }
    if (dm.hit_miss && is_miss)
    {
      atomicAdd(&dm.hit_miss[vid].y, 1u);  // NDT-TM miss count (RayMapperNdt.cpp:203-209)
    }
  }
}

__global__ void __launch_bounds__(256) ndtClampGaussians(DeviceMap dm, MapParams mp, Batch b)
{
  const uint32_t count = min(b.counters->gauss_count, b.gauss_capacity);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    const uint32_t vid = (uint32_t)(b.gauss_keys[i] >> 32);
    if (vid != kInvalidVoxel)
    {
      const float v = dm.occupancy[vid];
      if (v < mp.min_value)
      {
        dm.occupancy[vid] = mp.min_value;  // every writer of this voxel stores the same value
      }
    }
  }
}

// walkRegions for NDT maps: same counter tile, plus a bit per voxel saying "established Gaussian" (mean count >=
// sample threshold), staged from the mean layer when the work item starts.
template <bool kTraversal>
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) walkRegionsNdt(const __grid_constant__ DeviceMap dm, const __grid_constant__ Geom g,
                                                                  const __grid_constant__ MapParams mp, const __grid_constant__ Batch b,
                                                                  const __grid_constant__ TileLayout tl)
{
  extern __shared__ uint32_t tile[];
  __shared__ WorkItem items2[2];  // this work item and the next one (fetched during the walk)
  uint32_t parity = 1;
  __shared__ unsigned long long record_chunk[kWalkThreads / 32];
  __shared__ unsigned long long gauss_chunk[kWalkThreads / 32];
  __shared__ SegmentQueue queue;
  __shared__ MissLadder ladder;
  __shared__ uint32_t fold_ticket;  // next block of the fold (foldGroupsSole)
  const uint32_t words = tl.words;
  const uint32_t tile_base = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t bit_words = (g.vpr + 31u) >> 5;  // the persistent bits of a region, indexed by voxel
  const uint32_t kind_words = words >> 4;         // their copy in shared memory, indexed by counter position
  uint32_t *kind = tile + ((words + 3u) & ~3u);
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  if ((tid & 31u) == 0)
  {
    record_chunk[warp] = (unsigned long long)kRecordChunk;
    gauss_chunk[warp] = (unsigned long long)kRecordChunk;
  }
  if (tid == 32)
  {
    buildMissLadder(ladder, mp, 0u);  // RayMapperNdt applies no exclusion flags
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
      for (uint32_t w = tid; w < (words >> 2); w += blockDim.x)
      {
        tile4[w] = make_uint4(0, 0, 0, 0);
      }
    }
    {
      // established Gaussians: the persistent bit per voxel (mean count >= sample threshold), laid out like the tile
      const uint32_t fill = (mp.sample_threshold == 0) ? 0xFFFFFFFFu : 0u;
      for (uint32_t w = tid; w < kind_words; w += blockDim.x)
      {
        kind[w] = fill;
      }
    }
    __syncthreads();
    if (mp.sample_threshold != 0 && tl.fast)
    {
      // eight voxels = one byte in both layouts
      const uint8_t *bits8 = reinterpret_cast<const uint8_t *>(dm.voxel_bits + (size_t)slot * bit_words);
      uint8_t *kind8 = reinterpret_cast<uint8_t *>(kind);
      for (uint32_t c = tid; c < (g.vpr >> 3); c += blockDim.x)
      {
        kind8[groupPosition(tl, c) >> 3] = bits8[c];
      }
    }
    else if (mp.sample_threshold != 0)
    {
      const uint32_t *bits = dm.voxel_bits + (size_t)slot * bit_words;
      for (uint32_t w = tid; w < bit_words; w += blockDim.x)
      {
        uint32_t set = bits[w];
        while (set)
        {
          const uint32_t half = tileHalf(tl, (w << 5) + (uint32_t)__ffs(set) - 1u);
          set &= set - 1u;
          atomicOr(&kind[half >> 5], 1u << (half & 31u));
        }
      }
    }
    const uint32_t sample_end = b.sample_end[slot];
    for (uint32_t s = b.sample_begin[slot] + tid; s < sample_end; s += blockDim.x)
    {
      const uint32_t half = tileHalf(tl, b.keys_out[s] - vbase);
      atomicOr(&tile[half >> 1], kTileFlag << ((half & 1u) * 16u));
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
        auto count_visit = [&](uint32_t offset, uint32_t one) {
          if (tileAdd(offset, one) & (one << 15))
          {
            const uint32_t at = reserveRecord(&record_chunk[warp], &b.counters->record_count);
            if (at < b.record_capacity)
            {
              b.record_vid[at] = vbase + tileVoxel(tl, (offset - tile_base) >> 1);
              b.record_ray[at] = ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              atomicOr(&b.counters->overflow_seen, 1);
            }
          }
          else if ((kind[(offset - tile_base) >> 6] >> (((offset - tile_base) >> 1) & 31u)) & 1u)
          {
            // Established Gaussian: evaluated later, one thread per visit (ndtGaussianMisses).
            const uint32_t at = reserveRecord(&gauss_chunk[warp], &b.counters->gauss_count);
            if (at < b.gauss_capacity)
            {
              b.gauss_keys[at] = ((unsigned long long)(vbase + tileVoxel(tl, (offset - tile_base) >> 1)) << 32) | ray;
            }
            else
            {
              b.counters->record_overflow = 1;
              atomicOr(&b.counters->overflow_seen, 1);
            }
          }
        };
        if (kTraversal)
        {
          const int dx = g.dim[0], dxy = g.dim[0] * g.dim[1];
          resumeSegment<true>(sw.init, sw.delta, sw.local0, sw.total, sw.flags, sw.st, sw.visits, b.ray_length[ray], g,
                              [&](const int l[3], double t_enter, double t_exit, bool last_of_ray) {
                                const uint32_t idx = (uint32_t)(l[0] + l[1] * dx + l[2] * dxy);
                                count_visit(tile_base + 2u * (uint32_t)(l[0] + l[1] * tl.row + l[2] * tl.slab),
                                            (l[0] & 1) ? 0x10000u : 1u);
                                atomicAdd(&dm.traversal[vbase + idx], (float)(t_exit - t_enter));
                              });
        }
        else
        {
          resumeSegmentTile(sw.init, sw.delta, sw.entry, sw.total, sw.flags, sw.st, sw.visits, tl, tile_base, count_visit);
        }
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

    // Fold.  Plain voxels: k identical misses (RayMapperNdt applies no exclusion flags).  Gaussian voxels: the
    // adjustments are already in the slab; apply occupancyAdjustDown's clamp.
    foldLogOddsTile(tile, kind, tl, dm.occupancy + (size_t)vbase, dm.hit_miss ? dm.hit_miss + (size_t)vbase : nullptr, mp, 0u,
                    ladder, &fold_ticket, item.shared);
  }
}

// Sample-voxel updates of RayMapperNdt (RayMapperNdt.cpp:284-404) for the run (voxel) t: hits in ray order; the
// recorded misses of an interval are applied with the voxel state of that moment.  kWarp = false: one thread does it
// all.  kWarp = true: a whole warp executes this with the same t — every lane computes the (identical) hits, the
// recorded misses of an interval are evaluated 32 at a time, lane 0 stores.  Returns the number of samples applied.
template <bool kWarp>
__device__ __forceinline__ unsigned replayNdtRun(const DeviceMap &dm, const Geom &g, const MapParams &mp, const Batch &b,
                                                 uint32_t t)
{
  unsigned samples = 0;
  const uint32_t head = b.run_list[t];
  const uint32_t vid = b.keys_out[head];
  uint32_t k = 1;
  while (head + k < b.n && b.keys_out[head + k] == vid)
  {
    ++k;
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
  uint2 vm = dm.mean[vid];
  float cov[6];
#pragma unroll
  for (int c = 0; c < 6; ++c)
  {
    cov[c] = dm.covariance[(size_t)vid * 6 + c];
  }
  uint32_t incident = dm.incident ? dm.incident[vid] : 0;
  uint32_t touch = 0;
  bool touch_set = false;
  float traversal_add = 0.0f;
  const bool ndt_tm = mp.ndt_tm && dm.hit_miss && dm.intensity;
  uint2 hm = ndt_tm ? dm.hit_miss[vid] : make_uint2(0, 0);
  float2 im = ndt_tm ? dm.intensity[vid] : make_float2(0, 0);
  uint32_t pre_ray = 0;
  double pre_start[3] = { 0, 0, 0 }, pre_end[3] = { 0, 0, 0 };
  if (!kWarp)
  {
    pre_ray = b.vals_out[head];
    loadRay(b, pre_ray, pre_start, pre_end);
  }
  for (uint32_t j = 0; j <= k; ++j)
  {
    // the recorded misses of this interval, grouped by scatterRecords (j == k: the misses after the last hit)
    const uint32_t key = (j < k) ? head + j : b.n + head;
    const uint32_t first = b.interval_offset[key], last = b.interval_offset[key + 1u];
    if (first != last)
    {
      double mean[3];
      subVoxelToLocal(vm.x, g.res, mean);
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        mean[a] += centre[a];
      }
      if (!kWarp)
      {
        for (uint32_t at = first; at != last; ++at)
        {
          double sensor[3], sample[3];
          loadRay(b, b.sorted_rays[at], sensor, sample);
          unsigned filter_flags = 0;
          applyRayFilter(mp, sensor, sample, filter_flags);
          bool is_miss;
          value = ndtMissOnce(value, cov, sensor, sample, mean, vm.y, mp, is_miss);
          hm.y += is_miss ? 1u : 0u;
        }
      }
      else
      {
        // 32 records at a time: every lane evaluates the NDT term of one record (it depends on the voxel's Gaussian
        // and the ray only), then all lanes apply the 32 adjustments to the log-odds in the same order.
        const uint32_t lane = threadIdx.x & 31u;
        const bool gaussian = vm.y >= mp.sample_threshold;
        for (uint32_t base = first; base < last; base += 32u)
        {
          const uint32_t at = base + lane;
          float adj = 0.0f;
          bool valid = false, rec_miss = true;
          if (gaussian && at < last)
          {
            double sensor[3], sample[3];
            loadRay(b, b.sorted_rays[at], sensor, sample);
            unsigned filter_flags = 0;
            applyRayFilter(mp, sensor, sample, filter_flags);
            adj = ndtMissAdjustment(cov, sensor, sample, mean, mp.adaptation_rate, mp.sensor_noise, valid, rec_miss);
          }
          const uint32_t batch = min(32u, last - base);
          for (uint32_t i = 0; i < batch; ++i)
          {
            const float a = __shfl_sync(0xffffffffu, adj, i);
            const bool v = __shfl_sync(0xffffffffu, valid ? 1 : 0, i) != 0;
            const bool m = __shfl_sync(0xffffffffu, rec_miss ? 1 : 0, i) != 0;
            // ndtMissOnce with the NDT term already evaluated
            const float initial = value;
            float adjusted;
            bool is_miss = true;
            if (initial == INFINITY)
            {
              adjusted = mp.miss_value;
            }
            else if (!gaussian)
            {
              adjusted = initial + mp.miss_value;
            }
            else
            {
              adjusted = v ? initial + a : initial;
              is_miss = m;
            }
            value = adjustDown(initial, adjusted, mp);
            hm.y += is_miss ? 1u : 0u;
          }
        }
      }
    }
    if (j == k)
    {
      break;
    }
    // The sample ray of hit j.  Its load latency is kept off the (sequential) hit chain: a warp fetches the rays of
    // 32 hits at once and broadcasts them one by one; a single thread fetches the next ray before working on this one.
    uint32_t ray;
    double start[3], end[3];
    if (kWarp)
    {
      const uint32_t lane = threadIdx.x & 31u;
      if ((j & 31u) == 0)
      {
        pre_ray = (j + lane < k) ? b.vals_out[head + j + lane] : 0u;
        if (j + lane < k)
        {
          loadRay(b, pre_ray, pre_start, pre_end);
        }
      }
      ray = __shfl_sync(0xffffffffu, pre_ray, j & 31u);
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        start[a] = __shfl_sync(0xffffffffu, pre_start[a], j & 31u);
        end[a] = __shfl_sync(0xffffffffu, pre_end[a], j & 31u);
      }
    }
    else
    {
      ray = pre_ray;
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        start[a] = pre_start[a];
        end[a] = pre_end[a];
      }
      if (j + 1 < k)
      {
        pre_ray = b.vals_out[head + j + 1];
        loadRay(b, pre_ray, pre_start, pre_end);
      }
    }
    double mean[3];
    subVoxelToLocal(vm.x, g.res, mean);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      mean[a] += centre[a];
    }
    const float initial = value;
    float adjusted = initial;
    if (ndt_tm)
    {
      ndtHitMissOnHit(cov, adjusted, hm, start, end, mean, vm.y, mp);
      ndtIntensityOnHit(im, adjusted, b.intensities ? b.intensities[ray] : 0.0f, vm.y, mp);
    }
    const bool reset = ndtHit(cov, adjusted, end, mean, vm.y, mp.hit_value, (float)g.res, mp.reinit_threshold,
                              mp.reinit_count);
    value = adjustUp(initial, adjusted, mp);
    vm.y = reset ? 0u : vm.y;
    const double local_pt[3] = { end[0] - centre[0], end[1] - centre[1], end[2] - centre[2] };
    vm.x = subVoxelUpdate(vm.x, vm.y, local_pt, g.res);
    ++vm.y;
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
                                      (float)(start[2] - end[2]), vm.y - 1u);
    }
    ++samples;
  }
  if (!kWarp || (threadIdx.x & 31u) == 0)
  {
    dm.occupancy[vid] = value;
    dm.mean[vid] = vm;
    setVoxelBit(dm, g, vid, vm.y >= mp.sample_threshold);
  #pragma unroll
    for (int c = 0; c < 6; ++c)
    {
      dm.covariance[(size_t)vid * 6 + c] = cov[c];
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
    if (ndt_tm)
    {
      dm.hit_miss[vid] = hm;
      dm.intensity[vid] = im;
    }
  }
  return samples;
}

// A run is "heavy" when it holds b.heavy_run hits + recorded misses: it is replayed by a warp instead of a thread.

__global__ void __launch_bounds__(128) applySamplesNdt(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned samples = 0;
  if (t == 0 && !b.counters->segment_overflow)
  {
    b.counters->sample_voxels += b.counters->run_count;
  }
  if (t < b.counters->run_count && !b.counters->segment_overflow)
  {
    const uint32_t head = b.run_list[t];
    const uint32_t vid = b.keys_out[head];
    const uint32_t k = lowerBound(b.keys_out, b.n, vid + 1u) - head;
    const uint32_t records = (b.interval_offset[head + k] - b.interval_offset[head]) +
                             (b.interval_offset[b.n + head + 1u] - b.interval_offset[b.n + head]);
    if (k + records >= b.heavy_run)
    {
      b.run_head[atomicAdd(&b.counters->heavy_count, 1u)] = (int32_t)t;  // run_head[] is free on this path
    }
    else
    {
      samples = replayNdtRun<false>(dm, g, mp, b, t);
    }
  }
  __syncwarp();
  const unsigned s = __reduce_add_sync(0xffffffffu, samples);
  if ((threadIdx.x & 31) == 0 && s)
  {
    atomicAdd(&b.counters->sample_updates, (unsigned long long)s);
  }
}

// The heavy runs applySamplesNdt set aside, one warp each.
__global__ void __launch_bounds__(128) applySamplesNdtHeavy(DeviceMap dm, Geom g, MapParams mp, Batch b)
{
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = b.counters->heavy_count;
  unsigned samples = 0;
  for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += warps)
  {
    samples += replayNdtRun<true>(dm, g, mp, b, (uint32_t)b.run_head[w]);
  }
  if ((threadIdx.x & 31) == 0 && samples)
  {
    atomicAdd(&b.counters->sample_updates, (unsigned long long)samples);
  }
}
