// ohmb200_exchange.cuh — multi-GPU ray integration that routes WORK, not rays (included by ohmb200.cu).
//
// The map is sharded by region (owner = (rx + 2 ry + 4 rz) mod world, ohmb200_device.cuh).  The reference has no
// multi-device path; its nearest mechanism is the cut of long rays into clipped-end segments whose cut voxel is updated
// once, by the next piece (ohmgpu/GpuMap.cpp:747-795, ohmgpu/gpu/AdjustOccupancy.cl:13-18).  Here the cut is the exact
// region-boundary cut of the single-GPU path (enumerateSegments), so the union of the per-GPU maps is bit-identical to
// one GPU — or the CPU mapper — integrating the same rays in rank order.
//
// One step (every rank: ohmb200_exchange_send, then ohmb200_exchange_integrate):
//
//   sender  (its OWN rays only)                                         NVLink traffic
//     exPrepRays      filter, keys, RayRec, sample voxel + owner        -
//     exRouteSamples  one 96-byte sample record -> its owner's inbox    peer stores, 96 B x samples
//     broadcast       RayRec[] (+ rays / timestamps for NDT, ray        copy engines, 64 (+ 60) B x rays x peers,
//                     lengths for traversal) -> every peer              beside the kernels below
//     exPrepSegments  cut every ray at region boundaries; each 32-byte  peer stores, 32 B x segments
//                     segment record -> the inbox of its region's owner
//     exSignal        counts + a flag into every peer's mailbox         32 B x peers
//   owner   (what arrived, from every rank incl. itself)
//     exWait          spin on the mailbox flags of this step (one warp; bounded, never hangs the GPU)
//     exBinSamples    (side stream, as soon as the samples are in: beside everybody's cut) sample pairs (voxel id,
//                     global ray index) for the sort
//     exBinSegments   one thread per received record: region find-or-insert, per-region segment histogram, touched list
//     planRegions / exEmit / radix sort / markRuns, then the single-GPU walk and replay kernels unchanged
//
// No rank filters or cuts another rank's rays; nothing is all-gathered through the host.  Inboxes live in ONE device
// allocation per rank (the "arena"), exported as a CUDA IPC handle: a peer in another process maps it
// (cudaIpcOpenMemHandle), a peer in the same process uses the pointer (tests run `world` maps on one device that way).
// Everything a sender writes is double-buffered by step parity; the mailbox handshake orders reuse: a rank passes
// exWait(k) only when every peer has queued its records of step k, which each peer does after consuming step k - 1.
#pragma once

namespace ohmb200
{
constexpr int kMaxWorld = 16;
constexpr unsigned long long kExMagic = 0x6f686d6232303058ull;  // "ohmb200X"

struct WireSegment
{
  uint4 seg;              // the 16-byte Segment, ray = global ray index (rank * per + i)
  uint32_t key_lo, key_hi;  // packed region key
  uint32_t slot;          // written by the owner (exBin): the region's slot in ITS table
  uint32_t pad;
};
static_assert(sizeof(WireSegment) == 32, "WireSegment must be 32 bytes");

struct WireSample
{
  unsigned long long key;  // packed region key of the sample voxel
  uint32_t voxel;          // voxel index inside the region
  uint32_t ray;            // global ray index
  double last_exit;        // exit range of the last voxel the ray walked (traversal layer; see staleExit)
  double timestamp;
  float intensity;
  uint32_t pad0;
  double pts[6];           // the ray itself (unfiltered origin, sample): the owner replays the hit from it
  unsigned long long pad1;
};
static_assert(sizeof(WireSample) == 96, "WireSample must be 96 bytes");

struct ExMailbox
{
  uint32_t seg_count, sample_count, ray_count, overflow;
  uint32_t seg_flag;  // == step when the sender's segment records of that step are in place
  uint32_t ray_flag;  // == step when its per-ray broadcast is
  uint32_t smp_flag;  // == step when its sample records are (they leave first: the owner sorts them beside the cut)
  uint32_t pts_flag;  // == step when the rays / timestamps / intensities an NDT map's replays read have followed
  uint32_t barrier;   // ohmb200_exchange_barrier: the sender's barrier count (parity 0 only)
  uint32_t pad[3];
};
static_assert(sizeof(ExMailbox) == 48, "ExMailbox layout");

// One rank's arena, one parity.
struct ExView
{
  RayRec *recs;          // [world * per]
  double *rays;          // [world * per * 6]
  double *timestamps;    // [world * per]
  float *intensities;    // [world * per]
  double *ray_length;    // [world * per]
  WireSegment *seg_in;   // [world][seg_cap]
  WireSample *smp_in;    // [world][per]
  ExMailbox *mailbox;    // [world]
};

struct ExStep
{
  int rank, world;
  uint32_t per;      // ray slots per rank
  uint32_t seg_cap;  // segment records per (sender, owner) pair
  const uint32_t *step;  // the step number, counted on the device (a replayed graph carries no host-side number)
  uint32_t n_own;
  uint32_t *out_seg;  // [world] records this rank has sent to each owner this step
  uint32_t *out_smp;
  unsigned long long *smp_key;  // [per] region key of each own ray's sample voxel, its voxel index and owner, and the
  uint32_t *smp_voxel;          //       exit range of its last walked voxel: parked by exPrepRays for exRouteSamples
  uint32_t *smp_owner;
  double *smp_last_exit;        // (traversal layer only, else null)
  int *abort;         // set when a wait timed out: the step is dropped
  ExView peer[kMaxWorld];
};

// One 32-byte store (st.global.v8.b32, sm_100+): a record of the inboxes is written as whole sectors — lanes that
// drew neighbouring slots then fill whole 128-byte lines with one instruction, which is what a store over NVLink
// wants (two 16-byte halves per lane were measured at an eighth of the link rate).  `dst` is 32-byte aligned.
__device__ __forceinline__ void store32(void *dst, const uint4 &a, const uint4 &b)
{
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// Position among the lanes of the converged group that target the same `bucket`, after one atomic per group.
__device__ __forceinline__ uint32_t bucketAggregatedInc(uint32_t *counters, uint32_t bucket)
{
  return slotAggregatedInc(counters, bucket);
}
}  // namespace ohmb200

// Filter, walk constants and sample voxel of this rank's own rays.  The RayRec goes to the rank's own arena (the
// broadcast copies it on); the sample is parked in scratch until exRouteSamples.
// route_now: the sample record leaves from here (no traversal layer: nothing to wait for) — 1: the full 96-byte record,
// 2: the 16-byte record of an occupancy-only map; 0: exRouteSamples sends it once carryLastExit has run.
__global__ void __launch_bounds__(128) exPrepRays(DeviceMap dm, Geom g, MapParams mp, Batch own, ExStep ex, int mode, int copy_rays,
                                                  int route_now)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool accepted = false;
  if (i < ex.n_own)
  {
    const ExView &mine = ex.peer[ex.rank];
    const uint32_t gid = (uint32_t)ex.rank * ex.per + i;
    double start[3], end[3];
    loadRay(own, i, start, end);
    const double raw[6] = { start[0], start[1], start[2], end[0], end[1], end[2] };
    if (copy_rays)
    {
      // the consumers read rays by global index (ndtGaussianMisses, the replays): the raw ray, filtered again there
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        mine.rays[(size_t)gid * 6 + a] = start[a];
        mine.rays[(size_t)gid * 6 + 3 + a] = end[a];
      }
      if (own.timestamps)
      {
        mine.timestamps[gid] = own.timestamps[i];
      }
      if (own.intensities)
      {
        mine.intensities[gid] = own.intensities[i];
      }
    }
    unsigned filter_flags = 0;
    uint32_t voxel = kInvalidVoxel, owner = 0;
    unsigned long long key = 0;
    RayRec rec;
    rec.flags = 0;
    double last = nan("");
    if (applyRayFilter(mp, start, end, filter_flags))
    {
      accepted = true;
      const bool include_sample_in_ray = (filter_flags & kRffClippedEnd) || (own.ray_flags & OHMB200_RF_END_POINT_AS_FREE);
      bool hit = !include_sample_in_ray;
      if (mode == OHMB200_MODE_OCCUPANCY)
      {
        hit = hit && !(own.ray_flags & OHMB200_RF_EXCLUDE_SAMPLE);
      }
      Key ekey;
      if (hit && voxelKey(g, end, ekey))
      {
        voxel = voxelIndex(g, ekey);
        owner = (uint32_t)regionOwner(ekey.r[0], ekey.r[1], ekey.r[2], ex.world);
        key = packRegion(ekey.r[0], ekey.r[1], ekey.r[2]);
      }
      unsigned walk_flags = (!include_sample_in_ray) ? kExcludeEndVoxel : 0u;
      walk_flags |= (own.ray_flags & OHMB200_RF_EXCLUDE_ORIGIN) ? kExcludeStartVoxel : 0u;
      if (!(own.ray_flags & OHMB200_RF_EXCLUDE_RAY))
      {
        makeRayRec(rec, g, start, end, walk_flags);
      }
      if (dm.traversal)
      {
        const double d[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
        const double len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double length = (len2 > 1e-6) ? sqrt(len2) : 0;
        mine.ray_length[gid] = length;
        // closed-form exit range of the last walked voxel (see prepRays)
        if (rec.flags & kRecValid)
        {
          const int steps = (int)rec.total[0] + (int)rec.total[1] + (int)rec.total[2];
          if (!(rec.flags & kRecExcludeEnd))
          {
            last = length;
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
      }
    }
    if (route_now == 2)
    {
      // an occupancy-only map: the owner replays the hit from (voxel, ray order) alone — a 16-byte record
      if (voxel != kInvalidVoxel)
      {
        const uint32_t at = bucketAggregatedInc(ex.out_smp, owner);
        if (at < ex.per)
        {
          reinterpret_cast<uint4 *>(ex.peer[owner].smp_in)[(size_t)ex.rank * ex.per + at] =
            make_uint4((uint32_t)key, (uint32_t)(key >> 32), voxel, gid);
        }
      }
    }
    else if (route_now)
    {
      if (voxel != kInvalidVoxel)
      {
        const uint32_t at = bucketAggregatedInc(ex.out_smp, owner);
        if (at < ex.per)
        {
          char *dst = reinterpret_cast<char *>(ex.peer[owner].smp_in + (size_t)ex.rank * ex.per + at);
          const double ts = own.timestamps ? own.timestamps[i] : 0.0;
          const float intensity = own.intensities ? own.intensities[i] : 0.0f;
          store32(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), voxel, gid),
                  make_uint4(0u, 0u, (uint32_t)__double_as_longlong(ts), (uint32_t)(__double_as_longlong(ts) >> 32)));
          store32(dst + 32, make_uint4(__float_as_uint(intensity), 0u, (uint32_t)__double_as_longlong(raw[0]), (uint32_t)(__double_as_longlong(raw[0]) >> 32)),
                  make_uint4((uint32_t)__double_as_longlong(raw[1]), (uint32_t)(__double_as_longlong(raw[1]) >> 32),
                             (uint32_t)__double_as_longlong(raw[2]), (uint32_t)(__double_as_longlong(raw[2]) >> 32)));
          store32(dst + 64, make_uint4((uint32_t)__double_as_longlong(raw[3]), (uint32_t)(__double_as_longlong(raw[3]) >> 32),
                                       (uint32_t)__double_as_longlong(raw[4]), (uint32_t)(__double_as_longlong(raw[4]) >> 32)),
                  make_uint4((uint32_t)__double_as_longlong(raw[5]), (uint32_t)(__double_as_longlong(raw[5]) >> 32), 0u, 0u));
        }
      }
    }
    else
    {
      ex.smp_voxel[i] = voxel;
      ex.smp_owner[i] = owner;
      ex.smp_key[i] = key;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      reinterpret_cast<uint4 *>(mine.recs + gid)[k] = reinterpret_cast<const uint4 *>(&rec)[k];
    }
    if (ex.smp_last_exit)
    {
      ex.smp_last_exit[i] = last;
    }
  }
  __syncwarp();
  const unsigned n_acc = __reduce_add_sync(0xffffffffu, accepted ? 1u : 0u);
  if ((threadIdx.x & 31) == 0 && n_acc)
  {
    atomicAdd(&own.counters->rays_accepted, (unsigned long long)n_acc);
  }
}

// One sample record per own ray that hits, stored straight into the inbox of the sample voxel's owner (a peer store
// over NVLink, or a local store).  Runs after carryLastExit so that the record carries the final exit range.
__global__ void __launch_bounds__(128) exRouteSamples(Batch own, ExStep ex)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ex.n_own)
  {
    return;
  }
  const uint32_t voxel = ex.smp_voxel[i];
  if (voxel == kInvalidVoxel)
  {
    return;
  }
  const uint32_t owner = ex.smp_owner[i];
  const uint32_t at = bucketAggregatedInc(ex.out_smp, owner);
  if (at >= ex.per)
  {
    return;  // cannot happen: a rank sends at most n_own <= per samples
  }
  WireSample smp;
  smp.key = ex.smp_key[i];
  smp.voxel = voxel;
  smp.ray = (uint32_t)ex.rank * ex.per + i;
  smp.last_exit = ex.smp_last_exit ? ex.smp_last_exit[i] : 0.0;
  smp.timestamp = own.timestamps ? own.timestamps[i] : 0.0;
  smp.intensity = own.intensities ? own.intensities[i] : 0.0f;
  smp.pad0 = 0;
  smp.pad1 = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k)
  {
    smp.pts[k] = own.rays[(size_t)i * 6 + k];
  }
  char *dst = reinterpret_cast<char *>(ex.peer[owner].smp_in + (size_t)ex.rank * ex.per + at);
  const uint4 *src = reinterpret_cast<const uint4 *>(&smp);
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    store32(dst + 32 * k, src[2 * k], src[2 * k + 1]);
  }
}

// Cut this rank's own rays at region boundaries (enumerateSegments) and store each segment into the inbox of the
// region's owner.  No hash probe, no histogram here: the owner bins what it receives (exBin), one thread per record.
__global__ void __launch_bounds__(128) exPrepSegments(Geom g, Batch own, ExStep ex)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned visits = 0;
  bool overflow = false;
  if (i < ex.n_own)
  {
    const uint32_t ray = rayOfThread(i, ex.n_own);
    const uint32_t gid = (uint32_t)ex.rank * ex.per + ray;
    RayRec rec;
    loadRec(rec, ex.peer[ex.rank].recs + gid);
    if (rec.flags & kRecValid)
    {
      enumerateSegments(rec, g, [&](const int r[3], const int st[3], const int entry[3], int n) {
        const uint32_t owner = (uint32_t)regionOwner(r[0], r[1], r[2], ex.world);
        const uint32_t at = bucketAggregatedInc(ex.out_seg, owner);
        visits += (unsigned)n;
        if (at >= ex.seg_cap)
        {
          overflow = true;
          return;
        }
        const unsigned long long key = packRegion(r[0], r[1], r[2]);
        store32(ex.peer[owner].seg_in + (size_t)ex.rank * ex.seg_cap + at,
                make_uint4(gid, (uint32_t)st[0] | ((uint32_t)st[1] << 16), (uint32_t)st[2] | ((uint32_t)n << 16),
                           (uint32_t)entry[0] | ((uint32_t)entry[1] << 8) | ((uint32_t)entry[2] << 16)),
                make_uint4((uint32_t)key, (uint32_t)(key >> 32), 0xFFFFFFFFu, 0u));
      });
    }
  }
  __syncwarp();
  const unsigned n_vis = __reduce_add_sync(0xffffffffu, visits);
  if ((threadIdx.x & 31) == 0 && n_vis)
  {
    atomicAdd(&own.counters->voxel_visits, (unsigned long long)n_vis);
  }
  if (overflow)
  {
    atomicOr(&own.counters->overflow_seen, 2);
  }
}

__global__ void exBumpStep(uint32_t *step)
{
  *step += 1u;
}

// Tell every owner what this rank has put into its inbox.  stage 2: sample records (sent first); stage 0: segment
// records; stage 1: the walk constants of the per-ray broadcast; stage 3: its rays / timestamps / intensities (NDT).  The data was written by earlier work of the same stream(s); the system-scope fence orders it before
// the flag for a reader on another GPU.
__global__ void exSignal(ExStep ex, int stage)
{
  const int o = (int)threadIdx.x;
  if (o >= ex.world)
  {
    return;
  }
  ExMailbox *box = ex.peer[o].mailbox + ex.rank;
  if (stage == 0)
  {
    const uint32_t segs = ex.out_seg[o];
    box->seg_count = min(segs, ex.seg_cap);
    box->ray_count = ex.n_own;
    box->overflow = segs > ex.seg_cap ? 1u : 0u;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&box->seg_flag) = *ex.step;
  }
  else if (stage == 2)
  {
    box->sample_count = min(ex.out_smp[o], ex.per);
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&box->smp_flag) = *ex.step;
  }
  else if (stage == 3)
  {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&box->pts_flag) = *ex.step;
  }
  else
  {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&box->ray_flag) = *ex.step;
  }
}

// Wait (one warp, lane = sender) until every sender's flag of this step is in this rank's mailbox.  Bounded: after
// ~4 s of spinning the step is abandoned (abort flag: exBin / exEmit see empty inboxes, ohmb200_sync reports it) — a
// missing peer must never hang the GPU.
__global__ void exWait(ExMailbox *mailbox, int world, const uint32_t *step_counter, int stage, int *abort)
{
  const uint32_t step = *step_counter;
  const int s = (int)threadIdx.x;
  bool ok = true;
  if (s < world)
  {
    const volatile uint32_t *flag = stage == 0 ? &mailbox[s].seg_flag : (stage == 1 ? &mailbox[s].ray_flag : (stage == 2 ? &mailbox[s].smp_flag : &mailbox[s].pts_flag));
    const long long t0 = clock64();
    while (*flag != step)
    {
      if (clock64() - t0 > 8000000000ll)
      {
        ok = false;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
  if (!ok)
  {
    *abort = 1;
  }
}

// The owner's view of what arrived: per-sender counts and their prefix sums.
struct ExInbox
{
  uint32_t seg_first[kMaxWorld + 1];
  uint32_t smp_first[kMaxWorld + 1];
};

__device__ __forceinline__ void exLoadInbox(ExInbox &in, const ExStep &ex)
{
  const ExMailbox *box = ex.peer[ex.rank].mailbox;
  const bool dead = *ex.abort != 0;
  uint32_t segs = 0, smps = 0;
  for (int s = 0; s < ex.world; ++s)
  {
    in.seg_first[s] = segs;
    in.smp_first[s] = smps;
    segs += dead ? 0u : min(box[s].seg_count, ex.seg_cap);
    smps += dead ? 0u : min(box[s].sample_count, ex.per);
  }
  for (int s = ex.world; s <= kMaxWorld; ++s)
  {
    in.seg_first[s] = segs;
    in.smp_first[s] = smps;
  }
}

__device__ __forceinline__ int exSender(const uint32_t *first, int world, uint32_t j)
{
  int s = 0;
  while (s + 1 < world && j >= first[s + 1])
  {
    ++s;
  }
  return s;
}

// One thread per received segment record: region find-or-insert in THIS rank's table, per-region histogram, touched
// list (what prepSegments does per ray on the single-GPU path).  The lanes of a warp that hold the same region — the
// k-th segments of neighbouring rays — probe the table once: the first of them looks the region up (or creates it) and
// adds the whole group to the histogram.  (One probe and one compare-and-swap per LANE was 110 us on a fresh map: tens
// of thousands of lanes arrive at an empty slot of a hot region together.)
__global__ void __launch_bounds__(256) exBinSegments(DeviceMap dm, Batch b, ExStep ex)
{
  __shared__ ExInbox in;
  if (threadIdx.x == 0)
  {
    exLoadInbox(in, ex);
    if (*ex.abort)
    {
      b.counters->segment_overflow = 1;  // the sample replay kernels skip the step too
    }
    else
    {
      const ExMailbox *box = ex.peer[ex.rank].mailbox;
      for (int s = 0; s < ex.world; ++s)
      {
        if (box[s].overflow)
        {
          b.counters->segment_overflow = 1;  // a sender ran out of inbox space: drop the step here
          atomicOr(&b.counters->overflow_seen, 2);
        }
      }
    }
  }
  __syncthreads();
  const ExView &mine = ex.peer[ex.rank];
  const uint32_t total_segs = in.seg_first[kMaxWorld];
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned visits = 0;
  for (uint32_t base = blockIdx.x * blockDim.x; base < total_segs; base += stride)
  {
    const uint32_t j = base + threadIdx.x;
    if (j < total_segs)
    {
      const int s = exSender(in.seg_first, ex.world, j);
      WireSegment *ws = mine.seg_in + (size_t)s * ex.seg_cap + (j - in.seg_first[s]);
      const unsigned long long key = (unsigned long long)ws->key_lo | ((unsigned long long)ws->key_hi << 32);
      visits += ws->seg.z >> 16;
      const unsigned peers = __match_any_sync(__activemask(), key);
      const int leader = __ffs(peers) - 1;
      int slot = -1;
      if ((int)lane == leader)
      {
        slot = regionSlot(dm, key);
        if (slot >= 0 && atomicAdd(&b.seg_count[slot], (uint32_t)__popc(peers)) == 0u)
        {
          b.touched_list[atomicAdd(&b.counters->touched_count, 1u)] = (uint32_t)slot;
        }
      }
      slot = __shfl_sync(peers, slot, leader);
      ws->slot = (uint32_t)slot;
    }
  }
  visits = __reduce_add_sync(0xffffffffu, visits);
  if (lane == 0 && visits)
  {
    atomicAdd(&b.counters->owned_visits, (unsigned long long)visits);
  }
}

// One thread per received sample record: the (voxel id, global ray) pair at index `ray` — the pairs are then in global
// ray order, which the stable sort keeps inside a voxel — plus the ray itself where the rays were not broadcast.
__global__ void __launch_bounds__(256) exBinSamples(DeviceMap dm, Geom g, Batch b, ExStep ex, int rays_from_samples)
{
  __shared__ ExInbox in;
  if (threadIdx.x == 0)
  {
    exLoadInbox(in, ex);
  }
  __syncthreads();
  const ExView &mine = ex.peer[ex.rank];
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t total_smps = in.smp_first[kMaxWorld];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total_smps; j += stride)
  {
    const int s = exSender(in.smp_first, ex.world, j);
    const size_t at = (size_t)s * ex.per + (j - in.smp_first[s]);
    const WireSample *smp = mine.smp_in + at;
    uint32_t ray, voxel;
    unsigned long long key;
    if (rays_from_samples == 2)
    {
      const uint4 lite = reinterpret_cast<const uint4 *>(mine.smp_in)[at];  // {key, voxel, ray}: occupancy-only map
      key = (unsigned long long)lite.x | ((unsigned long long)lite.y << 32);
      voxel = lite.z;
      ray = lite.w;
    }
    else
    {
      key = smp->key;
      voxel = smp->voxel;
      ray = smp->ray;
    }
    // the lanes that hold the same region probe the table once (see exBinSegments)
    const unsigned peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1;
    int slot = -1;
    if ((int)(threadIdx.x & 31u) == leader)
    {
      slot = regionSlot(dm, key);
    }
    slot = __shfl_sync(peers, slot, leader);
    if (slot >= 0)
    {
      b.keys_in[ray] = (uint32_t)slot * g.vpr + voxel;
      b.vals_in[ray] = ray;
    }
    if (rays_from_samples == 1)
    {
#pragma unroll
      for (int k = 0; k < 6; ++k)
      {
        mine.rays[(size_t)ray * 6 + k] = smp->pts[k];
      }
      mine.timestamps[ray] = smp->timestamp;
      mine.intensities[ray] = smp->intensity;
    }
    if (b.last_exit)
    {
      b.last_exit[ray] = smp->last_exit;  // (a map with the traversal layer never takes the 16-byte records)
    }
  }
}

// Scatter the received segments into their region's range of the binned list (emitSegments' job on the single-GPU path).
__global__ void __launch_bounds__(256) exEmit(Batch b, ExStep ex)
{
  __shared__ ExInbox in;
  if (threadIdx.x == 0)
  {
    exLoadInbox(in, ex);
  }
  __syncthreads();
  if (b.counters->segment_overflow)
  {
    return;
  }
  const ExView &mine = ex.peer[ex.rank];
  const uint32_t total_segs = in.seg_first[kMaxWorld];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < total_segs; base += stride)
  {
    const uint32_t j = base + threadIdx.x;
    if (j < total_segs)
    {
      const int s = exSender(in.seg_first, ex.world, j);
      const WireSegment *ws = mine.seg_in + (size_t)s * ex.seg_cap + (j - in.seg_first[s]);
      const uint32_t slot = ws->slot;
      if (slot != 0xFFFFFFFFu)
      {
        const uint32_t at = b.seg_offset[slot] + slotAggregatedInc(b.seg_cursor, slot);
        if (at < b.seg_capacity)
        {
          reinterpret_cast<uint4 *>(b.segments)[at] = ws->seg;
        }
      }
    }
  }
}
