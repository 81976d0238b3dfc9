// ohmb200_exchange_host.inl — host side of the routed multi-GPU exchange (included inside ohmb200.cu's anonymous
// namespace; kernels and the protocol: ohmb200_exchange.cuh).

// Debug aid (OHMB200_TRACE_CAPTURE): reports the first operation after which a step's recording was invalidated.
#define EX_CAPCHK(m, what)                                                                                  \
  do                                                                                                        \
  {                                                                                                         \
    static const bool trace__ = getenv("OHMB200_TRACE_CAPTURE") != nullptr;                                 \
    if (trace__ && (m)->ex.capturing)                                                                       \
    {                                                                                                       \
      cudaStreamCaptureStatus st__ = cudaStreamCaptureStatusNone;                                           \
      cudaStreamIsCapturing((m)->stream, &st__);                                                            \
      fprintf(stderr, "[capture] %s:%d %s -> %d\n", __FILE__, __LINE__, what, (int)st__);                   \
    }                                                                                                       \
  } while (0)

struct ExHandle
{
  unsigned long long magic;
  int32_t pid, device, rank, world;
  uint32_t per, seg_cap;
  unsigned long long base, bytes;
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(ExHandle) <= sizeof(ohmb200_exchange_handle), "ohmb200_exchange_handle is too small");

size_t exAlign(size_t bytes)
{
  return (bytes + 255u) & ~(size_t)255u;
}

// Byte offsets of one parity of an arena.
struct ExLayout
{
  size_t recs, rays, timestamps, intensities, ray_length, seg_in, smp_in, mailbox, bytes;
};

ExLayout exLayout(int world, uint32_t per, uint32_t seg_cap)
{
  ExLayout l{};
  const size_t slots = (size_t)world * per;
  size_t at = 0;
  l.recs = at;
  at += exAlign(slots * sizeof(RayRec));
  l.rays = at;
  at += exAlign(slots * 6 * sizeof(double));
  l.timestamps = at;
  at += exAlign(slots * sizeof(double));
  l.intensities = at;
  at += exAlign(slots * sizeof(float));
  l.ray_length = at;
  at += exAlign(slots * sizeof(double));
  l.seg_in = at;
  at += exAlign((size_t)world * seg_cap * sizeof(WireSegment));
  l.smp_in = at;
  at += exAlign(slots * sizeof(WireSample));
  l.mailbox = at;
  at += exAlign((size_t)world * sizeof(ExMailbox));
  l.bytes = at;
  return l;
}

ExView exView(char *base, const ExLayout &l, size_t parity_bytes, int parity)
{
  char *p = base + (size_t)parity * parity_bytes;
  ExView v;
  v.recs = (RayRec *)(p + l.recs);
  v.rays = (double *)(p + l.rays);
  v.timestamps = (double *)(p + l.timestamps);
  v.intensities = (float *)(p + l.intensities);
  v.ray_length = (double *)(p + l.ray_length);
  v.seg_in = (WireSegment *)(p + l.seg_in);
  v.smp_in = (WireSample *)(p + l.smp_in);
  v.mailbox = (ExMailbox *)(p + l.mailbox);
  return v;
}

int exchangeClose(ohmb200_map *m)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  ohmb200_map::Exchange &x = m->ex;
  if (!x.open)
  {
    return OHMB200_OK;
  }
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  if (x.stream)
  {
    cudaStreamSynchronize(x.stream);
  }
  for (int r = 0; r < kMaxWorld; ++r)
  {
    if (x.peer_mapped[r] && x.peer_base[r])
    {
      cudaIpcCloseMemHandle(x.peer_base[r]);
    }
    x.peer_base[r] = nullptr;
    x.peer_mapped[r] = false;
  }
  cudaFree(x.arena);
  cudaFree(x.out_counts);
  cudaFree(x.smp_key);
  cudaFree(x.smp_voxel);
  cudaFree(x.smp_owner);
  cudaFree(x.smp_last_exit);
  cudaFree(x.abort);
  cudaFree(x.d_step);
  cudaFree(x.d_barrier);
  for (auto &g : x.graphs)
  {
    cudaGraphExecDestroy(g.exec);
  }
  if (x.bcast_done)
  {
    cudaEventDestroy(x.bcast_done);
  }
  if (x.stream)
  {
    cudaStreamDestroy(x.stream);
  }
  if (x.prepped)
  {
    cudaEventDestroy(x.prepped);
  }
  cudaGetLastError();
  x = ohmb200_map::Exchange();
  m->batch.recs = nullptr;  // (they pointed into the arena)
  m->batch.ray_length = nullptr;
  m->batch.rays = nullptr;
  m->dm.part_rank = 0;
  m->dm.part_world = 1;
  m->scratch_rays = 0;  // back to the single-GPU scratch layout at the next batch
  return OHMB200_OK;
}

int exchangeOpen(ohmb200_map *m, int rank, int world, size_t max_rays_per_rank, ohmb200_exchange_handle *handle)
{
  if (!m || !handle || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_rays_per_rank == 0 ||
      max_rays_per_rank * (size_t)world > 0x7FFFFFFFu)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_open: need 0 <= rank < world <= %d and 0 < max_rays_per_rank x world < 2^31",
                    kMaxWorld);
  }
  if (m->mode == OHMB200_MODE_TSDF || m->algo != 1)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_open: occupancy and NDT maps on the region-binned path only "
                                       "(a TSDF map is sharded with ohmb200_set_partition and fed every ray)");
  }
  cudaSetDevice(m->device);
  int rc = exchangeClose(m);
  if (rc)
  {
    return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  dropBatchGraphs(m);
  ohmb200_map::Exchange &x = m->ex;
  x.rank = rank;
  x.world = world;
  x.per = (uint32_t)((max_rays_per_rank + 2047u) & ~(size_t)2047u);
  // Segment records per (sender, owner) pair.  A sweep cuts into ~8 segments per ray and the owners share them about
  // evenly; the pair's inbox holds 8x the even share of that (at least 8 per ray slot), bounded by the 96 per ray of
  // the single-GPU list.  A sender that runs out flags the step (OHMB200_E_OVERFLOW), it never writes past the inbox.
  x.seg_cap = x.per * (uint32_t)std::min<size_t>(96u, std::max<size_t>(8u, 64u / (size_t)world));
  const ExLayout l = exLayout(world, x.per, x.seg_cap);
  x.parity_bytes = l.bytes;
  x.arena_bytes = 2 * l.bytes;
  void *arena = nullptr;
  if (cudaMalloc(&arena, x.arena_bytes) != cudaSuccess)
  {
    cudaGetLastError();
    x = ohmb200_map::Exchange();
    return setError(OHMB200_E_CUDA, "ohmb200_exchange_open: cannot allocate the %zu MiB exchange arena", x.arena_bytes >> 20);
  }
  x.arena = (char *)arena;
  bool ok = true;
  for (int p = 0; p < 2 && ok; ++p)
  {
    ok = cudaMemsetAsync(x.arena + p * l.bytes + l.mailbox, 0, exAlign((size_t)world * sizeof(ExMailbox)), m->stream) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&x.out_counts, sizeof(uint32_t) * 2 * kMaxWorld) == cudaSuccess;
  ok = ok && cudaMalloc(&x.smp_key, sizeof(unsigned long long) * x.per) == cudaSuccess;
  ok = ok && cudaMalloc(&x.smp_voxel, sizeof(uint32_t) * x.per) == cudaSuccess;
  ok = ok && cudaMalloc(&x.smp_owner, sizeof(uint32_t) * x.per) == cudaSuccess;
  if (m->dm.traversal)
  {
    ok = ok && cudaMalloc(&x.smp_last_exit, sizeof(double) * x.per) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&x.abort, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(x.abort, 0, sizeof(int), m->stream) == cudaSuccess;
  ok = ok && cudaMalloc(&x.d_step, sizeof(uint32_t)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(x.d_step, 0, sizeof(uint32_t), m->stream) == cudaSuccess;
  ok = ok && cudaMalloc(&x.d_barrier, sizeof(uint32_t)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(x.d_barrier, 0, sizeof(uint32_t), m->stream) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&x.bcast_done, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&x.prepped, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(m->stream) == cudaSuccess;
  ExHandle h{};
  h.magic = kExMagic;
  h.pid = (int32_t)getpid();
  h.device = m->device;
  h.rank = rank;
  h.world = world;
  h.per = x.per;
  h.seg_cap = x.seg_cap;
  h.base = (unsigned long long)(uintptr_t)x.arena;
  h.bytes = x.arena_bytes;
  if (ok && world > 1 && cudaIpcGetMemHandle(&h.ipc, x.arena) != cudaSuccess)
  {
    // no IPC on this platform: peers in the same process still work (they use the pointer)
    cudaGetLastError();
    memset(&h.ipc, 0, sizeof(h.ipc));
  }
  if (!ok)
  {
    const char *why = cudaGetErrorString(cudaGetLastError());
    x.open = true;
    exchangeClose(m);
    return setError(OHMB200_E_CUDA, "ohmb200_exchange_open failed: %s", why);
  }
  memset(handle, 0, sizeof(*handle));
  memcpy(handle, &h, sizeof(h));
  x.open = true;
  m->dm.part_rank = rank;
  m->dm.part_world = world;
  m->scratch_rays = 0;  // the batch scratch is laid out for the exchange from the next batch on
  return OHMB200_OK;
}

int exchangeConnect(ohmb200_map *m, const ohmb200_exchange_handle *handles, int count)
{
  if (!m || !handles)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_connect: null argument");
  }
  ohmb200_map::Exchange &x = m->ex;
  if (!x.open || count != x.world)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_connect: open the exchange first and pass one handle per rank (%d)",
                    x.open ? x.world : 0);
  }
  cudaSetDevice(m->device);
  for (int r = 0; r < x.world; ++r)
  {
    ExHandle h;
    memcpy(&h, &handles[r], sizeof(h));
    if (h.magic != kExMagic || h.rank != r || h.world != x.world || h.per != x.per || h.seg_cap != x.seg_cap ||
        h.bytes != x.arena_bytes)
    {
      return setError(OHMB200_E_INVALID, "ohmb200_exchange_connect: handle %d does not belong to this exchange (every rank "
                                         "must open with the same world and max_rays_per_rank, handles in rank order)", r);
    }
    if (r == x.rank)
    {
      x.peer_base[r] = x.arena;
      continue;
    }
    if (h.pid == (int32_t)getpid())
    {
      x.local_peers = true;
      // a peer map of this process: its pointer is ours too; another device needs peer access switched on
      if (h.device != m->device)
      {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->device, h.device);
        if (!can)
        {
          return setError(OHMB200_E_CUDA, "ohmb200_exchange_connect: device %d cannot access device %d", m->device, h.device);
        }
        const cudaError_t err = cudaDeviceEnablePeerAccess(h.device, 0);
        if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
        {
          return setError(OHMB200_E_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(err));
        }
        cudaGetLastError();
      }
      x.peer_base[r] = (char *)(uintptr_t)h.base;
      continue;
    }
    void *mapped = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&mapped, h.ipc, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess)
    {
      cudaGetLastError();
      return setError(OHMB200_E_CUDA, "ohmb200_exchange_connect: cannot map rank %d's arena (cudaIpcOpenMemHandle: %s)", r,
                      cudaGetErrorString(err));
    }
    x.peer_base[r] = (char *)mapped;
    x.peer_mapped[r] = true;
  }
  x.connected = true;
  return OHMB200_OK;
}

void exFillStep(ohmb200_map *m, ExStep &ex)
{
  const ohmb200_map::Exchange &x = m->ex;
  const ExLayout l = exLayout(x.world, x.per, x.seg_cap);
  memset(&ex, 0, sizeof(ex));
  ex.rank = x.rank;
  ex.world = x.world;
  ex.per = x.per;
  ex.seg_cap = x.seg_cap;
  ex.step = x.d_step;
  ex.n_own = (uint32_t)x.n_own;
  ex.out_seg = x.out_counts;
  ex.out_smp = x.out_counts + kMaxWorld;
  ex.smp_key = x.smp_key;
  ex.smp_voxel = x.smp_voxel;
  ex.smp_owner = x.smp_owner;
  ex.smp_last_exit = x.smp_last_exit;
  ex.abort = x.abort;
  for (int r = 0; r < x.world; ++r)
  {
    ex.peer[r] = exView(x.peer_base[r], l, x.parity_bytes, (int)(x.step & 1u));
  }
}

// An occupancy map with no per-sample layer: the owner replays a hit from (voxel, ray order) alone, so the sample
// records shrink from 96 to 16 bytes (hit value only: no mean, incident normal, touch time or traversal to update).
bool exLiteSamples(const ohmb200_map *m)
{
  return m->mode == OHMB200_MODE_OCCUPANCY && !m->dm.mean && !m->dm.traversal && !m->dm.touch_time && !m->dm.incident;
}

bool exNdt(const ohmb200_map *m)
{
  return m->mode == OHMB200_MODE_NDT || m->mode == OHMB200_MODE_NDT_TM;
}

// The owner's sample branch: wait for every rank's sample records, turn them into (voxel id, ray) pairs, sort, mark the
// runs — on the side stream (joined before the walk) unless per-kernel profiling serialises everything.
int exSampleBranch(ohmb200_map *m, const ExStep &ex)
{
  ohmb200_map::Exchange &x = m->ex;
  const ExView &mine = ex.peer[x.rank];
  const size_t n_total = (size_t)x.world * x.per;
  cudaStream_t s = m->stream;
  Batch &b = m->batch;
  b.rays = mine.rays;
  b.n = (uint32_t)n_total;
  b.counters = m->d_counters;
  const bool fork = !m->profiling;
  cudaStream_t ss = fork ? m->side_stream : s;
  if (fork)
  {
    CUDA_TRY(cudaEventRecord(m->fork_event, s));
    CUDA_TRY(cudaStreamWaitEvent(ss, m->fork_event, 0));
  }
  CUDA_TRY(cudaMemsetAsync(b.keys_in, 0xFF, sizeof(uint32_t) * n_total, ss));  // rays without a sample here: no voxel
  CUDA_TRY(cudaMemsetAsync(b.sample_begin, 0, sizeof(uint32_t) * 2 * m->dm.capacity, ss));
  EX_CAPCHK(m, "branch memsets");
  {
    KernelScope scope(m, kKExWait);
    exWait<<<1, 32, 0, ss>>>(mine.mailbox, x.world, x.d_step, 2, x.abort);
  }
  {
    KernelScope scope(m, kKExBinSamples);
    exBinSamples<<<(unsigned)m->sm_count * 4u, 256, 0, ss>>>(m->dm, m->geom, b, ex, exNdt(m) ? 0 : (exLiteSamples(m) ? 2 : 1));
  }
  {
    KernelScope scope(m, kKSort);
    size_t temp = m->cub_temp_bytes;
    cub::DeviceRadixSort::SortPairs(m->cub_temp, temp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n_total, 0, m->sort_bits, ss);
  }
  EX_CAPCHK(m, "branch sort");
  {
    KernelScope scope(m, kKMark);
    markRuns<<<(unsigned)((n_total + 127) / 128), 128, 0, ss>>>(m->dm, b, m->geom.vpr);
  }
  if (fork)
  {
    CUDA_TRY(cudaEventRecord(m->join_event, ss));
  }
  x.forked = fork;
  return OHMB200_OK;
}

// Phase 1 of a step: this rank's own rays (device memory) -> filter, cut, route.  Returns after queueing.
int exchangeSend(ohmb200_map *m, const double *d_rays, size_t element_count, const float *d_intensities,
                 const double *d_timestamps, unsigned ray_flags)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  ohmb200_map::Exchange &x = m->ex;
  const size_t n = element_count / 2;
  if (!x.open || !x.connected)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_send: the exchange is not open and connected");
  }
  if (x.pending)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_send: the previous step has not been integrated (ohmb200_exchange_integrate)");
  }
  if ((n && !d_rays) || n > x.per)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_send: %zu rays, the exchange was opened for at most %u per rank", n, x.per);
  }
  if (ray_flags & OHMB200_RF_STOP_ON_FIRST_OCCUPIED)
  {
    return setError(OHMB200_E_INVALID, "kRfStopOnFirstOccupied needs the whole map on one GPU: a ray's stop depends on voxels "
                                       "of every region it crosses");
  }
  cudaSetDevice(m->device);
  if (!(m->params.layers & (1u << OHMB200_LAYER_INTENSITY)))
  {
    d_intensities = nullptr;
  }
  if (!(m->params.layers & (1u << OHMB200_LAYER_TOUCH_TIME)))
  {
    d_timestamps = nullptr;
  }
  if (d_timestamps && m->first_ray_time < 0)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_send: set the time base first (ohmb200_set_first_ray_time, the same on "
                                       "every rank): the first timestamp of the whole batch is rank 0's");
  }
  int rc = ensureScratch(m, (size_t)x.world * x.per);
  rc = rc ? rc : ensureRoom(m);
  if (rc)
  {
    return rc;
  }
  ++x.step;
  x.n_own = n;
  x.ray_flags = ray_flags;
  x.has_timestamps = d_timestamps != nullptr;
  x.has_intensities = d_intensities != nullptr;
  x.replayed = false;
  x.capturing = false;
  cudaStream_t s = m->stream;
  // A step of the same shape as one met before — same ray buffers, count, flags, parity of the inboxes — is replayed as
  // one CUDA graph (send + integrate: ~45 stream operations); the second time a shape comes it is recorded (relaxed
  // capture mode: the recording spans two API calls, and what the caller does between them is none of its business).  Not with
  // peer maps in this process (their calls interleave with ours), per-kernel profiling or a paged-out map.
  if (m->use_graphs && !m->profiling && m->store.empty() && !x.local_peers)
  {
    const ohmb200_map::Exchange::StepGraph key{ d_rays, d_intensities, d_timestamps, n, ray_flags, (int)(x.step & 1u),
                                                 m->first_ray_time, nullptr, 0 };
    auto same = [&](const ohmb200_map::Exchange::StepGraph &g) {
      return g.rays == key.rays && g.intensities == key.intensities && g.timestamps == key.timestamps && g.n == key.n &&
             g.ray_flags == key.ray_flags && g.parity == key.parity && g.time_base == key.time_base;
    };
    for (auto &g : x.graphs)
    {
      if (same(g))
      {
        CUDA_TRY(cudaGraphLaunch(g.exec, s));
        m->launches += g.launches;
        x.replayed = true;
        x.pending = true;
        return OHMB200_OK;
      }
    }
    // (met before: whatever the parity was — the two recordings of a shape then fall into its second and third step)
    bool met = false;
    for (auto &g : x.seen)
    {
      met = met || (g.rays == key.rays && g.intensities == key.intensities && g.timestamps == key.timestamps && g.n == key.n &&
                    g.ray_flags == key.ray_flags && g.time_base == key.time_base);
    }
    if (!met)
    {
      if (x.seen.size() >= 16)
      {
        x.seen.erase(x.seen.begin());
      }
      x.seen.push_back(key);
    }
    else if (x.graphs.size() < 8 && cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess)
    {
      x.capturing = true;
      x.current = key;
      x.launches_before = m->launches;
    }
  }
  // A failure while the step is being recorded must not leave the stream in capture mode.
  struct CaptureGuard
  {
    ohmb200_map *m;
    bool armed = true;
    ~CaptureGuard()
    {
      if (armed && m->ex.capturing)
      {
        cudaGraph_t graph = nullptr;
        cudaStreamEndCapture(m->stream, &graph);
        if (graph)
        {
          cudaGraphDestroy(graph);
        }
        cudaGetLastError();
        m->ex.capturing = false;
        m->use_graphs = false;
      }
    }
  } capture_guard{ m };
  exBumpStep<<<1, 1, 0, s>>>(x.d_step);
  EX_CAPCHK(m, "bump");
  ExStep ex;
  exFillStep(m, ex);
  Batch own = m->batch;
  own.rays = d_rays;
  own.intensities = d_intensities;
  own.timestamps = d_timestamps;
  own.n = (uint32_t)n;
  own.ray_flags = ray_flags;
  own.time_base = m->first_ray_time;
  own.counters = m->d_counters;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  const bool broadcast_rays = exNdt(m);  // every owner evaluates NDT misses of every ray that crosses its regions
  const size_t n_total = (size_t)x.world * x.per;
  CUDA_TRY(cudaMemsetAsync(&m->d_counters->record_count, 0, sizeof(uint32_t) * kPerBatchCounterWords, s));
  CUDA_TRY(cudaMemsetAsync(x.out_counts, 0, sizeof(uint32_t) * 2 * kMaxWorld, s));
  if (n)
  {
    const bool route_now = x.smp_last_exit == nullptr;  // no traversal layer: no exit range to carry first
    {
      KernelScope scope(m, kKExPrepRays);
      exPrepRays<<<blocks, 128, 0, s>>>(m->dm, m->geom, m->mp, own, ex, m->mode, broadcast_rays ? 1 : 0,
                                        route_now ? (exLiteSamples(m) ? 2 : 1) : 0);
    }
    if (!route_now)
    {
      rc = carryLastExit(m, n, s, x.smp_last_exit);
      if (rc)
      {
        return rc;
      }
      KernelScope scope(m, kKExRoute);
      exRouteSamples<<<blocks, 128, 0, s>>>(own, ex);
    }
    EX_CAPCHK(m, "prep+carry+route");
  }
  exSignal<<<1, 32, 0, s>>>(ex, 2);  // the samples are out
  // The per-ray broadcast: copy engines over NVLink, on their own stream, beside the cut.
  const bool broadcast_first = true;  // (sending an occupancy map's 64-byte records AFTER the cut was measured at 8 GPUs: 0.72 vs 0.68 ms per step)
  auto broadcast = [&]() -> int {
  CUDA_TRY(cudaEventRecord(x.prepped, s));
  CUDA_TRY(cudaStreamWaitEvent(x.stream, x.prepped, 0));
  const ExView &mine = ex.peer[x.rank];
  const size_t first = (size_t)x.rank * x.per;
  // Peer (rank + k) in round k: every round is a permutation — each GPU receives from exactly one sender at a time.  (In
  // rank order 0, 1, 2, ... all ranks would copy to the same GPU at once: an incast that serialises the broadcast on one
  // GPU's ingress, eight times over: 0.67 -> 0.55 ms per step at 8 GPUs.)
  // First what the WALK needs (walk constants; ray lengths with a traversal layer), flagged as stage 1; then what only
  // an NDT map's Gaussian-miss and replay kernels read (rays, timestamps, intensities), flagged as stage 3 — the owners
  // wait for that after their walk.
  for (int k = 1; k < x.world && n; ++k)
  {
    const ExView &peer = ex.peer[(x.rank + k) % x.world];
    CUDA_TRY(cudaMemcpyAsync(peer.recs + first, mine.recs + first, n * sizeof(RayRec), cudaMemcpyDeviceToDevice, x.stream));
    if (m->dm.traversal)
    {
      CUDA_TRY(cudaMemcpyAsync(peer.ray_length + first, mine.ray_length + first, n * sizeof(double), cudaMemcpyDeviceToDevice, x.stream));
    }
  }
  exSignal<<<1, 32, 0, x.stream>>>(ex, 1);
  if (broadcast_rays)
  {
    for (int k = 1; k < x.world && n; ++k)
    {
      const ExView &peer = ex.peer[(x.rank + k) % x.world];
      CUDA_TRY(cudaMemcpyAsync(peer.rays + first * 6, mine.rays + first * 6, n * 6 * sizeof(double), cudaMemcpyDeviceToDevice, x.stream));
      if (x.has_timestamps)
      {
        CUDA_TRY(cudaMemcpyAsync(peer.timestamps + first, mine.timestamps + first, n * sizeof(double), cudaMemcpyDeviceToDevice, x.stream));
      }
      if (x.has_intensities)
      {
        CUDA_TRY(cudaMemcpyAsync(peer.intensities + first, mine.intensities + first, n * sizeof(float), cudaMemcpyDeviceToDevice, x.stream));
      }
    }
    exSignal<<<1, 32, 0, x.stream>>>(ex, 3);
  }
  CUDA_TRY(cudaEventRecord(x.bcast_done, x.stream));
  EX_CAPCHK(m, "broadcast");
  return OHMB200_OK;
  };
  if (broadcast_first)
  {
    rc = broadcast();
    if (rc)
    {
      return rc;
    }
  }

  // The sample branch of the OWNER side starts here, beside everybody's cut — unless a peer map lives in this process:
  // its send is queued by this same host thread AFTER this call, and a wait kernel queued now could sit in front of it
  // in a shared hardware queue.  Then ohmb200_exchange_integrate queues the branch (every send has been queued by then).
  x.forked = false;
  if (!x.local_peers)
  {
    rc = exSampleBranch(m, ex);
    if (rc)
    {
      return rc;
    }
    EX_CAPCHK(m, "sample branch");
  }

  if (n)
  {
    KernelScope scope(m, kKExSegments);
    exPrepSegments<<<blocks, 128, 0, s>>>(m->geom, own, ex);
  }
  exSignal<<<1, 32, 0, s>>>(ex, 0);
  m->launches += 3;
  EX_CAPCHK(m, "segments+signal");
  if (!broadcast_first)
  {
    rc = broadcast();
    if (rc)
    {
      return rc;
    }
  }
  CUDA_TRY(cudaGetLastError());
  x.pending = true;
  capture_guard.armed = false;
  return OHMB200_OK;
}

// The same from host memory: the rays are staged through the map's double-buffered input buffers (as ohmb200_integrate).
int exchangeSendHost(ohmb200_map *m, const double *rays, size_t element_count, const float *intensities, const double *timestamps,
                     unsigned ray_flags)
{
  if (!m || (!rays && element_count >= 2))
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_send: bad arguments");
  }
  cudaSetDevice(m->device);
  const size_t n = element_count / 2;
  const int buf = m->next_input;
  m->next_input ^= 1;
  if (!(m->params.layers & (1u << OHMB200_LAYER_INTENSITY)))
  {
    intensities = nullptr;
  }
  if (!(m->params.layers & (1u << OHMB200_LAYER_TOUCH_TIME)))
  {
    timestamps = nullptr;
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
      m->in_capacity[buf] = 0;
      return setError(OHMB200_E_CUDA, "input staging allocation failed");
    }
    m->in_capacity[buf] = cap;
  }
  cudaStreamWaitEvent(m->copy_stream, m->in_free[buf], 0);
  bool ok = n == 0 || cudaMemcpyAsync(m->d_rays[buf], rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, m->copy_stream) == cudaSuccess;
  if (intensities && n)
  {
    ok = ok && cudaMemcpyAsync(m->d_intensities[buf], intensities, sizeof(float) * n, cudaMemcpyHostToDevice, m->copy_stream) == cudaSuccess;
  }
  if (timestamps && n)
  {
    ok = ok && cudaMemcpyAsync(m->d_timestamps[buf], timestamps, sizeof(double) * n, cudaMemcpyHostToDevice, m->copy_stream) == cudaSuccess;
  }
  ok = ok && cudaEventRecord(m->in_ready[buf], m->copy_stream) == cudaSuccess;
  ok = ok && cudaStreamWaitEvent(m->stream, m->in_ready[buf], 0) == cudaSuccess;
  if (!ok)
  {
    return setError(OHMB200_E_CUDA, "ray upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  const int rc = exchangeSend(m, m->d_rays[buf], 2 * n, intensities ? m->d_intensities[buf] : nullptr,
                              timestamps ? m->d_timestamps[buf] : nullptr, ray_flags);
  // the staged rays are read by the send kernels only; their buffer is released when the step has been queued whole
  // (ohmb200_exchange_integrate: a step being recorded as a graph must not swallow the event)
  m->ex.host_buf = (rc == OHMB200_OK) ? buf : -1;
  if (rc != OHMB200_OK)
  {
    cudaEventRecord(m->in_free[buf], m->stream);
  }
  cudaEventSynchronize(m->in_ready[buf]);
  EX_CAPCHK(m, "upload wait");
  return rc;
}

// Phase 2: wait for every rank's records of this step and integrate what was routed here.  Returns after queueing.
int exchangeIntegrate(ohmb200_map *m)
{
  if (!m)
  {
    return setError(OHMB200_E_INVALID, "null map");
  }
  ohmb200_map::Exchange &x = m->ex;
  if (!x.open || !x.pending)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_integrate: no step pending (call ohmb200_exchange_send first)");
  }
  cudaSetDevice(m->device);
  x.pending = false;
  cudaStream_t s = m->stream;
  auto finish = [&]() {
    if (x.host_buf >= 0)
    {
      cudaEventRecord(m->in_free[x.host_buf], s);
      x.host_buf = -1;
    }
    m->rays_in += x.n_own;
    ++m->batches;
    snapshotRegionCount(m);
  };
  if (x.replayed)
  {
    x.replayed = false;
    finish();  // the step's graph holds both phases
    return OHMB200_OK;
  }
  ExStep ex;
  exFillStep(m, ex);
  const ExView &mine = ex.peer[x.rank];
  const size_t n_total = (size_t)x.world * x.per;
  Batch &b = m->batch;
  b.rays = mine.rays;
  b.recs = mine.recs;
  b.ray_length = m->dm.traversal ? mine.ray_length : nullptr;
  b.intensities = x.has_intensities ? mine.intensities : nullptr;
  b.timestamps = (x.has_timestamps && m->dm.touch_time) ? mine.timestamps : nullptr;
  b.n = (uint32_t)n_total;
  b.heavy_run = m->heavy_run;
  b.ray_flags = x.ray_flags;
  b.stamp = ++m->stamp;
  b.time_base = m->first_ray_time;
  b.counters = m->d_counters;
  b.stage_by_ray = 0;
  const unsigned bin_grid = (unsigned)m->sm_count * 8u;

  CUDA_TRY(cudaMemsetAsync(b.seg_count, 0, sizeof(uint32_t) * m->dm.capacity, s));
  CUDA_TRY(cudaMemsetAsync(b.seg_cursor, 0, sizeof(uint32_t) * m->dm.capacity, s));
  CUDA_TRY(cudaMemsetAsync(b.record_vid, 0xFF, sizeof(uint32_t) * b.record_capacity, s));
  CUDA_TRY(cudaMemsetAsync(b.run_head, 0xFF, sizeof(int32_t) * n_total, s));
  b.tail_overflow = b.interval_count + n_total;
  CUDA_TRY(cudaMemsetAsync(b.interval_count, 0, sizeof(uint32_t) * (2 * n_total + 1), s));
  if (x.local_peers)
  {
    int rc_branch = exSampleBranch(m, ex);
    if (rc_branch)
    {
      return rc_branch;
    }
  }
  {
    KernelScope scope(m, kKExWait);
    exWait<<<1, 32, 0, s>>>(mine.mailbox, x.world, x.d_step, 0, x.abort);
  }
  {
    KernelScope scope(m, kKExBin);
    exBinSegments<<<bin_grid, 256, 0, s>>>(m->dm, b, ex);
  }
  const bool fork = x.forked;
  if (!m->store.empty())
  {
    // regions this step brought back (the sample branch creates regions too): restore their chunks before anything
    // updates them
    if (fork)
    {
      CUDA_TRY(cudaStreamWaitEvent(s, m->join_event, 0));
    }
    int rc = pageInNewRegions(m);
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
    KernelScope scope(m, kKExEmit);
    exEmit<<<bin_grid, 256, 0, s>>>(b, ex);
  }
  if (fork)
  {
    CUDA_TRY(cudaStreamWaitEvent(s, m->join_event, 0));  // the sample branch (queued by ohmb200_exchange_send)
  }
  {
    KernelScope scope(m, kKExWait);
    exWait<<<1, 32, 0, s>>>(mine.mailbox, x.world, x.d_step, 1, x.abort);  // the walk constants of every rank's rays
  }
  x.mailbox_now = mine.mailbox;
  int rc = launchWalkAndReplay(m, b, s, n_total, true);
  CUDA_TRY(cudaStreamWaitEvent(s, x.bcast_done, 0));  // (long done: joins the broadcast stream for a recorded step)
  if (x.capturing)
  {
    x.capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(s, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == OHMB200_OK && end == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
    {
      cudaGraphDestroy(graph);
      x.current.exec = exec;
      x.current.launches = m->launches - x.launches_before;
      x.graphs.push_back(x.current);
      CUDA_TRY(cudaGraphLaunch(exec, s));
    }
    else
    {
      if (graph)
      {
        cudaGraphDestroy(graph);
      }
      cudaGetLastError();
      m->use_graphs = false;
      return setError(OHMB200_E_CUDA, "exchange: the step could not be recorded as a CUDA graph (graphs are now off; the step "
                                      "was NOT integrated: send it again)");
    }
  }
  if (rc)
  {
    return rc;
  }
  CUDA_TRY(cudaGetLastError());
  finish();
  return OHMB200_OK;
}

// A barrier between the ranks ON THE DEVICE: every rank's stream passes it only when every rank's stream has reached
// it (mailbox flag + bounded spin, like a step's handshake).  For callers that want the ranks to enter a step together
// without a host round trip — bench.py puts it in front of every timed step.
__global__ void exBarrierSignal(ExStep ex, uint32_t *counter)
{
  const int o = (int)threadIdx.x;
  if (o == 0)
  {
    *counter += 1u;
  }
  __syncthreads();
  if (o < ex.world)
  {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&ex.peer[o].mailbox[ex.rank].barrier) = *counter;
  }
}

__global__ void exBarrierWait(ExMailbox *mailbox, int world, const uint32_t *counter, int *abort)
{
  const int s = (int)threadIdx.x;
  const uint32_t want = *counter;
  bool ok = true;
  if (s < world)
  {
    const volatile uint32_t *flag = &mailbox[s].barrier;
    const long long t0 = clock64();
    while ((int32_t)(*flag - want) < 0)
    {
      if (clock64() - t0 > 8000000000ll)
      {
        ok = false;
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
  if (!ok)
  {
    *abort = 1;
  }
}

int exchangeBarrier(ohmb200_map *m)
{
  if (!m || !m->ex.open || !m->ex.connected || m->ex.pending)
  {
    return setError(OHMB200_E_INVALID, "ohmb200_exchange_barrier: needs an open, connected exchange with no step pending");
  }
  cudaSetDevice(m->device);
  ohmb200_map::Exchange &x = m->ex;
  // the barrier's flags live in parity 0 of the mailboxes (its own word: ExMailbox::barrier), whatever the step parity
  ExStep ex;
  exFillStep(m, ex);
  const ExLayout l = exLayout(x.world, x.per, x.seg_cap);
  for (int r = 0; r < x.world; ++r)
  {
    ex.peer[r] = exView(x.peer_base[r], l, x.parity_bytes, 0);
  }
  exBarrierSignal<<<1, 32, 0, m->stream>>>(ex, x.d_barrier);
  exBarrierWait<<<1, 32, 0, m->stream>>>(ex.peer[x.rank].mailbox, x.world, x.d_barrier, x.abort);
  m->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return OHMB200_OK;
}
