// GpuMapB200.cpp — the reference-side binding of libohmb200: ohm::GpuMap / GpuNdtMap / GpuTsdfMap, ohm::GpuCache and the
// gpumap:: free functions implemented over the C ABI (include/ohmb200.h), behind the reference's UNMODIFIED public
// headers ohmgpu/GpuMap.h:143-384, ohmgpu/GpuNdtMap.h:63-110, ohmgpu/GpuTsdfMap.h:37-80 — the translation unit a
// maintainer drops into ohmgpu/ in place of GpuMap.cpp, GpuNdtMap.cpp, GpuTsdfMap.cpp, GpuCache.cpp, GpuLayerCache.cpp
// (and the gputil / clu libraries under them).  OhmAppGpu (ohmapp/OhmAppGpu.cpp:187-268) then runs on it unchanged.
//
// TEST-SIDE: built here by tests/binding/Makefile against /root/reference's headers and oracle/_ref's ohm library and
// exercised by tests/test_gpu_binding.py.  It is not part of libohmb200.so and contains no reference code.
#include "GpuCacheB200.h"

#include <ohmgpu/GpuMap.h>
#include <ohmgpu/GpuNdtMap.h>
#include <ohmgpu/GpuTsdfMap.h>

#include <ohm/DefaultLayer.h>
#include <ohm/MapChunk.h>
#include <ohm/MapLayer.h>
#include <ohm/MapLayout.h>
#include <ohm/NdtMap.h>
#include <ohm/OccupancyMap.h>
#include <ohm/VoxelBuffer.h>
#include <ohm/private/OccupancyMapDetail.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

namespace ohm
{
// ---------------------------------------------------------------------------------------------------------------------
// GpuCache: MapRegionCache over the device map
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
int layerFromName(const char *name)
{
  using namespace default_layer;
  const std::string n(name);
  if (n == occupancyLayerName()) return OHMB200_LAYER_OCCUPANCY;
  if (n == meanLayerName()) return OHMB200_LAYER_MEAN;
  if (n == traversalLayerName()) return OHMB200_LAYER_TRAVERSAL;
  if (n == touchTimeLayerName()) return OHMB200_LAYER_TOUCH_TIME;
  if (n == incidentNormalLayerName()) return OHMB200_LAYER_INCIDENT;
  if (n == covarianceLayerName()) return OHMB200_LAYER_COVARIANCE;
  if (n == intensityLayerName()) return OHMB200_LAYER_INTENSITY;
  if (n == hitMissCountLayerName()) return OHMB200_LAYER_HIT_MISS;
  if (n == tsdfLayerName()) return OHMB200_LAYER_TSDF;
  if (n == secondarySamplesLayerName()) return OHMB200_LAYER_SECONDARY;
  return -1;
}
}  // namespace

GpuCache::GpuCache(OccupancyMap &map, size_t target_gpu_mem_size)
  : map_(map)
  , target_mem_(target_gpu_mem_size)
{
  pullMapParams();
}

GpuCache::~GpuCache()
{
  destroyDevice();
}

void GpuCache::destroyDevice()
{
  if (device_)
  {
    ohmb200_destroy(device_);
    device_ = nullptr;
  }
  dirty_ = false;
}

int GpuCache::b200Layer(unsigned host_layer) const
{
  const MapLayout &layout = map_.layout();
  if (host_layer >= layout.layerCount())
  {
    return -1;
  }
  const int layer = layerFromName(layout.layer(host_layer).name());
  return (layer >= 0 && (params_.layers & (1u << layer))) ? layer : -1;
}

void GpuCache::pullMapParams()
{
  // OccupancyMapDetail (ohm/private/OccupancyMapDetail.h:47-84) -> ohmb200_params; NDT / TSDF fields keep what the
  // mappers have set (GpuNdtMap::setSensorNoise, GpuTsdfMap::setTsdfOptions).
  ohmb200_params fresh;
  ohmb200_default_params(&fresh, map_.resolution());
  if (params_.resolution == 0)
  {
    params_ = fresh;
  }
  const glm::u8vec3 dim = map_.regionVoxelDimensions();
  const glm::dvec3 origin = map_.origin();
  params_.resolution = map_.resolution();
  for (int a = 0; a < 3; ++a)
  {
    params_.region_dim[a] = dim[a];
    params_.origin[a] = origin[a];
  }
  params_.hit_value = map_.hitValue();
  params_.miss_value = map_.missValue();
  params_.min_value = map_.minVoxelValue();
  params_.max_value = map_.maxVoxelValue();
  params_.threshold_value = map_.occupancyThresholdValue();
  params_.saturate_min = map_.saturateAtMinValue() ? 1 : 0;
  params_.saturate_max = map_.saturateAtMaxValue() ? 1 : 0;
  params_.layers = 0;
  const MapLayout &layout = map_.layout();
  for (size_t i = 0; i < layout.layerCount(); ++i)
  {
    const int layer = layerFromName(layout.layer(i).name());
    if (layer >= 0)
    {
      params_.layers |= 1u << layer;
    }
  }
}

void GpuCache::pushParams()
{
  const uint32_t layers = params_.layers;
  pullMapParams();
  params_.layers = layers;  // geometry and layers are fixed while the device map lives
  if (device_)
  {
    ohmb200_set_params(device_, &params_);
  }
}

void GpuCache::uploadHostChunks()
{
  // GpuLayerCache::upload (ohmgpu/GpuLayerCache.cpp:172-182) for every chunk the host map already holds: a slab slot is
  // byte-identical to the MapChunk layer block (x + y dx + z dx dy, ohm/MapChunk.h:47-50).
  std::vector<const MapChunk *> chunks;
  map_.enumerateRegions(chunks);
  const MapLayout &layout = map_.layout();
  for (const MapChunk *chunk : chunks)
  {
    const int16_t key[3] = { chunk->region.coord.x, chunk->region.coord.y, chunk->region.coord.z };
    for (unsigned host_layer = 0; host_layer < layout.layerCount(); ++host_layer)
    {
      const int layer = b200Layer(host_layer);
      if (layer < 0)
      {
        continue;
      }
      VoxelBuffer<const VoxelBlock> src(chunk->voxel_blocks[host_layer]);  // retains + uncompresses
      if (ohmb200_write_region(device_, key, layer, src.voxelMemory(), src.voxelMemorySize()) != OHMB200_OK)
      {
        throw std::runtime_error(std::string("ohmb200_write_region: ") + ohmb200_last_error());
      }
    }
  }
}

ohmb200_map *GpuCache::device(int mode)
{
  if (device_ && mode_ == mode)
  {
    return device_;
  }
  if (device_)
  {
    syncToHost();  // a mapper of another kind takes over: the host map is the hand-over point
    destroyDevice();
  }
  pullMapParams();
  device_ = ohmb200_create(&params_, mode, target_mem_, /*device=*/0);
  mode_ = mode;
  if (!device_)
  {
    return nullptr;  // gpuOk() == false: calls become no-ops (ohmgpu/GpuMap.cpp:548-551)
  }
  ohmb200_get_params(device_, &params_);
  if (map_.firstRayTime() >= 0)
  {
    ohmb200_set_first_ray_time(device_, map_.firstRayTime());
  }
  uploadHostChunks();
  return device_;
}

void GpuCache::syncToHost(const std::vector<int> *host_layers)
{
  // GpuLayerCache::syncToMainMemory (ohmgpu/GpuLayerCache.cpp:300-321,670-713) for every cached region.
  if (!device_)
  {
    return;
  }
  if (ohmb200_sync(device_) != OHMB200_OK)
  {
    throw std::runtime_error(std::string("ohmb200_sync: ") + ohmb200_last_error());
  }
  const size_t count = ohmb200_region_count(device_);
  std::vector<int16_t> keys(3 * std::max<size_t>(count, 1));
  const size_t listed = ohmb200_enumerate_regions(device_, keys.data(), count);
  const size_t n = std::min(count, listed);
  const MapLayout &layout = map_.layout();
  const uint64_t stamp = map_.touch();
  const glm::ivec3 dim = map_.regionVoxelDimensions();
  std::vector<char> staging;
  for (unsigned host_layer = 0; host_layer < layout.layerCount(); ++host_layer)
  {
    if (host_layers && std::find(host_layers->begin(), host_layers->end(), int(host_layer)) == host_layers->end())
    {
      continue;
    }
    const int layer = b200Layer(host_layer);
    if (layer < 0)
    {
      continue;
    }
    const size_t chunk_bytes = ohmb200_region_layer_bytes(device_, layer);
    // one gather + one copy per 256 regions, then into the (uncompressed, retained) voxel blocks
    const size_t piece = 256;
    staging.resize(piece * chunk_bytes);
    for (size_t first = 0; first < n; first += piece)
    {
      const size_t m = std::min(piece, n - first);
      if (ohmb200_read_regions(device_, layer, keys.data() + 3 * first, m, staging.data(), staging.size()) != OHMB200_OK)
      {
        throw std::runtime_error(std::string("ohmb200_read_regions: ") + ohmb200_last_error());
      }
      for (size_t i = 0; i < m; ++i)
      {
        const int16_t *k = keys.data() + 3 * (first + i);
        MapChunk *chunk = map_.region(glm::i16vec3(k[0], k[1], k[2]), /*allow_create=*/true);
        VoxelBuffer<VoxelBlock> dst(chunk->voxel_blocks[host_layer]);
        memcpy(dst.voxelMemory(), staging.data() + i * chunk_bytes, std::min(chunk_bytes, dst.voxelMemorySize()));
        chunk->touched_stamps[host_layer].store(stamp, std::memory_order_relaxed);
        chunk->dirty_stamp = stamp;
      }
    }
  }
  // onOccupancyLayerChunkSync (ohmgpu/private/GpuMapDetail.cpp:28-31): first_valid_index from the occupancy layer
  if (layout.occupancyLayer() >= 0 && (!host_layers || std::find(host_layers->begin(), host_layers->end(),
                                                                  layout.occupancyLayer()) != host_layers->end()))
  {
    for (size_t i = 0; i < n; ++i)
    {
      const int16_t *k = keys.data() + 3 * i;
      if (MapChunk *chunk = map_.region(glm::i16vec3(k[0], k[1], k[2]), false))
      {
        chunk->searchAndUpdateFirstValid(dim);
      }
    }
  }
  if (!host_layers)
  {
    dirty_ = false;
  }
}

void GpuCache::reinitialise()
{
  // The layout of the host map changed: what the device holds goes back first, then the device map is rebuilt for the
  // new layer set on its next use (ohmgpu/private/GpuMapDetail.cpp:59-216 reinitialiseGpuCache).
  const int mode = mode_;
  if (device_)
  {
    // the OLD layers are gone from the host layout: nothing to write back to; drop the device copy
    destroyDevice();
  }
  pullMapParams();
  mode_ = -1;
  (void)mode;
}

void GpuCache::flush()
{
  syncToHost();
}

void GpuCache::clear()
{
  if (device_)
  {
    ohmb200_clear(device_);
  }
  dirty_ = false;
}

void GpuCache::remove(const glm::i16vec3 &region_coord)
{
  if (device_)
  {
    const int16_t key[3] = { region_coord.x, region_coord.y, region_coord.z };
    ohmb200_remove_region(device_, key);  // OHMB200_E_NOT_FOUND: the region never reached the device
  }
}

bool GpuCache::syncLayerTo(MapChunk &dst_chunk, unsigned dst_layer, const MapChunk &src_chunk, unsigned src_layer)
{
  // GpuLayerCache::syncLayerTo -> syncToExternal (ohmgpu/GpuLayerCache.h:291): one region's layer into another chunk
  const int layer = b200Layer(src_layer);
  if (!device_ || layer < 0)
  {
    return false;
  }
  ohmb200_sync(device_);
  const int16_t key[3] = { src_chunk.region.coord.x, src_chunk.region.coord.y, src_chunk.region.coord.z };
  VoxelBuffer<VoxelBlock> dst(dst_chunk.voxel_blocks[dst_layer]);
  return ohmb200_read_region(device_, key, layer, dst.voxelMemory(), dst.voxelMemorySize()) == OHMB200_OK;
}

MapRegionCache *GpuCache::findLayerCache(unsigned layer)
{
  return (device_ && b200Layer(layer) >= 0) ? this : nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// gpumap:: free functions (ohmgpu/GpuMap.h:58-113)
// ---------------------------------------------------------------------------------------------------------------------
namespace gpumap
{
GpuCache *enableGpu(OccupancyMap &map)
{
  return enableGpu(map, GpuCache::kDefaultTargetMemSize, kGpuAllowMappedBuffers);
}

GpuCache *enableGpu(OccupancyMap &map, size_t target_gpu_mem_size, unsigned /*gpu_flags*/)
{
  OccupancyMapDetail &detail = *map.detail();
  if (!detail.gpu_cache)
  {
    detail.gpu_cache = new GpuCache(map, target_gpu_mem_size ? target_gpu_mem_size : GpuCache::kDefaultTargetMemSize);
  }
  return static_cast<GpuCache *>(detail.gpu_cache);  // owned by the map detail, deleted with it
}

void sync(OccupancyMap &map)
{
  if (GpuCache *cache = gpuCache(map))
  {
    cache->syncToHost();
  }
}

void sync(OccupancyMap &map, unsigned layer_index)
{
  if (GpuCache *cache = gpuCache(map))
  {
    const std::vector<int> layers{ int(layer_index) };
    cache->syncToHost(&layers);
  }
}

GpuCache *gpuCache(OccupancyMap &map)
{
  return static_cast<GpuCache *>(map.detail()->gpu_cache);
}
}  // namespace gpumap

// ---------------------------------------------------------------------------------------------------------------------
// GpuMap (ohmgpu/GpuMap.h:143-384)
// ---------------------------------------------------------------------------------------------------------------------
struct GpuMapDetail  // ohmgpu/private/GpuMapDetail.h, reduced to what this backend needs
{
  OccupancyMap *map = nullptr;
  bool borrowed_map = true;
  int mode = OHMB200_MODE_OCCUPANCY;
  RayFilterFunction ray_filter;  // set through GpuMap::setRayFilter
  bool custom_ray_filter = false;
  double ray_segment_length = 0;  // stored only: the walk is exact and resumable, nothing to segment
  bool grouped_rays = false;
  std::vector<glm::dvec3> run;    // host-side filter path: a run of filtered rays
  std::vector<float> run_intensities;
  std::vector<double> run_timestamps;
  virtual ~GpuMapDetail()
  {
    if (!borrowed_map)
    {
      delete map;
    }
  }
};

struct GpuNdtMapDetail : GpuMapDetail
{
  NdtMap ndt_map;
  GpuNdtMapDetail(OccupancyMap *map_in, bool borrowed, NdtMode ndt_mode)
    : ndt_map(map_in, true, ndt_mode)  // adds the mean / covariance (/ intensity, hit-miss) layers to the map
  {
    map = map_in;
    borrowed_map = borrowed;
    mode = ndt_mode == NdtMode::kTraversability ? OHMB200_MODE_NDT_TM : OHMB200_MODE_NDT;
  }
};

struct GpuTsdfMapDetail : GpuMapDetail
{
  TsdfOptions tsdf_options;
  GpuTsdfMapDetail(OccupancyMap *map_in, bool borrowed)
  {
    map = map_in;
    borrowed_map = borrowed;
    mode = OHMB200_MODE_TSDF;
  }
};

namespace
{
GpuCache *cacheOf(GpuMapDetail *imp)
{
  return imp->map ? gpumap::gpuCache(*imp->map) : nullptr;
}

ohmb200_map *deviceOf(GpuMapDetail *imp)
{
  GpuCache *cache = cacheOf(imp);
  return cache ? cache->device(imp->mode) : nullptr;
}

// Is `filter` the filter every OccupancyMap is born with (goodRayFilter, range 1e10: ohm/OccupancyMap.cpp:215-218)?
// Then the device filter does the job; any other std::function runs on the host.
bool isDefaultMapFilter(const OccupancyMap &map, const RayFilterFunction &filter)
{
  if (!filter)
  {
    return false;
  }
  static const OccupancyMap probe(1.0);
  return filter.target_type() == probe.rayFilter().target_type();
}
}  // namespace

GpuMap::GpuMap(GpuMapDetail *detail, unsigned expected_element_count, size_t gpu_mem_size)
  : imp_(detail)
{
  setMap(detail->map, detail->borrowed_map, expected_element_count, gpu_mem_size, true);
}

GpuMap::GpuMap(OccupancyMap *map, bool borrowed_map, unsigned expected_element_count, size_t gpu_mem_size)
  : imp_(new GpuMapDetail)
{
  setMap(map, borrowed_map, expected_element_count, gpu_mem_size, true);
}

GpuMap::~GpuMap()
{
  delete imp_;  // deletes the map too when it is not borrowed (ohmgpu/private/GpuMapDetail.cpp:34-40)
}

void GpuMap::setMap(OccupancyMap *map, bool borrowed_map, unsigned /*expected_element_count*/, size_t gpu_mem_size,
                    bool /*force_gpu_program_release*/)
{
  imp_->map = map;
  imp_->borrowed_map = borrowed_map;
  if (map)
  {
    gpumap::enableGpu(*map, gpu_mem_size, gpumap::kGpuAllowMappedBuffers);
    deviceOf(imp_);  // create the device map now: allocation failures surface at construction (GpuMap.h:159-160)
  }
}

bool GpuMap::gpuOk() const
{
  return imp_->map && deviceOf(imp_) != nullptr;
}

OccupancyMap &GpuMap::map()
{
  return *imp_->map;
}

const OccupancyMap &GpuMap::map() const
{
  return *imp_->map;
}

bool GpuMap::borrowedMap() const
{
  return imp_->borrowed_map;
}

void GpuMap::syncVoxels()
{
  if (GpuCache *cache = cacheOf(imp_))
  {
    cache->syncToHost();
    onSyncVoxels(0);
  }
}

void GpuMap::syncVoxels(const std::vector<int> &layer_indices)
{
  if (GpuCache *cache = cacheOf(imp_))
  {
    cache->syncToHost(&layer_indices);
    onSyncVoxels(0);
  }
}

void GpuMap::setRayFilter(const RayFilterFunction &ray_filter)
{
  imp_->ray_filter = ray_filter;
  imp_->custom_ray_filter = true;
}

const RayFilterFunction &GpuMap::rayFilter() const
{
  return imp_->ray_filter;
}

const RayFilterFunction &GpuMap::effectiveRayFilter() const
{
  return (imp_->custom_ray_filter || !imp_->map) ? imp_->ray_filter : imp_->map->rayFilter();  // GpuMap.cpp:356-359
}

void GpuMap::clearRayFilter()
{
  imp_->ray_filter = RayFilterFunction();
  imp_->custom_ray_filter = false;
}

float GpuMap::hitValue() const
{
  return imp_->map->hitValue();
}

void GpuMap::setHitValue(float value)
{
  imp_->map->setHitValue(value);
  cacheOf(imp_)->pushParams();
}

float GpuMap::missValue() const
{
  return imp_->map->missValue();
}

void GpuMap::setMissValue(float value)
{
  imp_->map->setMissValue(value);
  cacheOf(imp_)->pushParams();
}

double GpuMap::raySegmentLength() const
{
  return imp_->ray_segment_length;
}

void GpuMap::setRaySegmentLength(double length)
{
  imp_->ray_segment_length = length;
}

bool GpuMap::groupedRays() const
{
  return imp_->grouped_rays;
}

void GpuMap::setGroupedRays(bool group)
{
  imp_->grouped_rays = group;
}

GpuCache *GpuMap::gpuCache() const
{
  return cacheOf(imp_);
}

size_t GpuMap::integrateRays(const glm::dvec3 *rays, size_t element_count, const float *intensities,
                             const double *timestamps, unsigned region_update_flags)
{
  return integrateRays(rays, element_count, intensities, timestamps, region_update_flags, effectiveRayFilter());
}

size_t GpuMap::integrateRays(const glm::dvec3 *rays, size_t element_count, const float *intensities,
                             const double *timestamps, unsigned region_update_flags, const RayFilterFunction &filter)
{
  ohmb200_map *device = imp_->map ? deviceOf(imp_) : nullptr;
  if (!device || !rays || element_count < 2)
  {
    return 0u;  // GpuMap.cpp:543-551
  }
  GpuCache &cache = *cacheOf(imp_);
  static_assert(sizeof(glm::dvec3) == 3 * sizeof(double), "glm::dvec3 must be three packed doubles");
  imp_->map->touch();
  cache.pushParams();  // hit / miss / clamps may have been changed on the OccupancyMap since the last batch
  if (timestamps)
  {
    imp_->map->updateFirstRayTime(*timestamps);  // GpuMap.cpp:591-595
    ohmb200_set_first_ray_time(device, imp_->map->firstRayTime());
  }
  cache.markDirty();
  ohmb200_params &p = cache.params();
  if (isDefaultMapFilter(*imp_->map, filter) || !filter)
  {
    // the filter every map is born with, or none: the device does it
    const int kind = filter ? OHMB200_FILTER_GOOD_RAY : OHMB200_FILTER_NONE;
    if (p.filter_kind != kind || p.filter_range != 1e10)
    {
      p.filter_kind = kind;
      p.filter_range = 1e10;
      ohmb200_set_params(device, &p);
    }
    return ohmb200_integrate(device, &rays[0].x, element_count, intensities, timestamps, region_update_flags);
  }
  // An arbitrary std::function: run it here, ray by ray (GpuMap.cpp:738-745), and hand the device runs of consecutive
  // rays of one kind — plain, or clipped at the end (their sample voxel takes a miss: kRfEndPointAsFree does exactly
  // that) — in the original order.
  if (p.filter_kind != OHMB200_FILTER_NONE)
  {
    p.filter_kind = OHMB200_FILTER_NONE;
    ohmb200_set_params(device, &p);
  }
  size_t accepted = 0;
  bool run_clipped = false;
  auto flush = [&]() {
    if (!imp_->run.empty())
    {
      const unsigned flags = region_update_flags | (run_clipped ? unsigned(kRfEndPointAsFree) : 0u);
      ohmb200_integrate(device, &imp_->run[0].x, imp_->run.size(), intensities ? imp_->run_intensities.data() : nullptr,
                        timestamps ? imp_->run_timestamps.data() : nullptr, flags);
      imp_->run.clear();
      imp_->run_intensities.clear();
      imp_->run_timestamps.clear();
    }
  };
  for (size_t i = 0; i + 1 < element_count; i += 2)
  {
    glm::dvec3 start = rays[i], end = rays[i + 1];
    unsigned filter_flags = 0;
    if (!filter(&start, &end, &filter_flags))
    {
      continue;
    }
    const bool clipped = (filter_flags & kRffClippedEnd) != 0;
    if (clipped != run_clipped)
    {
      flush();
      run_clipped = clipped;
    }
    imp_->run.push_back(start);
    imp_->run.push_back(end);
    if (intensities)
    {
      imp_->run_intensities.push_back(intensities[i >> 1]);
    }
    if (timestamps)
    {
      imp_->run_timestamps.push_back(timestamps[i >> 1]);
    }
    ++accepted;
  }
  flush();
  return accepted * 2;  // GpuMap.cpp:874
}

// The reference's kernel plumbing: nothing to do on this backend (the virtuals must exist for the vtable).
void GpuMap::cacheGpuProgram(bool, bool, bool) {}
void GpuMap::releaseGpuProgram() {}
void GpuMap::waitOnPreviousOperation(int)
{
  if (ohmb200_map *device = deviceOf(imp_))
  {
    ohmb200_sync(device);
  }
}
void GpuMap::enqueueRegions(int, unsigned) {}
bool GpuMap::enqueueRegion(const glm::i16vec3 &, int)
{
  return true;
}
void GpuMap::finaliseBatch(unsigned) {}
int GpuMap::enableVoxelUpload(int, bool)
{
  return -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// GpuNdtMap (ohmgpu/GpuNdtMap.h:63-110)
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
void pushNdtParams(GpuMapDetail *imp, NdtMap &ndt)
{
  GpuCache *cache = cacheOf(imp);
  if (!cache)
  {
    return;
  }
  ohmb200_params &p = cache->params();
  p.sensor_noise = ndt.sensorNoise();
  p.adaptation_rate = ndt.adaptationRate();
  p.reinit_threshold = ndt.reinitialiseCovarianceThreshold();
  p.reinit_count = ndt.reinitialiseCovariancePointCount();
  p.sample_threshold = ndt.ndtSampleThreshold();
  p.initial_intensity_cov = ndt.initialIntensityCovariance();
  cache->pushParams();
}
}  // namespace

GpuNdtMap::GpuNdtMap(OccupancyMap *map, bool borrowed_map, unsigned expected_element_count, size_t gpu_mem_size,
                     NdtMode ndt_mode)
  : GpuMap(new GpuNdtMapDetail(map, borrowed_map, ndt_mode), expected_element_count, gpu_mem_size)
{
  pushNdtParams(imp_, detail()->ndt_map);
  setGroupedRays(true);
}

GpuNdtMap::~GpuNdtMap() = default;

void GpuNdtMap::setSensorNoise(float noise_range)
{
  detail()->ndt_map.setSensorNoise(noise_range);
  pushNdtParams(imp_, detail()->ndt_map);
}

float GpuNdtMap::sensorNoise() const
{
  return detail()->ndt_map.sensorNoise();
}

NdtMap &GpuNdtMap::ndtMap()
{
  return detail()->ndt_map;
}

const NdtMap &GpuNdtMap::ndtMap() const
{
  return detail()->ndt_map;
}

GpuNdtMapDetail *GpuNdtMap::detail()
{
  return static_cast<GpuNdtMapDetail *>(imp_);
}

const GpuNdtMapDetail *GpuNdtMap::detail() const
{
  return static_cast<const GpuNdtMapDetail *>(imp_);
}

void GpuNdtMap::cacheGpuProgram(bool, bool, bool) {}
void GpuNdtMap::finaliseBatch(unsigned) {}
void GpuNdtMap::releaseGpuProgram() {}

// ---------------------------------------------------------------------------------------------------------------------
// GpuTsdfMap (ohmgpu/GpuTsdfMap.h:37-80)
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
void pushTsdfParams(GpuMapDetail *imp, const TsdfOptions &o)
{
  GpuCache *cache = cacheOf(imp);
  if (!cache)
  {
    return;
  }
  ohmb200_params &p = cache->params();
  p.tsdf_max_weight = o.max_weight;
  p.tsdf_trunc = o.default_truncation_distance;
  p.tsdf_dropoff = o.dropoff_epsilon;
  p.tsdf_sparsity = o.sparsity_compensation_factor;
  cache->pushParams();
}
}  // namespace

GpuTsdfMap::GpuTsdfMap(OccupancyMap *map, bool borrowed_map, unsigned expected_element_count, size_t gpu_mem_size)
  : GpuMap(new GpuTsdfMapDetail(map, borrowed_map), expected_element_count, gpu_mem_size)
{
  pushTsdfParams(imp_, detail()->tsdf_options);
}

GpuTsdfMap::~GpuTsdfMap() = default;

void GpuTsdfMap::setTsdfOptions(const TsdfOptions &options)
{
  detail()->tsdf_options = options;
  pushTsdfParams(imp_, options);
}

const TsdfOptions &GpuTsdfMap::tsdfOptions() const
{
  return detail()->tsdf_options;
}

void GpuTsdfMap::setMaxWeight(float max_weight)
{
  detail()->tsdf_options.max_weight = max_weight;
  pushTsdfParams(imp_, detail()->tsdf_options);
}

float GpuTsdfMap::maxWeight() const
{
  return detail()->tsdf_options.max_weight;
}

void GpuTsdfMap::setDefaultTruncationDistance(float default_truncation_distance)
{
  detail()->tsdf_options.default_truncation_distance = default_truncation_distance;
  pushTsdfParams(imp_, detail()->tsdf_options);
}

float GpuTsdfMap::defaultTruncationDistance() const
{
  return detail()->tsdf_options.default_truncation_distance;
}

void GpuTsdfMap::setDropoffEpsilon(float dropoff_epsilon)
{
  detail()->tsdf_options.dropoff_epsilon = dropoff_epsilon;
  pushTsdfParams(imp_, detail()->tsdf_options);
}

float GpuTsdfMap::dropoffEpsilon() const
{
  return detail()->tsdf_options.dropoff_epsilon;
}

void GpuTsdfMap::setSparsityCompensationFactor(float sparsity_compensation_factor)
{
  detail()->tsdf_options.sparsity_compensation_factor = sparsity_compensation_factor;
  pushTsdfParams(imp_, detail()->tsdf_options);
}

float GpuTsdfMap::sparsityCompensationFactor() const
{
  return detail()->tsdf_options.sparsity_compensation_factor;
}

GpuTsdfMapDetail *GpuTsdfMap::detail()
{
  return static_cast<GpuTsdfMapDetail *>(imp_);
}

const GpuTsdfMapDetail *GpuTsdfMap::detail() const
{
  return static_cast<const GpuTsdfMapDetail *>(imp_);
}

void GpuTsdfMap::cacheGpuProgram(bool, bool, bool) {}
void GpuTsdfMap::finaliseBatch(unsigned) {}
}  // namespace ohm
