// GpuMap.hpp — C++ facade over the C ABI (include/ohmb200.h) with the public surface of ohm's GPU mappers.
//
// Mirrors, member for member, the calls OhmAppGpu and the reference tests make on:
//   ohm::RayMapper   ohm/RayMapper.h:22-65
//   ohm::GpuMap      ohmgpu/GpuMap.h:143-303   (gpuOk, valid, integrateRays, syncVoxels, hit/missValue, ...)
//   ohm::GpuNdtMap   ohmgpu/GpuNdtMap.h:63-110 (setSensorNoise, sensorNoise)
//   ohm::GpuTsdfMap  ohmgpu/GpuTsdfMap.h:37-80 (setTsdfOptions, maxWeight, ...)
//
// Inside the ohm source tree the same classes keep their `OccupancyMap *` constructors and forward to this facade;
// INTEGRATION.md shows that binding (chunk sync into MapChunk::voxel_blocks included).  Standalone, the map lives
// on the device and is read back with regionKeys()/readRegion().  Header-only; link with -lohmb200.
#ifndef OHMB200_GPUMAP_HPP
#define OHMB200_GPUMAP_HPP

#include "../ohmb200.h"

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__has_include)
#if __has_include(<glm/vec3.hpp>)
#include <glm/vec3.hpp>
#define OHMB200_HAVE_GLM 1
#endif
#endif

#ifndef OHMB200_HAVE_GLM
// Layout-compatible stand-ins for the three GLM types that appear in the RayMapper/GpuMap signatures.
namespace glm
{
struct dvec3
{
  double x, y, z;
  dvec3() : x(0), y(0), z(0) {}
  explicit dvec3(double s) : x(s), y(s), z(s) {}
  dvec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
struct u8vec3
{
  uint8_t x, y, z;
  u8vec3() : x(0), y(0), z(0) {}
  explicit u8vec3(uint8_t s) : x(s), y(s), z(s) {}
  u8vec3(uint8_t x_, uint8_t y_, uint8_t z_) : x(x_), y(y_), z(z_) {}
};
struct i16vec3
{
  int16_t x, y, z;
  i16vec3() : x(0), y(0), z(0) {}
  i16vec3(int16_t x_, int16_t y_, int16_t z_) : x(x_), y(y_), z(z_) {}
};
}  // namespace glm
#endif  // OHMB200_HAVE_GLM

namespace ohm
{
static_assert(sizeof(glm::dvec3) == 3 * sizeof(double), "glm::dvec3 must be three packed doubles");

/// ohm/RayFlag.h:16-60
enum RayFlag : unsigned
{
  kRfDefault = 0,
  kRfEndPointAsFree = (1u << 0u),
  kRfStopOnFirstOccupied = (1u << 1u),
  kRfExcludeOrigin = (1u << 2u),
  kRfExcludeSample = (1u << 3u),
  kRfExcludeRay = (1u << 4u),
  kRfExcludeUnobserved = (1u << 5u),
  kRfExcludeFree = (1u << 6u),
  kRfExcludeOccupied = (1u << 7u),
  kRfReverseWalk = (1u << 8u)
};

/// ohm/MapFlag.h — the layer-selecting subset.
enum class MapFlag : unsigned
{
  kNone = 0,
  kVoxelMean = (1u << 0u),
  kTraversal = (1u << 2u),
  kTouchTime = (1u << 3u),
  kIncidentNormal = (1u << 4u),
  kTsdf = (1u << 5u)
};
inline MapFlag operator|(MapFlag a, MapFlag b)
{
  return MapFlag(unsigned(a) | unsigned(b));
}
inline bool any(MapFlag a, MapFlag b)
{
  return (unsigned(a) & unsigned(b)) != 0u;
}

/// ohm/NdtMode.h
enum class NdtMode
{
  kNone,
  kOccupancy,
  kTraversability
};

/// ohm/VoxelTsdf.h:27-37
struct TsdfOptions
{
  float max_weight = 1e4f;
  float default_truncation_distance = 0.1f;
  float dropoff_epsilon = 0.0f;
  float sparsity_compensation_factor = 1.0f;
};

/// ohm/RayMapper.h:22-65
class RayMapper
{
public:
  RayMapper() = default;
  virtual ~RayMapper() = default;
  virtual bool valid() const = 0;
  virtual size_t integrateRays(const glm::dvec3 *rays, size_t element_count, const float *intensities,
                               const double *timestamps, unsigned ray_update_flags) = 0;
  virtual inline size_t integrateRays(const glm::dvec3 *rays, size_t element_count)
  {
    return integrateRays(rays, element_count, nullptr, nullptr, kRfDefault);
  }
};

/// ohmgpu/GpuMap.h:143-303 over a device-resident map.
class GpuMap : public RayMapper
{
public:
  /// Standalone construction: the OccupancyMap(resolution, region_dim, flags) + GpuMap(map, borrowed,
  /// expected_element_count, gpu_mem_size) pair of the reference (OccupancyMap.cpp:192, GpuMap.h:167-169).
  explicit GpuMap(double resolution, const glm::u8vec3 &region_voxel_dimensions = glm::u8vec3(0, 0, 0),
                  MapFlag flags = MapFlag::kNone, size_t gpu_mem_size = 0, int device = 0)
    : GpuMap(resolution, region_voxel_dimensions, flags, gpu_mem_size, device, OHMB200_MODE_OCCUPANCY)
  {}

  ~GpuMap() override { ohmb200_destroy(map_); }
  GpuMap(const GpuMap &) = delete;
  GpuMap &operator=(const GpuMap &) = delete;

  /// GpuMap.h:181-185. False when no sm_100 device is usable; every call is then a no-op (GpuMap.cpp:548-551).
  bool gpuOk() const { return map_ != nullptr; }
  bool valid() const override { return gpuOk(); }

  using RayMapper::integrateRays;
  /// GpuMap.cpp:540-875
  size_t integrateRays(const glm::dvec3 *rays, size_t element_count, const float *intensities,
                       const double *timestamps, unsigned ray_update_flags) override
  {
    if (!map_ || !rays || element_count < 2)
    {
      return 0u;
    }
    return ohmb200_integrate(map_, reinterpret_cast<const double *>(rays), element_count, intensities, timestamps,
                             ray_update_flags);
  }

  /// GpuMap.h:199 — wait for the device; voxel data stays resident (read with readRegion()).
  void syncVoxels()
  {
    if (map_)
    {
      ohmb200_sync(map_);
    }
  }

  float hitValue() const { return params_.hit_value; }
  float missValue() const { return params_.miss_value; }
  void setHitValue(float value)
  {
    params_.hit_value = value;
    push();
  }
  void setMissValue(float value)
  {
    params_.miss_value = value;
    push();
  }
  void setHitProbability(float p) { setHitValue(std::log(p / (1.0f - p))); }
  void setMissProbability(float p) { setMissValue(std::log(p / (1.0f - p))); }
  void setMinVoxelValue(float v)
  {
    params_.min_value = v;
    push();
  }
  void setMaxVoxelValue(float v)
  {
    params_.max_value = v;
    push();
  }
  void setSaturateAtMinValue(bool s)
  {
    params_.saturate_min = s;
    push();
  }
  void setSaturateAtMaxValue(bool s)
  {
    params_.saturate_max = s;
    push();
  }
  double resolution() const { return params_.resolution; }
  void setOrigin(const glm::dvec3 &origin)
  {
    params_.origin[0] = origin.x;
    params_.origin[1] = origin.y;
    params_.origin[2] = origin.z;
    push();
  }

  /// GpuMap.h:214-228: the filters ohm ships (ohm/RayFilter.h) by kind; arbitrary std::function filters run on the
  /// host in the in-tree binding before the call.
  void setRayFilterGoodRay(double max_range)
  {
    params_.filter_kind = OHMB200_FILTER_GOOD_RAY;
    params_.filter_range = max_range;
    push();
  }
  void setRayFilterClipRange(double max_length)
  {
    params_.filter_kind = OHMB200_FILTER_CLIP_RANGE;
    params_.filter_range = max_length;
    push();
  }
  void clearRayFilter()
  {
    params_.filter_kind = OHMB200_FILTER_NONE;
    push();
  }

  /// GpuMap.h:248-262.  Long rays need no host-side segmenting here (the walk can be resumed at any step from the
  /// per-axis step counts), so the value is stored for API compatibility only.
  double raySegmentLength() const { return ray_segment_length_; }
  void setRaySegmentLength(double length) { ray_segment_length_ = length; }
  /// GpuMap.h:271: grouping by sample voxel always happens on the device.
  bool groupedRays() const { return true; }

  double firstRayTime() const { return map_ ? ohmb200_first_ray_time(map_) : -1.0; }
  size_t regionCount() const { return map_ ? ohmb200_region_count(map_) : 0u; }
  std::vector<glm::i16vec3> regionKeys() const
  {
    std::vector<glm::i16vec3> keys(regionCount());
    if (!keys.empty())
    {
      keys.resize(ohmb200_enumerate_regions(map_, reinterpret_cast<int16_t *>(keys.data()), keys.size()));
    }
    return keys;
  }
  size_t regionLayerBytes(ohmb200_layer layer) const { return map_ ? ohmb200_region_layer_bytes(map_, layer) : 0u; }
  /// GpuLayerCache::syncToExternal (ohmgpu/GpuLayerCache.h:291)
  bool readRegion(const glm::i16vec3 &key, ohmb200_layer layer, void *dst, size_t bytes) const
  {
    return map_ && ohmb200_read_region(map_, reinterpret_cast<const int16_t *>(&key), layer, dst, bytes) == OHMB200_OK;
  }
  /// Paging (GpuLayerCache's eviction / upload on a miss, ohmgpu/GpuLayerCache.cpp:429-633): the regions that do not
  /// fit in device memory live in a host-side store inside the library; every read covers both halves.
  struct PagingStats
  {
    uint64_t resident = 0, stored = 0, evicted = 0, paged_in = 0;
  };
  PagingStats pagingStats() const
  {
    PagingStats s;
    if (map_)
    {
      ohmb200_paging_stats(map_, &s.resident, &s.stored, &s.evicted, &s.paged_in);
    }
    return s;
  }
  /// MapRegionCache::remove (ohm/MapRegionCache.h:52)
  bool removeRegion(const glm::i16vec3 &key)
  {
    return map_ && ohmb200_remove_region(map_, reinterpret_cast<const int16_t *>(&key)) == OHMB200_OK;
  }
  /// Free region slots guaranteed before every batch (the gpu_mem_size counterpart for regions created per batch).
  bool setRegionReserve(uint32_t free_slots) { return map_ && ohmb200_set_region_reserve(map_, free_slots) == OHMB200_OK; }
  /// GpuCache::clear + OccupancyMap::clear
  void clear()
  {
    if (map_)
    {
      ohmb200_clear(map_);
    }
  }

  ohmb200_map *handle() const { return map_; }
  static std::string lastError() { return ohmb200_last_error(); }

protected:
  GpuMap(double resolution, const glm::u8vec3 &dim, MapFlag flags, size_t gpu_mem_size, int device, int mode)
  {
    ohmb200_default_params(&params_, resolution);
    if (dim.x && dim.y && dim.z)
    {
      params_.region_dim[0] = dim.x;
      params_.region_dim[1] = dim.y;
      params_.region_dim[2] = dim.z;
    }
    params_.layers |= any(flags, MapFlag::kVoxelMean) ? (1u << OHMB200_LAYER_MEAN) : 0u;
    params_.layers |= any(flags, MapFlag::kTraversal) ? (1u << OHMB200_LAYER_TRAVERSAL) : 0u;
    params_.layers |= any(flags, MapFlag::kTouchTime) ? (1u << OHMB200_LAYER_TOUCH_TIME) : 0u;
    params_.layers |= any(flags, MapFlag::kIncidentNormal) ? (1u << OHMB200_LAYER_INCIDENT) : 0u;
    map_ = ohmb200_create(&params_, mode, gpu_mem_size, device);
    if (map_)
    {
      ohmb200_get_params(map_, &params_);
    }
  }
  void push()
  {
    if (map_)
    {
      ohmb200_set_params(map_, &params_);
    }
  }

  ohmb200_map *map_ = nullptr;
  ohmb200_params params_{};
  double ray_segment_length_ = 0;
};

/// ohmgpu/GpuNdtMap.h:63-110
class GpuNdtMap : public GpuMap
{
public:
  explicit GpuNdtMap(double resolution, const glm::u8vec3 &region_voxel_dimensions = glm::u8vec3(0, 0, 0),
                     MapFlag flags = MapFlag::kNone, size_t gpu_mem_size = 0, NdtMode mode = NdtMode::kOccupancy,
                     int device = 0)
    : GpuMap(resolution, region_voxel_dimensions, flags, gpu_mem_size, device,
             mode == NdtMode::kTraversability ? OHMB200_MODE_NDT_TM : OHMB200_MODE_NDT)
  {}
  void setSensorNoise(float noise_range)
  {
    params_.sensor_noise = noise_range;
    push();
  }
  float sensorNoise() const { return params_.sensor_noise; }
};

/// ohmgpu/GpuTsdfMap.h:37-80
class GpuTsdfMap : public GpuMap
{
public:
  explicit GpuTsdfMap(double resolution, const glm::u8vec3 &region_voxel_dimensions = glm::u8vec3(0, 0, 0),
                      size_t gpu_mem_size = 0, int device = 0)
    : GpuMap(resolution, region_voxel_dimensions, MapFlag::kNone, gpu_mem_size, device, OHMB200_MODE_TSDF)
  {}
  void setTsdfOptions(const TsdfOptions &o)
  {
    params_.tsdf_max_weight = o.max_weight;
    params_.tsdf_trunc = o.default_truncation_distance;
    params_.tsdf_dropoff = o.dropoff_epsilon;
    params_.tsdf_sparsity = o.sparsity_compensation_factor;
    push();
  }
  TsdfOptions tsdfOptions() const
  {
    TsdfOptions o;
    o.max_weight = params_.tsdf_max_weight;
    o.default_truncation_distance = params_.tsdf_trunc;
    o.dropoff_epsilon = params_.tsdf_dropoff;
    o.sparsity_compensation_factor = params_.tsdf_sparsity;
    return o;
  }
  void setMaxWeight(float v)
  {
    params_.tsdf_max_weight = v;
    push();
  }
  float maxWeight() const { return params_.tsdf_max_weight; }
  void setDefaultTruncationDistance(float v)
  {
    params_.tsdf_trunc = v;
    push();
  }
  float defaultTruncationDistance() const { return params_.tsdf_trunc; }
  void setDropoffEpsilon(float v)
  {
    params_.tsdf_dropoff = v;
    push();
  }
  float dropoffEpsilon() const { return params_.tsdf_dropoff; }
  void setSparsityCompensationFactor(float v)
  {
    params_.tsdf_sparsity = v;
    push();
  }
  float sparsityCompensationFactor() const { return params_.tsdf_sparsity; }
};

/// ohm::OccupancyType (ohm/OccupancyType.h:14-24).
enum OccupancyType
{
  kNull = -2,
  kUnobserved = -1,
  kFree = 0,
  kOccupied = 1
};

/// ohm::RaysQuery / ohm::RaysQueryGpu (ohm/RaysQuery.h:42-139, ohmgpu/RaysQueryGpu.h) on a GpuMap: for each ray, the range
/// to the first occupied voxel, the unobserved volume along it and the state of the voxel it ends in.  The map is
/// queried where it lives (no syncVoxels needed); execute() runs after every batch already handed to integrateRays().
class RaysQueryGpu
{
public:
  explicit RaysQueryGpu(GpuMap *map = nullptr) : map_(map) {}

  void setMap(GpuMap *map) { map_ = map; }
  GpuMap *map() const { return map_; }

  void setVolumeCoefficient(double coefficient) { volume_coefficient_ = coefficient; }
  double volumeCoefficient() const { return volume_coefficient_; }

  void setRays(const glm::dvec3 *rays, size_t element_count)
  {
    rays_.clear();
    addRays(rays, element_count);
  }
  void addRays(const glm::dvec3 *rays, size_t element_count)
  {
    rays_.insert(rays_.end(), rays, rays + (element_count & ~size_t(1)));
  }
  void addRay(const glm::dvec3 &origin, const glm::dvec3 &end_point)
  {
    rays_.push_back(origin);
    rays_.push_back(end_point);
  }
  void clearRays() { rays_.clear(); }
  const glm::dvec3 *rays(size_t *count = nullptr) const
  {
    if (count)
    {
      *count = rays_.size();
    }
    return rays_.data();
  }
  size_t numberOfRays() const { return rays_.size() / 2; }

  /// Query::execute(): false when there is no map or the device call fails.
  bool execute()
  {
    const size_t n = numberOfRays();
    ranges_.assign(n, 0.0);
    unobserved_volumes_.assign(n, 0.0);
    terminal_states_.assign(n, kNull);
    keys_.assign(6 * n, 0);
    if (!map_ || !map_->gpuOk())
    {
      return false;
    }
    static_assert(sizeof(OccupancyType) == sizeof(int), "terminal states are written as int");
    return n == 0 || ohmb200_rays_query(map_->handle(), &rays_[0].x, rays_.size(), volume_coefficient_, ranges_.data(),
                                        unobserved_volumes_.data(), reinterpret_cast<int *>(terminal_states_.data()),
                                        keys_.data()) == OHMB200_OK;
  }

  size_t numberOfResults() const { return ranges_.size(); }
  const double *ranges() const { return ranges_.data(); }
  const double *unobservedVolumes() const { return unobserved_volumes_.data(); }
  const OccupancyType *terminalOccupancyTypes() const { return terminal_states_.data(); }
  /// Terminal voxel keys, six int32 per ray: region x, y, z then local x, y, z (Query::intersectedVoxels()).
  const int32_t *intersectedVoxels() const { return keys_.data(); }

  void reset(bool hard_reset = false)
  {
    ranges_.clear();
    unobserved_volumes_.clear();
    terminal_states_.clear();
    keys_.clear();
    if (hard_reset)
    {
      rays_.clear();
    }
  }

private:
  GpuMap *map_;
  double volume_coefficient_ = 1.0;
  std::vector<glm::dvec3> rays_;
  std::vector<double> ranges_;
  std::vector<double> unobserved_volumes_;
  std::vector<OccupancyType> terminal_states_;
  std::vector<int32_t> keys_;
};
}  // namespace ohm

#endif  // OHMB200_GPUMAP_HPP
