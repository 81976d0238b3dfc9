// GpuCacheB200.h — ohm::GpuCache for the libohmb200 backend: the object ohm::OccupancyMap keeps in
// OccupancyMapDetail::gpu_cache (ohm/private/OccupancyMapDetail.h:96) and calls back through ohm::MapRegionCache
// (ohm/MapRegionCache.h:28-72) when the host map is cleared, culled, re-laid-out or copied.  It owns the device map.
// Replaces ohmgpu/GpuCache.h + ohmgpu/GpuLayerCache.h (which are built on gputil); ohmgpu/GpuMap.h only forward-declares
// `class GpuCache`, so this definition slots in behind the unmodified public header.
//
// TEST-SIDE binding: compiled against the reference's own headers by tests/binding/Makefile, never part of libohmb200.so.
#ifndef OHMB200_GPUCACHEB200_H
#define OHMB200_GPUCACHEB200_H

#include <ohm/MapRegionCache.h>

#include <ohmb200.h>

#include <glm/glm.hpp>

#include <cstddef>
#include <vector>

namespace ohm
{
class OccupancyMap;
struct MapChunk;

class GpuCache : public MapRegionCache
{
public:
  /// ohmgpu/GpuCache.h:50 kDefaultTargetMemSize
  static constexpr size_t kDefaultTargetMemSize = size_t(1) << 30;

  GpuCache(OccupancyMap &map, size_t target_gpu_mem_size);
  ~GpuCache() override;

  // ---- ohm::MapRegionCache ----
  void reinitialise() override;   ///< the map's layout changed (OccupancyMap::updateLayout, OccupancyMap.cpp:678)
  void flush() override;          ///< device -> host for every region and layer (GpuCache::flush)
  void clear() override;          ///< drop every device region without a download (OccupancyMap::clear, :638,1163)
  void remove(const glm::i16vec3 &region_coord) override;  ///< OccupancyMap::cullRegions... (:1217)
  bool syncLayerTo(MapChunk &dst_chunk, unsigned dst_layer, const MapChunk &src_chunk, unsigned src_layer) override;
  MapRegionCache *findLayerCache(unsigned layer) override;

  // ---- what GpuMap needs ----
  /// The device map in the given ohmb200_mode, created (and the host map's existing chunks uploaded) on first use; a
  /// change of mode syncs the host map and rebuilds the device map.
  ohmb200_map *device(int mode);
  ohmb200_map *device() const { return device_; }
  int mode() const { return mode_; }
  size_t targetGpuAllocSize() const { return target_mem_; }
  /// Device -> host: all layers, or the listed host layer indices.
  void syncToHost(const std::vector<int> *host_layers = nullptr);
  /// OccupancyMap parameters -> ohmb200_params (hit/miss/min/max/threshold/saturation; NDT / TSDF left to the mapper).
  ohmb200_params &params() { return params_; }
  void pushParams();
  /// Host layer index -> OHMB200_LAYER_* (-1: not a layer this backend integrates into).
  int b200Layer(unsigned host_layer) const;
  /// Set when the device holds updates the host map has not seen.
  void markDirty() { dirty_ = true; }

private:
  void destroyDevice();
  void pullMapParams();
  void uploadHostChunks();

  OccupancyMap &map_;
  size_t target_mem_;
  ohmb200_map *device_ = nullptr;
  int mode_ = -1;
  ohmb200_params params_{};
  bool dirty_ = false;
};
}  // namespace ohm

#endif  // OHMB200_GPUCACHEB200_H
