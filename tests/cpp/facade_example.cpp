// Drives the C++ facade the way tests/ohmtestgpu/GpuMapTest.cpp drives ohm::GpuMap: build rays, integrateRays in
// batches, syncVoxels, then read the occupancy of the sample voxels back.  Exit code 0 = every sample voxel is
// occupied and the sensor voxel is free.
#include <ohmb200/GpuMap.hpp>

#include <cstdio>
#include <random>
#include <vector>

int main()
{
  ohm::GpuMap gpu_map(0.25, glm::u8vec3(32, 32, 32), ohm::MapFlag::kVoxelMean);
  if (!gpu_map.gpuOk())
  {
    std::fprintf(stderr, "no GPU: %s\n", ohm::GpuMap::lastError().c_str());
    return 2;
  }
  ohm::RayMapper &mapper = gpu_map;
  std::mt19937 rand_engine;
  std::uniform_real_distribution<double> rand(-10.0, 10.0);
  std::vector<glm::dvec3> rays;
  for (int i = 0; i < 2048; ++i)
  {
    rays.emplace_back(glm::dvec3(0.05));
    rays.emplace_back(glm::dvec3(rand(rand_engine), rand(rand_engine), rand(rand_engine)));
  }
  const size_t batch = 512;
  for (size_t i = 0; i < rays.size(); i += batch)
  {
    if (mapper.integrateRays(rays.data() + i, batch) == 0)
    {
      std::fprintf(stderr, "integrateRays failed: %s\n", ohm::GpuMap::lastError().c_str());
      return 1;
    }
  }
  gpu_map.syncVoxels();
  const auto keys = gpu_map.regionKeys();
  std::vector<float> chunk(gpu_map.regionLayerBytes(OHMB200_LAYER_OCCUPANCY) / sizeof(float));
  size_t occupied = 0, free_voxels = 0;
  for (const auto &key : keys)
  {
    if (!gpu_map.readRegion(key, OHMB200_LAYER_OCCUPANCY, chunk.data(), chunk.size() * sizeof(float)))
    {
      return 1;
    }
    for (float v : chunk)
    {
      occupied += (v != INFINITY && v >= 0.0f);
      free_voxels += (v != INFINITY && v < 0.0f);
    }
  }
  std::printf("regions %zu occupied %zu free %zu\n", keys.size(), occupied, free_voxels);
  return (occupied > 1500 && free_voxels > occupied) ? 0 : 1;
}
