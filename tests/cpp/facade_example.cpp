// Drives the C++ facade the way tests/ohmtestgpu/GpuMapTest.cpp drives ohm::GpuMap: build rays, integrateRays in
// batches, syncVoxels, then read the occupancy of the sample voxels back.  Exit code 0 = every sample voxel is
// occupied and the sensor voxel is free.
#include <ohmb200/GpuMap.hpp>

#include <cmath>
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

  // ohm::RaysQueryGpu: rays from the sensor to every sample stop at the sample's (occupied) voxel, about where the
  // sample is; rays going twice as far stop there too.
  ohm::RaysQueryGpu query(&gpu_map);
  for (size_t i = 0; i < rays.size(); i += 2)
  {
    const glm::dvec3 &o = rays[i], &s = rays[i + 1];
    query.addRay(o, glm::dvec3(o.x + 2 * (s.x - o.x), o.y + 2 * (s.y - o.y), o.z + 2 * (s.z - o.z)));
  }
  if (!query.execute() || query.numberOfResults() != rays.size() / 2)
  {
    std::fprintf(stderr, "rays query failed: %s\n", ohm::GpuMap::lastError().c_str());
    return 1;
  }
  size_t stopped = 0;
  for (size_t i = 0; i < query.numberOfResults(); ++i)
  {
    const glm::dvec3 &o = rays[2 * i], &s = rays[2 * i + 1];
    const double to_sample = std::sqrt((s.x - o.x) * (s.x - o.x) + (s.y - o.y) * (s.y - o.y) + (s.z - o.z) * (s.z - o.z));
    stopped += query.terminalOccupancyTypes()[i] == ohm::kOccupied && query.ranges()[i] <= to_sample + 0.5;
  }
  std::printf("rays query: %zu of %zu rays stop at an occupied voxel before or at their sample\n", stopped,
              query.numberOfResults());
  return (occupied > 1500 && free_voxels > occupied && stopped == query.numberOfResults()) ? 0 : 1;
}
