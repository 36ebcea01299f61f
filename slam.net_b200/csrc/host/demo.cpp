// demo.cpp — the reference's usage pattern (Simulation/MainWindow.xaml.cs:69-72, 136-210) against the C++
// mirror: construct, feed scan segments, read Pose and HoleMap.  Deterministic inputs built from integer
// formulas so tests/test_host_mirror.py can rebuild them bit for bit and check the output with the oracle.
//   usage: coreslam_demo [scans] [rays] [threads] [iters]
#include <cstdio>
#include <cstdlib>

#include "coreslam.hpp"

static float hash01(uint32_t a, uint32_t b) {  // integer hash -> exactly representable float in [0,1)
  uint32_t h = a * 2654435761u ^ (b + 0x9E3779B9u + (a << 6) + (a >> 2));
  h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

int main(int argc, char** argv) {
  const int scans = argc > 1 ? atoi(argv[1]) : 10;
  const int rays = argc > 2 ? atoi(argv[2]) : 180;
  const int T = argc > 3 ? atoi(argv[3]) : 2;
  const int I = argc > 4 ? atoi(argv[4]) : 64;
  try {
    CoreSLAM::CoreSLAMProcessor slam(16.0f, 256, 64, {8.0f, 8.0f, 0.0f}, 0.1f, 0.1f, I, T);
    slam.HoleWidth(1.0f);
    std::vector<float> offsets((size_t)T * I * 3);
    for (int k = 0; k < scans; k++) {
      BaseSLAM::ScanSegment seg;
      seg.Pose = {8.0f + 0.03125f * (float)k, 8.0f - 0.015625f * (float)k, 0.0078125f * (float)k};
      seg.IsLast = true;
      for (int i = 0; i < rays; i++)  // a lumpy room: radius 3 m + 1 m * hash, angle = i * 2pi/rays (float)
        seg.Rays.emplace_back((float)i * (6.2831855f / (float)rays), 3.0f + hash01((uint32_t)i, 7u));
      for (size_t j = 0; j < offsets.size(); j++)
        offsets[j] = (hash01((uint32_t)j, (uint32_t)k + 100u) - 0.5f) * ((j % 3 == 2) ? 0.125f : 0.25f);
      slam.Update({seg}, offsets.data());
      const auto p = slam.Pose();
      printf("scan %d pose %a %a %a distance %d index %d\n", k, p.X, p.Y, p.Z, slam.LastResult().distance,
             slam.LastResult().index);
    }
    uint64_t sum = 0;
    cs_map_checksum(slam.Handle(), &sum);
    printf("map checksum %llu\n", (unsigned long long)sum);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
