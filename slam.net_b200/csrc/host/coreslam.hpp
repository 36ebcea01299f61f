// coreslam.hpp — C++ host-side mirror of the reference's public surface over the C ABI.
//
// Same class names, constructor arguments, properties and call order as
//   CoreSLAM/CoreSLAMProcessor.cs:18-775, CoreSLAM/HoleMap.cs:10-57,
//   BaseSLAM/ScanCloud.cs, BaseSLAM/ScanSegment.cs, BaseSLAM/Ray.cs
// of mikkleini/slam.net, so a C# maintainer can read it next to the originals.  Header-only; link with
// libcoreslam_b200.so.  Every call that fails throws std::runtime_error with cs_last_error().
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/coreslam_b200.h"

namespace BaseSLAM {

struct Vector2 { float X = 0, Y = 0; };
struct Vector3 { float X = 0, Y = 0, Z = 0; };

struct Ray {  // BaseSLAM/Ray.cs:10-32
  float Angle, Radius;
  Ray(float angle, float radius) : Angle(angle), Radius(radius) {}
};

struct ScanSegment {  // BaseSLAM/ScanSegment.cs:13-29
  std::vector<Ray> Rays;
  Vector3 Pose;
  bool IsLast = false;
};

struct ScanCloud {  // BaseSLAM/ScanCloud.cs:10-21
  Vector3 Pose;
  std::vector<Vector2> Points;
};

}  // namespace BaseSLAM

namespace CoreSLAM {

using BaseSLAM::Vector2;
using BaseSLAM::Vector3;

class CoreSLAMProcessor;

class HoleMap {  // CoreSLAM/HoleMap.cs:17-55
 public:
  int Size() const { return size_; }
  float Scale() const { return scale_; }
  // HoleMap.Pixels is a public field in the reference; here the map lives on the device and is pulled on demand
  std::vector<uint16_t> Pixels() const;
  std::vector<uint8_t> GetPackedPixels() const;  // HoleMap.cs:44-55

 private:
  friend class CoreSLAMProcessor;
  cs_processor* h_ = nullptr;
  int size_ = 0;
  float scale_ = 0;
};

class CoreSLAMProcessor {  // CoreSLAM/CoreSLAMProcessor.cs
 public:
  // :119-120 (obstacleMapSize is accepted for signature parity; the ObstacleMap stays on the host side)
  CoreSLAMProcessor(float physicalMapSize, int holeMapSize, int /*obstacleMapSize*/, Vector3 startPose, float sigmaXY,
                    float sigmaTheta, int iterationsPerThread, int numSearchThreads, int device = 0,
                    uint64_t seed = 0x5EED, int maxPoints = 0, uint32_t flags = 0)
      : PhysicalMapSize(physicalMapSize), SigmaXY(sigmaXY), SigmaTheta(sigmaTheta),
        SearchIterationsPerThread(iterationsPerThread), NumSearchThreads(numSearchThreads) {
    cs_config cfg{};
    cfg.physical_map_size = physicalMapSize;
    cfg.hole_map_size = holeMapSize;
    cfg.start_pose[0] = startPose.X; cfg.start_pose[1] = startPose.Y; cfg.start_pose[2] = startPose.Z;
    cfg.sigma_xy = sigmaXY;
    cfg.sigma_theta = sigmaTheta;
    cfg.iterations_per_thread = iterationsPerThread;
    cfg.num_search_threads = numSearchThreads;
    cfg.device = device;
    cfg.max_points = maxPoints;
    cfg.seed = seed;
    cfg.flags = flags;
    cs_status st = cs_create(&cfg, &h_);
    if (st != CS_OK) throw std::runtime_error(std::string("cs_create: ") + cs_last_error(nullptr));
    pose_ = startPose;
    map_.h_ = h_;
    cs_get_map_info(h_, &map_.size_, &map_.scale_);
  }
  CoreSLAMProcessor(const CoreSLAMProcessor&) = delete;
  CoreSLAMProcessor& operator=(const CoreSLAMProcessor&) = delete;
  ~CoreSLAMProcessor() { Dispose(); }

  // properties :40-106
  const float PhysicalMapSize, SigmaXY, SigmaTheta;
  const int SearchIterationsPerThread, NumSearchThreads;
  const HoleMap& Map() const { return map_; }  // "HoleMap" property
  Vector3 Pose() const { return pose_; }
  int Quality() const { return quality_; }
  void Quality(int q) { check(cs_set_quality(h_, q)); quality_ = q; }
  float HoleWidth() const { return holeWidth_; }
  void HoleWidth(float w) { check(cs_set_hole_width(h_, w)); holeWidth_ = w; }
  int PositionSearchBeginning() const { return psb_; }
  void PositionSearchBeginning(int n) { check(cs_set_position_search_beginning(h_, n)); psb_ = n; }
  const cs_result& LastResult() const { return last_; }

  void Reset() {  // :167-175
    check(cs_reset(h_));
    float p[3];
    check(cs_get_pose(h_, p));
    pose_ = {p[0], p[1], p[2]};
  }

  // ScanSegmentsToCloud, :187-207, host twin (Update() below runs it on the device through cs_update_segments)
  static void ScanSegmentsToCloud(const std::vector<BaseSLAM::ScanSegment>& segments, Vector3 odometryPose,
                                  BaseSLAM::ScanCloud& cloud) {
    cloud.Points.clear();
    for (const auto& seg : segments) {
      Vector3 pose{seg.Pose.X - odometryPose.X, seg.Pose.Y - odometryPose.Y, seg.Pose.Z - odometryPose.Z};
      for (const auto& r : seg.Rays)
        cloud.Points.push_back({pose.X + r.Radius * cosf(r.Angle + pose.Z), pose.Y + r.Radius * sinf(r.Angle + pose.Z)});
    }
  }

  // Update, :717-752.  candidateOffsets (T*I x 3 floats) switches on verification mode for this scan;
  // nullptr uses the on-device Philox stream.
  void Update(const std::vector<BaseSLAM::ScanSegment>& segments, const float* candidateOffsets = nullptr) {
    if (segments.empty()) throw std::invalid_argument("Sequence contains no elements");  // segments.Last(), :719
    // flatten List<ScanSegment>: rays (angle, radius) back to back, first-ray index and pose per segment; the cloud
    // (:723) is computed on the device, the odometry pose is the last segment's (:719)
    rays_.clear(); segFirst_.clear(); segPoses_.clear();
    for (const auto& seg : segments) {
      segFirst_.push_back((int32_t)(rays_.size() / 2));
      segPoses_.insert(segPoses_.end(), {seg.Pose.X, seg.Pose.Y, seg.Pose.Z});
      for (const auto& r : seg.Rays) { rays_.push_back(r.Angle); rays_.push_back(r.Radius); }
    }
    segFirst_.push_back((int32_t)(rays_.size() / 2));
    check(cs_update_segments(h_, rays_.data(), segFirst_.data(), segPoses_.data(), (int32_t)(rays_.size() / 2),
                             (int32_t)segments.size(), candidateOffsets, &last_));
    pose_ = {last_.pose[0], last_.pose[1], last_.pose[2]};
  }

  void Dispose() {  // :757-773
    if (h_) cs_destroy(h_);
    h_ = nullptr;
    map_.h_ = nullptr;
  }

  cs_processor* Handle() const { return h_; }

  // Candidate split over several GPUs (the cross-thread arg-min of ParallelMonteCarloSearch :694-705 at GPU granularity):
  // after attaching, Update() evaluates this rank's slice of the candidates and the arg-min is exchanged inside the search
  // kernels over peer memory.  One process driving several GPUs: AttachGroup(rank, {all processors of the group}); one
  // process per GPU: ExportGroupHandle() everywhere, exchange the 64-byte handles, AttachGroup(rank, handles).
  cs_ipc_handle ExportGroupHandle() { cs_ipc_handle out; check(cs_group_export(h_, &out)); return out; }
  void AttachGroup(int rank, const std::vector<cs_ipc_handle>& handles) { check(cs_group_attach(h_, rank, (int)handles.size(), handles.data())); }
  void AttachGroup(int rank, const std::vector<CoreSLAMProcessor*>& group) {
    std::vector<cs_processor*> hs;
    for (CoreSLAMProcessor* p : group) hs.push_back(p ? p->h_ : nullptr);
    check(cs_group_attach_local(h_, rank, (int)hs.size(), hs.data()));
  }
  void DetachGroup() { check(cs_group_detach(h_)); }

 private:
  void check(cs_status st) const {
    if (st != CS_OK) throw std::runtime_error(cs_last_error(h_));
  }
  cs_processor* h_ = nullptr;
  HoleMap map_;
  Vector3 pose_;
  std::vector<float> rays_, segPoses_;  // flattened List<ScanSegment> of the current Update
  std::vector<int32_t> segFirst_;
  cs_result last_{};
  int quality_ = 50, psb_ = 5;
  float holeWidth_ = 0.6f;
};

inline std::vector<uint16_t> HoleMap::Pixels() const {
  std::vector<uint16_t> px((size_t)size_ * size_);
  if (cs_map_download(h_, px.data()) != CS_OK) throw std::runtime_error(cs_last_error(h_));
  return px;
}

inline std::vector<uint8_t> HoleMap::GetPackedPixels() const {
  std::vector<uint8_t> out((size_t)size_ * size_ / 2);
  if (cs_map_packed(h_, out.data()) != CS_OK) throw std::runtime_error(cs_last_error(h_));
  return out;
}

}  // namespace CoreSLAM
