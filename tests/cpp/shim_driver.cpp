// Test driver for icpslam_b200/shims/b2icp_shims.hpp: replays raw float32 {x,y,z,w} cloud files through
// the IcpOdometer / OctreeMapper shims and prints one JSON line per scan.
//   shim_driver odom  leaf  a.bin b.bin ...      scan-to-scan odometry (IcpOdometer::laserCloudCallback)
//   shim_driver map   res   a.bin b.bin ...      + OctreeMapper::refineTransformAndGrowMap per accepted scan
//   shim_driver pose                              Pose6DOF algebra self-check (no GPU needed)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../icpslam_b200/shims/b2icp_ros_adapter.hpp"

static b2::Cloud::Ptr load(const char* path) {
  b2::Cloud::Ptr c(new b2::Cloud());
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  c->points.resize((size_t)bytes / sizeof(b2::PointXYZ));
  if (fread(c->points.data(), sizeof(b2::PointXYZ), c->points.size(), f) != c->points.size()) exit(2);
  fclose(f);
  return c;
}

static void print_pose(const char* key, const b2::Pose6DOF& p) {
  printf("\"%s\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g]", key, p.pos[0], p.pos[1], p.pos[2], p.rot[0], p.rot[1],
         p.rot[2], p.rot[3]);
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  if (mode == "pose") {
    const double T[16] = {0, -1, 0, 1, 1, 0, 0, 2, 0, 0, 1, 3, 0, 0, 0, 1};  // 90 deg yaw + translation
    b2::Pose6DOF a(T, 0.0), b(T, 1.0);
    b2::Pose6DOF c = a + b, d = c + c.inverse();
    printf("{"); print_pose("a", a); printf(", "); print_pose("ab", c); printf(", "); print_pose("ident", d); printf("}\n");
    return 0;
  }
  if (mode == "pc2") {  // fromROSMsg on a raw payload: pc2 <width> <point_step> <off_x> <off_y> <off_z> <payload file>
    if (argc < 8) return 2;
    try {
      b2::IcpOdometerParams op;
      op.mode = B2ICP_MODE_P2P_SVD;
      b2::IcpOdometer odo(op);
      FILE* f = fopen(argv[7], "rb");
      if (!f) return 2;
      std::vector<uint8_t> payload;
      uint8_t buf[65536];
      for (size_t k; (k = fread(buf, 1, sizeof(buf), f)) > 0;) payload.insert(payload.end(), buf, buf + k);
      fclose(f);
      b2::PointCloud2View v;
      v.data = payload.data();
      v.data_bytes = payload.size();
      v.width = (uint32_t)atoi(argv[2]);
      v.height = 1;
      v.point_step = (uint32_t)atoi(argv[3]);
      v.row_step = v.width * v.point_step;
      v.off_x = (uint32_t)atoi(argv[4]);
      v.off_y = (uint32_t)atoi(argv[5]);
      v.off_z = (uint32_t)atoi(argv[6]);
      b2::Cloud c;
      const int rc = b2::fromROSMsg(odo.engine(), v, c);
      printf("{\"status\": %d, \"xyzw\": [", rc);
      auto num = [](float v) {  // JSON has no nan: Python's json reads NaN
        if (v != v) printf("NaN"); else printf("%.9g", v);
      };
      for (size_t i = 0; i < c.size(); ++i) {
        if (i) printf(", ");
        num(c.points[i].x); printf(", "); num(c.points[i].y); printf(", "); num(c.points[i].z); printf(", "); num(c.points[i].w);
      }
      printf("]}\n");
    } catch (const std::exception& e) {
      fprintf(stderr, "shim_driver: %s\n", e.what());
      return 3;
    }
    return 0;
  }
  if (argc < 4) return 2;
  try {
    // B2_SHIM_MODE=p2p selects the point-to-point pipeline; the shims' own default is the reference's GICP
    const char* em = getenv("B2_SHIM_MODE");
    const bool p2p = em && std::string(em) == "p2p";
    b2::IcpOdometerParams op;
    op.voxel_leaf_size = mode == "odom" ? atof(argv[2]) : 0.0;
    if (p2p) op.mode = B2ICP_MODE_P2P_SVD;
    b2::IcpOdometer odo(op);
    b2::OctreeMapperParams mp;
    if (p2p) mp.mode = B2ICP_MODE_P2P_SVD;
    mp.pcl_octree = getenv("B2_SHIM_OCTREE") != nullptr;
    if (mode == "map") mp.octree_resolution = atof(argv[2]);
    std::unique_ptr<b2::OctreeMapper> mapper;
    if (mode == "map") mapper.reset(new b2::OctreeMapper(mp));
    odo.setInitialPose(b2::Pose6DOF());
    b2::Pose6DOF map_pose;  // pose of the latest accepted scan in the map frame (icpslam.cpp:136-140)
    for (int i = 3; i < argc; ++i) {
      b2::Cloud::Ptr cloud = load(argv[i]);
      odo.laserCloudCallback((double)(i - 3), cloud);
      printf("{\"scan\": %d, \"status\": %d, \"ready\": %d, \"iterations\": %d, \"converged\": %d, \"fitness\": %.17g, \"T\": [",
             i - 3, odo.last_status, (int)odo.isOdomReady(), odo.last_result.iterations, odo.last_result.converged,
             i == 3 ? 0.0 : odo.last_fitness);
      for (int k = 0; k < 16; ++k) printf("%s%.17g", k ? ", " : "", i == 3 ? (k % 5 == 0 ? 1.0 : 0.0) : odo.last_result.T[k]);
      printf("], ");
      print_pose("pose", odo.getLatestPose());
      if (mapper) {
        double stamp; b2::Cloud::Ptr c; b2::Pose6DOF tr, pose; bool fresh;
        odo.getEstimates(stamp, c, tr, pose, fresh);
        b2::Pose6DOF raw = map_pose + tr, refine;
        bool ok = mapper->refineTransformAndGrowMap(stamp, c, raw, refine);
        map_pose = ok ? raw + refine : raw;
        printf(", \"refined\": %d, \"map_points\": %zu, ", (int)ok, mapper->mapCloud()->size());
        print_pose("map_pose", map_pose);
      }
      printf("}\n");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_driver: %s\n", e.what());
    return 3;
  }
  return 0;
}
