#pragma once
namespace pcl { struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, pad = 1.0f; }; }
