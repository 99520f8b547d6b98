// TEST INFRASTRUCTURE ONLY.  The reference's own map writer -- PointCloudMap<PointT>::WritePLY and writeCustomData, sliced out of
// /root/reference/src/PointCloudMap.cc:304-437 at build time (oracle/ref_build.py::build_mapply) -- compiled against a stand-in for the PCL point type it
// writes (pcl::PointSurfelSegment's members as WritePLY touches them: x y z, PCL's {b, g, r, a} colour union, normal_x.., label, kfid) and the two
// compile-time switches of include/PointDefinitions.h:28-29.  Nothing of the reference's text is stored in the repository.
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define USE_NORMALS 1
#define USE_POINTSURFELSEGMENT 1

namespace Eigen { template <class T> using aligned_allocator = std::allocator<T>; }
namespace pcl {
namespace fields { struct kfid {}; }
namespace traits { template <class P, class F> struct has_field : std::true_type {}; }
}

struct PointStandIn {
    float x, y, z, pad0;
    float normal_x, normal_y, normal_z, pad1;
    union { struct { uint8_t b, g, r, a; }; float rgb; uint32_t rgba; };      // PCL_ADD_RGB
    uint32_t label, kfid;
};
template <class P> struct CloudStandIn { std::vector<P, Eigen::aligned_allocator<P> > points; };

namespace PLVS2 {
template <typename PointT>
class PointCloudMap {
public:
    typedef CloudStandIn<PointT> PointCloudT;
    bool WritePLY(PointCloudT& cloud, std::string filename, bool isMesh = true, bool binary = true);
};
#include "mapply_slices.inc"
}  // namespace PLVS2

extern "C" int ref_write_map_ply(const char* path, const float* xyz, const uint8_t* bgra, const float* normals, const uint32_t* label, const uint32_t* kfid, long long n,
                                 int is_mesh, int binary)
{
    CloudStandIn<PointStandIn> cloud;
    cloud.points.resize((size_t)n);
    for (long long i = 0; i < n; ++i) {
        PointStandIn& p = cloud.points[(size_t)i];
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.normal_x = normals[3 * i]; p.normal_y = normals[3 * i + 1]; p.normal_z = normals[3 * i + 2];
        p.b = bgra[4 * i]; p.g = bgra[4 * i + 1]; p.r = bgra[4 * i + 2]; p.a = bgra[4 * i + 3];
        p.label = label[i]; p.kfid = kfid[i];
    }
    PLVS2::PointCloudMap<PointStandIn> m;
    return m.WritePLY(cloud, path, is_mesh != 0, binary != 0) ? 0 : -1;
}
