// oracle/shim/cv/pcl/point_types.h -- TEST INFRASTRUCTURE ONLY.
// cvo::point_cloud carries two pcl clouds for visualisation / PCD export (include/data_type.h:68-69); nothing on the
// registration path reads them.  Empty stand-ins so that the reference's headers parse.
#ifndef CVO_ORACLE_SHIM_PCL_H
#define CVO_ORACLE_SHIM_PCL_H
#include <string>
#include <vector>
namespace pcl {
struct PointXYZRGBA {
    float x, y, z;
    unsigned char r, g, b, a;
};
struct PointXYZRGB {
    float x, y, z;
    unsigned char r, g, b;
};
template <class P> class PointCloud {
  public:
    std::vector<P> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    void push_back(const P& p) { points.push_back(p); }
    std::size_t size() const { return points.size(); }
    void clear() { points.clear(); }
};
namespace io {
template <class P> int savePCDFileASCII(const std::string&, const PointCloud<P>&) { return 0; }
template <class P> int savePCDFile(const std::string&, const PointCloud<P>&) { return 0; }
}  // namespace io
}  // namespace pcl
#endif
