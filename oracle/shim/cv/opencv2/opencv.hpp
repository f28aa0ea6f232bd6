// oracle/shim/cv/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY.
// The reference's hot path never touches an image: cv::Mat only appears in the signatures of set_pcd()/run_cvo()
// (src/cvo.cpp:319,422) and as members of cvo::frame (include/data_type.h:46-49), and is handed through to
// pcd_generator.  This stand-in carries an opaque payload pointer instead of pixels: the shim driver
// (oracle/refsrc_driver.cpp) passes the prepared point cloud of a test case through it, and its stand-in for
// pcd_generator::create_pointcloud unpacks it.  src/pcd_generator.cpp (the real image front end, which needs the real
// OpenCV) is NOT compiled.
#ifndef CVO_ORACLE_SHIM_OPENCV_HPP
#define CVO_ORACLE_SHIM_OPENCV_HPP
namespace cv {
class Mat {
  public:
    const void* payload = nullptr;
    int rows = 0, cols = 0;
    Mat() {}
    explicit Mat(const void* p) : payload(p) {}
};
}  // namespace cv
#endif
