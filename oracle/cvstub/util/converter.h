// oracle/cvstub/util/converter.h -- TEST INFRASTRUCTURE ONLY.  Shadows the reference's util/converter.h (which pulls in Eigen
// and g2o, absent from this image) when oracle/Makefile.ref compiles src/data/*.cpp: those files use only
// Converter::toDescriptorVector (frame.cpp:260, keyframe.cpp:129), declared here and defined in oracle/ref_wrap.cpp as
// util/converter.cpp:8-17 does (one 1 x 32 row header per descriptor).
#ifndef ORACLE_CVSTUB_CONVERTER_H_
#define ORACLE_CVSTUB_CONVERTER_H_
#include <opencv2/core/core.hpp>
#include <vector>
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};
#endif
