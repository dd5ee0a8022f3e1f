// dropin/orb_features/orb_extractor.h -- drop-in for the reference's src/orb_features/orb_extractor.h (:25-93).
// Put this directory before the reference's src/ on the include path: "orb_features/orb_extractor.h" (as src/data/frame.h:9
// includes it) then resolves here, and every reference file that uses the extractor -- Frame, Tracker, SlamSystem -- compiles
// unchanged against the B200 implementation.  class ORBextractor (same constructor, Compute(image, mask, keypoints,
// descriptors) + operator(), scale getters, GetImagePyramid) lives in include/orbfe_shim.hpp over the C ABI of include/orbfe.h.
// The reference's own orb_extractor.cpp is not compiled.
#ifndef ORB_EXTRACTOR_H_
#define ORB_EXTRACTOR_H_

#include <vector>
#include <list>
#include <opencv/cv.h>

#include "orbfe_shim.hpp"

#endif  // ORB_EXTRACTOR_H_
