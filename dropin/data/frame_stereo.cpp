// dropin/data/frame_stereo.cpp -- drop-in definition of Frame::ComputeStereoMatches (declared at the reference's
// src/data/frame.h:95, reference body src/data/frame.cpp:406-577).  Everything else of Frame stays the reference's: delete (or
// weaken, see dropin/Makefile) that one function in frame.cpp and add this file to the build.
//
// The row-band Hamming search, the 11x11 SAD refinement on the un-blurred pyramid levels, the parabola fit and the median
// cut run on the GPU (csrc/k_stereo.cuh) against the pyramids the two extractors left on the device when they processed this
// frame's images; only keypoints, descriptors and the two result vectors cross the boundary.
// baseline_: the reference reads it before the constructor assigns it (frame.cpp:436 vs :108); the value it holds from the
// second frame of a run on is baseline_fx_ / fx_, which is what is passed here (oracle-defined behaviour 2).
#include "data/frame.h"

void Frame::ComputeStereoMatches() {
  orbfe::ComputeStereoMatches(*left_orb_extractor_, *right_orb_extractor_, keypoints_, right_keypoints_, descriptors_,
                              right_descriptors_, baseline_fx_, baseline_fx_ / fx_, stereo_coords_, depths_);
}
