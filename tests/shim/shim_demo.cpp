// shim_demo.cpp -- exercises include/orbfe_shim.hpp the way the reference's Frame does
// (src/data/frame.cpp:61-111): two ORBextractor objects run Compute on two std::threads, then
// ComputeStereoMatches.  Reads two raw u8 images, writes keypoints / descriptors / stereo results as raw
// binary so that the Python test can compare them with the oracle.  usage:
//   shim_demo W H left.raw right.raw out_prefix      |      shim_demo --link-check
#include "orbfe_shim.hpp"

#include <cstdio>
#include <cstdlib>
#include <thread>

static std::vector<uint8_t> slurp(const char* path, size_t n) {
  std::vector<uint8_t> v(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}
template <class T>
static void dump(const std::string& path, const T* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(p, sizeof(T), n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc == 2 && std::string(argv[1]) == "--link-check") {
    printf("%s devices=%d\n", orbfe_version(), orbfe_device_count());
    return 0;
  }
  if (argc != 6) return 1;
  const int W = atoi(argv[1]), H = atoi(argv[2]);
  std::vector<uint8_t> l = slurp(argv[3], (size_t)W * H), r = slurp(argv[4], (size_t)W * H);
  cv::Mat left(H, W, CV_8UC1, l.data()), right(H, W, CV_8UC1, r.data());
  ORBextractor exL(2000, 1.2f, 8, 20, 7), exR(2000, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> kl, kr;
  cv::Mat dl, dr;
  std::thread tl([&]() { exL.Compute(left, cv::Mat(), kl, dl); });
  std::thread tr([&]() { exR(right, cv::Mat(), kr, dr); });
  tl.join();
  tr.join();
  std::vector<float> ur, depth;
  const float bf = 386.1448f, fx = 718.856f;
  orbfe::ComputeStereoMatches(exL, exR, kl, kr, dl, dr, bf, bf / fx, ur, depth);
  const std::string out(argv[5]);
  dump(out + ".kl", kl.data(), kl.size());
  dump(out + ".kr", kr.data(), kr.size());
  dump(out + ".dl", dl.data, (size_t)dl.rows * 32);
  dump(out + ".dr", dr.data, (size_t)dr.rows * 32);
  dump(out + ".ur", ur.data(), ur.size());
  dump(out + ".depth", depth.data(), depth.size());
  const std::vector<cv::Mat>& pyr = exL.GetImagePyramid();
  dump(out + ".pyr3", pyr[3].data, (size_t)pyr[3].rows * pyr[3].cols);
  printf("%zu %zu %d %d\n", kl.size(), kr.size(), pyr[3].cols, pyr[3].rows);
  return 0;
}
